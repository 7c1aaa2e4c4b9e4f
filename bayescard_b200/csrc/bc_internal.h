// Internal declarations shared by the translation units of libbayescard_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bayescard_b200.h"

#define BC_VERSION_STRING "bayescard_b200 0.1.0 (sm_100a)"
#define BC_CODEGEN_VERSION 7

// One record per node, copied to the device and (by K1) into shared memory.  32 bytes.
struct BcNodeRec {
    int32_t parent;      // topological index of the parent, -1 for the root
    int32_t card;        // number of states
    int32_t stride;      // row stride of T_v in floats
    int32_t lam_off;     // offset of this node's segment in a dense weight / lambda row (floats, multiple of 4)
    int64_t cpt_off;     // offset of T_v in the arena (floats)
    int32_t fan_off;     // offset of fanouts[v] in the fan arena (floats) or -1
    int32_t card_pa;     // card of the parent (1 for the root)
};

struct BcHostPipe;  // bc_api.cu

struct bc_model {
    int device = -1;     // -1: host-only model (code generation / ahead-of-time build)
    int n = 0;
    std::vector<BcNodeRec> nodes;
    std::vector<float> arena;        // host copy (code generator input)
    std::vector<float> fan;
    std::vector<uint16_t> ent_node;  // node id of every entry of a lambda row
    int lam_total = 0;               // floats per lambda / dense row
    int max_card = 0;
    int mask_words = 1;
    int64_t flops_dense = 0;
    // device side
    float* d_arena = nullptr;
    size_t arena_floats_padded = 0;
    float* d_fan = nullptr;
    BcNodeRec* d_nodes = nullptr;
    uint16_t* d_ent_node = nullptr;
    int sm_count = 0;
    int smem_optin = 0;
    // specialised kernel
    cudaLibrary_t spec_lib = nullptr;
    cudaKernel_t spec_range8 = nullptr;
    cudaKernel_t spec_dense = nullptr;
    int spec_threads = 0;
    int spec_min_blocks = 0;
    BcHostPipe* pipe = nullptr;
    std::mutex pipe_mu;
};

void bc_set_error(const char* fmt, ...);
void bc_count_launch(uint64_t n = 1);

#define BC_CUDA_CHECK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            bc_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return BC_ECUDA;                                                                     \
        }                                                                                        \
    } while (0)

static inline int64_t bc_round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// k1_generic.cu
int bc_k1_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask,
                 float* out, cudaStream_t stream);
// spec_codegen.cc
std::string bc_spec_generate(const bc_model& m);
uint64_t bc_spec_hash_of(const bc_model& m);
// spec_jit.cu
int bc_spec_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask,
                   float* out, cudaStream_t stream);
int bc_spec_attach(bc_model* m, const void* image, size_t bytes);
int bc_spec_build(bc_model* m, const char* cache_dir);
