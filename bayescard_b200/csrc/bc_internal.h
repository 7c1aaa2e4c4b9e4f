// Internal declarations shared by the translation units of libbayescard_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bayescard_b200.h"

#define BC_VERSION_STRING "bayescard_b200 0.1.0 (sm_100a)"
#define BC_CODEGEN_VERSION 10
#define BC_SPEC_CTR_SLOTS 256

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

// Second per-node table (kept out of BcNodeRec so that record stays 32 bytes): BITS rows.
struct BcBitsRec {
    int32_t bit_off;     // first bit of this node's state mask in a BITS row
    int32_t card;
};

struct BcHostPipe;  // bc_api.cu
struct BcUmmaPlan;  // k2_umma.cu
struct BcK3Plan;    // k3_fused.cu

// State of the batched large-domain path (K2), built on first use.
struct BcK2Plan {
    std::vector<char> is_internal;     // node has children
    std::vector<double*> d_prefix;     // leaves: fp64 column prefix sums of T_v, (card+1) x card_pa
    BcUmmaPlan* umma = nullptr;        // tensor-core operands (transposed hi / lo CPTs, tensor maps)
    // Lambda workspace, kept between calls (grow only): calls on one model are serialised on ws_event, so a call
    // on another stream cannot overwrite the workspace of one that is still running
    float* d_ws = nullptr;
    size_t ws_bytes = 0;
    cudaEvent_t ws_event = nullptr;
};

struct bc_model {
    int device = -1;     // -1: host-only model (code generation / ahead-of-time build)
    int n = 0;
    std::vector<BcNodeRec> nodes;
    std::vector<float> arena;        // host copy (code generator input)
    std::vector<float> fan;
    std::vector<uint16_t> ent_node;  // node id of every entry of a lambda row
    std::vector<BcBitsRec> bits;     // BITS descriptor geometry
    int bits_words = 0;              // 32-bit words per BITS row (multiple of 4)
    int lam_total = 0;               // floats per lambda / dense row
    int max_card = 0;
    int mask_words = 1;
    int64_t flops_dense = 0;
    // device side
    float* d_arena = nullptr;
    size_t arena_floats_padded = 0;
    float* d_fan = nullptr;
    BcNodeRec* d_nodes = nullptr;
    BcBitsRec* d_bits = nullptr;
    uint32_t* d_bits_default = nullptr;  // BITS row of the unconstrained query (every state selected)
    float* d_dense_default = nullptr;    // DENSE_F32 row of the unconstrained query (1 on every state, 0 on row padding)
    uint16_t* d_ent_node = nullptr;
    int sm_count = 0;
    int smem_optin = 0;
    // specialised kernel
    cudaLibrary_t spec_lib = nullptr;
    cudaKernel_t spec_range8 = nullptr;
    cudaKernel_t spec_dense = nullptr;
    cudaKernel_t spec_bits = nullptr;
    int spec_threads = 0;            // threads per CTA (from the image's bc_spec_meta)
    int spec_qpt = 1;                // queries per thread per loop trip
    int spec_blocks_bits = 1, spec_blocks_dense = 1, spec_blocks_range8 = 1;  // resident CTAs per SM
    std::vector<uint32_t> bits_default;
    // work counters of the dynamically scheduled specialised kernels: a ring of {u64 next query, u32 CTAs done, pad}
    // slots (one per launch, so launches of one model on different streams never share a counter); each kernel
    // leaves its slot zeroed
    uint64_t* d_spec_ctr = nullptr;
    std::atomic<uint32_t> spec_ctr_next{0};
    BcK2Plan* k2 = nullptr;
    std::mutex k2_mu;                // guards the lazy construction of k2 (pipe_mu may already be held)
    BcK3Plan* k3 = nullptr;          // fused tensor-core tree kernel: edge schedule, TMEM columns, operand images
    std::mutex k3_mu;
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
                 float* out, cudaStream_t stream, int32_t* out_exp = nullptr);
// bc_convert.cu
int bc_convert_launch(bc_model* m, const void* src, int src_fmt, void* dst, int dst_fmt, size_t nq, cudaStream_t stream);
int bc_expand_sparse_launch(bc_model* m, const uint32_t* row_off, const uint32_t* entries, size_t nq, void* dst_bits,
                            cudaStream_t stream);
int bc_expand_wsparse_launch(bc_model* m, const uint32_t* row_off, const uint32_t* words, size_t nq, float* dst_dense,
                             cudaStream_t stream);
#define BC_PACKED_BLOCK 128   // queries per blk_off entry of the PACKED wire format
void bc_packed_geometry(const bc_model* m, int* col_bits, int* state_bits);
// payload points at 32-bit word `word_base` of the bit stream (a slice of a larger batch keeps absolute entry indices)
int bc_expand_packed_launch(bc_model* m, const uint8_t* klen, const uint32_t* blk_off, const uint32_t* payload, unsigned long long word_base,
                            size_t nq, void* dst_bits, cudaStream_t stream);
// k2_batched.cu / k2_umma.cu
int bc_k2_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask, float* out, int use_umma,
                 cudaStream_t stream, int32_t* out_exp = nullptr);
void bc_k2_free(bc_model* m);
// one internal edge on the tensor cores; BC_ELIMIT = shape not served (caller falls back to FP32 SIMT)
int bc_k2_umma_edge(bc_model* m, const uint8_t* desc, size_t dstride, int fmt, size_t q0, int rows, int v, float* lam_v,
                    float* lam_v_lo, int ld_v, float* lam_pa, int ld_pa, int accumulate, cudaStream_t stream);
void bc_k2_umma_free(bc_model* m);
// k3_fused.cu: whole tree per 128-query tile on the tensor cores (BITS / DENSE_F32 rows); BC_ELIMIT = model not served
int bc_k3_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask, float* out, cudaStream_t stream,
                 int32_t* out_exp = nullptr);   // out_exp: scaled results (mantissa in out, exponent of two per query)
void bc_k3_free(bc_model* m);
// spec_codegen.cc
std::string bc_spec_generate(const bc_model& m);
uint64_t bc_spec_hash_of(const bc_model& m);
// spec_jit.cu
int bc_spec_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask,
                   float* out, cudaStream_t stream);
int bc_spec_attach(bc_model* m, const void* image, size_t bytes);
int bc_spec_build(bc_model* m, const char* cache_dir);
