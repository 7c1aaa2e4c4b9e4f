// K3 -- fused whole-tree sum-product on the tensor cores: one CTA carries a tile of 128 queries through EVERY edge
// of the tree without leaving the SM.
//
//     out[q] = sum_x prod_v w_v[x_v] * T_v[x_v, x_pa(v)]        (VariableEliminationJIT.query / .expectation,
//                                                                 reference Pgmpy/inference/ExactInference.py:112-287)
//
// evaluated leaf -> root as one small GEMM per edge v -> pa(v):
//
//     D[128 x N]  = U_v[128 x K] . T_v[K x N]        U_v = w_v (*) Lambda_v,  K = card(v), N = card(pa)
//     Lambda_pa  *= D
//
// Why it exists: the per-model straight-line kernel (K-spec) keeps the CPTs in the instruction stream and is
// instruction-fetch bound once a model has more than ~10k CPT entries (IMDB: 45 % of the FP32 peak with unit
// weights, 19 % with fractional weights / fan-out expectations); K2 runs one launch per edge and moves every Lambda
// through HBM.  Here the messages never leave the SM:
//
//   * Lambda_v of every live internal node sits in TENSOR MEMORY (one fp32 column per state, one lane per query);
//     the columns are assigned on the host by first fit over the nodes' lifetimes in the edge schedule;
//   * FOUR PRODUCER WARPS (thread = query = TMEM lane): per edge and per block of 16 child states they read Lambda_v
//     with tcgen05.ld, apply the query's weights (BITS mask, dense n_distinct weights, fan-out vector), split the
//     product into TF32 hi + lo and write both as the K-major, 64-byte-swizzled A operand straight into a 3-slot
//     shared-memory ring (fence.proxy.async + one mbarrier arrive per warp); the message-independent inputs of the
//     NEXT block (weights; finished 0/1 chunks for unit-weight leaves) are fetched one step ahead;
//   * a TMA WARP keeps a 4-slot ring of T_v^T blocks (hi and lo, pre-split and pre-swizzled once per model into an
//     "operand image") full: ONE bulk copy (cp.async.bulk ... mbarrier::complete_tx) per step;
//   * an ISSUER WARP (whole warp converged, elect.sync inside the asm, edge table in the kernel-parameter bank so that
//     every tcgen05 operand lives in uniform registers) issues tcgen05.mma.cta_group::1.kind::tf32, error compensated:
//     A_lo.B_hi + A_hi.B_lo first, then A_hi.B_hi (A_lo = 0 and is skipped for unit-weight leaves), into a TMEM
//     accumulator; tcgen05.commit frees the A and B slots and, after the edge's last block, releases the accumulator;
//   * the producer warps then multiply the accumulator into Lambda_pa with tcgen05.ld / tcgen05.st (32 columns per
//     round trip) -- no shared memory, no HBM traffic; the root is a dot product with T_root in registers;
//   * two CTAs per SM (256 TMEM columns and ~104 KB of shared memory each; one CTA with 512 columns when a model's live
//     messages need more) overlap one tile's operand building, MMA drain and epilogue with the other's MMAs.
//
// Algorithmic work per query = flops_dense(model) (every CPT entry once); HBM traffic = descriptor row + 4 B.
#include <algorithm>
#include <cstring>
#include <memory>

#include "bc_internal.h"

struct K3Edge {            // 64 bytes, copied to shared memory
    int32_t v, K, N, n_pad;
    int32_t lam_off;       // first float of column v in a DENSE row
    int32_t bit_off;       // first bit of column v in a BITS row
    int32_t fan_off;       // fanouts[v] in the fan arena, -1 if none
    int32_t col_v;         // TMEM column of Lambda_v, -1 for a leaf
    int32_t col_pa;        // TMEM column of Lambda_pa
    int32_t first;         // this edge is the first message into Lambda_pa
    int32_t nkb;           // blocks of 16 child states
    uint32_t idesc;        // tcgen05 instruction descriptor (M = 128, N = n_pad, TF32 x TF32 -> F32, K-major)
    uint64_t bimg_off;     // byte offset of the edge's operand images
    int32_t pad[2];
};
static_assert(sizeof(K3Edge) == 64, "K3Edge layout");

struct BcK3Plan {
    int failed = 0;
    std::vector<K3Edge> edges;
    uint8_t* d_bimg = nullptr;
    size_t bimg_bytes = 0;
    int npad_max = 16;
    int tmem_cols = 32;    // power of two
    int d_col = 0, n_dbuf = 1;
    int root_col = 0;
    int ctas_per_sm = 1;
    size_t smem = 0;
};

namespace {

constexpr int kTile = 128;    // queries per CTA tile = TMEM lanes = UMMA M
constexpr int kMaxEdges = 127; // trees of up to 128 columns
constexpr int kBK = 16;       // child states per ring step: one 64-byte swizzle row
constexpr int kStagesA = 3;   // A ring: U_v blocks written by the producer warps
constexpr int kStagesB = 4;   // B ring: T_v^T blocks fetched by TMA (deeper: an L2 round trip is longer than a step)
constexpr int kABytes = kTile * kBK * 4;   // 8 KB per half (hi or lo)
constexpr int kProducerWarps = 4;
constexpr int kThreads = 32 * (kProducerWarps + 2);   // + MMA issuer warp + TMA warp

struct K3Params {
    K3Edge edge[kMaxEdges];    // in the kernel parameter bank (8 KB of the 32 KB sm_100 allows): uniform loads, warp-uniform control flow
    int n_edges;
    const uint8_t* bimg;
    const uint8_t* desc;
    size_t dstride;
    const uint32_t* fan_mask;
    const float* fan;
    int fan_floats;            // multiple of 4 (shared-memory copy, zero padded)
    int fan_n;                 // floats in the fan arena
    const float* root_T;       // T_root in the arena
    int root_card, root_col, root_bit_off, root_lam_off, root_fan_off, root_has_children;
    float* out;
    size_t nq;
    long long n_tiles;
    int bits_words;
    int b_slot_bytes;          // 2 * npad_max * 64
    int d_col, d_stride, n_dbuf;
    int tmem_cols;
    int mask_words;            // fan-out mask words per query
    float debias_unit;         // expected relative truncation loss per accumulating MMA (1.1e-8 measured; BC_K3_DEBIAS overrides, 0 = off)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// K-major operands, 64-byte swizzle: rows of 64 B, 8-row groups 512 B apart (SBO), layout type 4, version 1 (sm_100).
// descriptors as 32-bit halves: the high word (SBO, version, swizzle mode) is the same for every operand.  Called by the
// WHOLE (converged) issuer warp; elect.sync inside picks the lane, so ptxas keeps every operand in uniform registers
// instead of wrapping each instruction in a divergence (ELECT / R2UR / BRA.U.ANY) loop.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint32_t a_lo32, uint32_t b_lo32, uint32_t desc_hi32, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p, q;\n.reg .b64 da, db;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %5, 0;\n"
        "mov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n}\n" ::"r"(tmem_c),
        "r"(a_lo32), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {   // whole warp, one elected lane
    asm volatile(
        "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// x = hi + lo, hi = x rounded to the nearest TF32 number (ties away from zero: two integer instructions); lo = x - hi is
// exact in fp32 and symmetric around zero, so the tensor core's truncation of its low bits is unbiased
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}

// selection bits of states [c0, c0 + 16) of a column for this thread's query (bits past the domain cleared)
__device__ __forceinline__ uint32_t bits16(const uint32_t* my_bits, int bits_words, int bit_off, int card, int c0) {
    const int b0 = bit_off + c0, idx = b0 >> 5, sh = b0 & 31;
    const uint32_t w0 = my_bits[idx * kTile];
    const uint32_t w1 = my_bits[(idx + 1 < bits_words ? idx + 1 : idx) * kTile];
    const int valid = card - c0;
    return __funnelshift_r(w0, w1, sh) & (valid >= 16 ? 0xFFFFu : ((1u << valid) - 1u));
}

// weights of 8 consecutive states [c0, c0 + 8) of a column (root only: the edges have their own code below)
template <int FMT>
__device__ __forceinline__ void load_weights8(const uint32_t* my_bits, int bits_words, const float* drow, const float* s_fan, uint32_t fm,
                                              int v, int lam_off, int bit_off, int fan_off, int card, int c0, float* w) {
    if (FMT == BC_DESC_BITS) {
        const uint32_t m = bits16(my_bits, bits_words, bit_off, card, c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = ((m >> j) & 1u) ? 1.f : 0.f;
    } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(drow + lam_off + c0));
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + 4 < card) b = __ldg(reinterpret_cast<const float4*>(drow + lam_off + c0 + 4));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c0 + j >= card) w[j] = 0.f;   // padding entries of a DENSE row never contribute
    }
    if (fan_off >= 0 && ((fm >> v) & 1u)) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c0 + j < card) w[j] *= s_fan[fan_off + c0 + j];
    }
}

template <int FMT>
__global__ void __launch_bounds__(kThreads, 2) k3_kernel(const __grid_constant__ K3Params P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (role dispatch without divergence)
    // carve-up: A ring | B ring | BITS rows of the tile [word][query] | fan arena | nibble table | barriers
    uint8_t* p = smem + (size_t)kStagesA * 2 * kABytes;
    const uint32_t a_ring = smem_u32(smem), b_ring = smem_u32(p);
    p += (size_t)kStagesB * P.b_slot_bytes;
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(p);
    p += (FMT == BC_DESC_BITS ? (size_t)P.bits_words * kTile * 4 : 0);
    float* s_fan = reinterpret_cast<float*>(p);
    p += (size_t)P.fan_floats * 4;
    float4* s_tab = reinterpret_cast<float4*>(p);   // nibble -> four 0/1 floats
    p += 256;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p);
    const uint32_t a_full0 = smem_u32(bars), a_empty0 = a_full0 + 8 * kStagesA, b_full0 = a_empty0 + 8 * kStagesA,
                   b_empty0 = b_full0 + 8 * kStagesB, d_full0 = b_empty0 + 8 * kStagesB, d_empty0 = d_full0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStagesA + 2 * kStagesB + 4);

    {   // tables -> shared memory
        for (int i = tid; i < P.fan_floats; i += kThreads) s_fan[i] = i < P.fan_n ? P.fan[i] : 0.f;
        if (tid < 16) s_tab[tid] = make_float4(tid & 1 ? 1.f : 0.f, tid & 2 ? 1.f : 0.f, tid & 4 ? 1.f : 0.f, tid & 8 ? 1.f : 0.f);
    }
    if (tid == 0) {
        for (int s = 0; s < kStagesA; ++s) {
            mbar_init(a_full0 + 8 * s, kProducerWarps);
            mbar_init(a_empty0 + 8 * s, 1);
        }
        for (int s = 0; s < kStagesB; ++s) {
            mbar_init(b_full0 + 8 * s, 1);
            mbar_init(b_empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(d_full0 + 8 * b, 1);
            mbar_init(d_empty0 + 8 * b, kProducerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler: tcgen05 operands live in uniform registers

    if (warp == kProducerWarps + 1) {
        // ================= TMA warp: keeps the B ring full; the (edge, block) sequence repeats for every tile
        if (lane == 0) {
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
                for (int e = 0; e < P.n_edges; ++e) {
                    const K3Edge& E = P.edge[e];
                    const unsigned bytes = (unsigned)E.n_pad * 128u;   // hi + lo, 64 B per row each
                    for (int kb = 0; kb < E.nkb; ++kb, ++it) {
                        const uint32_t s = it % kStagesB, par = (it / kStagesB) & 1u;
                        mbar_wait(b_empty0 + 8 * s, par ^ 1u);
                        mbar_expect_tx(b_full0 + 8 * s, bytes);
                        tma_bulk_g2s(b_ring + s * P.b_slot_bytes, P.bimg + E.bimg_off + (size_t)kb * bytes, bytes, b_full0 + 8 * s);
                    }
                }
        }
    } else if (warp == kProducerWarps) {
        // ================= MMA issuer warp: the whole warp runs the (uniform) loop, one elected lane issues
        uint32_t it = 0, ed = 0;   // ring step, edge counter (D buffer = ed % n_dbuf)
        const uint32_t desc_hi = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);   // SBO, version 1, 64-byte swizzle
        for (long long tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
            for (int e = 0; e < P.n_edges; ++e, ++ed) {
                const K3Edge& E = P.edge[e];
                const bool a_exact = FMT == BC_DESC_BITS && E.col_v < 0 && !(E.fan_off >= 0 && P.fan_mask != nullptr);
                const uint32_t db = P.n_dbuf == 2 ? (ed & 1u) : 0u, dpar = (P.n_dbuf == 2 ? (ed >> 1) : ed) & 1u;
                const uint32_t d = tmem + (uint32_t)(P.d_col + (int)db * P.d_stride);
                const uint32_t idesc = E.idesc, b_lo_off = (uint32_t)E.n_pad * 4u;   // n_pad * 64 B in 16-byte units
                const int K = E.K, nkb = E.nkb;
                mbar_wait(d_empty0 + 8 * db, dpar ^ 1u);   // the epilogue of the edge that used this accumulator is done
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t sa = it % kStagesA, pa = (it / kStagesA) & 1u;
                    const uint32_t sb = it % kStagesB, pb = (it / kStagesB) & 1u;
                    // low words of the four operand descriptors: start address >> 4 | LBO = 1
                    const uint32_t a_hi = (((a_ring + sa * 2 * kABytes) & 0x3FFFFu) >> 4) | (1u << 16), a_lo = a_hi + (kABytes >> 4);
                    const uint32_t b_hi = (((b_ring + sb * P.b_slot_bytes) & 0x3FFFFu) >> 4) | (1u << 16), b_lo = b_hi + b_lo_off;
                    const bool two = K - kb * kBK > 8;
                    mbar_wait(b_full0 + 8 * sb, pb);
                    mbar_wait(a_full0 + 8 * sa, pa);
                    tc_fence_after();
                    {
                        // error-compensated product, the small terms first; 8 TF32 = 32 bytes per k-step: +2 in the address field
                        if (!a_exact) {
                            umma_tf32(d, a_lo, b_hi, desc_hi, idesc, kb != 0);
                            umma_tf32(d, a_hi, b_lo, desc_hi, idesc, 1);
                            if (two) {
                                umma_tf32(d, a_lo + 2, b_hi + 2, desc_hi, idesc, 1);
                                umma_tf32(d, a_hi + 2, b_lo + 2, desc_hi, idesc, 1);
                            }
                        } else {
                            umma_tf32(d, a_hi, b_lo, desc_hi, idesc, kb != 0);
                            if (two) umma_tf32(d, a_hi + 2, b_lo + 2, desc_hi, idesc, 1);
                        }
                        umma_tf32(d, a_hi, b_hi, desc_hi, idesc, 1);
                        if (two) umma_tf32(d, a_hi + 2, b_hi + 2, desc_hi, idesc, 1);
                        umma_commit(a_empty0 + 8 * sa);
                        umma_commit(b_empty0 + 8 * sb);
                        if (kb == nkb - 1) umma_commit(d_full0 + 8 * db);
                    }
                }
            }
    } else {
        // ================= producer / epilogue warps: thread = query = TMEM lane
        const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
        const uint32_t sw = ((uint32_t)tid >> 1) & 3u;    // 64-byte swizzle: chunk j of row r lives at r * 64 + ((j ^ ((r >> 1) & 3)) << 4)
        const uint32_t* my_bits = s_bits + tid;
        uint32_t it = 0, ed = 0;
        for (long long tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
            const size_t q = (size_t)tile * kTile + tid;
            const size_t qc = q < P.nq ? q : P.nq - 1;
            // fan-out mask: one word per 32 columns; word 0 (the root's) is kept, the others are read per fan-out edge
            const uint32_t* fm_row = P.fan_mask ? P.fan_mask + qc * (size_t)P.mask_words : nullptr;
            const uint32_t fm = fm_row ? fm_row[0] : 0u;
            const float* drow = reinterpret_cast<const float*>(P.desc + qc * P.dstride);
            if (FMT == BC_DESC_BITS) {
                const uint32_t* grow = reinterpret_cast<const uint32_t*>(P.desc + qc * P.dstride);
                for (int w4 = 0; w4 < P.bits_words; w4 += 4) {   // bits_words is a multiple of 4
                    const uint4 x = __ldg(reinterpret_cast<const uint4*>(grow + w4));
                    s_bits[(w4 + 0) * kTile + tid] = x.x;
                    s_bits[(w4 + 1) * kTile + tid] = x.y;
                    s_bits[(w4 + 2) * kTile + tid] = x.z;
                    s_bits[(w4 + 3) * kTile + tid] = x.w;
                }
            }
            // Inputs of ring step (e, kb) that do not depend on any message: for unit-weight leaves the finished hi chunks
            // (0/1 floats from the nibble table), otherwise the weights w_v[c0 .. c0 + 16) (x fan-out).  They are fetched ONE
            // STEP AHEAD (also across edges), so the LDS / LDG latency overlaps the previous block's stores and barriers.
            auto fetch = [&](int e, int kb, float* pre) __attribute__((always_inline)) {
                const K3Edge& E = P.edge[e];
                const int c0 = kb * kBK;
                const bool exact = FMT == BC_DESC_BITS && E.col_v < 0 && !(E.fan_off >= 0 && P.fan_mask != nullptr);
                if (FMT == BC_DESC_BITS) {
                    const uint32_t m = bits16(my_bits, P.bits_words, E.bit_off, E.K, c0);
                    if (exact) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 t = s_tab[(m >> (4 * j)) & 15u];
                            pre[4 * j] = t.x; pre[4 * j + 1] = t.y; pre[4 * j + 2] = t.z; pre[4 * j + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pre[j] = ((m >> j) & 1u) ? 1.f : 0.f;
                    }
                } else {
                    const float4* src = reinterpret_cast<const float4*>(drow + E.lam_off + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c0 + 4 * j < E.K) t = __ldg(src + j);
                        pre[4 * j] = t.x; pre[4 * j + 1] = t.y; pre[4 * j + 2] = t.z; pre[4 * j + 3] = t.w;
                    }
                }
                // (the fan-out vector is applied where the block is consumed: a multiply here would wait for the loads just
                //  issued and undo the prefetch -- ncu: 22 % of the stall samples of the expectation rows sat on it)
            };
            // (two blocks in flight for DENSE rows were measured too: +3 % without fan-out columns, -4...-9 % on the fan-out
            //  weighted expectation factors this kernel exists for -- one step of lead it is; tools/runs/_gpu_run59.sh)
            float pre[16];
            fetch(0, 0, pre);
            for (int e = 0; e < P.n_edges; ++e, ++ed) {
                const K3Edge& E = P.edge[e];
                const bool leaf = E.col_v < 0;
                const bool a_exact = FMT == BC_DESC_BITS && leaf && !(E.fan_off >= 0 && P.fan_mask != nullptr);
                const bool fan_on = E.fan_off >= 0 && fm_row != nullptr && ((fm_row[E.v >> 5] >> (E.v & 31)) & 1u);
                const int K = E.K, nkb = E.nkb;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t sa = it % kStagesA, pa = (it / kStagesA) & 1u;
                    const uint32_t row = a_ring + sa * 2 * kABytes + (uint32_t)tid * 64u;
                    const int c0 = kb * kBK;
                    const int ks = (K - c0 > 8) ? 2 : 1;
                    float u[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) u[j] = pre[j];
                    // next step's inputs
                    if (kb + 1 < nkb) fetch(e, kb + 1, pre);
                    else if (e + 1 < P.n_edges) fetch(e + 1, 0, pre);
                    if (fan_on) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (c0 + 4 * j < K) {
                                const float4 f = *reinterpret_cast<const float4*>(s_fan + E.fan_off + c0 + 4 * j);
                                u[4 * j] *= f.x; u[4 * j + 1] *= f.y; u[4 * j + 2] *= f.z; u[4 * j + 3] *= f.w;
                            }
                    }
                    if (a_exact) {
                        // unit weights on a leaf: U is a 0/1 matrix, exact in TF32 -- no lo half
                        mbar_wait(a_empty0 + 8 * sa, pa ^ 1u);   // the MMAs that read this slot are done
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < 2 * ks) sts128(row + (((uint32_t)j ^ sw) << 4), u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
                    } else {
                        if (!leaf) {
                            float lv[16];
                            tmem_ld8(tlane + (uint32_t)(E.col_v + c0), lv);
                            if (ks == 2) tmem_ld8(tlane + (uint32_t)(E.col_v + c0 + 8), lv + 8);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) u[j] *= lv[j];
                        }
                        if (c0 + kBK > K) {   // last block: states >= K must be exact zeros (stale Lambda columns, row padding)
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j >= K) u[j] = 0.f;
                        }
                        float h[16], l[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) split_tf32(u[j], h[j], l[j]);
                        mbar_wait(a_empty0 + 8 * sa, pa ^ 1u);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < 2 * ks) {
                                const uint32_t a = row + (((uint32_t)j ^ sw) << 4);
                                sts128(a, h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
                                sts128(a + kABytes, l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
                            }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full0 + 8 * sa);
                }
                // ---- epilogue: Lambda_pa (*)= D, all in tensor memory, 32 columns per round trip
                const uint32_t db = P.n_dbuf == 2 ? (ed & 1u) : 0u, dpar = (P.n_dbuf == 2 ? (ed >> 1) : ed) & 1u;
                const uint32_t dcol = tlane + (uint32_t)(P.d_col + (int)db * P.d_stride), pcol = tlane + (uint32_t)E.col_pa;
                const bool first = E.first;
                // The tensor core adds into its fp32 accumulator with truncation (k2_umma.cu: UmmaCfg): every accumulating
                // tcgen05.mma of the edge loses half an ulp of the running sum on average, and the terms are non-negative, so
                // the loss never cancels.  The expected loss is added back here: (instructions - 1) * debias_unit relative;
                // debias_unit = 1.1e-8 is MEASURED (profiles/r1_k3_debias.txt: the mean signed error against the fp64 oracle crosses
                // zero there on all five IMDB models and DMV; small-product and early accumulations lose less than half an ulp).
                const int n_mma = ((int)E.K + 7) / 8 * (a_exact ? 2 : 3);
                const float debias = 1.f + (float)(n_mma - 1) * P.debias_unit;
                mbar_wait(d_full0 + 8 * db, dpar);
                tc_fence_after();
                const int n8 = (E.N + 7) & ~7;
                for (int j = 0; j < n8; j += 32) {
                    float dv[32], lv[32];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (j + 8 * c < n8) {
                            tmem_ld8(dcol + (uint32_t)(j + 8 * c), dv + 8 * c);
                            if (!first) tmem_ld8(pcol + (uint32_t)(j + 8 * c), lv + 8 * c);
                        }
                    tmem_ld_wait();
                    if (!first) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) dv[i] *= lv[i] * debias;
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) dv[i] *= debias;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (j + 8 * c < n8) tmem_st8(pcol + (uint32_t)(j + 8 * c), dv + 8 * c);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty0 + 8 * db);   // the accumulator may be overwritten
                tmem_st_wait();
            }

            // ---- root: sum_c w_0[c] * Lambda_0[c] * T_0[c]
            float res = 0.f;
            for (int c0 = 0; c0 < P.root_card; c0 += 8) {
                float lv[8], w[8];
                if (P.root_has_children) {
                    tmem_ld8(tlane + (uint32_t)(P.root_col + c0), lv);
                    tmem_ld_wait();
                }
                load_weights8<FMT>(my_bits, P.bits_words, drow, s_fan, fm, 0, P.root_lam_off, P.root_bit_off, P.root_fan_off, P.root_card, c0, w);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c0 + j < P.root_card) res = fmaf(P.root_has_children ? lv[j] * w[j] : w[j], __ldg(P.root_T + c0 + j), res);
            }
            if (q < P.nq) P.out[q] = res;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(P.tmem_cols) : "memory");
    }
}

uint32_t host_tf32_hi(float x) {
    uint32_t b;
    std::memcpy(&b, &x, 4);
    return (b + 0x1000u) & 0xFFFFE000u;
}

int k3_prepare(bc_model* m) {
    if (m->k3) {
        if (m->k3->failed) bc_set_error("fused tensor-core kernel (K3) does not serve this model");
        return m->k3->failed ? BC_ELIMIT : BC_OK;
    }
    // the plan is built locally and published to m->k3 only when it is complete: a structural "does not serve" verdict is
    // published (failed = 1, so it is not recomputed); a CUDA error while uploading the operand image publishes nothing, so
    // that the next call retries instead of launching with a NULL image
    std::unique_ptr<BcK3Plan> plan(new BcK3Plan());
    BcK3Plan* k = plan.get();
    auto fail = [&](const char* why) {
        k->failed = 1;
        m->k3 = plan.release();
        bc_set_error("fused tensor-core kernel (K3) does not serve this model: %s", why);
        return BC_ELIMIT;
    };
    const int n = m->n;
    if (n < 2) return fail("single-node model");
    if (n - 1 > kMaxEdges) return fail("more than 128 columns (the edge table travels in the kernel parameter bank)");
    if (m->arena.empty()) return fail("no host copy of the CPT arena");
    // ---- edge schedule: depth-first post order (children before parents, a node's edge right after its last child's),
    //      heaviest subtree first: a Lambda then only lives while its own subtree is being folded, so the live set is
    //      bounded by the depth of the tree instead of its width (the reverse topological order kept 5+ messages of a
    //      20-column synthetic tree alive at once and did not fit tensor memory)
    const int n_edges = n - 1;
    std::vector<std::vector<int>> kids(n);
    std::vector<long long> weight(n, 0);
    for (int v = n - 1; v >= 1; --v) {
        weight[v] += (long long)m->nodes[v].card * m->nodes[v].card_pa;
        weight[m->nodes[v].parent] += weight[v];
        kids[m->nodes[v].parent].push_back(v);
    }
    std::vector<int> sched;   // sched[e] = child node of edge e
    {
        std::vector<std::pair<int, size_t>> stack{{0, 0}};
        for (int v = 0; v < n; ++v)
            std::stable_sort(kids[v].begin(), kids[v].end(), [&](int a, int b) { return weight[a] > weight[b]; });
        while (!stack.empty()) {
            auto& [v, i] = stack.back();
            if (i < kids[v].size()) {
                const int c = kids[v][i++];
                stack.push_back({c, 0});
            } else {
                if (v != 0) sched.push_back(v);
                stack.pop_back();
            }
        }
    }
    std::vector<int> first_child_edge(n, -1), own_edge(n, -1);
    for (int e = 0; e < n_edges; ++e) {
        const int v = sched[e];
        own_edge[v] = e;
        const int pa = m->nodes[v].parent;
        if (first_child_edge[pa] < 0) first_child_edge[pa] = e;
    }
    int npad_max = 16;
    for (int v = 1; v < n; ++v) {
        const int np = (int)bc_round_up(m->nodes[v].card_pa, 16);
        if (np > 256) return fail("a parent domain exceeds 256 states");
        if (np > npad_max) npad_max = np;
    }
    // ---- TMEM columns: accumulator buffer(s) D first, then Lambda of every internal node by first fit over lifetimes
    //      [edge of its first child, its own edge] (the root lives to the end), in units of 8 columns.  Two accumulator
    //      buffers let the MMAs of the next edge start while the epilogue of this one runs; one buffer if that is
    //      what keeps a tile within 256 columns (two CTAs per SM).
    std::vector<int> col(n, -1);
    std::vector<int> order;   // internal nodes by start of lifetime
    for (int v = 0; v < n; ++v)
        if (first_child_edge[v] >= 0) order.push_back(v);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return first_child_edge[a] < first_child_edge[b]; });
    auto assign = [&](int n_dbuf) -> int {   // columns used, or -1 (col[] is only written when every node found a place)
        std::vector<int> place(n, -1);
        const int units_total = 512 / 8;
        std::vector<int> busy_until(units_total, -1);   // last edge index that uses the unit
        const int d_units = n_dbuf * npad_max / 8;
        if (d_units > units_total) return -1;
        for (int u = 0; u < d_units; ++u) busy_until[u] = 1 << 30;
        int units_used = d_units;
        for (int v : order) {
            const int need = (int)bc_round_up(m->nodes[v].card, 8) / 8;
            const int start = first_child_edge[v], end = v == 0 ? (1 << 30) : own_edge[v];
            int at = -1;
            for (int u0 = 0; u0 + need <= units_total && at < 0; ++u0) {
                bool ok = true;
                for (int u = u0; u < u0 + need; ++u)
                    if (busy_until[u] >= start) { ok = false; break; }
                if (ok) at = u0;
            }
            if (at < 0) return -1;
            for (int u = at; u < at + need; ++u) busy_until[u] = end;
            place[v] = at * 8;
            if (at + need > units_used) units_used = at + need;
        }
        col = place;
        return units_used * 8;
    };
    int n_dbuf = 2, used = assign(2);
    if (used < 0 || used > 256) {
        const int used1 = assign(1);
        if (used1 < 0) return fail("live messages exceed the 512 columns of tensor memory");
        if (used1 <= 256 || used < 0) { n_dbuf = 1; used = used1; }
        else used = assign(2);
    }
    if (const char* e = std::getenv("BC_K3_DBUF")) {   // experiments
        const int want = std::atoi(e);
        if ((want == 1 || want == 2) && assign(want) > 0) { n_dbuf = want; used = assign(want); }
    }
    int tmem_cols = 32;
    while (tmem_cols < used) tmem_cols <<= 1;
    k->tmem_cols = tmem_cols;
    k->npad_max = npad_max;
    k->d_col = 0;
    k->n_dbuf = n_dbuf;
    k->root_col = col[0];
    // ---- operand images: per edge and block of 16 child states, T_v^T hi then lo, 64-byte swizzled rows
    size_t total = 0;
    k->edges.resize(n_edges);
    for (int e = 0; e < n_edges; ++e) {
        const int v = sched[e];
        const BcNodeRec& nd = m->nodes[v];
        K3Edge& E = k->edges[e];
        std::memset(&E, 0, sizeof(E));
        E.v = v;
        E.K = nd.card;
        E.N = nd.card_pa;
        E.n_pad = (int)bc_round_up(nd.card_pa, 16);
        E.lam_off = nd.lam_off;
        E.bit_off = m->bits[v].bit_off;
        E.fan_off = nd.fan_off;
        E.col_v = col[v];
        E.col_pa = col[nd.parent];
        E.first = first_child_edge[nd.parent] == e;
        E.nkb = (nd.card + kBK - 1) / kBK;
        E.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(E.n_pad >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
        E.bimg_off = total;
        total += (size_t)E.nkb * E.n_pad * 128;
    }
    std::vector<uint8_t> img(total, 0);
    for (const K3Edge& E : k->edges) {
        const BcNodeRec& nd = m->nodes[E.v];
        const float* T = m->arena.data() + nd.cpt_off;
        for (int kb = 0; kb < E.nkb; ++kb) {
            uint8_t* hi = img.data() + E.bimg_off + (size_t)kb * E.n_pad * 128;
            uint8_t* lo = hi + (size_t)E.n_pad * 64;
            for (int p = 0; p < E.N; ++p)
                for (int kk = 0; kk < kBK; ++kk) {
                    const int c = kb * kBK + kk;
                    if (c >= E.K) continue;
                    const float x = T[(size_t)c * nd.stride + p];
                    const uint32_t hb = host_tf32_hi(x);
                    float h;
                    std::memcpy(&h, &hb, 4);
                    const float l = x - h;
                    const uint32_t lb = host_tf32_hi(l);
                    const size_t off = (size_t)p * 64 + ((size_t)((kk >> 2) ^ ((p >> 1) & 3)) << 4) + (size_t)(kk & 3) * 4;
                    std::memcpy(hi + off, &hb, 4);
                    std::memcpy(lo + off, &lb, 4);
                }
        }
    }
    k->bimg_bytes = total;
    // ---- shared memory / residency
    const size_t fan_floats = (size_t)bc_round_up((int64_t)m->fan.size(), 4);
    const size_t fixed = (size_t)m->bits_words * kTile * 4 + fan_floats * 4 + 256 /* nibble table */ +                          256 /* barriers */ + 1024 /* alignment */;
    k->smem = (size_t)kStagesA * 2 * kABytes + (size_t)kStagesB * npad_max * 128 + fixed;
    const size_t smem_optin = m->device >= 0 ? (size_t)m->smem_optin : (size_t)227 * 1024;   // host-only model: the sm_100 value
    if (k->smem > smem_optin) return fail("operand rings exceed shared memory");
    k->ctas_per_sm = (tmem_cols <= 256 && 2 * (k->smem + 1024) <= 228 * 1024) ? 2 : 1;
    if (k->ctas_per_sm == 1) k->smem = smem_optin;   // a second CTA would only spin in tcgen05.alloc
    if (m->device >= 0) {   // (a host-only model keeps the plan for inspection: bc_model_fused_plan)
        BC_CUDA_CHECK(cudaMalloc(&k->d_bimg, total));
        const cudaError_t ce = cudaMemcpy(k->d_bimg, img.data(), total, cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) {
            cudaFree(k->d_bimg);
            bc_set_error("upload of the K3 operand image failed: %s", cudaGetErrorString(ce));
            return BC_ECUDA;
        }
    }
    m->k3 = plan.release();
    return BC_OK;
}

template <int FMT>
int k3_launch_fmt(bc_model* m, const K3Params& P, int grid, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    BC_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        BC_CUDA_CHECK(cudaFuncSetAttribute(k3_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_optin));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    k3_kernel<FMT><<<grid, kThreads, m->k3->smem, st>>>(P);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

}  // namespace

int bc_k3_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask, float* out, cudaStream_t st) {
    if (nq == 0) return BC_OK;
    if (fmt != BC_DESC_BITS && fmt != BC_DESC_DENSE_F32) {
        bc_set_error("the fused kernel reads BITS or DENSE_F32 rows (convert range rows with bc_convert_desc)");
        return BC_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(desc) & 15u) {
        bc_set_error("descriptor rows are read with 16-byte loads: the buffer must be 16-byte aligned");
        return BC_EINVAL;
    }
    {
        std::lock_guard<std::mutex> g(m->k3_mu);
        int rc = k3_prepare(m);
        if (rc) return rc;
    }
    const BcK3Plan* k = m->k3;
    K3Params P{};
    P.n_edges = (int)k->edges.size();
    std::memcpy(P.edge, k->edges.data(), sizeof(K3Edge) * k->edges.size());
    P.bimg = k->d_bimg;
    P.desc = static_cast<const uint8_t*>(desc);
    P.dstride = (size_t)bc_model_desc_stride(m, fmt);
    P.fan_mask = fan_mask;
    P.fan = m->d_fan;
    P.fan_floats = (int)bc_round_up((int64_t)m->fan.size(), 4);
    P.fan_n = (int)m->fan.size();
    const BcNodeRec& r = m->nodes[0];
    P.root_T = m->d_arena + r.cpt_off;
    P.root_card = r.card;
    P.root_col = k->root_col;
    P.root_bit_off = m->bits[0].bit_off;
    P.root_lam_off = r.lam_off;
    P.root_fan_off = r.fan_off;
    P.root_has_children = k->root_col >= 0;
    P.out = out;
    P.nq = nq;
    P.n_tiles = (long long)((nq + kTile - 1) / kTile);
    P.bits_words = m->bits_words;
    P.b_slot_bytes = k->npad_max * 128;
    P.d_col = k->d_col;
    P.d_stride = k->npad_max;
    P.n_dbuf = k->n_dbuf;
    P.tmem_cols = k->tmem_cols;
    P.mask_words = m->mask_words;
    P.debias_unit = 1.1e-8f;
    if (const char* e = std::getenv("BC_K3_DEBIAS")) P.debias_unit = (float)std::atof(e);
    long long grid = (long long)m->sm_count * k->ctas_per_sm;
    if (grid > P.n_tiles) grid = P.n_tiles;
    if (fmt == BC_DESC_BITS) return k3_launch_fmt<BC_DESC_BITS>(m, P, (int)grid, st);
    return k3_launch_fmt<BC_DESC_DENSE_F32>(m, P, (int)grid, st);
}

extern "C" int bc_model_fused_plan(bc_model* m, int32_t* info, int32_t* edges, size_t edges_capacity) {
    if (!m || !info) { bc_set_error("model/info is NULL"); return BC_EINVAL; }
    {
        std::lock_guard<std::mutex> g(m->k3_mu);
        int rc = k3_prepare(m);
        if (rc) return rc;
    }
    const BcK3Plan* k = m->k3;
    info[0] = (int32_t)k->edges.size();
    info[1] = k->tmem_cols;
    info[2] = k->ctas_per_sm;
    info[3] = (int32_t)k->smem;
    info[4] = k->d_col;
    info[5] = k->npad_max * k->n_dbuf;   // accumulator columns [d_col, d_col + this)
    info[6] = k->root_col;
    info[7] = (int32_t)(k->bimg_bytes >> 10);
    if (edges) {
        if (edges_capacity < k->edges.size()) { bc_set_error("edges buffer too small"); return BC_EINVAL; }
        for (size_t e = 0; e < k->edges.size(); ++e) {
            const K3Edge& E = k->edges[e];
            int32_t* o = edges + 8 * e;
            o[0] = E.v; o[1] = E.K; o[2] = E.N; o[3] = E.n_pad; o[4] = E.col_v; o[5] = E.col_pa; o[6] = E.first; o[7] = E.nkb;
        }
    }
    return BC_OK;
}

void bc_k3_free(bc_model* m) {
    if (!m->k3) return;
    cudaFree(m->k3->d_bimg);
    delete m->k3;
    m->k3 = nullptr;
}
