// K3b -- single-CTA variant of the fused whole-tree tensor-core kernel (k3_fused.cu documents the algorithm and the
// operand layouts).  Same math, different schedule; built to attack what bounds K3 ("What bounds
// it"): a tcgen05.mma that accumulates into the columns of the previous one waits ~140 clk for it, and one tile per
// CTA serialises operand build -> MMA drain -> epilogue.
//
//   * ONE CTA per SM owns all 512 TMEM columns: per edge TWO accumulators (independent dependent-MMA chains), ring
//     steps alternate between them, and TWO issuer warps (one per chain) halve the per-step issue latency;
//   * [T_hi | T_lo] are concatenated along N in the operand image: one instruction computes the main product and the
//     first correction side by side (N = 2 n8), a second one adds U_lo . T_hi onto the correction half; unit-weight
//     leaves need ONE instruction per k-step instead of two, everything else two instead of three, and the epilogue
//     adds main + correction of both chains in round-to-nearest registers (less truncation bias than K3);
//   * TWO producer groups of four warps build alternate ring steps and split the epilogue's columns, so the
//     per-step latency of a producer (LDS / LDG -> split -> STS -> fence -> arrive, ~500 clk) is overlapped (three
//     groups = 480 threads hit the 128-register cap and spill: measured slower);
//   * one operand ring: a slot holds the U_v block (hi, lo) and the T_v^T block, ONE full barrier (four warp arrivals
//     + the TMA transaction) and ONE empty barrier (one tcgen05.commit) per slot.
#include <algorithm>
#include <cstring>

#include "bc_internal.h"

struct K3bEdge {           // 64 bytes, lives in the kernel parameter bank
    int16_t v, K, N, n8;   // child, card(v), card(pa), card(pa) rounded up to 8
    int32_t lam_off;       // first float of column v in a DENSE row
    int32_t bit_off;       // first bit of column v in a BITS row
    int32_t fan_off;       // fanouts[v] in the fan arena, -1 if none
    int16_t col_v;         // TMEM column of Lambda_v, -1 for a leaf
    int16_t col_pa;        // TMEM column of Lambda_pa
    int16_t d_col;         // TMEM column of the first accumulator of this edge
    int16_t d_stride;      // columns per accumulator: [0, n8) main product, [n8, 2 n8) correction products (+ spill)
    int8_t first;          // this edge is the first message into Lambda_pa
    int8_t next_reads;     // the next edge (or the root) reads a Lambda: the producer groups must meet after this epilogue
    int8_t pad0[2];
    int32_t nkb;           // blocks of 16 child states
    uint32_t idesc_cat;    // tcgen05 instruction descriptor, N = 2 n8: U_hi . [T_hi | T_lo]  (TF32 x TF32 -> F32, M = 128, K-major)
    uint32_t idesc_lo;     // N = n8 rounded up to 16: U_lo . T_hi
    uint32_t bimg_off16;   // offset of the edge's operand images in 16-byte units
    int32_t pad[4];
};
static_assert(sizeof(K3bEdge) == 64, "K3bEdge layout");

struct BcK3bPlan {
    int failed = 0;
    std::vector<K3bEdge> edges;
    uint8_t* d_bimg = nullptr;
    size_t bimg_bytes = 0;
    int b_slot_bytes = 0;
    int root_col = 0;
    size_t smem = 0;
};

namespace {

constexpr int kTile = 128;    // queries per CTA tile = TMEM lanes = UMMA M
constexpr int kBK = 16;       // child states per ring step: one 64-byte swizzle row
constexpr int kGroups = 2;    // producer groups of 4 warps (one warp per TMEM lane quarter): ring step i is built by group i % kGroups
constexpr int kIssuers = 2;   // issuer warp w owns accumulator w and the ring steps with step % 2 == w
constexpr int kStages = 6;
constexpr int kABytes = kTile * kBK * 4;   // 8 KB per half (hi or lo)
constexpr int kProducerWarps = 4 * kGroups;
constexpr int kProducerThreads = 32 * kProducerWarps;
constexpr int kThreads = 32 * (kProducerWarps + kIssuers + 1);   // + TMA warp

struct K3bParams {
    K3bEdge edge[31];          // in the kernel parameter bank: uniform loads, warp-uniform control flow
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
    int b_slot_bytes;          // largest operand image: 2 * n8 * 64, rounded up to 1 KB
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
__global__ void __launch_bounds__(kThreads, 1) k3b_kernel(const __grid_constant__ K3bParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (role dispatch without divergence)
    // carve-up: A ring | B ring | BITS rows of the tile [group][word][query] | fan arena | nibble table | barriers
    uint8_t* p = smem + (size_t)kStages * 2 * kABytes;
    const uint32_t a_ring = smem_u32(smem), b_ring = smem_u32(p);
    p += (size_t)kStages * P.b_slot_bytes;
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(p);
    p += (FMT == BC_DESC_BITS ? (size_t)kGroups * P.bits_words * kTile * 4 : 0);
    float* s_fan = reinterpret_cast<float*>(p);
    p += (size_t)P.fan_floats * 4;
    float4* s_tab = reinterpret_cast<float4*>(p);   // nibble -> four 0/1 floats
    p += 256;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kStages, d_full0 = empty0 + 8 * kStages, d_empty0 = d_full0 + 8 * kIssuers;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + kIssuers + 1);

    for (int i = tid; i < P.fan_floats; i += kThreads) s_fan[i] = i < P.fan_n ? P.fan[i] : 0.f;
    if (tid < 16) s_tab[tid] = make_float4(tid & 1 ? 1.f : 0.f, tid & 2 ? 1.f : 0.f, tid & 4 ? 1.f : 0.f, tid & 8 ? 1.f : 0.f);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 5);    // four warps of the producer group + the TMA thread's arrive.expect_tx
            mbar_init(empty0 + 8 * s, 1);   // one tcgen05.commit
        }
        for (int w = 0; w < kIssuers; ++w) mbar_init(d_full0 + 8 * w, 1);   // one commit per issuer and edge
        mbar_init(d_empty0, kProducerWarps);                                // every producer warp, once per edge
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform: tcgen05 operands live in uniform registers

    if (warp == kProducerWarps + kIssuers) {
        // ================= TMA warp: keeps the ring's T_v^T blocks ahead; the (edge, block) sequence repeats per tile
        if (lane == 0) {
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
                for (int e = 0; e < P.n_edges; ++e) {
                    const K3bEdge& E = P.edge[e];
                    const unsigned bytes = (unsigned)E.n8 * 128u;   // T_hi rows then T_lo rows, 64 B each
                    for (int kb = 0; kb < E.nkb; ++kb, ++it) {
                        const uint32_t s = it % kStages, par = (it / kStages) & 1u;
                        mbar_wait(empty0 + 8 * s, par ^ 1u);
                        mbar_expect_tx(full0 + 8 * s, bytes);
                        tma_bulk_g2s(b_ring + s * P.b_slot_bytes, P.bimg + (size_t)E.bimg_off16 * 16 + (size_t)kb * bytes, bytes, full0 + 8 * s);
                    }
                }
        }
    } else if (warp >= kProducerWarps) {
        // ================= issuer warp w: ring steps with step % 2 == w, accumulator w of every edge
        const uint32_t w = (uint32_t)(warp - kProducerWarps);
        uint32_t it = 0, ed = 0;
        const uint32_t desc_hi = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);   // SBO, version 1, 64-byte swizzle
        for (long long tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
            for (int e = 0; e < P.n_edges; ++e, ++ed) {
                const K3bEdge& E = P.edge[e];
                const bool a_exact = FMT == BC_DESC_BITS && E.col_v < 0 && !(E.fan_off >= 0 && P.fan_mask != nullptr);
                const uint32_t d = tmem + (uint32_t)E.d_col + w * (uint32_t)E.d_stride, lo_off = (uint32_t)E.n8;
                const uint32_t idesc_cat = E.idesc_cat, idesc_lo = E.idesc_lo;
                const int K = E.K, nkb = E.nkb;
                uint32_t acc = 0;   // the first instruction of this issuer in this edge overwrites its accumulator
                mbar_wait(d_empty0, (ed & 1u) ^ 1u);   // the epilogue of the previous edge has read the accumulators
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    if ((it & 1u) != w) continue;
                    const uint32_t s = it % kStages, par = (it / kStages) & 1u;
                    // low words of the operand descriptors: start address >> 4 | LBO = 1
                    const uint32_t a_hi = (((a_ring + s * 2 * kABytes) & 0x3FFFFu) >> 4) | (1u << 16), a_lo = a_hi + (kABytes >> 4);
                    const uint32_t b_cat = (((b_ring + s * P.b_slot_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
                    const bool two = K - kb * kBK > 8;
                    mbar_wait(full0 + 8 * s, par);
                    tc_fence_after();
                    umma_tf32(d, a_hi, b_cat, desc_hi, idesc_cat, acc);   // 8 TF32 = 32 bytes per k-step: +2 in the address field
                    if (!a_exact) umma_tf32(d + lo_off, a_lo, b_cat, desc_hi, idesc_lo, 1);
                    if (two) {
                        umma_tf32(d, a_hi + 2, b_cat + 2, desc_hi, idesc_cat, 1);
                        if (!a_exact) umma_tf32(d + lo_off, a_lo + 2, b_cat + 2, desc_hi, idesc_lo, 1);
                    }
                    acc = 1;
                    umma_commit(empty0 + 8 * s);
                }
                umma_commit(d_full0 + 8 * w);   // arrives at once when this issuer had no step in the edge
            }
    } else {
        // ================= producer / epilogue warps: row = query = TMEM lane; kGroups warps share a row
        const int grp = warp >> 2, row_id = tid & (kTile - 1);
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t sw = ((uint32_t)row_id >> 1) & 3u;    // 64-byte swizzle: chunk j of row r lives at r * 64 + ((j ^ ((r >> 1) & 3)) << 4)
        uint32_t* grp_bits = s_bits + (size_t)grp * P.bits_words * kTile;
        const uint32_t* my_bits = grp_bits + row_id;
        uint32_t it = 0, ed = 0;
        for (long long tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
            const size_t q = (size_t)tile * kTile + row_id;
            const size_t qc = q < P.nq ? q : P.nq - 1;
            const uint32_t fm = P.fan_mask ? P.fan_mask[qc] : 0u;
            const float* drow = reinterpret_cast<const float*>(P.desc + qc * P.dstride);
            if (FMT == BC_DESC_BITS) {   // each thread stages its own row in its group's copy: no cross-thread hazard
                const uint32_t* grow = reinterpret_cast<const uint32_t*>(P.desc + qc * P.dstride);
                for (int w4 = 0; w4 < P.bits_words; w4 += 4) {   // bits_words is a multiple of 4
                    const uint4 x = __ldg(reinterpret_cast<const uint4*>(grow + w4));
                    grp_bits[(w4 + 0) * kTile + row_id] = x.x;
                    grp_bits[(w4 + 1) * kTile + row_id] = x.y;
                    grp_bits[(w4 + 2) * kTile + row_id] = x.z;
                    grp_bits[(w4 + 3) * kTile + row_id] = x.w;
                }
            }
            // message-independent inputs of ring step (e, kb): finished hi chunks for unit-weight leaves, else the weights
            auto fetch = [&](int e, int kb, float* pre) __attribute__((always_inline)) {
                const K3bEdge& E = P.edge[e];
                const int c0 = kb * kBK, K = E.K;
                const bool exact = FMT == BC_DESC_BITS && E.col_v < 0 && !(E.fan_off >= 0 && P.fan_mask != nullptr);
                if (FMT == BC_DESC_BITS) {
                    const uint32_t m = bits16(my_bits, P.bits_words, E.bit_off, K, c0);
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
                        if (c0 + 4 * j < K) t = __ldg(src + j);
                        pre[4 * j] = t.x; pre[4 * j + 1] = t.y; pre[4 * j + 2] = t.z; pre[4 * j + 3] = t.w;
                    }
                }
                if (E.fan_off >= 0 && ((fm >> E.v) & 1u)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + 4 * j < K) {
                            const float4 f = *reinterpret_cast<const float4*>(s_fan + E.fan_off + c0 + 4 * j);
                            pre[4 * j] *= f.x; pre[4 * j + 1] *= f.y; pre[4 * j + 2] *= f.z; pre[4 * j + 3] *= f.w;
                        }
                }
            };
            // this group's next step: cursor (pe, pkb) runs kGroups steps ahead of the step being built
            int pe = 0, pkb = 0;
            auto advance = [&](int n) __attribute__((always_inline)) {
                for (int i = 0; i < n && pe < P.n_edges; ++i)
                    if (++pkb == P.edge[pe].nkb) { pkb = 0; ++pe; }
            };
            advance((int)((grp + kGroups - it % kGroups) % kGroups));   // first step of this tile that belongs to the group
            float pre[16];
            if (pe < P.n_edges) fetch(pe, pkb, pre);
            for (int e = 0; e < P.n_edges; ++e, ++ed) {
                const K3bEdge& E = P.edge[e];
                const bool leaf = E.col_v < 0;
                const bool a_exact = FMT == BC_DESC_BITS && leaf && !(E.fan_off >= 0 && P.fan_mask != nullptr);
                const int K = E.K, nkb = E.nkb;
                const uint32_t it0 = it;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    if ((int)(it % kGroups) != grp) continue;   // another group's block
                    const uint32_t s = it % kStages, par = (it / kStages) & 1u;
                    const uint32_t row = a_ring + s * 2 * kABytes + (uint32_t)row_id * 64u;
                    const int c0 = kb * kBK;
                    const int ks = (K - c0 > 8) ? 2 : 1;
                    float u[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) u[j] = pre[j];
                    advance(kGroups);
                    if (pe < P.n_edges) fetch(pe, pkb, pre);
                    if (a_exact) {
                        mbar_wait(empty0 + 8 * s, par ^ 1u);   // the MMAs that read this slot are done
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
                        mbar_wait(empty0 + 8 * s, par ^ 1u);
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
                    if (lane == 0) mbar_arrive(full0 + 8 * s);
                }
                // ---- epilogue: Lambda_pa (*)= sum over the accumulators of (main + correction); the groups take the
                //      16-column chunks round-robin; accumulator w was written iff an issuer-w step fell into this edge
                const bool use0 = nkb >= 2 || (it0 & 1u) == 0u, use1 = nkb >= 2 || (it0 & 1u) == 1u;
                const uint32_t dcol0 = tlane + (uint32_t)E.d_col, dcol1 = dcol0 + (uint32_t)E.d_stride, pcol = tlane + (uint32_t)E.col_pa;
                const int n8 = E.n8;
                const bool first = E.first;
                mbar_wait(d_full0, ed & 1u);
                mbar_wait(d_full0 + 8, ed & 1u);
                tc_fence_after();
                for (int j = 16 * grp; j < n8; j += 16 * kGroups) {
                    const bool two = j + 8 < n8;
                    float m0[16], c0v[16], m1[16], c1v[16], lv[16];
                    if (use0) {
                        tmem_ld8(dcol0 + (uint32_t)j, m0);
                        tmem_ld8(dcol0 + (uint32_t)(n8 + j), c0v);
                        if (two) {
                            tmem_ld8(dcol0 + (uint32_t)(j + 8), m0 + 8);
                            tmem_ld8(dcol0 + (uint32_t)(n8 + j + 8), c0v + 8);
                        }
                    }
                    if (use1) {
                        tmem_ld8(dcol1 + (uint32_t)j, m1);
                        tmem_ld8(dcol1 + (uint32_t)(n8 + j), c1v);
                        if (two) {
                            tmem_ld8(dcol1 + (uint32_t)(j + 8), m1 + 8);
                            tmem_ld8(dcol1 + (uint32_t)(n8 + j + 8), c1v + 8);
                        }
                    }
                    if (!first) {
                        tmem_ld8(pcol + (uint32_t)j, lv);
                        if (two) tmem_ld8(pcol + (uint32_t)(j + 8), lv + 8);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x = 0.f;
                        if (use0) x = m0[i] + c0v[i];
                        if (use1) x += m1[i] + c1v[i];
                        m0[i] = first ? x : x * lv[i];
                    }
                    tmem_st8(pcol + (uint32_t)j, m0);
                    if (two) tmem_st8(pcol + (uint32_t)(j + 8), m0 + 8);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty0);   // the accumulators may be overwritten
                tmem_st_wait();
                if (E.next_reads) {   // every group wrote a share of Lambda_pa and the next edge (or the root) reads a Lambda
                    tc_fence_before();
                    asm volatile("bar.sync 1, %0;" ::"r"(kProducerThreads) : "memory");
                    tc_fence_after();
                }
            }

            // ---- root: sum_c w_0[c] * Lambda_0[c] * T_0[c]   (group 0)
            if (grp == 0) {
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
            // the next tile's epilogues overwrite Lambda columns that group 0 may still be reading for the root
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"r"(kProducerThreads) : "memory");
            tc_fence_after();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

uint32_t host_tf32_hi(float x) {
    uint32_t b;
    std::memcpy(&b, &x, 4);
    return (b + 0x1000u) & 0xFFFFE000u;
}

int k3b_prepare(bc_model* m) {
    if (m->k3b) return m->k3b->failed ? BC_ELIMIT : BC_OK;
    BcK3bPlan* k = new BcK3bPlan();
    m->k3b = k;
    auto fail = [&](const char* why) {
        k->failed = 1;
        bc_set_error("fused tensor-core kernel (K3b) does not serve this model: %s", why);
        return BC_ELIMIT;
    };
    const int n = m->n;
    if (n < 2) return fail("single-node model");
    if (n > 32) return fail("more than 32 columns (one fan-out mask word per query)");
    if (m->arena.empty()) return fail("no host copy of the CPT arena");
    for (int v = 0; v < n; ++v)
        if (m->nodes[v].card > 128) return fail("a domain exceeds 128 states (main and correction product share one 256-column instruction)");
    // ---- edge schedule: reverse topological order (children before parents)
    const int n_edges = n - 1;
    std::vector<int> first_child_edge(n, -1), own_edge(n, -1);
    for (int v = n - 1, e = 0; v >= 1; --v, ++e) {
        own_edge[v] = e;
        const int pa = m->nodes[v].parent;
        if (first_child_edge[pa] < 0) first_child_edge[pa] = e;
    }
    k->edges.resize(n_edges);
    for (int v = n - 1, e = 0; v >= 1; --v, ++e) {
        const BcNodeRec& nd = m->nodes[v];
        K3bEdge& E = k->edges[e];
        std::memset(&E, 0, sizeof(E));
        E.v = (int16_t)v;
        E.K = (int16_t)nd.card;
        E.N = (int16_t)nd.card_pa;
        E.n8 = (int16_t)bc_round_up(nd.card_pa, 8);
        E.lam_off = nd.lam_off;
        E.bit_off = m->bits[v].bit_off;
        E.fan_off = nd.fan_off;
        E.first = first_child_edge[nd.parent] == e;
        E.nkb = (nd.card + kBK - 1) / kBK;
        const uint32_t n_cat = 2u * (uint32_t)E.n8, n_lo = (uint32_t)bc_round_up(E.n8, 16);
        E.idesc_cat = (1u << 4) | (2u << 7) | (2u << 10) | ((n_cat >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
        E.idesc_lo = (1u << 4) | (2u << 7) | (2u << 10) | ((n_lo >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
        // the correction instruction is N = n8 rounded up to 16: when n8 is not a multiple of 16 it spills 8 columns
        // (U_lo . first rows of T_lo: finite, never read) past the 2 n8 columns of the accumulator
        E.d_stride = (int16_t)(2 * E.n8 + (int)(n_lo - (uint32_t)E.n8));
    }
    // ---- TMEM columns by first fit over lifetimes, in units of 8 columns: Lambda of an internal node lives from the
    //      edge of its first child to its own edge (the root to the end); the two accumulators of edge e live during e.
    std::vector<int> col(n, -1), dcol(n_edges, -1);
    {
        const int units_total = 512 / 8;
        std::vector<int> busy_until(units_total, -1);   // last edge index that uses the unit
        auto place = [&](int need, int start, int end) -> int {
            for (int u0 = 0; u0 + need <= units_total; ++u0) {
                bool ok = true;
                for (int u = u0; u < u0 + need; ++u)
                    if (busy_until[u] >= start) { ok = false; break; }
                if (!ok) continue;
                for (int u = u0; u < u0 + need; ++u) busy_until[u] = end;
                return u0 * 8;
            }
            return -1;
        };
        for (int e = 0; e < n_edges; ++e) {
            // the two accumulators first (issuer w owns [d_col + w * d_stride, ...)), then the Lambda whose lifetime starts here
            dcol[e] = place(kIssuers * k->edges[e].d_stride / 8, e, e);
            if (dcol[e] < 0) return fail("live messages and accumulators exceed the 512 columns of tensor memory");
            for (int v = 0; v < n; ++v)
                if (first_child_edge[v] == e) {
                    col[v] = place((int)bc_round_up(m->nodes[v].card, 8) / 8, e, v == 0 ? (1 << 30) : own_edge[v]);
                    if (col[v] < 0) return fail("live messages and accumulators exceed the 512 columns of tensor memory");
                }
        }
    }
    k->root_col = col[0];
    // ---- operand images: per edge and block of 16 child states, rows [0, n8) = T_v^T hi, rows [n8, 2 n8) = lo,
    //      64-byte swizzled rows: ONE K-major operand of N = 2 n8 rows
    size_t total = 0;
    int b_slot = 0;
    for (int e = 0; e < n_edges; ++e) {
        K3bEdge& E = k->edges[e];
        const BcNodeRec& nd = m->nodes[E.v];
        E.col_v = (int16_t)col[E.v];
        E.col_pa = (int16_t)col[nd.parent];
        E.d_col = (int16_t)dcol[e];
        // the producer groups meet after this epilogue when the next edge builds U from a Lambda, or the root follows
        E.next_reads = (int8_t)(e + 1 == n_edges || col[k->edges[e + 1].v] >= 0);
        E.bimg_off16 = (uint32_t)(total / 16);
        total += (size_t)E.nkb * E.n8 * 128;
        if (E.n8 * 128 > b_slot) b_slot = E.n8 * 128;
    }
    std::vector<uint8_t> img(total + 1024, 0);   // slack: the correction instruction reads up to 8 rows past an image
    for (const K3bEdge& E : k->edges) {
        const BcNodeRec& nd = m->nodes[E.v];
        const float* T = m->arena.data() + nd.cpt_off;
        for (int kb = 0; kb < E.nkb; ++kb) {
            uint8_t* hi = img.data() + (size_t)E.bimg_off16 * 16 + (size_t)kb * E.n8 * 128;
            uint8_t* lo = hi + (size_t)E.n8 * 64;
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
                    // row r of the concatenated operand: chunk j lives at r * 64 + ((j ^ ((r >> 1) & 3)) << 4)
                    const int rh = p, rl = E.n8 + p;
                    const size_t oh = (size_t)rh * 64 + ((size_t)((kk >> 2) ^ ((rh >> 1) & 3)) << 4) + (size_t)(kk & 3) * 4;
                    const size_t ol = (size_t)rl * 64 + ((size_t)((kk >> 2) ^ ((rl >> 1) & 3)) << 4) + (size_t)(kk & 3) * 4;
                    std::memcpy(hi + oh, &hb, 4);
                    std::memcpy(hi + ol, &lb, 4);
                    (void)lo;
                }
        }
    }
    k->bimg_bytes = total + 1024;
    k->b_slot_bytes = (int)bc_round_up(b_slot, 1024);
    // ---- shared memory / residency
    const size_t fan_floats = (size_t)bc_round_up((int64_t)m->fan.size(), 4);
    const size_t fixed = (size_t)kGroups * m->bits_words * kTile * 4 + fan_floats * 4 + 256 /* nibble table */ + 256 /* barriers */ + 1024 /* alignment */;
    k->smem = (size_t)kStages * 2 * kABytes + (size_t)kStages * k->b_slot_bytes + fixed;
    if (k->smem > (size_t)m->smem_optin) return fail("operand rings exceed shared memory");
    k->smem = (size_t)m->smem_optin;
    BC_CUDA_CHECK(cudaMalloc(&k->d_bimg, k->bimg_bytes));
    BC_CUDA_CHECK(cudaMemcpy(k->d_bimg, img.data(), k->bimg_bytes, cudaMemcpyHostToDevice));
    return BC_OK;
}

template <int FMT>
int k3b_launch_fmt(bc_model* m, const K3bParams& P, int grid, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    BC_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        BC_CUDA_CHECK(cudaFuncSetAttribute(k3b_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_optin));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    k3b_kernel<FMT><<<grid, kThreads, m->k3b->smem, st>>>(P);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

}  // namespace

int bc_k3b_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask, float* out, cudaStream_t st) {
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
        std::lock_guard<std::mutex> g(m->k3b_mu);
        int rc = k3b_prepare(m);
        if (rc) return rc;
    }
    const BcK3bPlan* k = m->k3b;
    K3bParams P{};
    P.n_edges = (int)k->edges.size();
    std::memcpy(P.edge, k->edges.data(), sizeof(K3bEdge) * k->edges.size());
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
    P.b_slot_bytes = k->b_slot_bytes;
    long long grid = (long long)m->sm_count;
    if (grid > P.n_tiles) grid = P.n_tiles;
    return fmt == BC_DESC_BITS ? k3b_launch_fmt<BC_DESC_BITS>(m, P, (int)grid, st) : k3b_launch_fmt<BC_DESC_DENSE_F32>(m, P, (int)grid, st);
}

void bc_k3b_free(bc_model* m) {
    if (!m->k3b) return;
    cudaFree(m->k3b->d_bimg);
    delete m->k3b;
    m->k3b = nullptr;
}
