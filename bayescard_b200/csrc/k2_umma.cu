// K2 tensor-core edge kernel: one internal edge of the batched sum-product as a tcgen05 GEMM.
//
//     C[rows x N] = mask(Lambda_v)[rows x K] . T_v[K x N]         K = card(v), N = card(parent)
//     Lambda_pa   = accumulate ? Lambda_pa * C : C
//
// TF32 alone (10-bit mantissa) would miss the 1e-5 budget, so the product is error compensated
// ("3xTF32"): every fp32 operand x is split into hi = x rounded to the nearest TF32 number and
// lo = (x - hi) rounded to TF32, and
//
//     A.B  ~=  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi                  (fp32 accumulation in TMEM)
//
// which leaves a relative error of ~2^-21 per product (measured against the fp64 oracle in
// tests/test_gpu_parity.py::test_batched_large_domain_path).
//
// Shape of the kernel (sm_100a only):
//   * CTA tile 128 queries x BN parent states (BN = 32 / 64 / 128 / 256), K in blocks of 32 floats = one
//     128-byte swizzle atom; tiles are ordered so that co-resident CTAs share A panels and B panels through L2;
//   * warp 0: TMA producer -- four 2-D tiled bulk tensor loads per stage (A_hi, A_lo, B_hi, B_lo; 128-byte
//     swizzle, out-of-bounds rows zero filled) completing on an mbarrier (SASS UTMALDG);
//   * warp 1: allocates TMEM (two accumulator buffers) and issues tcgen05.mma.cta_group::1.kind::tf32 (SASS
//     UTCHMMA/UTC*MMA) from one thread, 12 MMAs (4 k-steps x 3 products) per stage in chunks of KS k-steps;
//     tcgen05.commit frees the stage and hands every chunk sum to the epilogue;
//   * warps 2..: epilogue -- tcgen05.ld (SASS LDTM) of every chunk sum into fp32 round-to-nearest register
//     accumulators (the tensor core truncates: see UmmaCfg), then the fused multiply into Lambda_pa;
//   * both operands K-major: Lambda rows are K-contiguous as produced by the previous edge, the CPT is
//     stored transposed (and pre-split) once per model.
// The range mask of column v and the hi/lo split of Lambda_v are one elementwise pass (k2_split_kernel).
#include <cuda.h>

#include "bc_internal.h"

struct BcUmmaPlan {
    std::vector<float*> d_tt_hi, d_tt_lo;  // per internal non-root node: T^T split, [N x ldk]
    std::vector<int> ldk;
    void* encode = nullptr;                // cuTensorMapEncodeTiled
    int failed = 0;
};

namespace {

constexpr int kBM = 128, kLdkAlign = 32;  // K is padded to 32 floats whatever the stage depth

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 -> nearest TF32 (round half to even on the 13 dropped mantissa bits); the result is a valid fp32 whose
// low 13 bits are zero, so the tensor core's own fp32->tf32 conversion is exact whatever its rounding mode
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t b = __float_as_uint(x);
    b += 0xFFFu + ((b >> 13) & 1u);
    return __uint_as_float(b & 0xFFFFE000u);
}
// x = hi + lo with both halves TF32 numbers (|x - hi - lo| <= 2^-23 |x|); rounding, not truncating, keeps
// the compensated product unbiased -- truncation gave a systematic -1e-5 over a 7-edge tree
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - hi);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// K-major shared-memory operand descriptor.  128-byte swizzle (BK = 32 floats): rows of 128 B, 8-row groups 1024 B
// apart (SBO), layout type 2.  64-byte swizzle (BK = 16 floats): rows of 64 B, 8-row groups 512 B apart, layout type 4.
// LBO = 1 (unused for swizzled K-major), version 1 (sm_100).
template <int BK>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    constexpr uint64_t sbo = (BK == 32 ? 1024 : 512) >> 4, layout = BK == 32 ? 2 : 4;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of fp32 from TMEM, NO wait: pair with tmem_ld_wait() before the registers are read
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Geometry of one kernel variant.  BN = parent states per CTA tile, KS = k-steps (of 8 TF32) accumulated inside
// TMEM before the partial sum is drained into fp32 registers.
//
// Why drain at all: tcgen05.mma adds into its fp32 accumulator with TRUNCATION, and every quantity of this path is
// a non-negative probability, so the truncation errors do not cancel -- measured on the first version of this
// kernel (one accumulator over the whole K loop): -2e-8 relative per MMA, i.e. -7e-6 per edge at K = 1000 and
// -3.3e-4 over the 51 internal edges of a 100-column tree, far outside the 1e-5 budget.  Here a TMEM accumulator
// only ever holds a chunk of KS k-steps; the chunk's two small correction products (A_lo.B_hi, A_hi.B_lo, 2^-11
// of the result) are issued FIRST, while the accumulator is still tiny, so only the KS main products truncate at
// full magnitude; the epilogue warps add the chunk sums in fp32 round-to-nearest registers while the tensor core
// fills the other TMEM buffer.
//
// BK = floats of K per pipeline stage (32: 128-byte swizzle, 16: 64-byte swizzle -- half the bytes per stage, so
// twice the stages in the same shared memory: what hides the TMA latency at BN = 256), EC = accumulator columns
// per epilogue thread, OCC = CTAs per SM the variant is sized for.
template <int BN_, int BK_, int KS_, int EC_, int OCC_>
struct UmmaCfg {
    static constexpr int BN = BN_, BK = BK_, KS = KS_, OCC = OCC_;
    static constexpr int EPI_COLS = EC_;                      // accumulator columns per epilogue thread (registers)
    static constexpr int EPI_WARPS = 4 * (BN / EPI_COLS);     // 4 TMEM lane quarters x column groups
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int A_BYTES = kBM * BK * 4, B_BYTES = BN * BK * 4, STAGE = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int BUDGET = (OCC == 1 ? 200 : 96) * 1024;
    static constexpr int STAGES = BUDGET / STAGE > 6 ? 6 : BUDGET / STAGE;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulator buffers; power of two >= 32
    static constexpr size_t SMEM = (size_t)STAGES * STAGE + 1024 /* alignment slack */ + 256 /* barriers */;
    static constexpr int CHUNKS_PER_KB = (BK / 8) / KS;
    static_assert(BK == 32 || BK == 16, "one swizzle atom per stage row");
    static_assert((BK / 8) % KS == 0 && STAGES >= 2 && BN % EC_ == 0 && TMEM_COLS * OCC <= 512, "bad variant");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::OCC)
k2_umma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, int rows, int N,
               int num_kb, int n_col_tiles, int n_row_tiles, float* __restrict__ lam_pa, int ld_pa, int accumulate) {
    constexpr int BN = Cfg::BN, KS = Cfg::KS, kBK = Cfg::BK, A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE = Cfg::STAGE;
    constexpr int kStg = Cfg::STAGES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the dynamic shared window is only 16 B aligned by contract: round up to the 1024 B the swizzle needs
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStg * STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStg + 4);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStg);
    const uint32_t acc_full0 = smem_u32(bars + 2 * kStg), acc_empty0 = smem_u32(bars + 2 * kStg + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile order: groups of up to 8 column tiles; inside a group the column index runs fastest, so the CTAs that
    // are resident together cover ~8 column tiles x ~18 row tiles: each A panel is read from HBM once and shared
    // through L2 by its column tiles, and a group's B panels stay L2 resident while the row tiles stream past
    int tile = blockIdx.x;
    const int group_cols = 8;
    const int group = tile / (group_cols * n_row_tiles);
    const int first_col = group * group_cols;
    const int cols_here = n_col_tiles - first_col < group_cols ? n_col_tiles - first_col : group_cols;
    tile -= group * group_cols * n_row_tiles;
    const int m0 = (tile / cols_here) * kBM, n0 = (first_col + tile % cols_here) * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStg; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full0 + 8 * b, 1);
            mbar_init(acc_empty0 + 8 * b, Cfg::EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: two accumulator buffers of BN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStg;
                mbar_wait(empty0 + 8 * s, ((kb / kStg) & 1) ^ 1);
                const uint32_t base = smem_u32(smem + s * STAGE), bar = full0 + 8 * s;
                mbar_expect_tx(bar, STAGE);
                tma_load_2d(base, &tm_a_hi, kb * kBK, m0, bar);
                tma_load_2d(base + A_BYTES, &tm_a_lo, kb * kBK, m0, bar);
                tma_load_2d(base + 2 * A_BYTES, &tm_b_hi, kb * kBK, n0, bar);
                tma_load_2d(base + 2 * A_BYTES + B_BYTES, &tm_b_lo, kb * kBK, n0, bar);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = BN, M = 128
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
            int chunk = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStg;
                mbar_wait(full0 + 8 * s, (kb / kStg) & 1);
                const uint32_t base = smem_u32(smem + s * STAGE);
                const uint64_t a_hi = umma_desc<kBK>(base), a_lo = umma_desc<kBK>(base + A_BYTES);
                const uint64_t b_hi = umma_desc<kBK>(base + 2 * A_BYTES), b_lo = umma_desc<kBK>(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
                for (int c = 0; c < Cfg::CHUNKS_PER_KB; ++c, ++chunk) {
                    const int buf = chunk & 1;
                    mbar_wait(acc_empty0 + 8 * buf, ((chunk >> 1) & 1) ^ 1);  // the epilogue has drained this buffer
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d = tmem + (uint32_t)(buf * BN);
                    // 8 TF32 = 32 bytes per MMA: +2 in the 16-byte address field of the descriptors
#pragma unroll
                    for (int kk = c * KS; kk < (c + 1) * KS; ++kk) {   // the small products first (see UmmaCfg)
                        const uint64_t o = (uint64_t)(kk * 2);
                        umma_tf32(d, a_lo + o, b_hi + o, idesc, kk != c * KS);
                        umma_tf32(d, a_hi + o, b_lo + o, idesc, 1);
                    }
#pragma unroll
                    for (int kk = c * KS; kk < (c + 1) * KS; ++kk) {
                        const uint64_t o = (uint64_t)(kk * 2);
                        umma_tf32(d, a_hi + o, b_hi + o, idesc, 1);
                    }
                    umma_commit(acc_full0 + 8 * buf);  // chunk sum complete -> epilogue
                }
                umma_commit(empty0 + 8 * s);  // the stage is free once these MMAs have read it
            }
        }
    } else {  // ---- epilogue warps: TMEM lane quarter = warp % 4 (hardware rule), column group = (warp - 2) / 4
        constexpr int EC = Cfg::EPI_COLS;
        const int quarter = warp & 3, cgrp = (warp - 2) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cgrp * EC);
        float acc[EC];
#pragma unroll
        for (int j = 0; j < EC; ++j) acc[j] = 0.f;
        const int n_chunks = num_kb * Cfg::CHUNKS_PER_KB;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int buf = chunk & 1;
            mbar_wait(acc_full0 + 8 * buf, (chunk >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t src = lane_base + (uint32_t)(buf * BN);
            float v0[32], v1[32];
            tmem_ld32_nowait(src, v0);
#pragma unroll
            for (int c0 = 0; c0 < EC; c0 += 64) {
                tmem_ld_wait();
                if (c0 + 32 < EC) tmem_ld32_nowait(src + (uint32_t)(c0 + 32), v1);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c0 + j] += v0[j];
                if (c0 + 32 < EC) {
                    tmem_ld_wait();
                    if (c0 + 64 < EC) tmem_ld32_nowait(src + (uint32_t)(c0 + 64), v0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + 32 + j] += v1[j];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty0 + 8 * buf);
        }
        const int r = m0 + quarter * 32 + lane;
        if (r < rows) {
            float* dst_row = lam_pa + (size_t)r * ld_pa + n0 + cgrp * EC;
#pragma unroll
            for (int j = 0; j < EC; j += 4) {
                const int n = n0 + cgrp * EC + j;
                float* d = dst_row + j;
                if (n + 3 < N) {  // ld_pa and n are multiples of 4: 16 B aligned
                    float4 x = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                    if (accumulate) {
                        const float4 o = *reinterpret_cast<const float4*>(d);
                        x.x *= o.x; x.y *= o.y; x.z *= o.z; x.w *= o.w;
                    }
                    *reinterpret_cast<float4*>(d) = x;
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (n + t < N) d[t] = accumulate ? d[t] * acc[j + t] : acc[j + t];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-CTA variant (cta_group::2): a cluster of two CTAs on an SM pair computes a 256-query x 256-state tile.
//
// Why: the single-CTA kernel above is bound by SHARED MEMORY bandwidth, not by the tensor pipe (ncu: tensor 56 %).
// Each of the three products of a k-step re-reads A (4 KB) and B (8 KB) from shared memory -- 96 B/clk of UMMA
// reads next to 62 B/clk of TMA writes, against 128 B/clk.  With cta_group::2 every CTA stages its own 128 rows of A
// and only HALF of the B tile (128 of the 256 parent states); the instruction reads both halves across the pair,
// so per CTA the UMMA reads drop to 64 B/clk and the TMA writes to 43 B/clk.
//
//   * both CTAs run the TMA producer for their own operands; every load completes on the LEADER's (rank 0) full
//     barrier (cp.async.bulk.tensor ... .cta_group::2 with the barrier address mapped to rank 0);
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 (M = 256: rows 0-127 accumulate in the leader's TMEM,
//     rows 128-255 in the peer's) and signals with tcgen05.commit ... .multicast::cluster to BOTH CTAs: empty[s]
//     frees the stage in both producers, acc_full[b] releases a chunk sum to both epilogues;
//   * each CTA's epilogue warps drain their own TMEM into RN registers (same chunked accumulation as above) and
//     arrive on the leader's acc_empty[b] (remote mbarrier arrive for the peer).
constexpr int k2S_BN = 256, k2S_BK = 32, k2S_STAGES = 3, k2S_EC = 128, k2S_EPI_WARPS = 8, k2S_THREADS = 64 + 32 * k2S_EPI_WARPS;
constexpr int k2S_A_BYTES = kBM * k2S_BK * 4, k2S_BH_BYTES = (k2S_BN / 2) * k2S_BK * 4;
constexpr int k2S_STAGE = 2 * k2S_A_BYTES + 2 * k2S_BH_BYTES;   // per CTA: 64 KB
constexpr size_t k2S_SMEM = (size_t)k2S_STAGES * k2S_STAGE + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Remote arrive on the leader's barrier.  RELAXED: the only data the arrival publishes are TMEM reads that have
// completed (tcgen05.wait::ld) and are ordered by tcgen05.fence::before_thread_sync; a release at cluster scope costs
// a full memory barrier per chunk and per warp (ncu: `membar` was the top stall of the first version).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {  // arrives on `bar` (same offset) in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

template <int KS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2S_THREADS, 1)
k2_umma2_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, int rows, int N,
                int num_kb, int n_col_tiles, int n_row_tiles, float* __restrict__ lam_pa, int ld_pa, int accumulate) {
    constexpr int BN = k2S_BN, kBK = k2S_BK, A_BYTES = k2S_A_BYTES, B_BYTES = k2S_BH_BYTES, STAGE = k2S_STAGE, kStg = k2S_STAGES;
    constexpr int CHUNKS_PER_KB = (kBK / 8) / KS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStg * STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStg + 4);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStg);
    const uint32_t acc_full0 = smem_u32(bars + 2 * kStg), acc_empty0 = smem_u32(bars + 2 * kStg + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    // tile order as in the single-CTA kernel, in units of 256-row pair tiles (n_row_tiles counts those)
    int tile = blockIdx.x >> 1;
    const int group_cols = 8;
    const int group = tile / (group_cols * n_row_tiles);
    const int first_col = group * group_cols;
    const int cols_here = n_col_tiles - first_col < group_cols ? n_col_tiles - first_col : group_cols;
    tile -= group * group_cols * n_row_tiles;
    const int m0 = (tile / cols_here) * (2 * kBM) + (int)rank * kBM;  // this CTA's 128 rows
    const int n0 = (first_col + tile % cols_here) * BN;                // the pair's 256 parent states
    const int nb0 = n0 + (int)rank * (BN / 2);                         // this CTA's half of the B tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStg; ++s) {
            mbar_init(full0 + 8 * s, 1);     // used in the leader: one arrive.expect_tx for the bytes of BOTH CTAs
            mbar_init(empty0 + 8 * s, 1);    // one multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full0 + 8 * b, 1);                      // one multicast commit
            mbar_init(acc_empty0 + 8 * b, 2 * k2S_EPI_WARPS);     // used in the leader: epilogue warps of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // both CTAs, same warp: 2 x 256 accumulator columns = all of TMEM
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // the peer's barriers exist before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer (both CTAs)
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStg;
                mbar_wait(empty0 + 8 * s, ((kb / kStg) & 1) ^ 1);
                const uint32_t base = smem_u32(smem + s * STAGE);
                const uint32_t bar = mapa_rank(full0 + 8 * s, 0);
                if (leader) mbar_expect_tx(full0 + 8 * s, 2 * STAGE);
                tma_load_2d_2sm(base, &tm_a_hi, kb * kBK, m0, bar);
                tma_load_2d_2sm(base + A_BYTES, &tm_a_lo, kb * kBK, m0, bar);
                tma_load_2d_2sm(base + 2 * A_BYTES, &tm_b_hi, kb * kBK, nb0, bar);
                tma_load_2d_2sm(base + 2 * A_BYTES + B_BYTES, &tm_b_lo, kb * kBK, nb0, bar);
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {  // ---- MMA issuer (leader only)
            // D = F32, A = B = TF32, K-major, N = 256, M = 256 (two CTAs x 128)
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * kBM) >> 4) << 24);
            int chunk = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStg;
                mbar_wait(full0 + 8 * s, (kb / kStg) & 1);
                const uint32_t base = smem_u32(smem + s * STAGE);
                const uint64_t a_hi = umma_desc<kBK>(base), a_lo = umma_desc<kBK>(base + A_BYTES);
                const uint64_t b_hi = umma_desc<kBK>(base + 2 * A_BYTES), b_lo = umma_desc<kBK>(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
                for (int c = 0; c < CHUNKS_PER_KB; ++c, ++chunk) {
                    const int buf = chunk & 1;
                    mbar_wait(acc_empty0 + 8 * buf, ((chunk >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d = tmem + (uint32_t)(buf * BN);
#pragma unroll
                    for (int kk = c * KS; kk < (c + 1) * KS; ++kk) {
                        const uint64_t o = (uint64_t)(kk * 2);
                        umma_tf32_2sm(d, a_lo + o, b_hi + o, idesc, kk != c * KS);
                        umma_tf32_2sm(d, a_hi + o, b_lo + o, idesc, 1);
                    }
#pragma unroll
                    for (int kk = c * KS; kk < (c + 1) * KS; ++kk) {
                        const uint64_t o = (uint64_t)(kk * 2);
                        umma_tf32_2sm(d, a_hi + o, b_hi + o, idesc, 1);
                    }
                    umma_commit_2sm(acc_full0 + 8 * buf);
                }
                umma_commit_2sm(empty0 + 8 * s);
            }
        }
    } else {  // ---- epilogue warps (both CTAs): lane quarter = warp % 4, column group = (warp - 2) / 4
        constexpr int EC = k2S_EC;
        const int quarter = warp & 3, cgrp = (warp - 2) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cgrp * EC);
        float acc[EC];
#pragma unroll
        for (int j = 0; j < EC; ++j) acc[j] = 0.f;
        const int n_chunks = num_kb * CHUNKS_PER_KB;
        const uint32_t leader_acc_empty = mapa_rank(acc_empty0, 0);
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int buf = chunk & 1;
            mbar_wait(acc_full0 + 8 * buf, (chunk >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t src = lane_base + (uint32_t)(buf * BN);
            float v0[32], v1[32];
            tmem_ld32_nowait(src, v0);
#pragma unroll
            for (int c0 = 0; c0 < EC; c0 += 64) {
                tmem_ld_wait();
                tmem_ld32_nowait(src + (uint32_t)(c0 + 32), v1);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c0 + j] += v0[j];
                tmem_ld_wait();
                if (c0 + 64 < EC) tmem_ld32_nowait(src + (uint32_t)(c0 + 64), v0);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c0 + 32 + j] += v1[j];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(acc_empty0 + 8 * buf);  // local, CTA scope
                else mbar_arrive_cluster(leader_acc_empty + 8 * buf);
            }
        }
        const int r = m0 + quarter * 32 + lane;
        if (r < rows) {
            float* dst_row = lam_pa + (size_t)r * ld_pa + n0 + cgrp * EC;
#pragma unroll
            for (int j = 0; j < EC; j += 4) {
                const int n = n0 + cgrp * EC + j;
                float* d = dst_row + j;
                if (n + 3 < N) {
                    float4 x = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                    if (accumulate) {
                        const float4 o = *reinterpret_cast<const float4*>(d);
                        x.x *= o.x; x.y *= o.y; x.z *= o.z; x.w *= o.w;
                    }
                    *reinterpret_cast<float4*>(d) = x;
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (n + t < N) d[t] = accumulate ? d[t] * acc[j + t] : acc[j + t];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other can still signal it or read its operands
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// Lambda_v: apply the range mask of column v, clear the padding columns, split into hi (in place) / lo
template <int FMT>
__global__ void __launch_bounds__(256) k2_split_kernel(const uint8_t* __restrict__ desc, size_t dstride, size_t q0, int rows,
                                                       int v, int card, float* __restrict__ hi_io, float* __restrict__ lo_out,
                                                       int ld) {
    const int ld4 = ld >> 2;
    const long long total = (long long)rows * ld4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / ld4), c = (int)(i - (long long)b * ld4) * 4;
        const uint8_t* row = desc + (q0 + b) * dstride;
        int lo, hi;
        if (FMT == BC_DESC_RANGE_U8) {
            const uint16_t p = reinterpret_cast<const uint16_t*>(row)[v];
            lo = p & 0xff; hi = p >> 8;
        } else {
            const uint32_t p = reinterpret_cast<const uint32_t*>(row)[v];
            lo = p & 0xffff; hi = p >> 16;
        }
        if (hi > card - 1) hi = card - 1;
        float4* ph = reinterpret_cast<float4*>(hi_io + (size_t)b * ld + c);
        float4 x = *ph;
        float xs[4] = {x.x, x.y, x.z, x.w}, hs[4], ls[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float val = (c + t >= lo && c + t <= hi) ? xs[t] : 0.f;  // also clears c >= card (uninitialised)
            split_tf32(val, hs[t], ls[t]);
        }
        *ph = make_float4(hs[0], hs[1], hs[2], hs[3]);
        *reinterpret_cast<float4*>(lo_out + (size_t)b * ld + c) = make_float4(ls[0], ls[1], ls[2], ls[3]);
    }
}

// T_v [K x N] (row stride) -> T^T split [N x ldk], zero padded
__global__ void k2_transpose_split_kernel(const float* __restrict__ T, int K, int N, int stride, float* __restrict__ hi,
                                          float* __restrict__ lo, int ldk) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int k = k0 + j, n = n0 + threadIdx.x;
        tile[j][threadIdx.x] = (k < K && n < N) ? T[(size_t)k * stride + n] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int n = n0 + j, k = k0 + threadIdx.x;
        if (n < N && k < ldk) {
            const float x = tile[threadIdx.x][j];
            float h, l;
            split_tf32(x, h, l);
            hi[(size_t)n * ldk + k] = h;
            lo[(size_t)n * ldk + k] = l;
        }
    }
}

int encode_map(EncodeTiledFn fn, CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
               uint32_t box_rows, int bk) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld_elems * 4};
    cuuint32_t box[2] = {(cuuint32_t)bk, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        bc_set_error("cuTensorMapEncodeTiled failed (%d) for a %llu x %llu fp32 tensor", (int)r, (unsigned long long)outer,
                     (unsigned long long)inner);
        return BC_ECUDA;
    }
    return BC_OK;
}

template <class Cfg>
int launch_edge(EncodeTiledFn fn, const float* a_hi, const float* a_lo, int ld_a, const float* b_hi, const float* b_lo, int ldk,
                int rows, int N, float* lam_pa, int ld_pa, int accumulate, cudaStream_t st) {
    static bool attr_set[64] = {};  // per device (the attribute lives in the device's context) and instantiation
    int dev = 0;
    BC_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        BC_CUDA_CHECK(cudaFuncSetAttribute(k2_umma_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    CUtensorMap maps[4];
    int rc;
    if ((rc = encode_map(fn, &maps[0], a_hi, (uint64_t)ldk, (uint64_t)rows, (uint64_t)ld_a, kBM, Cfg::BK))) return rc;
    if ((rc = encode_map(fn, &maps[1], a_lo, (uint64_t)ldk, (uint64_t)rows, (uint64_t)ld_a, kBM, Cfg::BK))) return rc;
    if ((rc = encode_map(fn, &maps[2], b_hi, (uint64_t)ldk, (uint64_t)N, (uint64_t)ldk, (uint32_t)Cfg::BN, Cfg::BK))) return rc;
    if ((rc = encode_map(fn, &maps[3], b_lo, (uint64_t)ldk, (uint64_t)N, (uint64_t)ldk, (uint32_t)Cfg::BN, Cfg::BK))) return rc;
    const int n_col_tiles = (N + Cfg::BN - 1) / Cfg::BN, n_row_tiles = (rows + kBM - 1) / kBM;
    k2_umma_kernel<Cfg><<<n_col_tiles * n_row_tiles, Cfg::THREADS, Cfg::SMEM, st>>>(
        maps[0], maps[1], maps[2], maps[3], rows, N, ldk / Cfg::BK, n_col_tiles, n_row_tiles, lam_pa, ld_pa, accumulate);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

template <int KS>
int launch_edge_2sm(EncodeTiledFn fn, const float* a_hi, const float* a_lo, int ld_a, const float* b_hi, const float* b_lo, int ldk,
                    int rows, int N, float* lam_pa, int ld_pa, int accumulate, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    BC_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        BC_CUDA_CHECK(cudaFuncSetAttribute(k2_umma2_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2S_SMEM));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    CUtensorMap maps[4];
    int rc;
    if ((rc = encode_map(fn, &maps[0], a_hi, (uint64_t)ldk, (uint64_t)rows, (uint64_t)ld_a, kBM, k2S_BK))) return rc;
    if ((rc = encode_map(fn, &maps[1], a_lo, (uint64_t)ldk, (uint64_t)rows, (uint64_t)ld_a, kBM, k2S_BK))) return rc;
    if ((rc = encode_map(fn, &maps[2], b_hi, (uint64_t)ldk, (uint64_t)N, (uint64_t)ldk, (uint32_t)(k2S_BN / 2), k2S_BK))) return rc;
    if ((rc = encode_map(fn, &maps[3], b_lo, (uint64_t)ldk, (uint64_t)N, (uint64_t)ldk, (uint32_t)(k2S_BN / 2), k2S_BK))) return rc;
    const int n_col_tiles = (N + k2S_BN - 1) / k2S_BN, n_pair_tiles = (rows + 2 * kBM - 1) / (2 * kBM);
    k2_umma2_kernel<KS><<<2 * n_col_tiles * n_pair_tiles, k2S_THREADS, k2S_SMEM, st>>>(
        maps[0], maps[1], maps[2], maps[3], rows, N, ldk / k2S_BK, n_col_tiles, n_pair_tiles, lam_pa, ld_pa, accumulate);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

int umma_prepare(bc_model* m) {
    BcK2Plan* k2 = m->k2;
    if (k2->umma) return k2->umma->failed ? BC_ELIMIT : BC_OK;
    BcUmmaPlan* u = new BcUmmaPlan();
    k2->umma = u;
    u->d_tt_hi.assign(m->n, nullptr);
    u->d_tt_lo.assign(m->n, nullptr);
    u->ldk.assign(m->n, 0);
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &u->encode, cudaEnableDefault, &q) != cudaSuccess || !u->encode ||
        q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        u->failed = 1;  // driver without tensor maps: K2 stays on its SIMT GEMM
        return BC_ELIMIT;
    }
    for (int v = 1; v < m->n; ++v) {
        if (!k2->is_internal[v]) continue;
        const BcNodeRec& nd = m->nodes[v];
        const int K = nd.card, N = nd.card_pa;
        if (K < 64 || N < 16) continue;  // tiny edges stay on the SIMT kernel
        const int ldk = (int)bc_round_up(K, kLdkAlign);
        u->ldk[v] = ldk;
        BC_CUDA_CHECK(cudaMalloc(&u->d_tt_hi[v], (size_t)N * ldk * 4));
        BC_CUDA_CHECK(cudaMalloc(&u->d_tt_lo[v], (size_t)N * ldk * 4));
        dim3 grid((ldk + 31) / 32, (N + 31) / 32), block(32, 8);
        k2_transpose_split_kernel<<<grid, block>>>(m->d_arena + nd.cpt_off, K, N, nd.stride, u->d_tt_hi[v], u->d_tt_lo[v], ldk);
        BC_CUDA_CHECK(cudaGetLastError());
        bc_count_launch();
    }
    BC_CUDA_CHECK(cudaDeviceSynchronize());
    return BC_OK;
}

}  // namespace

int bc_k2_umma_edge(bc_model* m, const uint8_t* desc, size_t dstride, int fmt, size_t q0, int rows, int v, float* lam_v,
                    float* lam_lo, int ld_v, float* lam_pa, int ld_pa, int accumulate, cudaStream_t st) {
    {   // the caller (bc_k2_launch) holds m->k2_mu
        int rc = umma_prepare(m);
        if (rc) return rc;
    }
    BcUmmaPlan* u = m->k2->umma;
    if (!u->d_tt_hi[v]) return BC_ELIMIT;
    const BcNodeRec& nd = m->nodes[v];
    const int K = nd.card, N = nd.card_pa, ldk = u->ldk[v];
    if (ld_v < ldk) return BC_ELIMIT;
    const long long total = (long long)rows * (ld_v / 4);
    long long sgrid = (total + 255) / 256;
    if (sgrid > (long long)m->sm_count * 16) sgrid = (long long)m->sm_count * 16;
    if (fmt == BC_DESC_RANGE_U8)
        k2_split_kernel<BC_DESC_RANGE_U8><<<(int)sgrid, 256, 0, st>>>(desc, dstride, q0, rows, v, K, lam_v, lam_lo, ld_v);
    else
        k2_split_kernel<BC_DESC_RANGE_U16><<<(int)sgrid, 256, 0, st>>>(desc, dstride, q0, rows, v, K, lam_v, lam_lo, ld_v);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(u->encode);
    // Variant: the widest tile the edge fills.  Letters select the same kernels by hand (BC_K2_UMMA_VARIANT, for
    // the sweeps in profiles/): A = 128x256 tile, 32-float stages (2 stages); B = 128x256, 16-float stages (4);
    // C = 128x128, 16-float stages (6); D = 128x128, 32-float stages (3).  Measured (profiles/r1_k2_umma_v3_variants.txt):
    // B == A at 1000 bins (L2 resident operands) and 19 % slower at 10k bins (64-byte rows from HBM), so the deeper
    // pipeline is not what the kernel waits for -- ncu puts the tensor pipe at 56 % with both operands read from
    // shared memory 3x per k-step (96 B/clk of UMMA reads + 62 B/clk of TMA writes against 128 B/clk of shared memory).
    char variant = N > 128 ? 'A' : N > 64 ? 'D' : N > 32 ? 'E' : 'F';
    if (const char* e = std::getenv("BC_K2_UMMA_VARIANT"))
        if (((*e >= 'A' && *e <= 'D') || *e == 'T') && N > 64) variant = *e;
    const float *bh = u->d_tt_hi[v], *bl = u->d_tt_lo[v];
    switch (variant) {
        case 'T': {
            const char* ks = std::getenv("BC_K2_UMMA_KS");
            if (ks && std::atoi(ks) == 4)
                return launch_edge_2sm<4>(fn, lam_v, lam_lo, ld_v, bh, bl, ldk, rows, N, lam_pa, ld_pa, accumulate, st);
            return launch_edge_2sm<2>(fn, lam_v, lam_lo, ld_v, bh, bl, ldk, rows, N, lam_pa, ld_pa, accumulate, st);
        }
        case 'A': return launch_edge<UmmaCfg<256, 32, 2, 128, 1>>(fn, lam_v, lam_lo, ld_v, bh, bl, ldk, rows, N, lam_pa, ld_pa, accumulate, st);
        case 'B': return launch_edge<UmmaCfg<256, 16, 2, 128, 1>>(fn, lam_v, lam_lo, ld_v, bh, bl, ldk, rows, N, lam_pa, ld_pa, accumulate, st);
        case 'C': return launch_edge<UmmaCfg<128, 16, 2, 128, 1>>(fn, lam_v, lam_lo, ld_v, bh, bl, ldk, rows, N, lam_pa, ld_pa, accumulate, st);
        case 'D': return launch_edge<UmmaCfg<128, 32, 2, 128, 1>>(fn, lam_v, lam_lo, ld_v, bh, bl, ldk, rows, N, lam_pa, ld_pa, accumulate, st);
        case 'E': return launch_edge<UmmaCfg<64, 32, 2, 64, 1>>(fn, lam_v, lam_lo, ld_v, bh, bl, ldk, rows, N, lam_pa, ld_pa, accumulate, st);
        default: return launch_edge<UmmaCfg<32, 32, 2, 32, 1>>(fn, lam_v, lam_lo, ld_v, bh, bl, ldk, rows, N, lam_pa, ld_pa, accumulate, st);
    }
}

void bc_k2_umma_free(bc_model* m) {
    if (!m->k2 || !m->k2->umma) return;
    for (float* p : m->k2->umma->d_tt_hi) cudaFree(p);
    for (float* p : m->k2->umma->d_tt_lo) cudaFree(p);
    delete m->k2->umma;
    m->k2->umma = nullptr;
}
