// K1 -- generic (runtime-tree) sum-product kernel: one warp per query.
//
// Computes  out[q] = sum_x prod_v w_v[x_v] * T_v[x_v, x_pa(v)]  leaf -> root, which is what
// VariableEliminationJIT.query / .expectation compute (reference
// Pgmpy/inference/ExactInference.py:112-287) once the predicate bins / n_distinct weights /
// fan-out vectors are written as dense per-column weights w_v.
//
// Mapping
//   * persistent CTAs, W warps each; warp w takes queries w, w+total_warps, ...
//   * the CPT arena is staged ONCE per CTA into shared memory with TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx) when it fits; otherwise T is read through L1/L2.
//   * per warp a "lambda row" in shared memory: lambda_v[c] = w_v[c] * prod_{children k} m_k[c].
//   * per edge v -> pa(v):  m_v[p] = sum_c lambda_v[c] * T_v[c][p], lanes over p (coalesced,
//     conflict-free rows of T), lambda_v broadcast as float4; for card(pa) <= 16 the warp splits
//     into 2 or 4 groups over c and finishes with shuffle reductions.
//   * like the reference's Steiner-tree pruning (ExactInference.py:55-75) subtrees without any
//     constrained / fan-out column are skipped (their message is the column sums of T = 1), and
//     for range descriptors the c loop only visits [lo, hi].
#include "bc_internal.h"

namespace {

constexpr int kWarp = 32;

struct K1Params {
    const BcNodeRec* nodes;
    int n;
    const float* arena;
    unsigned arena_bytes;  // multiple of 16 (only used by the shared-memory variant)
    const float* fan;
    const uint16_t* ent_node;
    const BcBitsRec* bits;  // BITS rows only
    int lam_total;
    const uint8_t* desc;
    size_t desc_stride;  // bytes
    const uint32_t* fan_mask;
    int mask_words;
    float* out;
    int32_t* out_exp;  // scaled results: out[q] * 2^out_exp[q] (nullptr: plain fp32 results)
    size_t nq;
    // shared memory carve-up (bytes from the base, all multiples of 16)
    unsigned off_arena, off_nodes, off_ent, off_bits, off_warp, warp_bytes, off_lam, off_desc, off_act;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Scaled results (fp32 range): after a message has been multiplied into lambda_pa the row is renormalised to a maximum in
// [1, 2) -- an exact power-of-two scale -- and the exponent is carried per query.  The reference computes in fp64
// (Pgmpy/inference/ExactInference.py:157-177); ten narrow predicates on wide domains go below 1e-38.
__device__ __forceinline__ int k1_renormalise(float* row, int card, int lane) {
    float mx = 0.f;
    for (int c = lane; c < card; c += kWarp) mx = fmaxf(mx, fabsf(row[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (!(mx > 0.f) || !isfinite(mx)) return 0;
    const int e = ilogbf(mx);
    if (e == 0) return 0;
    const float s = ldexpf(1.f, -e > 126 ? 126 : -e), s2 = -e > 126 ? ldexpf(1.f, -e - 126) : 1.f;   // (2^-e itself may not be a normal float)
    for (int c = lane; c < card; c += kWarp) row[c] = row[c] * s * s2;
    return e;
}

template <int FMT, bool ARENA_SMEM>
__global__ void __launch_bounds__(512) k1_kernel(const K1Params P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* s_arena = reinterpret_cast<float*>(smem + P.off_arena);
    BcNodeRec* s_nodes = reinterpret_cast<BcNodeRec*>(smem + P.off_nodes);
    uint16_t* s_ent = reinterpret_cast<uint16_t*>(smem + P.off_ent);
    int32_t* s_bitoff = reinterpret_cast<int32_t*>(smem + P.off_bits);

    // ---- prologue: node tables by plain loads, arena by TMA bulk copies -----------------------
    if (ARENA_SMEM && threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.nodes);
        uint4* dst = reinterpret_cast<uint4*>(s_nodes);
        for (int i = threadIdx.x; i < P.n * 2; i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < P.lam_total; i += blockDim.x) s_ent[i] = P.ent_node[i];
        if (FMT == BC_DESC_BITS)
            for (int i = threadIdx.x; i < P.n; i += blockDim.x) s_bitoff[i] = P.bits[i].bit_off;
    }
    __syncthreads();
    if (ARENA_SMEM) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, P.arena_bytes);
            const unsigned chunk = 32768;
            for (unsigned o = 0; o < P.arena_bytes; o += chunk) {
                unsigned b = P.arena_bytes - o < chunk ? P.arena_bytes - o : chunk;
                tma_bulk_g2s(reinterpret_cast<unsigned char*>(s_arena) + o,
                             reinterpret_cast<const unsigned char*>(P.arena) + o, b, bar);
            }
        }
        mbar_wait(bar, 0);
    }
    const float* __restrict__ T0 = ARENA_SMEM ? s_arena : P.arena;

    unsigned char* wbase = smem + P.off_warp + (size_t)warp * P.warp_bytes;
    float* lam = reinterpret_cast<float*>(wbase + P.off_lam);
    uint32_t* sdesc32 = reinterpret_cast<uint32_t*>(wbase + P.off_desc);
    volatile unsigned char* act = wbase + P.off_act;

    const size_t total_warps = (size_t)gridDim.x * nwarps;
    for (size_t q = (size_t)blockIdx.x * nwarps + warp; q < P.nq; q += total_warps) {
        const uint32_t* fm = P.fan_mask ? P.fan_mask + q * P.mask_words : nullptr;
        // ---- 1. weights -> lambda row ------------------------------------------------------
        if (FMT == BC_DESC_DENSE_F32) {
            const float* row = reinterpret_cast<const float*>(P.desc + q * P.desc_stride);
            for (int e = lane; e < P.lam_total; e += kWarp) {
                const int v = s_ent[e];
                const BcNodeRec& nd = s_nodes[v];
                const int c = e - nd.lam_off;
                float w = (c < nd.card) ? row[e] : 0.f;  // padding entries never contribute
                if (fm && ((fm[v >> 5] >> (v & 31)) & 1u) && nd.fan_off >= 0 && c < nd.card)
                    w *= P.fan[nd.fan_off + c];
                lam[e] = w;
            }
            for (int v = lane; v < P.n; v += kWarp) act[v] = 1;
        } else if (FMT == BC_DESC_BITS) {
            const uint32_t* grow = reinterpret_cast<const uint32_t*>(P.desc + q * P.desc_stride);
            const int words = (int)(P.desc_stride >> 2);
            for (int i = lane; i < words; i += kWarp) sdesc32[i] = grow[i];
            for (int v = lane; v < P.n; v += kWarp) act[v] = 0;
            __syncwarp();
            for (int e = lane; e < P.lam_total; e += kWarp) {
                const int v = s_ent[e];
                const BcNodeRec& nd = s_nodes[v];
                const int c = e - nd.lam_off;
                const int b = s_bitoff[v] + c;
                float w = (c < nd.card && ((sdesc32[b >> 5] >> (b & 31)) & 1u)) ? 1.f : 0.f;
                if (fm && ((fm[v >> 5] >> (v & 31)) & 1u) && nd.fan_off >= 0 && c < nd.card)
                    w *= P.fan[nd.fan_off + c];
                lam[e] = w;
            }
            // Steiner pruning: a column is constrained when any of its bits is cleared
            for (int v = lane; v < P.n; v += kWarp) {
                const BcNodeRec& nd = s_nodes[v];
                bool constrained = false;
                for (int c = 0, b = s_bitoff[v]; c < nd.card; ++c, ++b)
                    constrained |= !((sdesc32[b >> 5] >> (b & 31)) & 1u);
                if (fm && ((fm[v >> 5] >> (v & 31)) & 1u) && nd.fan_off >= 0) constrained = true;
                if (constrained) {
                    // lanes climbing from different evidence nodes meet on the way up: the mark is set with an atomic OR on the
                    // byte's word (act is 16-byte aligned per warp), the first lane to arrive goes on, the others stop
                    int u = v;
                    while (u >= 0) {
                        const unsigned bit = 1u << (8 * (u & 3));
                        const unsigned old = atomicOr(reinterpret_cast<unsigned*>(const_cast<unsigned char*>(act)) + (u >> 2), bit);
                        if (old & (0xFFu << (8 * (u & 3)))) break;
                        u = s_nodes[u].parent;
                    }
                }
            }
        } else {
            const uint32_t* grow = reinterpret_cast<const uint32_t*>(P.desc + q * P.desc_stride);
            const int words = (int)(P.desc_stride >> 2);
            for (int i = lane; i < words; i += kWarp) sdesc32[i] = grow[i];
            for (int v = lane; v < P.n; v += kWarp) act[v] = 0;
            __syncwarp();
            for (int e = lane; e < P.lam_total; e += kWarp) {
                const int v = s_ent[e];
                const BcNodeRec& nd = s_nodes[v];
                const int c = e - nd.lam_off;
                int lo, hi;
                if (FMT == BC_DESC_RANGE_U8) {
                    const unsigned char* b = reinterpret_cast<const unsigned char*>(sdesc32);
                    lo = b[2 * v];
                    hi = b[2 * v + 1];
                } else {
                    const unsigned short* b = reinterpret_cast<const unsigned short*>(sdesc32);
                    lo = b[2 * v];
                    hi = b[2 * v + 1];
                }
                float w = (c >= lo && c <= hi && c < nd.card) ? 1.f : 0.f;
                if (fm && ((fm[v >> 5] >> (v & 31)) & 1u) && nd.fan_off >= 0 && c < nd.card)
                    w *= P.fan[nd.fan_off + c];
                lam[e] = w;
            }
            // Steiner pruning: mark every constrained node and its ancestors.
            for (int v = lane; v < P.n; v += kWarp) {
                const BcNodeRec& nd = s_nodes[v];
                int lo, hi;
                if (FMT == BC_DESC_RANGE_U8) {
                    const unsigned char* b = reinterpret_cast<const unsigned char*>(sdesc32);
                    lo = b[2 * v];
                    hi = b[2 * v + 1];
                } else {
                    const unsigned short* b = reinterpret_cast<const unsigned short*>(sdesc32);
                    lo = b[2 * v];
                    hi = b[2 * v + 1];
                }
                bool constrained = lo > 0 || hi < nd.card - 1;
                if (fm && ((fm[v >> 5] >> (v & 31)) & 1u) && nd.fan_off >= 0) constrained = true;
                if (constrained) {
                    // lanes climbing from different evidence nodes meet on the way up: the mark is set with an atomic OR on the
                    // byte's word (act is 16-byte aligned per warp), the first lane to arrive goes on, the others stop
                    int u = v;
                    while (u >= 0) {
                        const unsigned bit = 1u << (8 * (u & 3));
                        const unsigned old = atomicOr(reinterpret_cast<unsigned*>(const_cast<unsigned char*>(act)) + (u >> 2), bit);
                        if (old & (0xFFu << (8 * (u & 3)))) break;
                        u = s_nodes[u].parent;
                    }
                }
            }
        }
        __syncwarp();

        // ---- 2. edges in reverse topological order -----------------------------------------
        int esum = 0;
        for (int v = P.n - 1; v >= 1; --v) {
            if (!act[v]) continue;  // warp uniform
            const BcNodeRec nd = s_nodes[v];
            const float* lam_v = lam + nd.lam_off;
            float* lam_p = lam + s_nodes[nd.parent].lam_off;
            int c_lo = 0, c_hi = nd.card - 1;
            if (FMT == BC_DESC_RANGE_U8) {
                const unsigned char* b = reinterpret_cast<const unsigned char*>(sdesc32);
                c_lo = b[2 * v];
                c_hi = min((int)b[2 * v + 1], nd.card - 1);
            } else if (FMT == BC_DESC_RANGE_U16) {
                const unsigned short* b = reinterpret_cast<const unsigned short*>(sdesc32);
                c_lo = b[2 * v];
                c_hi = min((int)b[2 * v + 1], nd.card - 1);
            }
            c_lo &= ~3;
            const float* __restrict__ T = T0 + nd.cpt_off;
            const int st = nd.stride;
            const int cpa = nd.card_pa;
            if (cpa <= 8) {
                // 4 groups of 8 lanes: group g takes rows c+g of every float4 of lambda
                const int g = lane >> 3, p = lane & 7;
                float acc = 0.f;
                for (int c = c_lo; c <= c_hi; c += 4) {
                    const float4 l = *reinterpret_cast<const float4*>(lam_v + c);
                    const float lg = g == 0 ? l.x : g == 1 ? l.y : g == 2 ? l.z : l.w;
                    acc = fmaf(lg, T[(size_t)(c + g) * st + p], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 8);
                acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                if (lane < cpa) lam_p[lane] *= acc;
            } else if (cpa <= 16) {
                const int g = lane >> 4, p = lane & 15;
                float acc = 0.f;
                for (int c = c_lo; c <= c_hi; c += 4) {
                    const float4 l = *reinterpret_cast<const float4*>(lam_v + c);
                    const float l0 = g ? l.z : l.x, l1 = g ? l.w : l.y;
                    const float* r = T + (size_t)(c + 2 * g) * st + p;
                    acc = fmaf(l0, r[0], acc);
                    acc = fmaf(l1, r[st], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                if (lane < cpa) lam_p[lane] *= acc;
            } else {
                for (int p0 = 0; p0 < cpa; p0 += 64) {
                    const int pa = p0 + lane;
                    if (cpa - p0 > 32) {
                        float a0 = 0.f, a1 = 0.f;
                        for (int c = c_lo; c <= c_hi; c += 4) {
                            const float4 l = *reinterpret_cast<const float4*>(lam_v + c);
                            const float* r = T + (size_t)c * st + pa;
                            a0 = fmaf(l.x, r[0], a0);
                            a1 = fmaf(l.x, r[32], a1);
                            a0 = fmaf(l.y, r[st], a0);
                            a1 = fmaf(l.y, r[st + 32], a1);
                            a0 = fmaf(l.z, r[2 * st], a0);
                            a1 = fmaf(l.z, r[2 * st + 32], a1);
                            a0 = fmaf(l.w, r[3 * st], a0);
                            a1 = fmaf(l.w, r[3 * st + 32], a1);
                        }
                        if (pa < cpa) lam_p[pa] *= a0;
                        if (pa + 32 < cpa) lam_p[pa + 32] *= a1;
                    } else {
                        float a0 = 0.f;
                        for (int c = c_lo; c <= c_hi; c += 4) {
                            const float4 l = *reinterpret_cast<const float4*>(lam_v + c);
                            const float* r = T + (size_t)c * st + pa;
                            a0 = fmaf(l.x, r[0], a0);
                            a0 = fmaf(l.y, r[st], a0);
                            a0 = fmaf(l.z, r[2 * st], a0);
                            a0 = fmaf(l.w, r[3 * st], a0);
                        }
                        if (pa < cpa) lam_p[pa] *= a0;
                    }
                }
            }
            __syncwarp();
            if (P.out_exp != nullptr) {
                esum += k1_renormalise(lam_p, cpa, lane);
                __syncwarp();
            }
        }

        // ---- 3. root: sum_c lambda_0[c] * T_0[c] -----------------------------------------------
        {
            const BcNodeRec nd = s_nodes[0];
            const float* __restrict__ T = T0 + nd.cpt_off;
            float acc = 0.f;
            for (int c = lane; c < nd.card; c += kWarp) acc = fmaf(lam[nd.lam_off + c], T[c], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                P.out[q] = acc;
                if (P.out_exp != nullptr) P.out_exp[q] = esum;
            }
        }
        __syncwarp();
    }
}

template <int FMT>
int launch_fmt(bc_model* m, K1Params& P, bool arena_smem, int threads, int grid, size_t smem,
               cudaStream_t stream) {
    if (arena_smem) {
        BC_CUDA_CHECK(cudaFuncSetAttribute(k1_kernel<FMT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        k1_kernel<FMT, true><<<grid, threads, smem, stream>>>(P);
    } else {
        BC_CUDA_CHECK(cudaFuncSetAttribute(k1_kernel<FMT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        k1_kernel<FMT, false><<<grid, threads, smem, stream>>>(P);
    }
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

}  // namespace

int bc_k1_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask, float* out,
                 cudaStream_t stream, int32_t* out_exp) {
    if (nq == 0) return BC_OK;
    K1Params P{};
    P.nodes = m->d_nodes;
    P.n = m->n;
    P.arena = m->d_arena;
    P.fan = m->d_fan;
    P.ent_node = m->d_ent_node;
    P.bits = m->d_bits;
    P.lam_total = m->lam_total;
    P.desc = static_cast<const uint8_t*>(desc);
    P.desc_stride = (size_t)bc_model_desc_stride(m, fmt);
    P.fan_mask = fan_mask;
    P.mask_words = m->mask_words;
    P.out = out;
    P.out_exp = out_exp;
    P.nq = nq;

    const size_t arena_bytes = (size_t)bc_round_up((int64_t)m->arena_floats_padded * 4, 16);
    const size_t nodes_bytes = (size_t)bc_round_up((int64_t)m->n * sizeof(BcNodeRec), 16);
    const size_t ent_bytes = (size_t)bc_round_up((int64_t)m->lam_total * 2, 16);
    const size_t bits_bytes = fmt == BC_DESC_BITS ? (size_t)bc_round_up((int64_t)m->n * 4, 16) : 0;
    const size_t lam_bytes = (size_t)m->lam_total * 4;  // lam_total is a multiple of 4
    const size_t desc_bytes = fmt == BC_DESC_DENSE_F32 ? 0 : (size_t)bc_round_up((int64_t)P.desc_stride, 16);
    const size_t act_bytes = (size_t)bc_round_up(m->n, 16);
    const size_t warp_bytes = lam_bytes + desc_bytes + act_bytes;
    const size_t budget = (size_t)m->smem_optin;

    // choose warps per CTA and whether the arena lives in shared memory
    int warps = 8;
    const size_t fixed_smem = 16 + arena_bytes + nodes_bytes + ent_bytes + bits_bytes;
    bool arena_smem = arena_bytes < (1u << 20) && fixed_smem + 4 * warp_bytes <= budget;
    size_t fixed = arena_smem ? fixed_smem : 16 + nodes_bytes + ent_bytes + bits_bytes;
    if (arena_smem && arena_bytes > 48 * 1024) warps = 16;  // one big CTA per SM shares the arena
    while (warps > 1 && fixed + (size_t)warps * warp_bytes > budget) warps >>= 1;
    if (fixed + (size_t)warps * warp_bytes > budget) {
        bc_set_error("model too large for the generic kernel: %zu B of shared memory per warp (sum of "
                     "domain sizes %d); use BC_KERNEL_GEMM",
                     warp_bytes, m->lam_total);
        return BC_ELIMIT;
    }
    const size_t smem = fixed + (size_t)warps * warp_bytes;
    P.off_arena = 16;
    P.off_nodes = (unsigned)(arena_smem ? 16 + arena_bytes : 16);
    P.off_ent = (unsigned)(P.off_nodes + nodes_bytes);
    P.off_bits = (unsigned)(P.off_ent + ent_bytes);
    P.off_warp = (unsigned)(P.off_bits + bits_bytes);
    P.warp_bytes = (unsigned)warp_bytes;
    P.off_lam = 0;
    P.off_desc = (unsigned)lam_bytes;
    P.off_act = (unsigned)(lam_bytes + desc_bytes);
    P.arena_bytes = (unsigned)arena_bytes;

    const int threads = warps * 32;
    int ctas_per_sm = (int)((228 * 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm * threads > 2048) ctas_per_sm = 2048 / threads;
    long long grid = (long long)m->sm_count * ctas_per_sm;
    const long long needed = (long long)((nq + warps - 1) / warps);
    if (grid > needed) grid = needed;

    switch (fmt) {
        case BC_DESC_RANGE_U8: return launch_fmt<BC_DESC_RANGE_U8>(m, P, arena_smem, threads, (int)grid, smem, stream);
        case BC_DESC_RANGE_U16: return launch_fmt<BC_DESC_RANGE_U16>(m, P, arena_smem, threads, (int)grid, smem, stream);
        case BC_DESC_DENSE_F32: return launch_fmt<BC_DESC_DENSE_F32>(m, P, arena_smem, threads, (int)grid, smem, stream);
        case BC_DESC_BITS: return launch_fmt<BC_DESC_BITS>(m, P, arena_smem, threads, (int)grid, smem, stream);
    }
    bc_set_error("unknown descriptor format %d", fmt);
    return BC_EINVAL;
}
