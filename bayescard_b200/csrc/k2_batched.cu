// K2 -- batched per-edge path for LARGE domains (BASELINE.json config 4: 10-100 columns, 10-10k bins).
//
// For domains of hundreds to thousands of states neither K-spec (code size = number of CPT entries)
// nor K1 (re-reads every CPT once per query) is the right shape.  Over a TILE of queries the
// per-edge step of the sum-product (reference Pgmpy/inference/ExactInference.py:157-177, one np.dot
// per query) is a real dense contraction
//
//     M_v[b, p] = sum_c (w_v[b, c] * Lambda_v[b, c]) * T_v[c, p]          (tile x card_v) . (card_v x card_pa)
//     Lambda_pa[b, p] *= M_v[b, p]
//
// and the CPT is reused by every query of the tile.  Three device steps, all stream ordered:
//
//   leaf edges      w_v is a range [lo, hi] and Lambda_v = 1, so M_v[b, p] = S_v[hi+1, p] - S_v[lo, p] with
//                   S_v the column prefix sums of T_v, kept in fp64 (a difference of two prefix sums in
//                   fp32 would lose the 1e-5 budget on narrow ranges).  O(tile x card_pa), HBM bound.
//   internal edges  the GEMM above; the range mask of v is applied while the A tile is staged.
//                   FP32 SIMT kernel here (128 x BN x 8 tiles, 8 x TN register blocks); the tcgen05
//                   3xTF32 kernel (k2_umma.cu) replaces it for large aligned domains.
//   root            out[b] = sum_c w_0[b, c] * Lambda_0[b, c] * T_0[c].
//
// The first message into a node WRITES Lambda (no fill pass); later ones multiply in place.
#include "bc_internal.h"

namespace {

template <int FMT>
__device__ __forceinline__ void load_range(const uint8_t* __restrict__ desc, size_t stride, size_t q, int v, int card,
                                           int& lo, int& hi) {
    const uint8_t* row = desc + q * stride;
    if (FMT == BC_DESC_RANGE_U8) {
        const uint16_t p = reinterpret_cast<const uint16_t*>(row)[v];
        lo = p & 0xff;
        hi = p >> 8;
    } else {
        const uint32_t p = reinterpret_cast<const uint32_t*>(row)[v];
        lo = p & 0xffff;
        hi = p >> 16;
    }
    if (hi > card - 1) hi = card - 1;
}

// ---- prefix sums of one CPT, column-wise, fp64: S[k][p] = sum_{c<k} T[c][p], (card+1) x card_pa
__global__ void k2_prefix_kernel(const float* __restrict__ T, int card, int card_pa, int stride, double* __restrict__ S) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= card_pa) return;
    double acc = 0.0;
    S[p] = 0.0;
    for (int c = 0; c < card; ++c) {
        acc += (double)T[(size_t)c * stride + p];
        S[(size_t)(c + 1) * card_pa + p] = acc;
    }
}

// ---- leaf edge: Lambda_pa[b][p] (*)= S[hi+1][p] - S[lo][p]
template <int FMT>
__global__ void __launch_bounds__(256) k2_leaf_kernel(const uint8_t* __restrict__ desc, size_t dstride, size_t q0,
                                                      int rows, int v, int card, int card_pa,
                                                      const double* __restrict__ S, const float* __restrict__ T, int ld_t,
                                                      float* __restrict__ lam_pa, int ld_pa, int accumulate) {
    const long long total = (long long)rows * card_pa;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / card_pa), p = (int)(i - (long long)b * card_pa);
        int lo, hi;
        load_range<FMT>(desc, dstride, q0 + b, v, card, lo, hi);
        float* dst = lam_pa + (size_t)b * ld_pa + p;
        if (lo == 0 && hi == card - 1) {  // unconstrained leaf: the columns of T sum to 1
            if (!accumulate) *dst = 1.f;
            continue;
        }
        float val = 0.f;
        if (hi - lo < 16) {
            // narrow ranges (equality predicates above all) are summed directly: a difference of two prefix sums carries their
            // ABSOLUTE rounding error (~1e-16), which on a rare state (T ~ 1e-12) is 1e-4 relative
            double acc = 0.0;
            for (int c = lo; c <= hi; ++c) acc += (double)T[(size_t)c * ld_t + p];
            val = (float)acc;
        } else {
            val = (float)(S[(size_t)(hi + 1) * card_pa + p] - S[(size_t)lo * card_pa + p]);
        }
        *dst = accumulate ? *dst * val : val;
    }
}

// ---- internal edge, FP32 SIMT:  C = mask(A) . T ;  Lambda_pa (*)= C
// A: rows x K (row-major, ld_a), T: K x N (row-major, ld_t), Lambda_pa: rows x N (ld_pa)
template <int FMT, int BN, int TN>
__global__ void __launch_bounds__(256) k2_gemm_simt_kernel(const uint8_t* __restrict__ desc, size_t dstride, size_t q0,
                                                           int rows, int v, int K, int N,
                                                           const float* __restrict__ A, int ld_a,
                                                           const float* __restrict__ T, int ld_t,
                                                           float* __restrict__ lam_pa, int ld_pa, int accumulate) {
    constexpr int BM = 128, BK = 8, TM = 8;
    constexpr int TX = BN / TN;             // threads along N
    static_assert(TX * (BM / TM) == 256, "256 threads per CTA");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];
    __shared__ int s_lo[BM], s_hi[BM];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (tid < BM) {
        int lo = 1, hi = 0;
        if (m0 + tid < rows) load_range<FMT>(desc, dstride, q0 + m0 + tid, v, K, lo, hi);
        s_lo[tid] = lo;
        s_hi[tid] = hi;
    }
    __syncthreads();
    const int tx = tid % TX, ty = tid / TX;  // thread tile: rows ty*8.., cols tx*TN..
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // global -> shared mapping.  A tile: 128 rows x 8 k = 256 float4 (2 per row); B tile: 8 k x BN.
    const int a_row = tid >> 1, a_k = (tid & 1) * 4;
    const int lo_r = s_lo[a_row], hi_r = s_hi[a_row];
    const bool a_ok = m0 + a_row < rows;
    const float* a_src = A + (size_t)(m0 + a_row) * ld_a + a_k;
    auto load_a = [&](int k0, float4& r) {
        r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_ok && k0 + a_k < K) {
            // ld_a is a multiple of 4 and rows are 16 B aligned; the padding beyond K is masked below
            r = *reinterpret_cast<const float4*>(a_src + k0);
            const int c = k0 + a_k;
            if (c + 0 < lo_r || c + 0 > hi_r) r.x = 0.f;
            if (c + 1 < lo_r || c + 1 > hi_r) r.y = 0.f;
            if (c + 2 < lo_r || c + 2 > hi_r) r.z = 0.f;
            if (c + 3 < lo_r || c + 3 > hi_r) r.w = 0.f;
        }
    };
    auto store_a = [&](int buf, const float4& r) {
        As[buf][a_k + 0][a_row] = r.x;
        As[buf][a_k + 1][a_row] = r.y;
        As[buf][a_k + 2][a_row] = r.z;
        As[buf][a_k + 3][a_row] = r.w;
    };
    constexpr int B_ELEMS = BK * BN / 256;  // floats per thread of the B tile (4 for BN=128, 1 for BN=32)
    float b_reg[B_ELEMS];
    auto load_b = [&](int k0) {
#pragma unroll
        for (int e = 0; e < B_ELEMS; ++e) {
            const int idx = tid + e * 256, k = idx / BN, n = idx % BN;
            b_reg[e] = (k0 + k < K && n0 + n < N) ? __ldg(T + (size_t)(k0 + k) * ld_t + n0 + n) : 0.f;
        }
    };
    auto store_b = [&](int buf) {
#pragma unroll
        for (int e = 0; e < B_ELEMS; ++e) {
            const int idx = tid + e * 256;
            Bs[buf][idx / BN][idx % BN] = b_reg[e];
        }
    };

    float4 a_reg;
    load_a(0, a_reg);
    load_b(0);
    store_a(0, a_reg);
    store_b(0);
    __syncthreads();
    const int nk = (K + BK - 1) / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            load_a((kt + 1) * BK, a_reg);
            load_b((kt + 1) * BK);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[buf][k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_a(buf ^ 1, a_reg);
            store_b(buf ^ 1);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = m0 + ty * TM + i;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int c = n0 + tx * TN + j;
            if (c >= N) continue;
            float* dst = lam_pa + (size_t)r * ld_pa + c;
            *dst = accumulate ? *dst * acc[i][j] : acc[i][j];
        }
    }
}

// ---- root: out[b] = sum_c mask * Lambda_0[b][c] * T_0[c]   (Lambda_0 == nullptr: the root is the only node)
template <int FMT>
__global__ void __launch_bounds__(256) k2_root_kernel(const uint8_t* __restrict__ desc, size_t dstride, size_t q0,
                                                      int rows, int card, const float* __restrict__ lam, int ld,
                                                      const float* __restrict__ T0, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b = warp; b < rows; b += nwarps) {
        int lo, hi;
        load_range<FMT>(desc, dstride, q0 + b, 0, card, lo, hi);
        float acc = 0.f;
        for (int c = lo + lane; c <= hi; c += 32) acc = fmaf(lam ? lam[(size_t)b * ld + c] : 1.f, T0[c], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[q0 + b] = acc;
    }
}

// Scaled results (fp32 range): after an edge has been multiplied into Lambda_pa every row is renormalised to a maximum in
// [1, 2) by an exact power of two, the exponent is accumulated per query.  One warp per row.
__global__ void __launch_bounds__(256) k2_rownorm_kernel(float* __restrict__ lam, int ld, int rows, int card, int32_t* __restrict__ exps) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b = warp; b < rows; b += nwarps) {
        float* row = lam + (size_t)b * ld;
        float mx = 0.f;
        for (int c = lane; c < card; c += 32) mx = fmaxf(mx, fabsf(row[c]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (!(mx > 0.f) || !isfinite(mx)) continue;
        const int e = ilogbf(mx);
        if (e == 0) continue;
        const float s = ldexpf(1.f, -e > 126 ? 126 : -e), s2 = -e > 126 ? ldexpf(1.f, -e - 126) : 1.f;
        for (int c = lane; c < card; c += 32) row[c] = row[c] * s * s2;
        if (lane == 0) exps[b] += e;
    }
}

template <int FMT>
int run_tile(bc_model* m, const uint8_t* desc, size_t dstride, size_t q0, int rows, size_t tile_rows, float* out,
             float* const* lam, const int* ld, int use_umma, cudaStream_t st, int32_t* out_exp) {
    const int n = m->n;
    std::vector<char> written(n, 0);
    for (int v = n - 1; v >= 1; --v) {
        const BcNodeRec& nd = m->nodes[v];
        const int pa = nd.parent;
        const int acc = written[pa];
        if (!m->k2->is_internal[v]) {
            const long long total = (long long)rows * nd.card_pa;
            long long grid = (total + 255) / 256;
            if (grid > (long long)m->sm_count * 16) grid = (long long)m->sm_count * 16;
            k2_leaf_kernel<FMT><<<(int)grid, 256, 0, st>>>(desc, dstride, q0, rows, v, nd.card, nd.card_pa, m->k2->d_prefix[v],
                                                           m->d_arena + nd.cpt_off, nd.stride, lam[pa], ld[pa], acc);
        } else {
            const float* T = m->d_arena + nd.cpt_off;
            int rc = BC_ELIMIT;
            // with use_umma the low halves of Lambda_v sit right behind the high halves in the workspace
            if (use_umma)
                rc = bc_k2_umma_edge(m, desc, dstride, FMT, q0, rows, v, lam[v], lam[v] + tile_rows * (size_t)ld[v], ld[v],
                                     lam[pa], ld[pa], acc, st);
            if (rc == BC_ELIMIT) {  // shape not served by the tensor-core kernel: FP32 SIMT
                if (nd.card_pa > 32) {
                    dim3 grid((nd.card_pa + 127) / 128, (rows + 127) / 128);
                    k2_gemm_simt_kernel<FMT, 128, 8><<<grid, 256, 0, st>>>(desc, dstride, q0, rows, v, nd.card, nd.card_pa,
                                                                        lam[v], ld[v], T, nd.stride, lam[pa], ld[pa], acc);
                } else {
                    dim3 grid((nd.card_pa + 31) / 32, (rows + 127) / 128);
                    k2_gemm_simt_kernel<FMT, 32, 2><<<grid, 256, 0, st>>>(desc, dstride, q0, rows, v, nd.card, nd.card_pa,
                                                                       lam[v], ld[v], T, nd.stride, lam[pa], ld[pa], acc);
                }
            } else if (rc != BC_OK) {
                return rc;
            }
        }
        BC_CUDA_CHECK(cudaGetLastError());
        bc_count_launch();
        written[pa] = 1;
        if (out_exp) {
            long long g = ((long long)rows * 32 + 255) / 256;
            if (g > (long long)m->sm_count * 16) g = (long long)m->sm_count * 16;
            k2_rownorm_kernel<<<(int)g, 256, 0, st>>>(lam[pa], ld[pa], rows, nd.card_pa, out_exp + q0);
            BC_CUDA_CHECK(cudaGetLastError());
            bc_count_launch();
        }
    }
    const BcNodeRec& r0 = m->nodes[0];
    long long grid = ((long long)rows * 32 + 255) / 256;
    if (grid > (long long)m->sm_count * 8) grid = (long long)m->sm_count * 8;
    k2_root_kernel<FMT><<<(int)grid, 256, 0, st>>>(desc, dstride, q0, rows, r0.card, written[0] ? lam[0] : nullptr, ld[0],
                                                   m->d_arena + r0.cpt_off, out);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

}  // namespace

void bc_k2_free(bc_model* m) {
    if (!m->k2) return;
    for (double* p : m->k2->d_prefix) cudaFree(p);
    if (m->k2->d_ws) cudaFree(m->k2->d_ws);
    if (m->k2->ws_event) cudaEventDestroy(m->k2->ws_event);
    bc_k2_umma_free(m);
    delete m->k2;
    m->k2 = nullptr;
}

static int k2_prepare(bc_model* m) {
    if (m->k2) return BC_OK;
    BcK2Plan* k = new BcK2Plan();
    k->is_internal.assign(m->n, 0);
    k->d_prefix.assign(m->n, nullptr);
    for (int v = 1; v < m->n; ++v) k->is_internal[m->nodes[v].parent] = 1;
    m->k2 = k;
    for (int v = 1; v < m->n; ++v) {
        if (k->is_internal[v]) continue;
        const BcNodeRec& nd = m->nodes[v];
        BC_CUDA_CHECK(cudaMalloc(&k->d_prefix[v], (size_t)(nd.card + 1) * nd.card_pa * sizeof(double)));
        k2_prefix_kernel<<<(nd.card_pa + 127) / 128, 128>>>(m->d_arena + nd.cpt_off, nd.card, nd.card_pa, nd.stride,
                                                          k->d_prefix[v]);
        BC_CUDA_CHECK(cudaGetLastError());
        bc_count_launch();
    }
    BC_CUDA_CHECK(cudaDeviceSynchronize());
    return BC_OK;
}

int bc_k2_launch(bc_model* m, const void* desc, size_t nq, int fmt, const uint32_t* fan_mask, float* out, int use_umma,
                 cudaStream_t stream, int32_t* out_exp) {
    if (fmt != BC_DESC_RANGE_U8 && fmt != BC_DESC_RANGE_U16) {
        bc_set_error("the batched large-domain kernel takes RANGE_U8 / RANGE_U16 descriptors");
        return BC_EINVAL;
    }
    if (fan_mask) {
        bc_set_error("the batched large-domain kernel does not implement fan-out columns");
        return BC_EINVAL;
    }
    if (nq == 0) return BC_OK;
    std::lock_guard<std::mutex> lock(m->k2_mu);  // held while the call is enqueued (the workspace is shared)
    {
        int rc = k2_prepare(m);
        if (rc) return rc;
    }
    // Workspace: Lambda matrices for a tile of queries.  Lambda_v is first written by the edge of v's highest-numbered
    // child and last read by edge v itself (edges run v = n-1 .. 1, stream ordered), so the matrices share a few
    // SLOTS by lifetime -- O(depth) of them instead of one per internal node -- and the tile can be that much taller.
    const int n = m->n;
    std::vector<int> ld(n, 0), slot(n, -1);
    int max_ld = 0;
    for (int v = 0; v < n; ++v)
        if (m->k2->is_internal[v]) {
            ld[v] = (int)bc_round_up(m->nodes[v].card, 32);  // 128 B rows: what the TMA path wants
            if (ld[v] > max_ld) max_ld = ld[v];
        }
    int n_slots = 0;
    {
        std::vector<int> free_slots;
        for (int v = n - 1; v >= 1; --v) {
            const int pa = m->nodes[v].parent;
            if (slot[pa] < 0) {
                if (free_slots.empty()) slot[pa] = n_slots++;
                else { slot[pa] = free_slots.back(); free_slots.pop_back(); }
            }
            if (slot[v] >= 0) free_slots.push_back(slot[v]);  // edge v was the last reader of Lambda_v
        }
    }
    const size_t slot_floats = (size_t)max_ld * (use_umma ? 2 : 1);  // + the low halves for 3xTF32
    const size_t floats_per_query = slot_floats * (size_t)n_slots;
    size_t budget = (size_t)32 << 30;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t have = free_b + m->k2->ws_bytes;  // what is free once the cached workspace is counted
            if (budget > have / 2) budget = have / 2;
        } else {
            cudaGetLastError();
        }
    }
    if (const char* env = std::getenv("BC_K2_WORKSPACE_MB")) {
        const long long v = std::atoll(env);
        if (v >= 16) budget = (size_t)v << 20;
    }
    size_t tile = floats_per_query ? budget / (floats_per_query * 4) : nq;
    if (tile > 262144) tile = 262144;
    if (tile >= nq) {
        tile = (size_t)bc_round_up((int64_t)nq, 128);
    } else {
        // equal tiles, each a whole number of 128-row CTA tiles and -- when tall enough -- of 148-SM waves
        const size_t wave = (size_t)m->sm_count * 128;
        tile = tile >= wave ? tile / wave * wave : (tile / 128 ? tile / 128 * 128 : 128);
        const size_t n_tiles = (nq + tile - 1) / tile;
        const size_t even = (size_t)bc_round_up((int64_t)((nq + n_tiles - 1) / n_tiles), 128);
        if (even < tile) tile = even;
    }
    BcK2Plan* k2 = m->k2;
    const size_t need = tile * floats_per_query * 4;
    if (!k2->ws_event) BC_CUDA_CHECK(cudaEventCreateWithFlags(&k2->ws_event, cudaEventDisableTiming));
    if (need > k2->ws_bytes) {
        BC_CUDA_CHECK(cudaEventSynchronize(k2->ws_event));  // nothing may still be using the old workspace
        if (k2->d_ws) cudaFree(k2->d_ws);
        k2->d_ws = nullptr;
        k2->ws_bytes = 0;
        BC_CUDA_CHECK(cudaMalloc(&k2->d_ws, need));
        k2->ws_bytes = need;
    }
    BC_CUDA_CHECK(cudaStreamWaitEvent(stream, k2->ws_event, 0));
    float* ws = k2->d_ws;
    std::vector<float*> lam(n, nullptr);
    for (int v = 0; v < n; ++v)
        if (slot[v] >= 0) lam[v] = ws + (size_t)slot[v] * tile * slot_floats;
    const size_t dstride = (size_t)bc_model_desc_stride(m, fmt);
    const uint8_t* d = static_cast<const uint8_t*>(desc);
    int rc = BC_OK;
    if (out_exp) BC_CUDA_CHECK(cudaMemsetAsync(out_exp, 0, nq * sizeof(int32_t), stream));
    for (size_t q0 = 0; q0 < nq && rc == BC_OK; q0 += tile) {
        const int rows = (int)(nq - q0 < tile ? nq - q0 : tile);
        rc = fmt == BC_DESC_RANGE_U8
                 ? run_tile<BC_DESC_RANGE_U8>(m, d, dstride, q0, rows, tile, out, lam.data(), ld.data(), use_umma, stream, out_exp)
                 : run_tile<BC_DESC_RANGE_U16>(m, d, dstride, q0, rows, tile, out, lam.data(), ld.data(), use_umma, stream, out_exp);
    }
    BC_CUDA_CHECK(cudaEventRecord(k2->ws_event, stream));
    return rc;
}
