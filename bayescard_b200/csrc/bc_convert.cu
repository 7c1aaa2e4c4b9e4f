// Descriptor conversion kernels: RANGE_U8 / RANGE_U16 rows and SPARSE (CSR) rows -> BITS rows.
//
// A BITS row is the per-query selection mask the specialised kernel consumes: one bit per
// (column, state), columns in topological order, tightly packed, row padded to 16 bytes.  It is
// the direct device form of what Bayescard_BN.query_decoding (reference Models/Bayescard_BN.py:279-325)
// returns when every n_distinct weight is 1: the set of selected bins per predicated column; a
// column without a predicate has all of its bits set.
//
// SPARSE is the compact host-facing form of the same thing (the reference's query is a sparse
// dict {column: bins}, Evaluation/cardinality_estimation.py:60-111): per query a CSR slice of 32-bit
// entries  col[0:15) | cont<<15 | lo<<16 | hi<<24 .  An entry selects states lo..hi of column
// col; cont = 1 ORs the range into the mask built so far for that column (IN lists), cont = 0
// replaces the column's mask.  Only PCIe sees this form: it is expanded to BITS rows on the device.
//
// Both kernels are pure byte/bit movers and HBM-bound: each thread owns one query, builds the row
// in a word-interleaved shared-memory tile (conflict-free: word w of thread t at [w][t]) and writes
// it back with 16-byte stores.
#include "bc_internal.h"

namespace {

__device__ __forceinline__ uint32_t bits_between(int a, int b) {  // bits a..b, 0 <= a <= b <= 31
    return (0xffffffffu >> (31 - b)) & (0xffffffffu << a);
}

// Sets states lo..hi of a column (bit_off, card) in a row held at row[w * stride]; clears the
// column's other bits first unless `keep`.
__device__ __forceinline__ void row_select(uint32_t* row, int stride, int bit_off, int card, int lo, int hi,
                                           bool keep) {
    const int n0 = bit_off, n1 = bit_off + card - 1;
    const int s0 = bit_off + lo, s1 = bit_off + (hi < card - 1 ? hi : card - 1);
    for (int w = n0 >> 5; w <= (n1 >> 5); ++w) {
        const int wb = w << 5;
        const uint32_t nm = bits_between(n0 > wb ? n0 - wb : 0, n1 - wb < 31 ? n1 - wb : 31);
        uint32_t sm = 0;
        if (s1 >= s0 && s1 >= wb && s0 <= wb + 31) sm = bits_between(s0 > wb ? s0 - wb : 0, s1 - wb < 31 ? s1 - wb : 31);
        uint32_t x = row[w * stride];
        if (!keep) x &= ~nm;
        row[w * stride] = x | (sm & nm);
    }
}

template <int FMT>  // BC_DESC_RANGE_U8 or BC_DESC_RANGE_U16
__global__ void __launch_bounds__(128) ranges_to_bits_kernel(const BcBitsRec* __restrict__ bits, int n,
                                                            const uint32_t* __restrict__ dflt, int words,
                                                            const uint8_t* __restrict__ src, size_t src_stride,
                                                            uint32_t* __restrict__ dst, size_t nq) {
    extern __shared__ uint32_t smem[];
    uint32_t* tile = smem;                                   // [words][blockDim.x]
    BcBitsRec* s_bits = reinterpret_cast<BcBitsRec*>(smem + (size_t)words * blockDim.x);
    for (int v = threadIdx.x; v < n; v += blockDim.x) s_bits[v] = bits[v];
    __syncthreads();
    const int T = blockDim.x;
    uint32_t* row = tile + threadIdx.x;
    for (size_t q = (size_t)blockIdx.x * T + threadIdx.x; q < nq; q += (size_t)gridDim.x * T) {
        for (int w = 0; w < words; ++w) row[w * T] = dflt[w];
        const uint8_t* r8 = src + q * src_stride;
        for (int v = 0; v < n; ++v) {
            int lo, hi;
            if (FMT == BC_DESC_RANGE_U8) {
                const uint16_t p = reinterpret_cast<const uint16_t*>(r8)[v];
                lo = p & 0xff;
                hi = p >> 8;
            } else {
                const uint32_t p = reinterpret_cast<const uint32_t*>(r8)[v];
                lo = p & 0xffff;
                hi = p >> 16;
            }
            const BcBitsRec b = s_bits[v];
            if (lo > 0 || hi < b.card - 1) row_select(row, T, b.bit_off, b.card, lo, hi, false);
        }
        uint4* out = reinterpret_cast<uint4*>(dst + q * words);
        for (int w = 0; w < words; w += 4) out[w >> 2] = make_uint4(row[w * T], row[(w + 1) * T], row[(w + 2) * T], row[(w + 3) * T]);
    }
}

__global__ void __launch_bounds__(128) sparse_to_bits_kernel(const BcBitsRec* __restrict__ bits, int n,
                                                            const uint32_t* __restrict__ dflt, int words,
                                                            const uint32_t* __restrict__ row_off,
                                                            const uint32_t* __restrict__ entries,
                                                            uint32_t* __restrict__ dst, size_t nq) {
    extern __shared__ uint32_t smem[];
    uint32_t* tile = smem;
    BcBitsRec* s_bits = reinterpret_cast<BcBitsRec*>(smem + (size_t)words * blockDim.x);
    for (int v = threadIdx.x; v < n; v += blockDim.x) s_bits[v] = bits[v];
    __syncthreads();
    const int T = blockDim.x;
    uint32_t* row = tile + threadIdx.x;
    for (size_t q = (size_t)blockIdx.x * T + threadIdx.x; q < nq; q += (size_t)gridDim.x * T) {
        for (int w = 0; w < words; ++w) row[w * T] = dflt[w];
        const uint32_t e0 = row_off[q], e1 = row_off[q + 1];
        for (uint32_t e = e0; e < e1; ++e) {
            const uint32_t x = __ldg(entries + e);
            const int col = x & 0x7fff;
            if (col >= n) continue;  // ignored, like a column outside the root component
            const BcBitsRec b = s_bits[col];
            row_select(row, T, b.bit_off, b.card, (x >> 16) & 0xff, x >> 24, (x >> 15) & 1);
        }
        uint4* out = reinterpret_cast<uint4*>(dst + q * words);
        for (int w = 0; w < words; w += 4) out[w >> 2] = make_uint4(row[w * T], row[(w + 1) * T], row[(w + 2) * T], row[(w + 3) * T]);
    }
}

// PACKED -> BITS.  PACKED is the densest wire form of what SPARSE carries (unit-weight range / IN-list queries), made for the
// one link that bounds the end-to-end rate, PCIe:
//     klen[q]     uint8    number of entries of query q (<= 255)
//     blk_off[b]  uint32   index of the first entry of block b of kPackedBlock = 128 queries
//     payload     bit stream, entry e at bits [e * w, (e + 1) * w), little endian in 32-bit words;
//                 w = cb + 2 * sb, cb = bits(n_nodes - 1), sb = bits(max_card - 1);  entry = col | lo << cb | hi << (cb + sb)
// Entries of a query are sorted by column; an entry whose column equals its predecessor's ORs into the column's mask (IN
// lists), the first entry of a column replaces it.  Census: 17 bits per entry, 17 B per query against 35 B as SPARSE.
// One CTA per block of 128 queries: a warp-shuffle scan of klen gives every thread its first entry.
__global__ void __launch_bounds__(128) packed_to_bits_kernel(const BcBitsRec* __restrict__ bits, int n, const uint32_t* __restrict__ dflt,
                                                            int words, const uint8_t* __restrict__ klen, const uint32_t* __restrict__ blk_off,
                                                            const uint32_t* __restrict__ payload, unsigned long long word_base, int cb, int sb,
                                                            uint32_t* __restrict__ dst, size_t nq) {
    extern __shared__ uint32_t smem[];
    uint32_t* tile = smem;
    BcBitsRec* s_bits = reinterpret_cast<BcBitsRec*>(smem + (size_t)words * blockDim.x);
    __shared__ uint32_t s_warp[4];
    for (int v = threadIdx.x; v < n; v += blockDim.x) s_bits[v] = bits[v];
    __syncthreads();
    const int T = blockDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* row = tile + threadIdx.x;
    const int w = cb + 2 * sb;
    const uint32_t cmask = (1u << cb) - 1u, smask = (1u << sb) - 1u;
    for (size_t blk = blockIdx.x; blk * T < nq; blk += gridDim.x) {
        const size_t q = blk * T + threadIdx.x;
        const uint32_t k = q < nq ? klen[q] : 0u;
        uint32_t incl = k;   // inclusive scan over the CTA
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (int i = 0; i < warp; ++i) before += s_warp[i];
        __syncthreads();
        if (q >= nq) continue;
        const unsigned long long e0 = (unsigned long long)blk_off[blk] + before + incl - k;
        for (int i = 0; i < words; ++i) row[i * T] = dflt[i];
        int prev = -1;
        for (uint32_t j = 0; j < k; ++j) {
            const unsigned long long bit = (e0 + j) * (unsigned long long)w;
            const unsigned long long wi = (bit >> 5) - word_base;
            const uint32_t sh = (uint32_t)(bit & 31u);
            const uint32_t a = __ldg(payload + wi), b = __ldg(payload + wi + 1);   // (the stream is padded by 8 bytes)
            const uint32_t x = __funnelshift_r(a, b, sh);
            const int col = (int)(x & cmask);
            if (col < n) {   // ignored otherwise, like a column outside the root component
                const BcBitsRec r = s_bits[col];
                row_select(row, T, r.bit_off, r.card, (int)((x >> cb) & smask), (int)((x >> (cb + sb)) & smask), col == prev);
            }
            prev = col;
        }
        uint4* out = reinterpret_cast<uint4*>(dst + q * words);
        for (int i = 0; i < words; i += 4) out[i >> 2] = make_uint4(row[i * T], row[(i + 1) * T], row[(i + 2) * T], row[(i + 3) * T]);
    }
}

int grid_for(const bc_model* m, size_t nq, int threads) {
    long long grid = (long long)((nq + threads - 1) / threads);
    const long long cap = (long long)m->sm_count * 16;
    return (int)(grid > cap ? cap : grid);
}

}  // namespace

int bc_convert_launch(bc_model* m, const void* src, int src_fmt, void* dst, int dst_fmt, size_t nq,
                      cudaStream_t stream) {
    if (dst_fmt != BC_DESC_BITS || (src_fmt != BC_DESC_RANGE_U8 && src_fmt != BC_DESC_RANGE_U16)) {
        bc_set_error("bc_convert_desc supports RANGE_U8 / RANGE_U16 -> BITS only");
        return BC_EINVAL;
    }
    if (nq == 0) return BC_OK;
    const int threads = 128;
    const size_t smem = ((size_t)m->bits_words * threads) * 4 + (size_t)m->n * sizeof(BcBitsRec);
    if (smem > (size_t)m->smem_optin) {
        bc_set_error("model too large for the descriptor conversion kernel (%zu B of shared memory)", smem);
        return BC_ELIMIT;
    }
    const size_t stride = (size_t)bc_model_desc_stride(m, src_fmt);
    if (src_fmt == BC_DESC_RANGE_U8) {
        BC_CUDA_CHECK(cudaFuncSetAttribute(ranges_to_bits_kernel<BC_DESC_RANGE_U8>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ranges_to_bits_kernel<BC_DESC_RANGE_U8><<<grid_for(m, nq, threads), threads, smem, stream>>>(
            m->d_bits, m->n, m->d_bits_default, m->bits_words, static_cast<const uint8_t*>(src), stride,
            static_cast<uint32_t*>(dst), nq);
    } else {
        BC_CUDA_CHECK(cudaFuncSetAttribute(ranges_to_bits_kernel<BC_DESC_RANGE_U16>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ranges_to_bits_kernel<BC_DESC_RANGE_U16><<<grid_for(m, nq, threads), threads, smem, stream>>>(
            m->d_bits, m->n, m->d_bits_default, m->bits_words, static_cast<const uint8_t*>(src), stride,
            static_cast<uint32_t*>(dst), nq);
    }
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

// WSPARSE -> DENSE_F32: one warp per query.  A row is a list of runs, each a header word
//   column (bits 0-14) | continuation (bit 15) | first state (bits 16-23) | count (bits 24-31)
// followed by `count` fp32 weights for the states first .. first + count - 1.  The first run of a column clears the
// column (unlisted states get weight 0, the meaning of a predicate); a continuation run adds more states to it.
// Columns without a run keep weight 1 everywhere -- the shape of the reference's sparse (query, n_distinct) dicts
// (Models/Bayescard_BN.py:279-325) with fractional weights, ~10-40x smaller than the DENSE row it expands to.
__global__ void __launch_bounds__(256) wsparse_to_dense_kernel(const BcNodeRec* __restrict__ nodes, int n, const float* __restrict__ dflt,
                                                                int lam_total, const uint32_t* __restrict__ row_off,
                                                                const uint32_t* __restrict__ words, float* __restrict__ dst, size_t nq) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t q = warp; q < nq; q += nwarps) {
        float* row = dst + q * (size_t)lam_total;
        for (int e = lane * 4; e < lam_total; e += 128)   // lam_total is a multiple of 4
            *reinterpret_cast<float4*>(row + e) = *reinterpret_cast<const float4*>(dflt + e);
        __syncwarp();
        uint32_t i = row_off[q];
        const uint32_t end = row_off[q + 1];
        while (i < end) {
            const uint32_t h = words[i];
            const int col = (int)(h & 0x7fffu), first = (int)((h >> 16) & 0xffu), cnt = (int)(h >> 24);
            if (col < n) {
                const BcNodeRec nd = nodes[col];
                if (!((h >> 15) & 1u)) {
                    for (int c = lane; c < nd.card; c += 32) row[nd.lam_off + c] = 0.f;
                    __syncwarp();
                }
                for (int j = lane; j < cnt; j += 32)
                    if (first + j < nd.card && i + 1 + j < end) row[nd.lam_off + first + j] = __uint_as_float(words[i + 1 + j]);
                __syncwarp();
            }
            i += 1 + (uint32_t)cnt;
        }
    }
}

int bc_expand_wsparse_launch(bc_model* m, const uint32_t* row_off, const uint32_t* words, size_t nq, float* dst_dense,
                             cudaStream_t stream) {
    if (nq == 0) return BC_OK;
    if (m->max_card > 256 || m->n > 32767) {
        bc_set_error("WSPARSE runs hold 8-bit state indices and 15-bit column ids (max domain %d, %d columns)", m->max_card, m->n);
        return BC_ELIMIT;
    }
    const int threads = 256;
    long long grid = (long long)((nq * 32 + threads - 1) / threads);
    const long long cap = (long long)m->sm_count * 16;
    if (grid > cap) grid = cap;
    wsparse_to_dense_kernel<<<(int)grid, threads, 0, stream>>>(m->d_nodes, m->n, m->d_dense_default, m->lam_total, row_off, words,
                                                               dst_dense, nq);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

static int bits_for(int x) {   // bits needed for values 0 .. x
    int b = 1;
    while ((1 << b) <= x) ++b;
    return b;
}
void bc_packed_geometry(const bc_model* m, int* col_bits, int* state_bits) {
    *col_bits = bits_for(m->n - 1);
    *state_bits = bits_for(m->max_card - 1);
}

int bc_expand_packed_launch(bc_model* m, const uint8_t* klen, const uint32_t* blk_off, const uint32_t* payload, unsigned long long word_base,
                            size_t nq, void* dst_bits, cudaStream_t stream) {
    if (nq == 0) return BC_OK;
    int cb, sb;
    bc_packed_geometry(m, &cb, &sb);
    if (cb + 2 * sb > 31) {
        bc_set_error("PACKED entries hold at most 31 bits (%d columns, domains up to %d states need %d)", m->n, m->max_card, cb + 2 * sb);
        return BC_ELIMIT;
    }
    const int threads = BC_PACKED_BLOCK;
    const size_t smem = ((size_t)m->bits_words * threads) * 4 + (size_t)m->n * sizeof(BcBitsRec);
    if (smem > (size_t)m->smem_optin) {
        bc_set_error("model too large for the packed expansion kernel (%zu B of shared memory)", smem);
        return BC_ELIMIT;
    }
    BC_CUDA_CHECK(cudaFuncSetAttribute(packed_to_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    packed_to_bits_kernel<<<grid_for(m, nq, threads), threads, smem, stream>>>(m->d_bits, m->n, m->d_bits_default, m->bits_words, klen, blk_off,
                                                                              payload, word_base, cb, sb, static_cast<uint32_t*>(dst_bits), nq);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}

int bc_expand_sparse_launch(bc_model* m, const uint32_t* row_off, const uint32_t* entries, size_t nq, void* dst_bits,
                            cudaStream_t stream) {
    if (nq == 0) return BC_OK;
    if (m->max_card > 256 || m->n > 32767) {
        bc_set_error("SPARSE entries hold 8-bit state bounds and 15-bit column ids (max domain %d, %d columns)",
                     m->max_card, m->n);
        return BC_ELIMIT;
    }
    const int threads = 128;
    const size_t smem = ((size_t)m->bits_words * threads) * 4 + (size_t)m->n * sizeof(BcBitsRec);
    if (smem > (size_t)m->smem_optin) {
        bc_set_error("model too large for the sparse expansion kernel (%zu B of shared memory)", smem);
        return BC_ELIMIT;
    }
    BC_CUDA_CHECK(cudaFuncSetAttribute(sparse_to_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sparse_to_bits_kernel<<<grid_for(m, nq, threads), threads, smem, stream>>>(
        m->d_bits, m->n, m->d_bits_default, m->bits_words, row_off, entries, static_cast<uint32_t*>(dst_bits), nq);
    BC_CUDA_CHECK(cudaGetLastError());
    bc_count_launch();
    return BC_OK;
}
