// CPT fitting for a FIXED tree: the step right before the inference path.
//
// Replaces `self.model.fit(discrete_table)` of Bayescard_BN.build_from_data (Models/Bayescard_BN.py:108-110), i.e.
// BayesianModel.fit (Pgmpy/models/BayesianModel.py:278-323) -> MaximumLikelihoodEstimator.estimate_cpd
// (Pgmpy/estimators/MLE.py:61-104) -> state_counts (Pgmpy/estimators/base.py:63-135): for every node a 2-D histogram
// of (own state, parent state) over the discretised table, then column normalisation (an all-zero column becomes
// uniform, MLE.py:77-79).  The reference spends 140 s (DMV, 11.6 M rows) to 717 s (Census) in pandas group-bys
// (logs/dmv_20210112-101604.log:6, logs/census_20210112-103814.log:6).
//
// Here the table (n_rows x n_cols bin ids, row-major, columns in topological order, uint8 or uint16) is read ONCE:
// each thread walks rows, every CTA keeps a private copy of ALL the count tables in shared memory (DMV 10 882
// counters = 43 KB, IMDB <= 122 KB) and adds into it with shared-memory atomics; tables of <= 64 counters (boolean
// columns) get one copy per lane so that the 32 lanes of a warp never collide on them.  At the
// end every CTA adds its private tables into the global 64-bit counters.  Models whose tables do not fit shared
// memory count straight into the global counters.  The counts are integers, so the result is bit-exact whatever the
// order of the atomics; the normalisation (a few thousand divisions) is done by the caller in fp64, exactly as the
// reference does it.  Bound: HBM (n_rows x n_cols x element size bytes, read once).
#include "bc_internal.h"

namespace {

struct FitNode {
    int parent;     // topological index of the parent column, -1 for the root
    int card;
    int card_pa;    // 1 for the root
    int off;        // first counter of this node; counter of (c, p) = off + c * card_pa + p
    int rep;        // tiny tables: first word of the lane-replicated copy in shared memory, else -1
};

constexpr int kMaxParamNodes = 160;  // the node table travels as a kernel parameter (3.2 KB) up to this many columns
struct FitParams {
    int n_nodes;
    FitNode nodes[kMaxParamNodes];
};

// Lane-replicated counters for tiny tables: a boolean-by-boolean edge has four counters that all 32 lanes of a warp
// hammer; with one copy per lane (counter i of a small table lives at rep + i * 32 + lane) there is no conflict.
constexpr int kSmallTable = 64;

constexpr int kRowsPerIter = 4;  // rows in flight per thread: their byte loads overlap

template <typename T, bool SHARED>
__global__ void __launch_bounds__(256) fit_count_kernel(const T* __restrict__ table, size_t n_rows, size_t row_stride,
                                                         const __grid_constant__ FitParams prm, const FitNode* __restrict__ nodes_dev,
                                                         int n_nodes, int n_counters, int n_shared, unsigned long long* __restrict__ counts) {
    unsigned long long* bad_rows = counts + n_counters;  // one extra 64-bit counter behind the tables
    // dynamic shared memory: [node table (5 ints per node)] [private counters + lane replicas]
    extern __shared__ __align__(16) unsigned int smem_u[];
    int* s_nodes = reinterpret_cast<int*>(smem_u);
    unsigned int* s_cnt = smem_u + ((n_nodes * 5 + 3) & ~3);
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n_nodes * 5; i += blockDim.x)
        s_nodes[i] = nodes_dev ? reinterpret_cast<const int*>(nodes_dev)[i] : reinterpret_cast<const int*>(prm.nodes)[i];
    if (SHARED)
        for (int i = threadIdx.x; i < n_shared; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t r0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r0 < n_rows; r0 += stride * kRowsPerIter) {
        const T* row[kRowsPerIter];
        int badv[kRowsPerIter];  // first node whose bin id is outside its domain (n_nodes: none); -1: no such row
#pragma unroll
        for (int k = 0; k < kRowsPerIter; ++k) {
            const size_t r = r0 + (size_t)k * stride;
            row[k] = table + (r < n_rows ? r : r0) * row_stride;
            badv[k] = r < n_rows ? n_nodes : -1;
        }
        for (int v = 0; v < n_nodes; ++v) {
            const int parent = s_nodes[5 * v], card = s_nodes[5 * v + 1], card_pa = s_nodes[5 * v + 2];
            const int off = s_nodes[5 * v + 3], rep = s_nodes[5 * v + 4];
            int c[kRowsPerIter], p[kRowsPerIter];
#pragma unroll
            for (int k = 0; k < kRowsPerIter; ++k) {
                c[k] = (int)row[k][v];
                p[k] = parent >= 0 ? (int)row[k][parent] : 0;  // validated earlier: parents come first
            }
#pragma unroll
            for (int k = 0; k < kRowsPerIter; ++k) {
                if (v >= badv[k]) continue;
                if (c[k] >= card) {  // rare: the row is skipped and reported, its earlier increments are undone below
                    badv[k] = v;
                    continue;
                }
                const int li = c[k] * card_pa + p[k];
                if (SHARED) atomicAdd(&s_cnt[rep >= 0 ? rep + (li << 5) + lane : off + li], 1u);
                else atomicAdd(&counts[off + li], 1ull);
            }
        }
#pragma unroll
        for (int k = 0; k < kRowsPerIter; ++k) {
            if (badv[k] < 0 || badv[k] >= n_nodes) continue;
            atomicAdd(bad_rows, 1ull);
            for (int u = 0; u < badv[k]; ++u) {
                const int parent = s_nodes[5 * u], card_pa = s_nodes[5 * u + 2], off = s_nodes[5 * u + 3], rep = s_nodes[5 * u + 4];
                const int li = (int)row[k][u] * card_pa + (parent >= 0 ? (int)row[k][parent] : 0);
                if (SHARED) atomicSub(&s_cnt[rep >= 0 ? rep + (li << 5) + lane : off + li], 1u);
                else atomicAdd(&counts[off + li], ~0ull);
            }
        }
    }
    if (SHARED) {
        __syncthreads();
        // private tables -> global 64-bit counters; the lane replicas of a tiny table are summed first
        for (int v = 0; v < n_nodes; ++v) {
            const int card = s_nodes[5 * v + 1], card_pa = s_nodes[5 * v + 2], off = s_nodes[5 * v + 3], rep = s_nodes[5 * v + 4];
            if (rep < 0) continue;
            for (int li = threadIdx.x; li < card * card_pa; li += blockDim.x) {
                unsigned int x = 0;
                for (int l = 0; l < 32; ++l) x += s_cnt[rep + (li << 5) + l];
                s_cnt[off + li] = x;  // the plain slot of a replicated table is otherwise unused
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_counters; i += blockDim.x) {
            const unsigned int x = s_cnt[i];
            if (x) atomicAdd(&counts[i], (unsigned long long)x);
        }
    }
}

}  // namespace

extern "C" int bc_fit_counts(int device, int n_nodes, const int32_t* parent, const int32_t* card, const void* table_dev,
                             int elem_bytes, size_t n_rows, size_t row_stride_elems, unsigned long long* counts_dev,
                             size_t n_counters, uint32_t* bad_rows_host, void* stream) {
    if (n_nodes <= 0 || !parent || !card || (!table_dev && n_rows) || !counts_dev || (elem_bytes != 1 && elem_bytes != 2) ||
        row_stride_elems < (size_t)n_nodes) {
        bc_set_error("bc_fit_counts: bad arguments");
        return BC_EINVAL;
    }
    std::vector<FitNode> nodes(n_nodes);
    long long total = 0;
    for (int v = 0; v < n_nodes; ++v) {
        if (card[v] <= 0 || card[v] > (elem_bytes == 1 ? 256 : 65536) || parent[v] >= v || (v == 0) != (parent[v] < 0)) {
            bc_set_error("bc_fit_counts: node %d: card %d, parent %d (topological order, one root, card within the element type)",
                         v, card[v], parent[v]);
            return BC_EINVAL;
        }
        nodes[v].parent = parent[v];
        nodes[v].card = card[v];
        nodes[v].card_pa = parent[v] >= 0 ? card[parent[v]] : 1;
        nodes[v].off = (int)total;
        nodes[v].rep = -1;
        total += (long long)card[v] * nodes[v].card_pa;
        if (total > (1ll << 30)) {
            bc_set_error("bc_fit_counts: more than 2^30 counters");
            return BC_ELIMIT;
        }
    }
    if ((size_t)total != n_counters) {
        bc_set_error("bc_fit_counts: counts buffer holds %zu counters, the tree needs %lld", n_counters, total);
        return BC_EINVAL;
    }
    BC_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int sm_count = 0, smem_optin = 0;
    BC_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    BC_CUDA_CHECK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    // shared-memory plan: all private tables + lane replicas of the tiny ones
    size_t n_shared = (size_t)total;
    for (int v = 0; v < n_nodes; ++v)
        if (nodes[v].card * nodes[v].card_pa <= kSmallTable) {
            nodes[v].rep = (int)n_shared;
            n_shared += (size_t)nodes[v].card * nodes[v].card_pa * 32;
        }
    bool shared = (n_shared + (size_t)n_nodes * 5 + 4) * 4 + 1024 <= (size_t)smem_optin;
    if (!shared && (size_t)total * 4 + 1024 <= (size_t)smem_optin) {  // the tables fit, their replicas do not
        for (int v = 0; v < n_nodes; ++v) nodes[v].rep = -1;
        n_shared = (size_t)total;
        shared = true;
    }
    const size_t node_words = ((size_t)n_nodes * 5 + 3) & ~(size_t)3;
    const size_t smem = (node_words + (shared ? n_shared : 0)) * 4;
    // the node table travels as a kernel parameter; only trees wider than that need a device copy
    const bool by_param = n_nodes <= kMaxParamNodes;
    FitNode* d_nodes = nullptr;
    if (!by_param) {
        BC_CUDA_CHECK(cudaMallocAsync(&d_nodes, sizeof(FitNode) * n_nodes, st));
        // pageable source: the copy is staged by the runtime before the call returns
        BC_CUDA_CHECK(cudaMemcpyAsync(d_nodes, nodes.data(), sizeof(FitNode) * n_nodes, cudaMemcpyHostToDevice, st));
    }
    BC_CUDA_CHECK(cudaMemsetAsync(counts_dev, 0, sizeof(unsigned long long) * (n_counters + 1), st));
    if (n_rows) {
        static FitParams prm;  // filled under the lock below; passed by value at launch
        static std::mutex prm_mu;
        std::lock_guard<std::mutex> lock(prm_mu);
        prm.n_nodes = n_nodes;
        if (by_param) std::copy(nodes.begin(), nodes.end(), prm.nodes);
        // persistent grid: as many 256-thread CTAs per SM as the private tables allow (at most 4)
        int per_sm = shared ? (int)((size_t)smem_optin / (smem + 1024)) : 4;
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        long long grid = (long long)sm_count * per_sm;
        const long long needed = (long long)((n_rows + 256 * kRowsPerIter - 1) / (256 * kRowsPerIter));
        if (grid > needed) grid = needed;
#define BC_FIT_LAUNCH(T, SH)                                                                                          \
    do {                                                                                                              \
        static int attr_bytes[64] = {};  /* per device: raise the dynamic shared-memory limit only when it grows */       \
        if (device >= 64 || attr_bytes[device] < (int)smem) {                                                         \
            BC_CUDA_CHECK(cudaFuncSetAttribute(fit_count_kernel<T, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            if (device < 64) attr_bytes[device] = (int)smem;                                                          \
        }                                                                                                             \
        fit_count_kernel<T, SH><<<(unsigned)grid, 256, smem, st>>>(static_cast<const T*>(table_dev), n_rows, row_stride_elems, prm, \
                                                                   d_nodes, n_nodes, (int)total, (int)n_shared, counts_dev);         \
    } while (0)
        if (elem_bytes == 1) {
            if (shared) BC_FIT_LAUNCH(uint8_t, true); else BC_FIT_LAUNCH(uint8_t, false);
        } else {
            if (shared) BC_FIT_LAUNCH(uint16_t, true); else BC_FIT_LAUNCH(uint16_t, false);
        }
#undef BC_FIT_LAUNCH
        BC_CUDA_CHECK(cudaGetLastError());
        bc_count_launch();
    }
    if (d_nodes) cudaFreeAsync(d_nodes, st);
    if (bad_rows_host) {
        unsigned long long bad = 0;
        BC_CUDA_CHECK(cudaMemcpyAsync(&bad, counts_dev + n_counters, sizeof(bad), cudaMemcpyDeviceToHost, st));
        BC_CUDA_CHECK(cudaStreamSynchronize(st));
        *bad_rows_host = (uint32_t)(bad > 0xffffffffull ? 0xffffffffull : bad);
    }
    return BC_OK;
}
