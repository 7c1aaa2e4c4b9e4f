// What stalls the tcgen05.mma issuer?  One warp issues groups of G MMAs (kind::tf32, M128 x N96 x K8, A in tensor memory, all
// into the same accumulator) and between two groups executes one of: nothing | tcgen05.fence::after_thread_sync | a
// tcgen05.commit to an mbarrier | an mbarrier.try_wait on an already completed phase | commit + try_wait + fence (what a
// producer/consumer ring needs per stage).  Reported: clocks per group against the floor G * 48.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/umma_issue umma_issue.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int G, int MODE>
__global__ void __launch_bounds__(64, 1) k_issue(int groups, long long* res) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[18];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < (256 * 64) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0u;
    if (tid == 0) {
        for (int i = 0; i < 18; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[17])) : "memory");   // phase 0 of bar[17] is complete
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t b_base = smem_u32(smem);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t b_lo32 = ((b_base & 0x3FFFFu) >> 4) | (1u << 16), hi32 = (512u >> 4) | (1u << 14) | (4u << 29);
    if (warp == 0) {
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
#pragma unroll
            for (int c = 0; c < G; ++c)
                asm volatile(
                    "{\n.reg .pred p, q;\n.reg .b64 db;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, 1, 0;\nmov.b64 db, {%2, %3};\n"
                    "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n}\n" ::"r"(tmem),
                    "r"(tmem + 480u), "r"(b_lo32), "r"(hi32), "r"(idesc)
                    : "memory");
            if (MODE == 2 || MODE == 4)
                asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(
                                 smem_u32(&bar[g & 15]))
                             : "memory");
            if (MODE == 3 || MODE == 4)
                asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
                                 smem_u32(&bar[17]))
                             : "memory");
            if (MODE == 1 || MODE == 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (MODE == 5) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(
                         smem_u32(&bar[16]))
                     : "memory");
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar[16]))
                     : "memory");
        if (lane == 0) res[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <int G, int MODE>
void run(const char* what, long long* d) {
    const int groups = 1024;
    const size_t smem = 256 * 64 + 2048;
    cudaFuncSetAttribute(k_issue<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_issue<G, MODE><<<1, 64, smem>>>(groups, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("G %d  %-38s: %7.1f clk per group (floor %d)%s\n", G, what, (double)h / groups, G * 48, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
#define ALL(G)                                              \
    run<G, 0>("nothing", d);                                \
    run<G, 1>("tcgen05.fence::after_thread_sync", d);       \
    run<G, 5>("tcgen05.fence::before_thread_sync", d);      \
    run<G, 2>("tcgen05.commit", d);                         \
    run<G, 3>("mbarrier.try_wait (complete phase)", d);     \
    run<G, 4>("commit + try_wait + fence::after", d);
    ALL(1) ALL(2) ALL(4) ALL(6)
    printf("# status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
