// tcgen05.mma kind::tf32 (M = 128, K = 8) issue/latency behaviour on sm_100a, the facts the fused tree kernel (K3) is
// scheduled around:
//   1. clocks per MMA when every instruction accumulates into the SAME tensor-memory columns (dependent chain) for
//      N = 16 .. 256, against C = 2, 3, 4 independent accumulators issued round-robin;
//   2. tcgen05.ld bandwidth of four warps (tensor memory -> registers), alone and while the MMA chain runs;
//   3. A operand from tensor memory instead of shared memory.
// One CTA per SM; operands are zeros (timing only).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/umma_chain umma_chain.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Res { long long clk_mma, clk_ld; };

// mode bit 0: A from tensor memory; ld_warps: warps 1..4 stream tcgen05.ld during the MMA loop (ld_iters x 32 columns each)
template <int N, int C, int a_tmem>
__global__ void __launch_bounds__(192, 1) k_chain(int n_mma, int ld_iters, int mma_on, Res* res) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < (8192 + 256 * 64) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 8192;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0 && mma_on) {
        // the whole (converged) warp runs the loop; elect.sync inside the asm picks the issuing lane, so every operand stays
        // in uniform registers (a divergent single lane costs ~200 clk of scalar code per instruction)
        const long long t0 = clock64();
        const uint32_t a_lo32 = ((a_base & 0x3FFFFu) >> 4) | (1u << 16), b_lo32 = ((b_base & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t hi32 = (512u >> 4) | (1u << 14) | (4u << 29);
        for (int i = 0; i < n_mma; i += C) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const uint32_t d = tmem + (uint32_t)(c * N);
                if (a_tmem) {
                    asm volatile(
                        "{\n.reg .pred p, q;\n.reg .b64 db;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, 1, 0;\nmov.b64 db, {%2, %3};\n"
                        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n}\n" ::"r"(d),
                        "r"(tmem + 480u), "r"(b_lo32), "r"(hi32), "r"(idesc)
                        : "memory");
                } else {
                    asm volatile(
                        "{\n.reg .pred p, q;\n.reg .b64 da, db;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %3};\n"
                        "mov.b64 db, {%2, %3};\n@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n}\n" ::"r"(d),
                        "r"(a_lo32), "r"(b_lo32), "r"(hi32), "r"(idesc)
                        : "memory");
                }
            }
        }
        asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(
                         smem_u32(&bar))
                     : "memory");
        asm volatile(
            "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar))
            : "memory");
        if (lane == 0) res[blockIdx.x].clk_mma = clock64() - t0;
    } else if (warp >= 1 && warp <= 4 && ld_iters > 0) {
        const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int i = 0; i < ld_iters; ++i) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(t + (uint32_t)((i & 7) * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
        }
        if (lane == 0 && warp == 1) res[blockIdx.x].clk_ld = clock64() - t0;
        if (acc == 0x12345u) res[blockIdx.x].clk_ld = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    Res* d;
    cudaMalloc(&d, sizeof(Res) * sms);
    Res* h = new Res[sms];
    const size_t smem = 8192 + 256 * 64 + 2048;
    auto launch = [&](int N, int C, int a_tmem, int n, int ld_iters, int mma_on, int grid) {
#define CASE(NN, CC, AA)                                                                                              \
    if (N == NN && C == CC && a_tmem == AA) {                                                                         \
        cudaFuncSetAttribute(k_chain<NN, CC, AA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        k_chain<NN, CC, AA><<<grid, 192, smem>>>(n, ld_iters, mma_on, d);                                             \
        return;                                                                                                       \
    }
#define CASES(NN) CASE(NN, 1, 0) CASE(NN, 2, 0) CASE(NN, 3, 0) CASE(NN, 4, 0) CASE(NN, 1, 1) CASE(NN, 2, 1) CASE(NN, 3, 1) CASE(NN, 4, 1)
        CASES(16) CASES(32) CASES(48) CASES(64) CASES(96) CASE(128, 1, 0) CASE(128, 2, 0) CASE(128, 3, 0) CASE(128, 1, 1) CASE(128, 2, 1)
        CASE(128, 3, 1) CASE(192, 1, 0) CASE(192, 2, 0) CASE(192, 1, 1) CASE(192, 2, 1) CASE(256, 1, 0) CASE(256, 1, 1)
        printf("# no instance N %d C %d\n", N, C);
    };
    auto run = [&](int N, int C, int n, int a_tmem, int ld_iters, int mma_on, int grid) {
        cudaMemset(d, 0, sizeof(Res) * sms);
        launch(N, C, a_tmem, n, ld_iters, mma_on, grid);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("# error: %s\n", cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(h, d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
        long long m = 0, l = 0;
        for (int i = 0; i < grid; ++i) { m += h[i].clk_mma; l += h[i].clk_ld; }
        h[0].clk_mma = m / grid;
        h[0].clk_ld = l / grid;
    };
    const int n = 4092;  // multiple of 1, 2, 3, 4
    printf("# tcgen05.mma kind::tf32 M128 K8, %d instructions per CTA; floor = N/2 clk per instruction\n", n);
    for (int grid : {sms})
        for (int a_tmem : {0, 1})
            for (int N : {16, 32, 48, 64, 96, 128, 192, 256})
                for (int C : {1, 2, 3, 4}) {
                    if (C * N > 480) continue;
                    run(N, C, n, a_tmem, 0, 1, grid);
                    printf("grid %3d A %s N %3d chains %d: %7.1f clk/mma (floor %5.1f, pipe %4.1f %%)\n", grid, a_tmem ? "tmem" : "smem", N, C,
                           (double)h[0].clk_mma / n, N / 2.0, 100.0 * (N / 2.0) * n / (double)h[0].clk_mma);
                }
    printf("# latency: issue n MMAs (N = 96, A in tensor memory) + tcgen05.commit, spin on the mbarrier: clk from first issue to wake-up\n");
    for (int nn : {1, 2, 4, 6, 8, 16, 32}) {
        run(96, 1, nn, 1, 0, 1, sms);
        printf("n %2d: %lld clk (floor %d)\n", nn, h[0].clk_mma, nn * 48);
    }
    printf("# tcgen05.ld 32x32b.x32 by four warps (each instruction: 4 KB per warp)\n");
    run(96, 1, n, 0, 2048, 0, sms);
    printf("ld alone              : %7.1f clk per x32 load per warp -> %.1f B/clk/SM\n", (double)h[0].clk_ld / 2048, 4.0 * 4096 * 2048 / (double)h[0].clk_ld);
    for (int N : {96, 192})
        for (int C : {1, 2, 3}) {
            if (C * N > 480) continue;
            run(N, C, n, 0, 2048, 1, sms);
            printf("ld + mma N %3d chains %d: %7.1f clk/mma, ld %.1f B/clk/SM (ld loop %lld clk, mma loop %lld clk)\n", N, C, (double)h[0].clk_mma / n,
                   4.0 * 4096 * 2048 / (double)h[0].clk_ld, h[0].clk_ld, h[0].clk_mma);
        }
    printf("# status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
