// mma.sync (legacy warp-level tensor path) issue rate on sm_100a: m16n8k8 tf32 and m16n8k16 bf16.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void __launch_bounds__(256) k_mma(float* out, int iters) {
    float d[8][4];
    unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x ^ 5u, 11u};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 123.456f) out[0] = s;
}
int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* d_out; cudaMalloc(&d_out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int kind = 0; kind < 2; ++kind)
        for (int bps : {1, 2, 4}) {
            const int iters = 20000;
            float best = 1e30f;
            for (int r = 0; r < 3; ++r) {
                cudaEventRecord(e0);
                if (kind == 0) k_mma<0><<<sms * bps, 256>>>(d_out, iters); else k_mma<1><<<sms * bps, 256>>>(d_out, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (r && ms < best) best = ms;
            }
            double instr = 8.0 * iters * sms * bps * 8;  // warp instructions
            double flop = instr * 2.0 * 16 * 8 * (kind == 0 ? 8 : 16);
            double clk = best * 1e-3 * khz * 1e3;
            printf("%s warps/SM %d: %.3f ms, %.1f TFLOP/s, %.4f mma/clk/SMSP\n", kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16", bps * 8, best,
                   flop / (best * 1e-3) / 1e12, instr / clk / (sms * 4));
        }
    printf("# status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
