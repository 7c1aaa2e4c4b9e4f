// Micro-benchmark for the K3 (fused tree, warp-tile) inner loop: one warp multiplies a WQ x K block of
// child messages U (shared memory, [K][WQ]) with a K x N CPT (shared memory, [K][ldt]) into a register tile.
// Answers before the real kernel is written: 3-register FFMA vs FFMA2 (fma.rn.f32x2) issue rate, and which
// lane mapping keeps the FMA pipe fed from LDS.128 fragments.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/k3_proto k3_proto.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// lanes: LQ along the queries (TQ queries each, WQ = LQ*TQ), LP = 32/LQ along the parent states
// (NG float4 groups each, interleaved: group g = pi + LP*j).
template <int LQ, int TQ, int NG, bool F2>
__global__ void __launch_bounds__(512) k3_loop(float* out, int K, int N, int ldt, int reps) {
    extern __shared__ __align__(16) float smem[];
    constexpr int WQ = LQ * TQ, LP = 32 / LQ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* T = smem;
    float* U = smem + K * ldt + warp * K * WQ;
    for (int i = threadIdx.x; i < K * ldt; i += blockDim.x) T[i] = 1.f / (1 + (i % 97));
    for (int i = lane; i < K * WQ; i += 32) U[i] = 1.f / (3 + (i % 89));
    __syncthreads();
    const int qi = lane % LQ, pi = lane / LQ;
    const int gmax = (N + 3) / 4 - 1;
    int goff[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) {
        int g = pi + LP * j;
        goff[j] = 4 * (g > gmax ? gmax : g);
    }
    float acc[TQ][NG * 4];
    unsigned long long acc2[TQ][NG * 2];
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int p = 0; p < NG * 4; ++p) acc[q][p] = 0.f;
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int p = 0; p < NG * 2; ++p) acc2[q][p] = 0ull;

    for (int r = 0; r < reps; ++r) {
        const float* up = U + qi * TQ;
        const float* tp = T;
        if (F2) {
            // explicit register double buffering: the fragments of row c+1 are loaded before the FMAs of row c
            ulonglong2 tA[NG], tB[NG];
            float uA[4], uB[4];
            auto load = [&](ulonglong2* t, float* u, const float* up_, const float* tp_) {
                if (TQ == 4) {
                    float4 v = *reinterpret_cast<const float4*>(up_);
                    u[0] = v.x; u[1] = v.y; u[2] = v.z; u[3] = v.w;
                } else {
                    float2 v = *reinterpret_cast<const float2*>(up_);
                    u[0] = v.x; u[1] = v.y;
                }
#pragma unroll
                for (int j = 0; j < NG; ++j) t[j] = *reinterpret_cast<const ulonglong2*>(tp_ + goff[j]);
            };
            auto fmas = [&](const ulonglong2* t, const float* u) {
#pragma unroll
                for (int j = 0; j < NG; ++j)
#pragma unroll
                    for (int q = 0; q < TQ; ++q) {
                        const unsigned long long uu = pack2(u[q], u[q]);
                        ffma2(acc2[q][2 * j], uu, t[j].x);
                        ffma2(acc2[q][2 * j + 1], uu, t[j].y);
                    }
            };
            load(tA, uA, up, tp);
            int c = 0;
            for (; c + 2 < K; c += 2) {
                load(tB, uB, up + WQ, tp + ldt);
                fmas(tA, uA);
                up += 2 * WQ;
                tp += 2 * ldt;
                load(tA, uA, up, tp);
                fmas(tB, uB);
            }
            if (c + 1 < K) {
                load(tB, uB, up + WQ, tp + ldt);
                fmas(tA, uA);
                fmas(tB, uB);
            } else {
                fmas(tA, uA);
            }
        } else {
#pragma unroll 2
            for (int c = 0; c < K; ++c, up += WQ, tp += ldt) {
                float u[TQ];
                if (TQ == 4) {
                    float4 v = *reinterpret_cast<const float4*>(up);
                    u[0] = v.x; u[1] = v.y; u[TQ > 2 ? 2 : 0] = v.z; u[TQ > 3 ? 3 : 0] = v.w;
                } else {
                    float2 v = *reinterpret_cast<const float2*>(up);
                    u[0] = v.x; u[1] = v.y;
                }
#pragma unroll
                for (int j = 0; j < NG; ++j) {
                    const float4 t = *reinterpret_cast<const float4*>(tp + goff[j]);
#pragma unroll
                    for (int q = 0; q < TQ; ++q) {
                        acc[q][4 * j + 0] = fmaf(u[q], t.x, acc[q][4 * j + 0]);
                        acc[q][4 * j + 1] = fmaf(u[q], t.y, acc[q][4 * j + 1]);
                        acc[q][4 * j + 2] = fmaf(u[q], t.z, acc[q][4 * j + 2]);
                        acc[q][4 * j + 3] = fmaf(u[q], t.w, acc[q][4 * j + 3]);
                    }
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int p = 0; p < NG * 4; ++p) s += acc[q][p];
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int p = 0; p < NG * 2; ++p) {
            float2 v = unpack2(acc2[q][p]);
            s += v.x + v.y;
        }
    if (s == 123.456f) out[0] = s;
}

template <int LQ, int TQ, int NG, bool F2>
void run(const char* name, int K, int N, float* d_out, int sms, int khz) {
    constexpr int WQ = LQ * TQ;
    const int ldt = (N + 3) / 4 * 4;
    for (int warps : {4, 8, 12, 16}) {
        size_t smem = (size_t)(K * ldt + warps * K * WQ) * 4;
        if (smem > 227 * 1024) continue;
        cudaFuncSetAttribute(k3_loop<LQ, TQ, NG, F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const int reps = 2000;
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            k3_loop<LQ, TQ, NG, F2><<<sms, warps * 32, smem>>>(d_out, K, N, ldt, reps);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        double useful = 2.0 * WQ * K * N * (double)reps * warps * sms;
        double executed = 2.0 * WQ * K * (NG * 4 * (32 / LQ)) * (double)reps * warps * sms;
        printf("%s, K=%d N=%d, warps/SM %d, smem %zu KB, %.3f ms, useful %.2f TFLOP/s, executed %.2f TFLOP/s %s\n", name, K, N,
               warps, smem / 1024, best, useful / (best * 1e-3) / 1e12, executed / (best * 1e-3) / 1e12,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* d_out;
    cudaMalloc(&d_out, 4);
    printf("# sms=%d clock=%d kHz; peak FFMA = %.1f TFLOP/s\n", sms, khz, sms * 128.0 * 2 * khz * 1e3 / 1e12);
    // 16-query warp tile, lanes 4q x 8p, thread tile 4 x 12
    run<4, 4, 3, false>("q16 4x8 t4x12 ffma ", 84, 84, d_out, sms, khz);
    run<4, 4, 3, true>("q16 4x8 t4x12 ffma2", 84, 84, d_out, sms, khz);
    // 16-query warp tile, lanes 8q x 4p, thread tile 2 x 24
    run<8, 2, 6, false>("q16 8x4 t2x24 ffma ", 84, 84, d_out, sms, khz);
    run<8, 2, 6, true>("q16 8x4 t2x24 ffma2", 84, 84, d_out, sms, khz);
    // 32-query warp tile, lanes 8q x 4p, thread tile 4 x 24
    run<8, 4, 6, false>("q32 8x4 t4x24 ffma ", 84, 84, d_out, sms, khz);
    run<8, 4, 6, true>("q32 8x4 t4x24 ffma2", 84, 84, d_out, sms, khz);
    // 32-query warp tile, lanes 4q... 8 x 4 with 8 queries per thread: lanes 4q x 8p, thread tile 8 x 12? (TQ=4 max here)
    // narrower parents
    run<4, 4, 2, true>("q16 4x8 t4x8  ffma2", 76, 63, d_out, sms, khz);
    run<4, 4, 1, true>("q16 4x8 t4x4  ffma2", 52, 32, d_out, sms, khz);
    run<8, 2, 4, true>("q16 8x4 t2x16 ffma2", 76, 63, d_out, sms, khz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("# status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
