#!/usr/bin/env python
"""Generates pipes.cu: FMA-pipe / ALU-pipe issue micro-benchmarks for sm_100a (B200).

Questions it answers (numbers land in DESIGN.md section "What bounds K-spec"):
  * FFMA with a 32-bit immediate vs FFMA2 (fma.rn.f32x2, Blackwell) with a broadcast immediate:
    flop/clk/SM of each, i.e. which one is the FP32 roofline denominator;
  * how many FMA-pipe slots an ALU-pipe instruction (FSEL / LOP3) costs when interleaved;
  * cost of straight-line code that does not fit the instruction caches.
"""
import sys

N_ACC = 8
BODY = 256  # instructions per unrolled body


def imm(i):
    return "%.9gf" % (0.5 + (i % 97) / 256.0)


def kernel(name, kind, alu_every=0, body=BODY, loop=True):
    """kind: 'ffma' | 'ffma2'.  alu_every=k inserts one FSEL after every k FMA instructions."""
    s = []
    s.append(f'extern "C" __global__ void __launch_bounds__(256) {name}(float* out, const float* in, int iters, unsigned sel)\n{{')
    if kind == "ffma":
        s.append("  float u = in[threadIdx.x & 31];")
        for a in range(N_ACC):
            s.append(f"  float a{a} = u * {a + 1}.0f;")
    else:
        s.append("  float2 u = make_float2(in[threadIdx.x & 31], in[(threadIdx.x + 1) & 31]);")
        for a in range(N_ACC):
            s.append(f"  float2 a{a} = make_float2(u.x * {a + 1}.0f, u.y * {a + 2}.0f);")
    s.append("  float w0 = u%s, w1 = w0 + 1.f, w2 = w0 + 2.f, w3 = w0 + 3.f;" % ("" if kind == "ffma" else ".x"))
    if loop:
        s.append("  for (int it = 0; it < iters; ++it) {")
    n_alu = 0
    for i in range(body):
        a = i % N_ACC
        if kind == "ffma":
            s.append(f"    a{a} = fmaf(u, {imm(i)}, a{a});")
        else:
            s.append(f"    a{a} = ffma2(u, make_float2({imm(i)}, {imm(i)}), a{a});")
        if alu_every and (i + 1) % alu_every == 0:
            w = n_alu % 4
            # FSEL on a loop-invariant predicate bit: ALU pipe, independent of the FMA chains
            s.append(f"    w{w} = (sel & {1 << (n_alu % 31)}u) ? w{(w + 1) % 4} : w{w};")
            n_alu += 1
    if loop:
        s.append("  }")
    if kind == "ffma":
        s.append("  float r = " + " + ".join(f"a{a}" for a in range(N_ACC)) + " + w0 + w1 + w2 + w3;")
    else:
        s.append("  float r = " + " + ".join(f"a{a}.x + a{a}.y" for a in range(N_ACC)) + " + w0 + w1 + w2 + w3;")
    s.append("  if (r == 123.456f) out[0] = r;\n}\n")
    return "\n".join(s), body, n_alu


def main():
    out = ['#include <cstdio>\n#include <cuda_runtime.h>\n',
           '__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {\n'
           '  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),\n'
           '                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;\n'
           '  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));\n'
           '  return *reinterpret_cast<float2*>(&rd);\n}\n']
    table = []
    for kind in ("ffma", "ffma2"):
        for alu in (0, 8, 4, 2, 1):
            name = f"k_{kind}_alu{alu}"
            src, nf, na = kernel(name, kind, alu)
            out.append(src)
            table.append((name, kind, nf, na, 1))
    # straight-line code of growing size, executed once per outer iteration (no inner reuse)
    for kind in ("ffma", "ffma2"):
        for body in (1024, 2048, 4096, 8192):
            name = f"k_{kind}_line{body}"
            src, nf, na = kernel(name, kind, 0, body)
            out.append(src)
            table.append((name, kind, nf, na, 0))
    out.append("struct K { const char* name; void (*fn)(float*, const float*, int, unsigned); int lanes; int nf; int na; };\n")
    out.append("static K ks[] = {\n" + "".join(
        f'  {{"{n}", {n}, {1 if k == "ffma" else 2}, {nf}, {na}}},\n' for n, k, nf, na, _ in table) + "};\n")
    out.append(r'''
int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float *d_out, *d_in;
  cudaMalloc(&d_out, 4); cudaMalloc(&d_in, 128); cudaMemset(d_in, 0, 128);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("# sms=%d clock=%d kHz\n# name, warps/SM, TFLOP/s, fma_instr/clk/SMSP, all_instr/clk/SMSP\n", sms, khz);
  for (auto& k : ks) {
    for (int blocks_per_sm : {1, 2, 4, 8}) {
      const int threads = 256, grid = sms * blocks_per_sm;
      const long long target = 1LL << 22;   // FMA instructions per thread
      int iters = (int)(target / k.nf);
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k.fn<<<grid, threads>>>(d_out, d_in, iters, 0x55555555u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      double fma_instr = (double)k.nf * iters * grid * threads / 32.0;       // warp instructions
      double all_instr = (double)(k.nf + k.na) * iters * grid * threads / 32.0;
      double clk = best * 1e-3 * khz * 1e3;
      double tf = fma_instr * 32 * 2 * k.lanes / (best * 1e-3) / 1e12;
      printf("%s, %d, %.2f, %.3f, %.3f\n", k.name, blocks_per_sm * 8, tf, fma_instr / clk / (sms * 4), all_instr / clk / (sms * 4));
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("# status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
''')
    open(sys.argv[1] if len(sys.argv) > 1 else "pipes.cu", "w").write("\n".join(out))


if __name__ == "__main__":
    main()
