// micro-probe: issue cost of the packed fp32 instructions of sm_100 (FFMA2 / FADD2) against scalar FFMA, and of a mix with FMNMX
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  uint64_t p0, p1, p2, p3, pa, pb;
  asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
  asm("mov.b64 %0, {%1,%1};" : "=l"(pa) : "f"(a));
  asm("mov.b64 %0, {%1,%1};" : "=l"(pb) : "f"(b));
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {  // 8 scalar FFMA
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
      } else if (MODE == 1) {  // 4 FFMA2 = the same flops
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
      } else if (MODE == 2) {  // 8 FFMA + 8 FMNMX (two pipes)
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        x0 = fminf(x0, x1); x1 = fmaxf(x1, x2); x2 = fminf(x2, x3); x3 = fmaxf(x3, x4);
        x4 = fminf(x4, x5); x5 = fmaxf(x5, x6); x6 = fminf(x6, x7); x7 = fmaxf(x7, a);
      } else if (MODE == 3) {  // 4 FFMA2 + 8 FMNMX
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
        x0 = fminf(x0, x1); x1 = fmaxf(x1, x2); x2 = fminf(x2, x3); x3 = fmaxf(x3, x4);
        x4 = fminf(x4, x5); x5 = fmaxf(x5, x6); x6 = fminf(x6, x7); x7 = fmaxf(x7, a);
      } else if (MODE == 4) {  // 8 FMNMX only
        x0 = fminf(x0, x1); x1 = fmaxf(x1, x2); x2 = fminf(x2, x3); x3 = fmaxf(x3, x4);
        x4 = fminf(x4, x5); x5 = fmaxf(x5, x6); x6 = fminf(x6, x7); x7 = fmaxf(x7, a);
      } else if (MODE == 5) {  // 4 FADD2
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pa));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pa));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pa));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pa));
      }
    }
  }
  float lo, hi, s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0)); s += lo + hi;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p1)); s += lo + hi;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p2)); s += lo + hi;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p3)); s += lo + hi;
  if (s == 12345.678f) out[0] = s;
}
template <int MODE>
void run(const char* name, int per_iter) {
  float* out; cudaMalloc(&out, 4);
  const int iters = 20000, blocks = 148 * 2, threads = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double winst = (double)blocks * threads / 32 * iters * 8.0 * per_iter;
  printf("%-28s %8.3f ms  %7.1f G warp-inst/s  (%.2f inst/clk/SM at 1.965 GHz)\n", name, ms, winst / ms / 1e6, winst / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  run<0>("8 FFMA", 8); run<1>("4 FFMA2", 4); run<2>("8 FFMA + 8 FMNMX", 16); run<3>("4 FFMA2 + 8 FMNMX", 12); run<4>("8 FMNMX", 8); run<5>("4 FADD2", 4);
  return 0;
}
