// Micro-benchmark: issue throughput of FFMA vs FFMA2 (fma.rn.f32x2) vs HADD2.F32 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float b = 1.0001f, c = 0.5f;
  unsigned long long p0, p1, p2, p3, bb, cc;
  { float2 t = make_float2(a0, a1); p0 = *reinterpret_cast<unsigned long long*>(&t); }
  { float2 t = make_float2(a2, a3); p1 = *reinterpret_cast<unsigned long long*>(&t); }
  { float2 t = make_float2(a4, a5); p2 = *reinterpret_cast<unsigned long long*>(&t); }
  { float2 t = make_float2(a6, a7); p3 = *reinterpret_cast<unsigned long long*>(&t); }
  { float2 t = make_float2(b, b); bb = *reinterpret_cast<unsigned long long*>(&t); }
  { float2 t = make_float2(c, c); cc = *reinterpret_cast<unsigned long long*>(&t); }
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(bb), "l"(cc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(bb), "l"(cc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(bb), "l"(cc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(bb), "l"(cc));
      }
    }
  }
  float2 q0 = *reinterpret_cast<float2*>(&p0), q1 = *reinterpret_cast<float2*>(&p1), q2 = *reinterpret_cast<float2*>(&p2), q3 = *reinterpret_cast<float2*>(&p3);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + q0.x + q0.y + q1.x + q1.y + q2.x + q2.y + q3.x + q3.y;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 4, 512>>>(out, iters); else k<1><<<148 * 4, 512>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)148 * 4 * 512 * iters * 64.0;   // scalar FMAs (mode 1: 32 instr x 2)
      if (rep) printf("%s: %.3f ms, %.1f scalar FMA/clk/SM at 1.9 GHz (%s)\n", mode ? "FFMA2" : "FFMA ", ms,
                      fma / (ms * 1e-3) / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
