// Does moving part of the softmax exponentials from the MUFU to the FMA pipe pay?  Per-element mix of the attention
// kernel's softmax loop (FFMA, exponential, row-sum FADD, F2FP per pair, FMNMX3 per pair), 128 elements per thread
// per iteration; PK of every 8 exponentials are a Cody-Waite polynomial exp2 (clip, 3 FADD, DEG FFMA, 1 LEA)
// instead of MUFU.EX2; SUM = 0 drops the row-sum FADD (as if the sum came from an extra MMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I simplesfm_b200/csrc -I include \
//        tools/micro/exp_poly_bench.cu -o tools/micro/exp_poly_bench
#include <cstdio>
#include "tc_common.cuh"
using namespace sfm::tc;
namespace sfm { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*, const char*, int) { return -2; } void count_launches(long long) {} }

template <int DEG>
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.f);
  const float xf = x + 12582912.f;            // round to nearest integer in the low mantissa bits
  const float f = x - (xf - 12582912.f);      // [-0.5, 0.5]
  float p;
  if (DEG == 3) {
    p = fmaf(0.05550357f, f, 0.24022651f);
    p = fmaf(p, f, 0.69314720f);
  } else {
    p = fmaf(0.24022651f, f, 0.69314720f);
  }
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// packed variant: x, the row sum and the polynomial in f32x2 instructions (two elements per issue slot)
template <int PK, int SUM, int DEG>
__global__ void __launch_bounds__(256, 1) bench2(int iters, long long* cycles, float* sink, const float* src) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float sv[128];
#pragma unroll
  for (int i = 0; i < 128; i++) sv[i] = src[(threadIdx.x * 128 + i) & 1023];
  float acc = 0.f;
  uint32_t xacc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const float c = 0.18033688f, mc = (float)(it & 3);
    const u64 c2 = pk2(c, c), mc2 = pk2(-mc, -mc), magic2 = pk2(12582912.f, 12582912.f), nmagic2 = pk2(-12582912.f, -12582912.f);
    const u64 m1 = pk2(-1.f, -1.f), k3 = pk2(0.05550357f, 0.05550357f), k2 = pk2(0.24022651f, 0.24022651f), k1 = pk2(0.69314720f, 0.69314720f), one2 = pk2(1.f, 1.f);
    u64 ls[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
    float mxa[8];
#pragma unroll
    for (int q = 0; q < 8; q++) mxa[q] = fmaxf(sv[q], sv[8 + q]);
#pragma unroll
    for (int i = 16; i < 128; i += 16)
#pragma unroll
      for (int q = 0; q < 8; q++) mxa[q] = fmax3(mxa[q], sv[i + q], sv[i + 8 + q]);
    acc += fmaxf(fmax3(mxa[0], mxa[1], mxa[2]), fmax3(fmax3(mxa[3], mxa[4], mxa[5]), mxa[6], mxa[7]));
#pragma unroll
    for (int i = 0; i < 128; i += 8) {
#pragma unroll
      for (int q = 0; q < 8; q += 2) {
        const u64 x2 = fma2(pk2(sv[i + q], sv[i + q + 1]), c2, mc2);
        const bool poly = PK > 0 && ((q / 2) * PK / 4) != ((q / 2 + 1) * PK / 4);   // PK of 4 pairs
        u64 p2;
        float p0, p1;
        if (poly) {
          float x0, x1;
          up2(x2, x0, x1);
          const u64 xc = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
          const u64 xf = add2(xc, magic2);
          const u64 f = fma2(add2(xf, nmagic2), m1, xc);
          u64 pp = DEG == 3 ? fma2(fma2(k3, f, k2), f, k1) : fma2(k2, f, k1);
          pp = fma2(pp, f, one2);
          float q0, q1, e0, e1;
          up2(pp, q0, q1);
          up2(xf, e0, e1);
          p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(e0) << 23));
          p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(e1) << 23));
        } else {
          float x0, x1;
          up2(x2, x0, x1);
          p0 = fast_exp2(x0);
          p1 = fast_exp2(x1);
        }
        p2 = pk2(p0, p1);
        if (SUM) ls[(q >> 1) & 1] = add2(ls[(q >> 1) & 1], p2);
        xacc ^= pack_bf16(p0, p1);
      }
    }
    float a, b, c_, d;
    up2(ls[0], a, b);
    up2(ls[1], c_, d);
    acc += (a + b) + (c_ + d);
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(xacc & 0x3fffffffu);
}

template <int PK, int SUM, int DEG>
__global__ void __launch_bounds__(256, 1) bench(int iters, long long* cycles, float* sink, const float* src) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float sv[128];
#pragma unroll
  for (int i = 0; i < 128; i++) sv[i] = src[(threadIdx.x * 128 + i) & 1023];
  float acc = 0.f;
  uint32_t xacc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const float c = 0.18033688f, mc = (float)(it & 3);
    float ls[4] = {0.f, 0.f, 0.f, 0.f};
    float mxa[8];
#pragma unroll
    for (int q = 0; q < 8; q++) mxa[q] = fmaxf(sv[q], sv[8 + q]);
#pragma unroll
    for (int i = 16; i < 128; i += 16)
#pragma unroll
      for (int q = 0; q < 8; q++) mxa[q] = fmax3(mxa[q], sv[i + q], sv[i + 8 + q]);
    acc += fmaxf(fmax3(mxa[0], mxa[1], mxa[2]), fmax3(fmax3(mxa[3], mxa[4], mxa[5]), mxa[6], mxa[7]));
#pragma unroll
    for (int i = 0; i < 128; i += 8) {
      float p[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const float x = fmaf(sv[i + q], c, -mc);
        // spread the polynomial ones evenly over the group of 8
        const bool poly = PK > 0 && (q * PK / 8) != ((q + 1) * PK / 8);
        p[q] = poly ? poly_exp2<DEG>(x) : fast_exp2(x);
        if (SUM) ls[q & 3] += p[q];
      }
#pragma unroll
      for (int q = 0; q < 8; q += 2) xacc ^= pack_bf16(p[q], p[q + 1]);
    }
    acc += (ls[0] + ls[1]) + (ls[2] + ls[3]);
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(xacc & 0x3fffffffu);
}

template <int PK, int SUM, int DEG, int PACKED = 0>
void run(long long* cyc, float* sink, const float* src) {
  const int iters = 2000;
  for (int W : {4, 8}) {
    if (PACKED) bench2<PK, SUM, DEG><<<148, 32 * W>>>(iters, cyc, sink, src);
    else bench<PK, SUM, DEG><<<148, 32 * W>>>(iters, cyc, sink, src);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return; }
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < W; w++) mx = h[w] > mx ? h[w] : mx;
    printf("%s poly %d/%d deg %d sum %d  warps/SMSP %d: %7.1f clk per 128 elements per warp  (%5.2f clk / element / SMSP)\n",
           PACKED ? "f32x2 " : "scalar", PK, PACKED ? 4 : 8, DEG, SUM, W / 4, (double)mx / iters, (double)mx / iters / 128.0 / (4.0 / W));
  }
}

int main() {
  long long* cyc; float *sink, *src;
  cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
  cudaMalloc(&sink, 148 * 512 * sizeof(float));
  cudaMalloc(&src, 1024 * sizeof(float));
  cudaMemset(src, 0, 1024 * sizeof(float));
  run<0, 1, 3>(cyc, sink, src);
  run<0, 0, 3>(cyc, sink, src);
  run<1, 1, 3>(cyc, sink, src);
  run<2, 1, 3>(cyc, sink, src);
  run<3, 1, 3>(cyc, sink, src);
  run<2, 0, 3>(cyc, sink, src);
  run<3, 0, 3>(cyc, sink, src);
  run<4, 0, 3>(cyc, sink, src);
  run<2, 1, 2>(cyc, sink, src);
  run<3, 1, 2>(cyc, sink, src);
  run<3, 0, 2>(cyc, sink, src);
  run<0, 1, 3, 1>(cyc, sink, src);
  run<1, 1, 3, 1>(cyc, sink, src);
  run<2, 1, 3, 1>(cyc, sink, src);
  run<1, 1, 2, 1>(cyc, sink, src);
  run<2, 1, 2, 1>(cyc, sink, src);
  run<1, 0, 3, 1>(cyc, sink, src);
  run<2, 0, 3, 1>(cyc, sink, src);
  return 0;
}
