// What paces the softmax inner loop?  128 elements per thread per iteration, W warps per CTA, variants of the
// per-element instruction mix (all operands in registers):
//   0: FFMA + MUFU.EX2 + FADD                      3: mix 0 + PRMT pack (truncating bf16 pair)
//   1: mix 0 + F2FP.BF16.PACK_AB per two elements  4: F2FP only            5: mix 1 + 64 FMNMX3
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I simplesfm_b200/csrc -I include \
//        tools/micro/exp_mix_bench.cu -o tools/micro/exp_mix_bench
#include <cstdio>
#include "tc_common.cuh"
using namespace sfm::tc;
namespace sfm { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*, const char*, int) { return -2; } void count_launches(long long) {} }

template <int MODE>
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
    if (MODE == 5) {
      float mxa[8];
#pragma unroll
      for (int q = 0; q < 8; q++) mxa[q] = fmaxf(sv[q], sv[8 + q]);
#pragma unroll
      for (int i = 16; i < 128; i += 16)
#pragma unroll
        for (int q = 0; q < 8; q++) mxa[q] = fmax3(mxa[q], sv[i + q], sv[i + 8 + q]);
      acc += fmaxf(fmax3(mxa[0], mxa[1], mxa[2]), fmax3(fmax3(mxa[3], mxa[4], mxa[5]), mxa[6], mxa[7]));
    }
#pragma unroll
    for (int i = 0; i < 128; i += 4) {
      float p0, p1, p2, p3;
      if (MODE == 4) {
        p0 = sv[i] + mc; p1 = sv[i + 1]; p2 = sv[i + 2]; p3 = sv[i + 3];
      } else {
        p0 = fast_exp2(fmaf(sv[i], c, -mc)); p1 = fast_exp2(fmaf(sv[i + 1], c, -mc));
        p2 = fast_exp2(fmaf(sv[i + 2], c, -mc)); p3 = fast_exp2(fmaf(sv[i + 3], c, -mc));
        ls[0] += p0; ls[1] += p1; ls[2] += p2; ls[3] += p3;
      }
      if (MODE == 1 || MODE == 4 || MODE == 5) { xacc ^= pack_bf16(p0, p1); xacc ^= pack_bf16(p2, p3); }
      if (MODE == 3) {
        xacc ^= __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
        xacc ^= __byte_perm(__float_as_uint(p2), __float_as_uint(p3), 0x7632);
      }
    }
    acc += (ls[0] + ls[1]) + (ls[2] + ls[3]);
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(xacc & 0x3fffffffu);
}

template <int MODE>
void run(const char* name, long long* cyc, float* sink, const float* src) {
  const int iters = 2000;
  for (int W : {4, 8}) {
    bench<MODE><<<148, 32 * W>>>(iters, cyc, sink, src);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return; }
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < W; w++) mx = h[w] > mx ? h[w] : mx;
    printf("%-44s warps/SMSP %d: %7.1f clk per 128 elements per warp  (%5.2f clk / element / SMSP)\n", name, W / 4,
           (double)mx / iters, (double)mx / iters / 128.0 / (8.0 / W) );
  }
}

int main() {
  long long* cyc; float *sink, *src;
  cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
  cudaMalloc(&sink, 148 * 512 * sizeof(float));
  cudaMalloc(&src, 1024 * sizeof(float));
  cudaMemset(src, 0, 1024 * sizeof(float));
  run<0>("FFMA + MUFU.EX2 + FADD", cyc, sink, src);
  run<1>("FFMA + MUFU.EX2 + FADD + F2FP/2", cyc, sink, src);
  run<3>("FFMA + MUFU.EX2 + FADD + PRMT/2", cyc, sink, src);
  run<4>("F2FP/2 only", cyc, sink, src);
  run<5>("FFMA + MUFU.EX2 + FADD + F2FP/2 + 64 FMNMX3", cyc, sink, src);
  return 0;
}
