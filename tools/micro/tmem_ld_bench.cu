// TMEM -> register load throughput on sm_100a (tcgen05.ld.32x32b.x32), the path that feeds the attention
// softmax: W warps per CTA each load 128 columns x 32 lanes x 4 B = 16 KB per iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I simplesfm_b200/csrc -I include \
//        tools/micro/tmem_ld_bench.cu -o tools/micro/tmem_ld_bench && tools/micro/tmem_ld_bench
#include <cstdio>
#include "tc_common.cuh"
using namespace sfm::tc;
namespace sfm { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*, const char*, int) { return -2; } void count_launches(long long) {} }

// mode 0: loads only; mode 1: loads + 128 ex2 per thread per iteration; mode 2: ex2 only
__global__ void __launch_bounds__(512, 1) bench(int iters, int mode, long long* cycles, float* sink) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tbase);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  float acc = 0.f;
  float x = lane * 0.001f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    uint32_t r[4][32];
    if (mode != 2) {
#pragma unroll
      for (int ch = 0; ch < 4; ch++) tmem_ld_32x32(lane_addr + ch * 32, r[ch]);
      tmem_wait_ld();
#pragma unroll
      for (int ch = 0; ch < 4; ch++) {
        uint32_t x4 = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) x4 ^= r[ch][i];     // every loaded register is consumed (static indices)
        acc += __uint_as_float(x4 & 0x3fffffffu);
      }
    }
    if (mode != 0) {
      float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float y = -(float)(it & 7);
#pragma unroll
      for (int i = 0; i < 128; i++) a8[i & 7] += fast_exp2(fmaf(x, (float)i, y));   // 128 independent ex2
#pragma unroll
      for (int i = 0; i < 8; i++) acc += a8[i];
    }
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
  cudaMalloc(&sink, 148 * 512 * sizeof(float));
  const int iters = 2000;
  for (int mode = 0; mode < 3; mode++)
    for (int W : {1, 4, 8, 16}) {
      bench<<<148, 32 * W>>>(iters, mode, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[16];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < W; w++) mx = h[w] > mx ? h[w] : mx;
      const double per_it = (double)mx / iters;
      printf("mode %d (%s) warps %2d: %8.1f clk / iteration / warp", mode, mode == 0 ? "ld only" : mode == 1 ? "ld + 128 ex2" : "128 ex2 only", W, per_it);
      if (mode != 2) printf("  -> %7.1f B/clk/SM TMEM read", W * 16384.0 / per_it);
      if (mode != 0) printf("  -> %5.2f ex2/clk/SM", W * 32 * 128.0 / per_it);
      printf("\n");
    }
  return 0;
}
