// fp32 SIMT GEMM / implicit-GEMM 3x3 convolution (the exact-parity arithmetic path).
// tcgen05 has no fp32 MMA kind (tf32/f16/f8f6f4 only), so the "fp32" precision mode that must
// reproduce the reference's mutual matches bit-for-bit runs on FFMA; the bf16 throughput path
// lives in tc_*.cu.
#pragma once
#include "common.cuh"

namespace sfm {

enum : int { GEMM_RELU = 1, GEMM_ACCUM = 2 };

struct GemmF32 {
  const float* A = nullptr;   // [M, K] row-major, leading dim lda  (columns k < K1)
  int lda = 0;
  const float* A2 = nullptr;  // optional second operand for columns k >= K1 (concat along K)
  int lda2 = 0;
  int K1 = 0;                 // split point (multiple of 16); ignored when A2 == nullptr
  const float* B = nullptr;   // NT: [N, K] row-major (ldb);  NN: [K, N] row-major (ldb)
  int ldb = 0;
  float* C = nullptr;         // [M, N] row-major, ldc
  int ldc = 0;
  const float* bias = nullptr;  // [N] or nullptr
  float alpha = 1.f;
  int M = 0, N = 0, K = 0;
  int flags = 0;
  // batching: blockIdx.z = z -> (z / nb1, z % nb1)
  int batch = 1;
  int nb1 = 1;
  long long sA0 = 0, sA1 = 0, sB0 = 0, sB1 = 0, sC0 = 0, sC1 = 0;
  // implicit 3x3/pad-1 convolution over NHWC input: A = input [n_img, H, W, Cin], M = n_img*H*W,
  // K = 9*Cin ordered (ky, kx, ci), B = repacked weights [Cout, 9*Cin]
  int convH = 0, convW = 0, convCin = 0;
};

int gemm_f32_nt(const GemmF32& p, cudaStream_t st);   // C = A * B^T
int gemm_f32_nn(const GemmF32& p, cudaStream_t st);   // C = A * B
int conv3x3_f32(const GemmF32& p, cudaStream_t st);   // implicit GEMM, NHWC

// conv1a: single input channel, 3x3, pad 1, + bias + ReLU.  in [n,H,W], w [64][9], out NHWC [n,H,W,64]
// writes fp32 NHWC `out`, or bf16 NHWC `out_h` when that is non-null
int conv1a_f32(const float* img, const float* w9, const float* bias, float* out, __nv_bfloat16* out_h, int n_img,
               int H, int W, cudaStream_t st);
// 2x2/2 max pool on NHWC fp32
int maxpool2_nhwc_f32(const float* in, float* out, int n_img, int H, int W, int C, cudaStream_t st);

}  // namespace sfm
