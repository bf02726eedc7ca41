// Element-wise / reduction helpers of the bf16 path: casts, train-mode BatchNorm statistics
// (superglue.py:56-57 as the reference runs it) and the fused normalise + ReLU.
#include "tc_common.cuh"
#include "tc_path.cuh"

namespace sfm {
namespace tc {

namespace {

constexpr int STAT_ROWS = 128;  // rows per partial block

__global__ void cast_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = in[i];
  out[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}

// grid.x = row blocks of group 0 followed by row blocks of group 1; block = C/2 threads (2 columns each)
__global__ void stats_partial_kernel(const __nv_bfloat162* __restrict__ X, int C2, long long T0, long long T1,
                                     int blocks0, float* __restrict__ partial) {
  const int blk = blockIdx.x;
  const int g = blk >= blocks0 ? 1 : 0;
  const long long base = g ? T0 : 0, n = g ? T1 : T0;
  const long long r0 = base + (long long)(g ? blk - blocks0 : blk) * STAT_ROWS;
  const long long r1 = min(base + n, r0 + STAT_ROWS);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (int c = threadIdx.x; c < C2; c += blockDim.x) {
    s0 = s1 = q0 = q1 = 0.f;
    for (long long r = r0; r < r1; r++) {
      float2 v = __bfloat1622float2(X[r * C2 + c]);
      s0 += v.x; s1 += v.y;
      q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1);
    }
    float4* dst = reinterpret_cast<float4*>(partial + ((long long)blk * C2 + c) * 4);
    *dst = make_float4(s0, s1, q0, q1);
  }
}

// grid.x = 2 groups; one thread per column pair: deterministic fixed-order reduction of the partials
__global__ void stats_final_kernel(const float* __restrict__ partial, int C2, long long T0, long long T1, int blocks0,
                                   int blocks1, float* __restrict__ mean, float* __restrict__ var) {
  const int g = blockIdx.x;
  const int b0 = g ? blocks0 : 0, nb = g ? blocks1 : blocks0;
  const float n = (float)(g ? T1 : T0);
  for (int c = threadIdx.x; c < C2; c += blockDim.x) {
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    for (int b = 0; b < nb; b++) {
      float4 v = *reinterpret_cast<const float4*>(partial + ((long long)(b0 + b) * C2 + c) * 4);
      s0 += v.x; s1 += v.y; q0 += v.z; q1 += v.w;
    }
    float m0 = s0 / n, m1 = s1 / n;
    mean[g * 2 * C2 + 2 * c] = m0;
    mean[g * 2 * C2 + 2 * c + 1] = m1;
    var[g * 2 * C2 + 2 * c] = fmaxf(q0 / n - m0 * m0, 0.f);
    var[g * 2 * C2 + 2 * c + 1] = fmaxf(q1 / n - m1 * m1, 0.f);
  }
}

__global__ void bn_relu_bf16_kernel(uint4* __restrict__ X, int C8, long long T0, long long total8,
                                    const float* __restrict__ mean, const float* __restrict__ var,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = i % C8;
  const long long row = i / C8;
  const int g = row >= T0 ? 1 : 0;
  const int C = C8 * 8;
  uint4 raw = X[i];
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float2 v = __bfloat1622float2(h[k]);
    const int c = c8 * 8 + 2 * k;
    float a0 = gamma[c] * rsqrtf(var[g * C + c] + eps), a1 = gamma[c + 1] * rsqrtf(var[g * C + c + 1] + eps);
    float y0 = fmaxf(fmaf(v.x - mean[g * C + c], a0, beta[c]), 0.f);
    float y1 = fmaxf(fmaf(v.y - mean[g * C + c + 1], a1, beta[c + 1]), 0.f);
    o[k] = pack_bf16(y0, y1);
  }
  X[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace

int f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t st) {
  SFM_REQUIRE(n % 4 == 0, "f32_to_bf16: n must be a multiple of 4");
  if (n == 0) return SFM_OK;
  cast_kernel<<<cdiv(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out),
                                                n / 4);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int stat_blocks_for(long long T0, long long T1) { return cdiv(T0, STAT_ROWS) + cdiv(T1, STAT_ROWS); }

int col_stats_bf16(const __nv_bfloat16* X, int C, long long T0, long long T1, float* partial, float* mean, float* var,
                   cudaStream_t st) {
  SFM_REQUIRE(C % 2 == 0, "col_stats_bf16: C must be even");
  const int b0 = cdiv(T0, STAT_ROWS), b1 = cdiv(T1, STAT_ROWS);
  if (b0 + b1 == 0) return SFM_OK;
  stats_partial_kernel<<<b0 + b1, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat162*>(X), C / 2, T0, T1, b0, partial);
  stats_final_kernel<<<2, 256, 0, st>>>(partial, C / 2, T0, T1, b0, b1, mean, var);
  SFM_LAUNCH_CHECK_N(2);
  return SFM_OK;
}

int bn_relu_bf16(__nv_bfloat16* X, int C, long long T0, long long T1, const float* mean, const float* var,
                 const float* gamma, const float* beta, float eps, cudaStream_t st) {
  SFM_REQUIRE(C % 8 == 0, "bn_relu_bf16: C must be a multiple of 8");
  const long long total8 = (T0 + T1) * (C / 8);
  if (total8 == 0) return SFM_OK;
  bn_relu_bf16_kernel<<<cdiv(total8, 256), 256, 0, st>>>(reinterpret_cast<uint4*>(X), C / 8, T0, total8, mean, var,
                                                         gamma, beta, eps);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

}  // namespace tc
}  // namespace sfm
