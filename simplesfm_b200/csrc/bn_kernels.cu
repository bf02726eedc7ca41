#include "bn_kernels.cuh"

namespace sfm {

namespace {

constexpr int RB = 64;  // rows per partial block

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// Per (row block, column): mean and centred second moment of up to RB rows (two passes over a
// block that stays in L1/L2).  256 threads = (256 / min(C,256)) row lanes x columns.
template <typename T>
__global__ void __launch_bounds__(256) bn_partial_kernel(const T* __restrict__ X, int ld, int C, BnGroups gr, int nb0,
                                                         int nb1, float* __restrict__ partial) {
  __shared__ float red[256];
  __shared__ float smean[256];
  const int blk = blockIdx.x;
  const int side = blk >= gr.G * nb0 ? 1 : 0;
  const int local = side ? blk - gr.G * nb0 : blk;
  const int nb = side ? nb1 : nb0;
  const int grp = local / nb, bi = local % nb;
  const long long rows_g = (side ? gr.T1 : gr.T0) / gr.G;
  const long long base = (side ? gr.T0 : 0) + grp * rows_g;
  const long long r0 = base + (long long)bi * RB;
  const int rows = (int)min((long long)RB, base + rows_g - r0);
  const int cw = C < 256 ? C : 256;       // columns handled per sweep
  const int lanes = 256 / cw;             // row lanes
  const int cl = threadIdx.x % cw, rl = threadIdx.x / cw;
  for (int c0 = 0; c0 < C; c0 += cw) {
    const int c = c0 + cl;
    const T* col = X + r0 * ld + c;
    float s = 0.f;
    if (rl < lanes)
      for (int r = rl; r < rows; r += lanes) s += ldf(col + (long long)r * ld);
    red[threadIdx.x] = s;
    __syncthreads();
    if (rl == 0) {
      float t = 0.f;
      for (int k = 0; k < lanes; k++) t += red[k * cw + cl];
      smean[cl] = t / (float)rows;
    }
    __syncthreads();
    const float mu = smean[cl];
    s = 0.f;
    if (rl < lanes)
      for (int r = rl; r < rows; r += lanes) {
        float d = ldf(col + (long long)r * ld) - mu;
        s = fmaf(d, d, s);
      }
    __syncthreads();
    red[threadIdx.x] = s;
    __syncthreads();
    if (rl == 0) {
      float t = 0.f;
      for (int k = 0; k < lanes; k++) t += red[k * cw + cl];
      partial[((long long)blk * C + c) * 2] = mu;
      partial[((long long)blk * C + c) * 2 + 1] = t;
    }
    __syncthreads();
  }
}

// bf16 fast path (C % 8 == 0, C <= 2048): one pass, 16-byte loads, thread = 8 consecutive channels,
// 256 / (C/8) row lanes; (sum, sum of squares) in fp32 over <= RB rows -> (mean, centred M2).
__global__ void __launch_bounds__(256) bn_partial_bf16_kernel(const __nv_bfloat16* __restrict__ X, int ld, int C,
                                                              BnGroups gr, int nb0, int nb1,
                                                              float* __restrict__ partial) {
  extern __shared__ float red[];  // [lanes][C][2]
  const int blk = blockIdx.x;
  const int side = blk >= gr.G * nb0 ? 1 : 0;
  const int local = side ? blk - gr.G * nb0 : blk;
  const int nb = side ? nb1 : nb0;
  const int grp = local / nb, bi = local % nb;
  const long long rows_g = (side ? gr.T1 : gr.T0) / gr.G;
  const long long base = (side ? gr.T0 : 0) + grp * rows_g;
  const long long r0 = base + (long long)bi * RB;
  const int rows = (int)min((long long)RB, base + rows_g - r0);
  const int tpr = C / 8;             // threads per row
  const int lanes = 256 / tpr;       // row lanes
  const int cl = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; k++) s[k] = q[k] = 0.f;
  if (rl < lanes) {
    const uint4* src = reinterpret_cast<const uint4*>(X + r0 * ld + cl * 8);
    const int ld16 = ld / 8;
    for (int r = rl; r < rows; r += lanes) {
      uint4 raw = src[(long long)r * ld16];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float2 v = __bfloat1622float2(h[k]);
        s[2 * k] += v.x; s[2 * k + 1] += v.y;
        q[2 * k] = fmaf(v.x, v.x, q[2 * k]); q[2 * k + 1] = fmaf(v.y, v.y, q[2 * k + 1]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      red[(rl * C + cl * 8 + k) * 2] = s[k];
      red[(rl * C + cl * 8 + k) * 2 + 1] = q[k];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float ts = 0.f, tq = 0.f;
    for (int l = 0; l < lanes; l++) {
      ts += red[(l * C + c) * 2];
      tq += red[(l * C + c) * 2 + 1];
    }
    const float mu = ts / (float)rows;
    partial[((long long)blk * C + c) * 2] = mu;
    partial[((long long)blk * C + c) * 2 + 1] = fmaxf(tq - ts * mu, 0.f);
  }
}

// Chan's pairwise combination of the block moments (fixed tree => deterministic), one warp per
// (group, channel); then scale = gamma / sqrt(var + eps), shift = beta - mean * scale.
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float qb) {
  if (nb == 0.f) return;
  const float tot = n + nb;
  const float d = mb - mean;
  mean += d * (nb / tot);
  m2 += qb + d * d * (n * nb / tot);
  n = tot;
}

__global__ void __launch_bounds__(256) bn_final_kernel(const float* __restrict__ partial, int C, BnGroups gr, int nb0,
                                                       int nb1, int rb, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       float* __restrict__ scale, float* __restrict__ shift) {
  const int g = blockIdx.y;   // side * G + group
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const int side = g >= gr.G ? 1 : 0, grp = side ? g - gr.G : g;
  const int nb = side ? nb1 : nb0;
  const int b0 = side ? gr.G * nb0 + grp * nb1 : grp * nb0;
  const long long n_total = (side ? gr.T1 : gr.T0) / gr.G;
  float mean = 0.f, m2 = 0.f, n = 0.f;
  for (int b = lane; b < nb; b += 32) {
    const float rows = (float)min((long long)rb, n_total - (long long)b * rb);
    const float2 pm = *reinterpret_cast<const float2*>(partial + ((long long)(b0 + b) * C + c) * 2);
    chan_merge(n, mean, m2, rows, pm.x, pm.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float n2 = __shfl_xor_sync(0xffffffffu, n, o);
    const float me2 = __shfl_xor_sync(0xffffffffu, mean, o);
    const float q2 = __shfl_xor_sync(0xffffffffu, m2, o);
    // merge in a lane-symmetric order so every lane ends with bit-identical totals
    if (lane & o) {
      float tn = n2, tm = me2, tq = q2;
      chan_merge(tn, tm, tq, n, mean, m2);
      n = tn; mean = tm; m2 = tq;
    } else {
      chan_merge(n, mean, m2, n2, me2, q2);
    }
  }
  if (lane == 0) {
    const float var = n > 0.f ? m2 / n : 0.f;
    const float a = gamma[c] * (1.0f / sqrtf(var + eps));
    scale[g * C + c] = a;
    shift[g * C + c] = beta[c] - mean * a;
  }
}

// Same result for partials whose blocks all hold exactly `rb` rows (what the GEMM epilogue writes): one
// coalesced pass, no divisions in the loop.  CTA = (32 channels, group); lane = channel, 32 warps stride the
// blocks.  total M2 = sum M2_b + rb * (sum mean_b^2 - (sum mean_b)^2 / nb): only the (small) between-block term
// is formed by subtraction.
__global__ void __launch_bounds__(1024) bn_final_equal_kernel(const float* __restrict__ partial, int C, BnGroups gr,
                                                              int nb0, int nb1, int rb, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps,
                                                              float* __restrict__ scale, float* __restrict__ shift) {
  __shared__ float red[3][32][33];
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int side = g >= gr.G ? 1 : 0, grp = side ? g - gr.G : g;
  const int nb = side ? nb1 : nb0;
  const long long b0 = side ? (long long)gr.G * nb0 + (long long)grp * nb1 : (long long)grp * nb0;
  const float2* src = reinterpret_cast<const float2*>(partial) + b0 * C + c;
  float s = 0.f, ss = 0.f, q = 0.f;
#pragma unroll 4
  for (int b = w; b < nb; b += 32) {
    const float2 pm = __ldg(src + (long long)b * C);
    s += pm.x;
    ss = fmaf(pm.x, pm.x, ss);
    q += pm.y;
  }
  red[0][w][lane] = s;
  red[1][w][lane] = ss;
  red[2][w][lane] = q;
  __syncthreads();
  if (w == 0) {
    s = ss = q = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k++) {
      s += red[0][k][lane];
      ss += red[1][k][lane];
      q += red[2][k][lane];
    }
    const float mean = s / (float)nb;
    const float m2 = q + (float)rb * fmaxf(ss - s * mean, 0.f);
    const float var = m2 / ((float)nb * (float)rb);
    const float a = gamma[c] * (1.0f / sqrtf(var + eps));
    scale[g * C + c] = a;
    shift[g * C + c] = beta[c] - mean * a;
  }
}

__device__ __forceinline__ int group_of(const BnGroups& gr, long long row) {
  if (row >= gr.T0) return gr.G + (int)((row - gr.T0) / (gr.T1 / gr.G));
  return (int)(row / (gr.T0 / gr.G));
}

__global__ void bn_apply_f32_kernel(float4* __restrict__ X, int ld4, int C4, BnGroups gr, long long total4,
                                    const float4* __restrict__ scale, const float4* __restrict__ shift) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = i % C4;
  const long long row = i / C4;
  const int g = group_of(gr, row);
  float4 x = X[row * ld4 + c4];
  const float4 a = scale[g * C4 + c4], b = shift[g * C4 + c4];
  x.x = fmaxf(fmaf(x.x, a.x, b.x), 0.f);
  x.y = fmaxf(fmaf(x.y, a.y, b.y), 0.f);
  x.z = fmaxf(fmaf(x.z, a.z, b.z), 0.f);
  x.w = fmaxf(fmaf(x.w, a.w, b.w), 0.f);
  X[row * ld4 + c4] = x;
}

__global__ void bn_apply_bf16_kernel(uint4* __restrict__ X, int ld8, int C8, BnGroups gr, long long total8,
                                     const float4* __restrict__ scale, const float4* __restrict__ shift) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = i % C8;
  const long long row = i / C8;
  const int g = group_of(gr, row);
  uint4 raw = X[row * ld8 + c8];
  const float4 a0 = scale[(g * C8 + c8) * 2], a1 = scale[(g * C8 + c8) * 2 + 1];
  const float4 b0 = shift[(g * C8 + c8) * 2], b1 = shift[(g * C8 + c8) * 2 + 1];
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  float2 v0 = __bfloat1622float2(h[0]), v1 = __bfloat1622float2(h[1]), v2 = __bfloat1622float2(h[2]),
         v3 = __bfloat1622float2(h[3]);
  h[0] = __floats2bfloat162_rn(fmaxf(fmaf(v0.x, a0.x, b0.x), 0.f), fmaxf(fmaf(v0.y, a0.y, b0.y), 0.f));
  h[1] = __floats2bfloat162_rn(fmaxf(fmaf(v1.x, a0.z, b0.z), 0.f), fmaxf(fmaf(v1.y, a0.w, b0.w), 0.f));
  h[2] = __floats2bfloat162_rn(fmaxf(fmaf(v2.x, a1.x, b1.x), 0.f), fmaxf(fmaf(v2.y, a1.y, b1.y), 0.f));
  h[3] = __floats2bfloat162_rn(fmaxf(fmaf(v3.x, a1.z, b1.z), 0.f), fmaxf(fmaf(v3.y, a1.w, b1.w), 0.f));
  X[row * ld8 + c8] = raw;
}

__global__ void bn_apply_f32_to_bf16_kernel(const float4* __restrict__ X, int ld4, uint2* __restrict__ out, int ldo4,
                                            int C4, BnGroups gr, long long total4, const float4* __restrict__ scale,
                                            const float4* __restrict__ shift) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = i % C4;
  const long long row = i / C4;
  const int g = group_of(gr, row);
  const float4 x = X[row * ld4 + c4];
  const float4 a = scale[g * C4 + c4], b = shift[g * C4 + c4];
  __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(fmaf(x.x, a.x, b.x), 0.f), fmaxf(fmaf(x.y, a.y, b.y), 0.f));
  __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(fmaf(x.z, a.z, b.z), 0.f), fmaxf(fmaf(x.w, a.w, b.w), 0.f));
  out[row * ldo4 + c4] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

}  // namespace

static int blocks_per_group(long long rows_total, int G) { return cdiv(rows_total / G, RB); }

int bn_stat_blocks(const BnGroups& g) { return g.G * (blocks_per_group(g.T0, g.G) + blocks_per_group(g.T1, g.G)); }

template <typename T>
int bn_train_stats(const T* X, int ld, int C, const BnGroups& g, const float* gamma, const float* beta, float eps,
                   float* partial, float* scale, float* shift, cudaStream_t st) {
  SFM_REQUIRE(C <= 256 ? (256 % C == 0) : (C % 256 == 0), "bn_train_stats: unsupported channel count %d", C);
  SFM_REQUIRE(g.G >= 1 && g.T0 % g.G == 0 && g.T1 % g.G == 0, "bn_train_stats: rows not divisible into %d groups", g.G);
  const int nb0 = blocks_per_group(g.T0, g.G), nb1 = blocks_per_group(g.T1, g.G);
  const int total = g.G * (nb0 + nb1);
  if (total == 0) return SFM_OK;
  if (sizeof(T) == 2 && C % 8 == 0 && 256 % (C / 8) == 0 && ld % 8 == 0) {
    const int lanes = 256 / (C / 8);
    bn_partial_bf16_kernel<<<total, 256, (size_t)lanes * C * 2 * sizeof(float), st>>>(
        reinterpret_cast<const __nv_bfloat16*>(X), ld, C, g, nb0, nb1, partial);
  } else {
    bn_partial_kernel<T><<<total, 256, 0, st>>>(X, ld, C, g, nb0, nb1, partial);
  }
  bn_final_kernel<<<dim3(cdiv(C, 8), 2 * g.G), 256, 0, st>>>(partial, C, g, nb0, nb1, RB, gamma, beta, eps, scale, shift);
  SFM_LAUNCH_CHECK_N(2);
  return SFM_OK;
}

int bn_finalize(const float* partial, int rb, int C, const BnGroups& g, const float* gamma, const float* beta, float eps,
                float* scale, float* shift, cudaStream_t st) {
  SFM_REQUIRE(g.G >= 1 && (g.T0 / g.G) % rb == 0 && (g.T1 / g.G) % rb == 0 && g.T0 % g.G == 0 && g.T1 % g.G == 0,
              "bn_finalize: statistic groups must be whole %d-row blocks", rb);
  const int nb0 = (int)(g.T0 / g.G / rb), nb1 = (int)(g.T1 / g.G / rb);
  if (g.G * (nb0 + nb1) == 0) return SFM_OK;
  if (C % 32 == 0 && nb0 > 0 && nb1 > 0)
    bn_final_equal_kernel<<<dim3(C / 32, 2 * g.G), 1024, 0, st>>>(partial, C, g, nb0, nb1, rb, gamma, beta, eps, scale,
                                                                  shift);
  else
    bn_final_kernel<<<dim3(cdiv(C, 8), 2 * g.G), 256, 0, st>>>(partial, C, g, nb0, nb1, rb, gamma, beta, eps, scale, shift);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

template <>
int bn_apply_relu<float>(float* X, int ld, int C, const BnGroups& g, const float* scale, const float* shift,
                         cudaStream_t st) {
  SFM_REQUIRE(C % 4 == 0 && ld % 4 == 0, "bn_apply_relu: C and ld must be multiples of 4");
  const long long total = (g.T0 + g.T1) * (C / 4);
  if (total == 0) return SFM_OK;
  bn_apply_f32_kernel<<<cdiv(total, 256), 256, 0, st>>>(reinterpret_cast<float4*>(X), ld / 4, C / 4, g, total,
                                                        reinterpret_cast<const float4*>(scale),
                                                        reinterpret_cast<const float4*>(shift));
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

template <>
int bn_apply_relu<__nv_bfloat16>(__nv_bfloat16* X, int ld, int C, const BnGroups& g, const float* scale,
                                 const float* shift, cudaStream_t st) {
  SFM_REQUIRE(C % 8 == 0 && ld % 8 == 0, "bn_apply_relu: C and ld must be multiples of 8");
  const long long total = (g.T0 + g.T1) * (C / 8);
  if (total == 0) return SFM_OK;
  bn_apply_bf16_kernel<<<cdiv(total, 256), 256, 0, st>>>(reinterpret_cast<uint4*>(X), ld / 8, C / 8, g, total,
                                                         reinterpret_cast<const float4*>(scale),
                                                         reinterpret_cast<const float4*>(shift));
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int bn_apply_relu_f32_to_bf16(const float* X, int ld, __nv_bfloat16* out, int ld_out, int C, const BnGroups& g,
                              const float* scale, const float* shift, cudaStream_t st) {
  SFM_REQUIRE(C % 4 == 0 && ld % 4 == 0 && ld_out % 4 == 0, "bn_apply_relu_f32_to_bf16: C, ld must be multiples of 4");
  const long long total = (g.T0 + g.T1) * (C / 4);
  if (total == 0) return SFM_OK;
  bn_apply_f32_to_bf16_kernel<<<cdiv(total, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(X), ld / 4,
                                                                reinterpret_cast<uint2*>(out), ld_out / 4, C / 4, g,
                                                                total, reinterpret_cast<const float4*>(scale),
                                                                reinterpret_cast<const float4*>(shift));
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

template int bn_train_stats<float>(const float*, int, int, const BnGroups&, const float*, const float*, float, float*,
                                   float*, float*, cudaStream_t);
template int bn_train_stats<__nv_bfloat16>(const __nv_bfloat16*, int, int, const BnGroups&, const float*, const float*,
                                           float, float*, float*, float*, cudaStream_t);

}  // namespace sfm
