#include "sg_kernels.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace sfm {

namespace {

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frontend_kernel(const float* __restrict__ desc, long long img_stride, int ldk,
                                                       const float* __restrict__ kpts,
                                                       const float* __restrict__ scores, int kcap,
                                                       const int* __restrict__ idx, int K, int imgH, int imgW,
                                                       float* __restrict__ X, float* __restrict__ kin) {
  __shared__ float tile[32][33];
  const int b = blockIdx.y;
  const int img = idx ? idx[b] : b;
  const int k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* d = desc + (long long)img * img_stride;
  for (int cb = 0; cb < 8; cb++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      int c = cb * 32 + ty * 4 + r;
      int k = k0 + tx;
      tile[ty * 4 + r][tx] = k < K ? d[(long long)c * ldk + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
      int k = k0 + ty * 4 + r;
      if (k < K) X[((long long)b * K + k) * 256 + cb * 32 + tx] = tile[tx][ty * 4 + r];
    }
    __syncthreads();
  }
  if (kin != nullptr && threadIdx.x < 32) {
    int k = k0 + threadIdx.x;
    if (k < K) {
      // normalize_keypoints: (kp - size/2) / (0.7 * max(W, H))   superglue.py:62-69
      float w = (float)imgW, h = (float)imgH;
      float scaling = __fmul_rn(fmaxf(w, h), 0.7f);
      float x = kpts[((long long)img * kcap + k) * 2 + 0];
      float y = kpts[((long long)img * kcap + k) * 2 + 1];
      float4 o;
      o.x = __fdiv_rn(__fsub_rn(x, __fdiv_rn(w, 2.f)), scaling);
      o.y = __fdiv_rn(__fsub_rn(y, __fdiv_rn(h, 2.f)), scaling);
      o.z = scores[(long long)img * kcap + k];
      o.w = 0.f;
      reinterpret_cast<float4*>(kin)[(long long)b * K + k] = o;
    }
  }
}

__global__ void softmax_rows_kernel(float* __restrict__ S, long long rows, int M, int ld) {
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = S + row * ld;
  float m = -INFINITY;
  for (int j = lane; j < M; j += 32) m = fmaxf(m, p[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < M; j += 32) {
    float e = expf(p[j] - m);
    p[j] = e;
    s += e;
  }
  s = warp_sum(s);
  for (int j = lane; j < M; j += 32) p[j] = p[j] / s;
  for (int j = M + lane; j < ld; j += 32) p[j] = 0.f;   // keep the row padding at zero (GEMM reads float4)
}

// ------------------------------------------------------------------------------------------
// Sinkhorn.  Z = [[S, a], [a, a]] is never built: the dustbin row/column are the constant a.
// FAST = true: ex2.approx based (bf16 precision mode); false: IEEE-accurate expf (fp32 parity mode)
template <bool FAST>
__device__ __forceinline__ float ex(float x) {
  return FAST ? __expf(x) : expf(x);
}
template <bool FAST = false>
__device__ __forceinline__ void lse_push(float& m, float& s, float x) {
  if (x > m) {
    s = s * ex<FAST>(m - x) + 1.f;
    m = x;
  } else {
    s += ex<FAST>(x - m);
  }
}
template <bool FAST = false>
__device__ __forceinline__ void lse_join(float& m, float& s, float m2, float s2) {
  float mm = fmaxf(m, m2);
  if (mm == -INFINITY) {
    s = 0.f;
    m = mm;
    return;
  }
  s = s * ex<FAST>(m - mm) + s2 * ex<FAST>(m2 - mm);
  m = mm;
}
// fold four values into a running (max, sum) with one rescale
template <bool FAST>
__device__ __forceinline__ void lse_push4(float& m, float& s, float x0, float x1, float x2, float x3) {
  const float cm = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
  if (cm == -INFINITY) return;
  if (cm > m) {
    s *= ex<FAST>(m - cm);
    m = cm;
  }
  s += (ex<FAST>(x0 - m) + ex<FAST>(x1 - m)) + (ex<FAST>(x2 - m) + ex<FAST>(x3 - m));
}

// u_i = log_mu_i - LSE_j(Z_ij + v_j); one warp per row (rows 0..N, row N is the dustbin row)
template <bool FAST>
__global__ void __launch_bounds__(256) sk_row_kernel(const float* __restrict__ S, int N, int M, float alpha,
                                                     const float* __restrict__ v, float* __restrict__ u, float norm,
                                                     float log_mu_bin) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row > N) return;
  const float* vb = v + (long long)b * (M + 1);
  float m = -INFINITY, s = 0.f;
  if (row < N) {
    const float* sr = S + ((long long)b * N + row) * M;
    int j = lane;
    for (; j + 224 < M; j += 256) {   // 8 independent loads in flight per lane
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; q++) x[q] = sr[j + 32 * q] + vb[j + 32 * q];
      lse_push4<FAST>(m, s, x[0], x[1], x[2], x[3]);
      lse_push4<FAST>(m, s, x[4], x[5], x[6], x[7]);
    }
    for (; j + 96 < M; j += 128)
      lse_push4<FAST>(m, s, sr[j] + vb[j], sr[j + 32] + vb[j + 32], sr[j + 64] + vb[j + 64], sr[j + 96] + vb[j + 96]);
    for (; j < M; j += 32) lse_push<FAST>(m, s, sr[j] + vb[j]);
  } else {
    for (int j = lane; j < M; j += 32) lse_push<FAST>(m, s, alpha + vb[j]);
  }
  if (lane == 0) lse_push<FAST>(m, s, alpha + vb[M]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_join<FAST>(m, s, m2, s2);
  }
  if (lane == 0) u[(long long)b * (N + 1) + row] = (row < N ? norm : log_mu_bin) - (m + logf(s));
}

// column pass.  MODE 0: v_j = log_nu_j - LSE_i(Z_ij + u_i).  MODE 1: column argmax of
// ((S_ij + u_i) + v_j) - norm over i < N (first index on ties).
// grid (colblocks(128), R, B), 256 threads: warp w takes rows r0 + w, r0 + w + 8, ...
template <int MODE, bool FAST>
__global__ void __launch_bounds__(256) sk_col_kernel(const float* __restrict__ S, int N, int M, float alpha,
                                                     const float* __restrict__ u, const float* __restrict__ vin,
                                                     float* __restrict__ out_v, float* __restrict__ part,
                                                     int* __restrict__ counters, int R, float norm, float log_nu_bin,
                                                     float* __restrict__ max1, int* __restrict__ idx1) {
  __shared__ float sm_m[8][128];
  __shared__ float sm_s[8][128];   // MODE 1: bit pattern of the row index
  __shared__ int is_last;
  const int b = blockIdx.z, split = blockIdx.y, cb = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per = (N + R - 1) / R;
  const int r0 = split * rows_per, r1 = min(N, r0 + rows_per);
  const float* ub = u + (long long)b * (N + 1);
  const float* Sb = S + (long long)b * N * M;
  const int c0 = cb * 128;
  float m[4], s[4];
  int bi[4];
  float vj[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    m[q] = -INFINITY;
    s[q] = 0.f;
    bi[q] = 0x7fffffff;
    int j = c0 + lane + 32 * q;
    vj[q] = (MODE == 1 && j < M) ? vin[(long long)b * (M + 1) + j] : 0.f;
    if (MODE == 2) m[q] = -1.f;
  }
  if (MODE == 0) {
    // each warp folds 4 consecutive rows per step: 16 independent loads in flight per lane
    for (int i = r0 + warp * 4; i < r1; i += 32) {
      float x[4][4];
#pragma unroll
      for (int rr = 0; rr < 4; rr++) {
        const bool rok = i + rr < r1;
        const float ui = rok ? ub[i + rr] : 0.f;
        const float* sr = Sb + (long long)(i + rr) * M + c0 + lane;
#pragma unroll
        for (int q = 0; q < 4; q++) x[rr][q] = (rok && c0 + lane + 32 * q < M) ? sr[32 * q] + ui : -INFINITY;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) lse_push4<FAST>(m[q], s[q], x[0][q], x[1][q], x[2][q], x[3][q]);
    }
  } else {
    for (int i = r0 + warp; i < r1; i += 8) {
      const float ui = ub[i];
      const float* sr = Sb + (long long)i * M + c0 + lane;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        int j = c0 + lane + 32 * q;
        if (j < M) {
          float z = MODE == 2 ? sr[32 * q] * ui : __fsub_rn(__fadd_rn(__fadd_rn(sr[32 * q], ui), vj[q]), norm);
          if (z > m[q]) {
            m[q] = z;
            bi[q] = i;
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    sm_m[warp][lane + 32 * q] = m[q];
    sm_s[warp][lane + 32 * q] = MODE == 0 ? s[q] : __int_as_float(bi[q]);
  }
  __syncthreads();
  const int pstride = (M + 1) * 2;
  float* pb = part + ((long long)b * R + split) * pstride;
  if (threadIdx.x < 128) {
    int j = c0 + threadIdx.x;
    if (j < M) {
      float mm = sm_m[0][threadIdx.x], ss = sm_s[0][threadIdx.x];
      for (int w = 1; w < 8; w++) {
        float m2 = sm_m[w][threadIdx.x], s2 = sm_s[w][threadIdx.x];
        if (MODE == 0) {
          lse_join<FAST>(mm, ss, m2, s2);
        } else {
          int i1 = __float_as_int(ss), i2 = __float_as_int(s2);
          if (m2 > mm || (m2 == mm && i2 < i1)) {
            mm = m2;
            ss = s2;
          }
        }
      }
      __stcg(&pb[2 * j], mm);
      __stcg(&pb[2 * j + 1], ss);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    int prev = atomicAdd(&counters[b * gridDim.x + cb], 1);
    is_last = (prev == R - 1);
    if (is_last) counters[b * gridDim.x + cb] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float* p0 = part + (long long)b * R * pstride;
  if (threadIdx.x < 128) {
    int j = c0 + threadIdx.x;
    if (j < M) {
      float mm = __ldcg(&p0[2 * j]), ss = __ldcg(&p0[2 * j + 1]);
      for (int r = 1; r < R; r++) {
        float m2 = __ldcg(&p0[(long long)r * pstride + 2 * j]), s2 = __ldcg(&p0[(long long)r * pstride + 2 * j + 1]);
        if (MODE == 0) {
          lse_join<FAST>(mm, ss, m2, s2);
        } else {
          int i1 = __float_as_int(ss), i2 = __float_as_int(s2);
          if (m2 > mm || (m2 == mm && i2 < i1)) {
            mm = m2;
            ss = s2;
          }
        }
      }
      if (MODE == 0) {
        lse_join<FAST>(mm, ss, alpha + ub[N], 1.f);   // dustbin row
        out_v[(long long)b * (M + 1) + j] = norm - (mm + logf(ss));
      } else {
        int bidx = __float_as_int(ss);
        max1[(long long)b * M + j] = mm;
        idx1[(long long)b * M + j] = bidx == 0x7fffffff ? 0 : bidx;
      }
    }
  }
  if (MODE == 0 && cb == 0) {
    // dustbin column: LSE_i(alpha + u_i), i = 0..N
    float mm = -INFINITY, ss = 0.f;
    for (int i = threadIdx.x; i <= N; i += 256) lse_push<FAST>(mm, ss, alpha + ub[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, mm, o), s2 = __shfl_xor_sync(0xffffffffu, ss, o);
      lse_join<FAST>(mm, ss, m2, s2);
    }
    __syncthreads();
    if (lane == 0) {
      sm_m[0][warp] = mm;
      sm_s[0][warp] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; w++) lse_join<FAST>(mm, ss, sm_m[0][w], sm_s[0][w]);
      out_v[(long long)b * (M + 1) + M] = log_nu_bin - (mm + logf(ss));
    }
  }
}

// ------------------------------------------------------------------------------------------
// Scaling-form Sinkhorn (throughput mode).  The log-space iteration of the reference
//   u_i = log_mu_i - LSE_j(Z_ij + v_j),  v_j = log_nu_j - LSE_i(Z_ij + u_i)        (superglue.py:142-148)
// is algebraically the matrix-scaling iteration on K_ij = exp(S_ij - c_i) (c_i = row maximum, which keeps
// every row representable), A_i = exp(u_i + c_i), B_j = exp(v_j):
//   A_i = mu_i / (sum_j K_ij B_j + Kb_i B_M),   B_j = nu_j / (sum_i K_ij A_i + e^alpha A_N)
// with the dustbin column Kb_i = exp(alpha - c_i) and dustbin row e^alpha handled analytically.
// One exponential per element ONCE, then every iteration is two FMAs per element and the matrix is read
// from HBM a single time per iteration (fused row + column sweep).  fp32 throughout: the result differs
// from the log-space path only by rounding (<< the bf16 noise of the descriptors feeding it).
constexpr int FUSED_RB = 64;

// prologue: c_i = max_j S_ij ; S_ij <- exp(S_ij - c_i) in place ; kb_i = exp(alpha - c_i).  Warp per row.
// With k16 != nullptr the iterations run on a half-precision copy K16_ij = fp16(2^14 K_ij) (half the HBM bytes
// per sweep; the 2^14 keeps entries down to e^-19 of the row maximum in fp16's normal range, 11-bit mantissa);
// the final arg-max passes still read the fp32 K.
constexpr float K16_SCALE = 16384.f, K16_INV = 1.f / 16384.f;
// With rowmax != nullptr (and k16) the score matrix is left untouched: only the fp16 copy and c_i are written, the
// final row arg-max works on S + log B directly.
__global__ void __launch_bounds__(256) sks_prologue_kernel(float* __restrict__ S, int N, int M, float alpha,
                                                           float* __restrict__ kb, __half* __restrict__ k16,
                                                           float* __restrict__ rowmax) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float* sr = S + ((long long)b * N + row) * M;
  float m = -INFINITY;
  for (int j = lane; j < M; j += 32) m = fmaxf(m, sr[j]);
  m = warp_max(m);
  if (k16 != nullptr) {
    __half* hr = k16 + ((long long)b * N + row) * M;
    for (int j = lane * 2; j < M; j += 64) {   // M % 4 == 0
      const float2 sv = *reinterpret_cast<const float2*>(sr + j);
      const float e0 = __expf(sv.x - m), e1 = __expf(sv.y - m);
      if (rowmax == nullptr) *reinterpret_cast<float2*>(sr + j) = make_float2(e0, e1);
      *reinterpret_cast<__half2*>(hr + j) = __floats2half2_rn(e0 * K16_SCALE, e1 * K16_SCALE);
    }
  } else {
    for (int j = lane; j < M; j += 32) sr[j] = __expf(sr[j] - m);
  }
  if (lane == 0) {
    kb[(long long)b * N + row] = __expf(alpha - m);
    if (rowmax != nullptr) rowmax[(long long)b * N + row] = m;
  }
}

// fused sweep: A for the rows of this block, then the block's column sums of K_ij A_i
// Sum R per-lane values (one per row) across the 32 lanes of a warp with R - 1 + log2(32 / R) shuffles instead of
// 5 R: every round the lanes swap half of their values with a partner and keep the other half.  On return v[0]
// of lane L is the warp total of row L >> log2(32 / R).
template <int R>
__device__ __forceinline__ float transpose_reduce(float (&v)[R], int lane) {
  int off = 16;
#pragma unroll
  for (int n = R; n > 1; n >>= 1, off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < n / 2; k++) {
      const float send = hi ? v[k] : v[k + n / 2];
      const float keep = hi ? v[k + n / 2] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

// LAST (final iteration): the sweep also tracks, per column, the largest K_ij A_i of its 64 rows and the row it
// came from -- the column arg-max of the assignment (first index on ties) -- so that matching needs no separate
// pass over the matrix for it.  part then holds [sums | maxima | indices], each [B, nblk, M].
template <int G, bool H16, bool LAST>
__global__ void __launch_bounds__(256, 2) sks_fused_kernel(const void* __restrict__ Kp, int N, int M, float ealpha,
                                                        const float* __restrict__ kb, const float* __restrict__ Bv,
                                                        float* __restrict__ Av, float* __restrict__ part, float mu,
                                                        float mu_bin) {
  // rows per step: the fp16 matrix is kept PACKED in registers (converted once for the row sums and once for
  // the column sums), so twice as many rows are in flight for the same register budget -- the sweep is bound
  // by bytes in flight per SM, not by arithmetic
  constexpr int RS = (H16 ? 32 : 16) / G;
  constexpr int XW = H16 ? 2 : 4;   // registers per 4 matrix elements
  __shared__ float red[2][8][RS];
  __shared__ float sai[2][RS];
  __shared__ float red_bin[8];
  const int b = blockIdx.y, blk = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Kb_ = reinterpret_cast<const float*>(Kp) + (long long)b * N * M;          // !H16
  const __half* Kh_ = reinterpret_cast<const __half*>(Kp) + (long long)b * N * M;        // H16
  const float* Bb = Bv + (long long)b * (M + 1);
  float* Ab = Av + (long long)b * (N + 1);
  const float* kbb = kb + (long long)b * N;
  float bj[4 * G];
#pragma unroll
  for (int g = 0; g < G; g++) {
    const int col = (g * 256 + tid) * 4;
#pragma unroll
    for (int e = 0; e < 4; e++) bj[4 * g + e] = col < M ? Bb[col + e] : 0.f;   // (M+1)-strided: scalar loads
  }
  const float bbin = Bb[M];
  if (blk == 0) {   // dustbin row: A_N = mu_bin / (e^alpha (sum_j B_j + B_M))
    float sacc = 0.f;
#pragma unroll
    for (int c = 0; c < 4 * G; c++) sacc += bj[c];
    sacc = warp_sum(sacc);
    if (lane == 0) red_bin[warp] = sacc;
    __syncthreads();
    if (tid == 0) {
      float tot = bbin;
#pragma unroll
      for (int w = 0; w < 8; w++) tot += red_bin[w];
      Ab[N] = mu_bin / (ealpha * tot);
    }
  }
  float cs[4 * G];
  [[maybe_unused]] float bv[LAST ? 4 * G : 1];
  [[maybe_unused]] int bidx[LAST ? 4 * G : 1];
#pragma unroll
  for (int c = 0; c < 4 * G; c++) cs[c] = 0.f;
  if (LAST) {
#pragma unroll
    for (int c = 0; c < 4 * G; c++) {
      bv[LAST ? c : 0] = 0.f;            // K A > 0 for every real entry
      bidx[LAST ? c : 0] = 0x7fffffff;
    }
  }
  const int row0 = blk * FUSED_RB;
  int par = 0;
  auto unpack = [](const uint32_t* xr, float* f) {   // 4 matrix elements of one (row, g) slot
    if (H16) {
      // volatile: the two uses of a slot (row sums, column sums) must each re-convert from the packed
      // registers instead of keeping 4x as many floats live across the barrier
      asm volatile("{.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;}"
                   : "=f"(f[0]), "=f"(f[1]) : "r"(xr[0]));
      asm volatile("{.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;}"
                   : "=f"(f[2]), "=f"(f[3]) : "r"(xr[1]));
    } else {
      f[0] = __uint_as_float(xr[0]); f[1] = __uint_as_float(xr[1]);
      f[2] = __uint_as_float(xr[2]); f[3] = __uint_as_float(xr[3]);
    }
  };
  for (int step = 0; step < FUSED_RB / RS; step++, par ^= 1) {
    const int i0 = row0 + step * RS;
    if (i0 >= N) break;   // block-uniform
    uint32_t x[RS][G][XW];
#pragma unroll
    for (int r = 0; r < RS; r++) {
      const bool rok = i0 + r < N;
#pragma unroll
      for (int g = 0; g < G; g++) {
        const int col = (g * 256 + tid) * 4;
        if (H16) {
          const uint2 t = (rok && col < M) ? *reinterpret_cast<const uint2*>(Kh_ + (long long)(i0 + r) * M + col)
                                           : make_uint2(0u, 0u);
          x[r][g][0] = t.x; x[r][g][1] = t.y;
        } else {
          const uint4 t = (rok && col < M) ? *reinterpret_cast<const uint4*>(Kb_ + (long long)(i0 + r) * M + col)
                                           : make_uint4(0u, 0u, 0u, 0u);
          x[r][g][0] = t.x; x[r][g][1] = t.y; x[r][g][XW - 2] = t.z; x[r][g][XW - 1] = t.w;
        }
      }
    }
    float rs[RS];
#pragma unroll
    for (int r = 0; r < RS; r++) {
      float sacc = 0.f;
#pragma unroll
      for (int g = 0; g < G; g++) {
        float f[4];
        unpack(x[r][g], f);
#pragma unroll
        for (int e = 0; e < 4; e++) sacc = fmaf(f[e], bj[4 * g + e], sacc);
      }
      rs[r] = sacc;
    }
    {
      const float tot = transpose_reduce<RS>(rs, lane);
      constexpr int LPR = 32 / RS;   // lanes that end up with the same row
      if ((lane & (LPR - 1)) == 0) red[par][warp][lane / LPR] = tot;
    }
    __syncthreads();
    if (tid < RS) {   // one thread per row: A_i (the 8 warp partials, the dustbin term, one division)
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; w++) tot += red[par][w][tid];
      if (H16) tot *= K16_INV;
      const bool rok = i0 + tid < N;
      const float ai = rok ? mu / fmaf(kbb[rok ? i0 + tid : 0], bbin, tot) : 0.f;
      sai[par][tid] = ai;
      if (rok) Ab[i0 + tid] = ai;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS; r++) {
      const float ai = sai[par][r];
#pragma unroll
      for (int g = 0; g < G; g++) {
        float f[4];
        unpack(x[r][g], f);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          cs[4 * g + e] = fmaf(f[e], ai, cs[4 * g + e]);
          if (LAST) {
            const float z = f[e] * ai;
            if (z > bv[LAST ? 4 * g + e : 0]) {   // strict: rows ascend, so ties keep the first row
              bv[LAST ? 4 * g + e : 0] = z;
              bidx[LAST ? 4 * g + e : 0] = i0 + r;
            }
          }
        }
      }
    }
  }
  if (LAST) {
    const long long plane = (long long)gridDim.y * gridDim.x * M;
    float* pm = part + plane + ((long long)b * gridDim.x + blk) * M;
    int* pi = reinterpret_cast<int*>(part + 2 * plane) + ((long long)b * gridDim.x + blk) * M;
#pragma unroll
    for (int g = 0; g < G; g++) {
      const int col = (g * 256 + tid) * 4;
      if (col < M) {
        *reinterpret_cast<float4*>(pm + col) = make_float4(bv[LAST ? 4 * g : 0], bv[LAST ? 4 * g + 1 : 0],
                                                           bv[LAST ? 4 * g + 2 : 0], bv[LAST ? 4 * g + 3 : 0]);
        *reinterpret_cast<int4*>(pi + col) = make_int4(bidx[LAST ? 4 * g : 0], bidx[LAST ? 4 * g + 1 : 0],
                                                       bidx[LAST ? 4 * g + 2 : 0], bidx[LAST ? 4 * g + 3 : 0]);
      }
    }
  }
  float* pb = part + ((long long)b * gridDim.x + blk) * M;
#pragma unroll
  for (int g = 0; g < G; g++) {
    const int col = (g * 256 + tid) * 4;
    const float sc = H16 ? K16_INV : 1.f;
    if (col < M)
      *reinterpret_cast<float4*>(pb + col) =
          make_float4(cs[4 * g] * sc, cs[4 * g + 1] * sc, cs[4 * g + 2] * sc, cs[4 * g + 3] * sc);
  }
}

// B_j = nu / (sum over row blocks + e^alpha A_N);  last block: B_M = nu_bin / (sum_i kb_i A_i + e^alpha A_N)
__global__ void __launch_bounds__(256) sks_combine_kernel(const float* __restrict__ part, int nblk, int N, int M,
                                                          float ealpha, const float* __restrict__ kb,
                                                          const float* __restrict__ Av, float* __restrict__ Bv,
                                                          float nu, float nu_bin, int* __restrict__ col_idx) {
  __shared__ float red[8];
  const int b = blockIdx.y;
  const float* Ab = Av + (long long)b * (N + 1);
  float* Bb = Bv + (long long)b * (M + 1);
  const float an = Ab[N];
  if (blockIdx.x == gridDim.x - 1) {
    const float* kbb = kb + (long long)b * N;
    float sacc = 0.f;
    for (int i = threadIdx.x; i < N; i += 256) sacc = fmaf(kbb[i], Ab[i], sacc);
    sacc = warp_sum(sacc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = ealpha * an;
#pragma unroll
      for (int w = 0; w < 8; w++) tot += red[w];
      Bb[M] = nu_bin / tot;
    }
    return;
  }
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= M) return;
  const float* p = part + (long long)b * nblk * M + j;
  float tot = ealpha * an;
  for (int k = 0; k < nblk; k++) tot += p[(long long)k * M];
  Bb[j] = nu / tot;
  if (col_idx != nullptr) {   // last iteration: column arg-max over the row blocks (blocks ascend: first index wins ties)
    const long long plane = (long long)gridDim.y * nblk * M;
    const float* pm = p + plane;
    const int* pi = reinterpret_cast<const int*>(part + 2 * plane) + (long long)b * nblk * M + j;
    float best = 0.f;
    int bi = 0x7fffffff;
    for (int k = 0; k < nblk; k++) {
      const float v = pm[(long long)k * M];
      if (v > best) {
        best = v;
        bi = pi[(long long)k * M];
      }
    }
    col_idx[(long long)b * M + j] = bi == 0x7fffffff ? 0 : bi;
  }
}

// row argmax of K_ij B_j (== argmax of the log assignment); matching score = K A B (N + M)
__global__ void __launch_bounds__(256) sks_row_argmax_kernel(const float* __restrict__ K, int N, int M,
                                                             const float* __restrict__ Av,
                                                             const float* __restrict__ Bv, float scale,
                                                             float* __restrict__ max0, int* __restrict__ idx0) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* sr = K + ((long long)b * N + row) * M;
  const float* Bb = Bv + (long long)b * (M + 1);
  float best = -1.f;
  int bi = 0x7fffffff;
  for (int j = lane; j < M; j += 32) {
    const float z = sr[j] * Bb[j];
    if (z > best) {
      best = z;
      bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float b2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (b2 > best || (b2 == best && i2 < bi)) {
      best = b2;
      bi = i2;
    }
  }
  if (lane == 0) {
    max0[(long long)b * N + row] = best * Av[(long long)b * (N + 1) + row] * scale;   // probability, not log
    idx0[(long long)b * N + row] = bi == 0x7fffffff ? 0 : bi;
  }
}

// Same on the untouched score matrix: argmax_j K_ij B_j = argmax_j (S_ij + log B_j); cmax holds c_i on entry and is
// overwritten with the matching score exp(S_ij + log B_j - c_i) A_i (N + M).  logB: [B, M + 1].
__global__ void log_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = logf(in[i]);
}
__global__ void __launch_bounds__(256) sks_row_argmax_s_kernel(const float* __restrict__ S, int N, int M,
                                                               const float* __restrict__ Av,
                                                               const float* __restrict__ logB, float scale,
                                                               float* __restrict__ cmax_max0, int* __restrict__ idx0) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* sr = S + ((long long)b * N + row) * M;
  const float* lb = logB + (long long)b * (M + 1);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane * 4; j < M; j += 128) {   // M % 4 == 0
    const float4 sv = *reinterpret_cast<const float4*>(sr + j);
    const float z0 = sv.x + lb[j], z1 = sv.y + lb[j + 1], z2 = sv.z + lb[j + 2], z3 = sv.w + lb[j + 3];
    if (z0 > best) { best = z0; bi = j; }
    if (z1 > best) { best = z1; bi = j + 1; }
    if (z2 > best) { best = z2; bi = j + 2; }
    if (z3 > best) { best = z3; bi = j + 3; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float b2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (b2 > best || (b2 == best && i2 < bi)) {
      best = b2;
      bi = i2;
    }
  }
  if (lane == 0) {
    const long long r = (long long)b * N + row;
    cmax_max0[r] = __expf(best - cmax_max0[r]) * Av[(long long)b * (N + 1) + row] * scale;
    idx0[r] = bi == 0x7fffffff ? 0 : bi;
  }
}

// row argmax of ((S_ij + u_i) + v_j) - norm over j < M, first index on ties; one warp per row
__global__ void __launch_bounds__(256) sk_row_argmax_kernel(const float* __restrict__ S, int N, int M,
                                                            const float* __restrict__ u, const float* __restrict__ v,
                                                            float norm, float* __restrict__ max0,
                                                            int* __restrict__ idx0) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* sr = S + ((long long)b * N + row) * M;
  const float* vb = v + (long long)b * (M + 1);
  const float ui = u[(long long)b * (N + 1) + row];
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < M; j += 32) {
    float z = __fsub_rn(__fadd_rn(__fadd_rn(sr[j], ui), vb[j]), norm);
    if (z > best) {
      best = z;
      bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float b2 = __shfl_xor_sync(0xffffffffu, best, o);
    int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (b2 > best || (b2 == best && i2 < bi)) {
      best = b2;
      bi = i2;
    }
  }
  if (lane == 0) {
    max0[(long long)b * N + row] = best;
    idx0[(long long)b * N + row] = bi == 0x7fffffff ? 0 : bi;
  }
}

__global__ void mutual_kernel(const float* __restrict__ max0, const int* __restrict__ idx0,
                              const int* __restrict__ idx1, int N, int M, float thr, int is_prob,
                              long long* __restrict__ m0, long long* __restrict__ m1, float* __restrict__ ms0,
                              float* __restrict__ ms1) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const float* mx = max0 + (long long)b * N;
  const int* i0 = idx0 + (long long)b * N;
  const int* i1 = idx1 + (long long)b * M;
  if (t < N) {
    int j = i0[t];
    bool mutual = (i1[j] == t);
    float sc = mutual ? (is_prob ? mx[t] : expf(mx[t])) : 0.f;
    bool valid = mutual && sc > thr;
    m0[(long long)b * N + t] = valid ? (long long)j : -1ll;
    ms0[(long long)b * N + t] = sc;
  }
  if (t < M) {
    int i = i1[t];
    bool mutual1 = (i0[i] == t);
    bool mutual0 = (i1[i0[i]] == i);
    float sc0 = mutual0 ? (is_prob ? mx[i] : expf(mx[i])) : 0.f;
    bool valid0 = mutual0 && sc0 > thr;
    m1[(long long)b * M + t] = (mutual1 && valid0) ? (long long)i : -1ll;
    ms1[(long long)b * M + t] = mutual1 ? sc0 : 0.f;
  }
}

__global__ void write_Z_kernel(const float* __restrict__ S, int N, int M, float alpha, const float* __restrict__ u,
                               const float* __restrict__ v, float norm, float* __restrict__ Z) {
  const int b = blockIdx.z;
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > M) return;
  float c = (i < N && j < M) ? S[((long long)b * N + i) * M + j] : alpha;
  float z = __fsub_rn(__fadd_rn(__fadd_rn(c, u[(long long)b * (N + 1) + i]), v[(long long)b * (M + 1) + j]), norm);
  Z[((long long)b * (N + 1) + i) * (M + 1) + j] = z;
}

__global__ void fill_kernel(float* p, long long n, float v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// one block per pair: ordered compaction of valid rows
__global__ void __launch_bounds__(256) compact_kernel(const long long* __restrict__ m0, int N,
                                                      unsigned* __restrict__ tables, int* __restrict__ lengths) {
  const int b = blockIdx.x;
  const long long* m = m0 + (long long)b * N;
  unsigned* out = tables + (long long)b * N * 2;
  __shared__ int wsums[8];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 256) {
    int i = base + threadIdx.x;
    long long mv = i < N ? m[i] : -1ll;
    bool keep = mv > -1;
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wsums[wid] = __popc(bal);
    __syncthreads();
    int off = running;
    for (int k = 0; k < wid; k++) off += wsums[k];
    off += __popc(bal & ((1u << lane) - 1));
    if (keep) {
      out[2 * off] = (unsigned)i;
      out[2 * off + 1] = (unsigned)mv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < 8; k++) t += wsums[k];
      running += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) lengths[b] = running;
}

// X[b * K + k, :] = store[idx[b] * store_k + k, :]   (256 floats per row; one warp per row, float4 x 2 per lane)
__global__ void __launch_bounds__(256) gather_tokens_kernel(const float4* __restrict__ store, int store_k,
                                                            const int* __restrict__ idx, int K,
                                                            float4* __restrict__ X, long long rows) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long b = row / K;
  const int k = (int)(row - b * K);
  const long long img = idx ? idx[b] : b;
  const float4* src = store + (img * store_k + k) * 64;
  float4* dst = X + row * 64;
  dst[lane] = src[lane];
  dst[lane + 32] = src[lane + 32];
}

// rows [0, lengths[p]) of every pair's (kcap, 2) table -> one flat (sum lengths, 2) array at offsets[p]
__global__ void __launch_bounds__(256) pack_tables_kernel(const uint2* __restrict__ tables, int kcap,
                                                          const int* __restrict__ lengths,
                                                          const long long* __restrict__ offsets,
                                                          uint2* __restrict__ out) {
  const int p = blockIdx.x;
  const int n = lengths[p];
  const uint2* src = tables + (long long)p * kcap;
  uint2* dst = out + offsets[p];
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// Exhaustive pair schedule on the device (matcher.py:143-144: `itertools.combinations(ids, 2)`): entry q of the
// lexicographic list of (i < j) over n images, for q in [first, first + count) -- a rank of the multi-GPU partition
// writes exactly its own range.  Row i starts at offset i (2n - i - 1) / 2; i comes from the closed form in double
// precision, corrected by at most one step.
__global__ void __launch_bounds__(256) pair_list_kernel(int n, long long first, long long count, int* __restrict__ idx0,
                                                        int* __restrict__ idx1) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const long long q = first + k;
  const double b = 2.0 * n - 1.0;
  long long i = (long long)floor((b - sqrt(b * b - 8.0 * (double)q)) * 0.5);
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  auto row_start = [n](long long r) { return r * (2ll * n - r - 1) / 2; };
  while (i > 0 && row_start(i) > q) i--;
  while (i < n - 2 && row_start(i + 1) <= q) i++;
  idx0[k] = (int)i;
  idx1[k] = (int)(q - row_start(i) + i + 1);
}

// SimpleMatcher: S <- -sqrt(2 - 2 clamp(S, -1, 1))   (simple_matcher.py:89-90; negated so argmax == argmin)
__global__ void neg_l2_kernel(float* __restrict__ S, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float d = fminf(fmaxf(S[i], -1.f), 1.f);
    S[i] = -sqrtf(__fsub_rn(2.f, __fmul_rn(2.f, d)));
  }
}

// keep i when dist_i < thr and the nearest neighbour is mutual; ordered compaction into (3, L)
__global__ void __launch_bounds__(256) simple_mutual_kernel(const float* __restrict__ max0,
                                                            const int* __restrict__ idx0,
                                                            const int* __restrict__ idx1, int N, float thr,
                                                            float* __restrict__ out, int ld, int* __restrict__ count) {
  __shared__ int wsums[8];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 256) {
    int i = base + threadIdx.x;
    bool keep = false;
    int j = 0;
    float d = 0.f;
    if (i < N) {
      j = idx0[i];
      d = -max0[i];
      keep = (d < thr) && (idx1[j] == i);
    }
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wsums[wid] = __popc(bal);
    __syncthreads();
    int off = running;
    for (int k = 0; k < wid; k++) off += wsums[k];
    off += __popc(bal & ((1u << lane) - 1));
    if (keep && off < ld) {
      out[off] = (float)i;
      out[ld + off] = (float)j;
      out[2 * ld + off] = d;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < 8; k++) t += wsums[k];
      running += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = running;
}

}  // namespace

int transpose_to_tokens(const float* desc, int ld, int N, float* X, cudaStream_t st) {
  return sg_frontend(desc, 0, ld, nullptr, nullptr, 0, nullptr, 1, N, 0, 0, X, nullptr, st);
}

int neg_l2_from_dot(float* S, long long n, cudaStream_t st) {
  if (n == 0) return SFM_OK;
  neg_l2_kernel<<<cdiv(n, 256), 256, 0, st>>>(S, n);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int simple_mutual(const float* max0, const int* idx0, const int* idx1, int N, float thr, float* out, int ld,
                  int* count, cudaStream_t st) {
  simple_mutual_kernel<<<1, 256, 0, st>>>(max0, idx0, idx1, N, thr, out, ld, count);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int sg_frontend(const float* desc, long long desc_img_stride, int ldk, const float* kpts, const float* scores,
                int kcap, const int* idx, int B, int K, int imgH, int imgW, float* X, float* kin, cudaStream_t st) {
  if (B == 0 || K == 0) return SFM_OK;
  dim3 grid(cdiv(K, 32), B);
  frontend_kernel<<<grid, 256, 0, st>>>(desc, desc_img_stride, ldk, kpts, scores, kcap, idx, K, imgH, imgW, X, kin);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

namespace {
// [batch][R][C] (row stride ldi, batch stride = bsi0 * (z / nb1) + bsi1 * (z % nb1)) -> [batch][C][R] likewise
__global__ void transpose_batched_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C,
                                         long long ldi, long long ldo, int nb1, long long bsi0, long long bsi1,
                                         long long bso0, long long bso1) {
  __shared__ float tile[32][33];
  const int z = blockIdx.z;
  const float* src = in + bsi0 * (z / nb1) + bsi1 * (z % nb1);
  float* dst = out + bso0 * (z / nb1) + bso1 * (z % nb1);
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[r * ldi + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[c * ldo + r] = tile[threadIdx.x][i];
  }
}
__global__ void copy_rows_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int M,
                                 int ldi, int ldo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * M) return;
  long long r = i / M;
  int c = (int)(i - r * M);
  out[r * ldo + c] = in[r * ldi + c];
}
}  // namespace

int transpose_batched_f32(const float* in, float* out, int batch, int nb1, int R, int C, long long ldi,
                          long long ldo, long long bsi0, long long bsi1, long long bso0, long long bso1,
                          cudaStream_t st) {
  if (batch == 0 || R == 0 || C == 0) return SFM_OK;
  SFM_REQUIRE(batch <= 65535 && cdiv(R, 32) <= 65535, "transpose: batch/rows too large");
  dim3 grid(cdiv(C, 32), cdiv(R, 32), batch), block(32, 8);
  transpose_batched_kernel<<<grid, block, 0, st>>>(in, out, R, C, ldi, ldo, nb1, bsi0, bsi1, bso0, bso1);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int copy_rows_f32(const float* in, float* out, long long rows, int M, int ldi, int ldo, cudaStream_t st) {
  if (rows == 0 || M == 0) return SFM_OK;
  copy_rows_kernel<<<(unsigned)cdiv(rows * M, 256), 256, 0, st>>>(in, out, rows, M, ldi, ldo);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int softmax_rows_f32(float* S, long long rows, int M, int ld, cudaStream_t st) {
  if (rows == 0 || M == 0) return SFM_OK;
  softmax_rows_kernel<<<cdiv(rows * 32, 256), 256, 0, st>>>(S, rows, M, ld);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

// Pairs swept together per launch.  Measured on B200 (profiles/README.md): running all iterations on
// L2-sized groups of 1/2/4 pairs is SLOWER than sweeping the whole batch (launch-latency bound,
// 42 % / 29 % / 23 % vs 20 % of the step), so the default is the whole batch; SFM_SK_SUB overrides.
int sinkhorn_sub_batch(int B, int N, int M) {
  (void)N; (void)M;
  int sub = B;
  if (const char* e = getenv("SFM_SK_SUB")) sub = atoi(e);
  if (sub < 1) sub = 1;
  if (sub > B) sub = B;
  return sub;
}

static int pick_R(int B, int N, int M) {
  const int sub = sinkhorn_sub_batch(B, N, M);
  int colblocks = cdiv(M, 128);
  long long base = (long long)sub * colblocks;
  int R = (int)((148 * 16 + base - 1) / base);   // >= 16 CTAs per SM in flight + queued: balances the tail
  int maxR = (N + 63) / 64;
  if (R > maxR) R = maxR;
  if (R < 1) R = 1;
  return R;
}

size_t sinkhorn_partial_floats(int B, int N, int M, int* R_out) {
  int R = pick_R(B, N, M);
  if (R_out) *R_out = R;
  return (size_t)B * R * (M + 1) * 2;
}
size_t sinkhorn_fused_floats(int B, int N, int M) { return (size_t)3 * B * cdiv(N, FUSED_RB) * M; }   // sums | maxima | rows
bool sinkhorn_use_scaling(const SinkhornArgs& a) {
  return a.fast && a.fused_part != nullptr && a.kb != nullptr && a.M % 4 == 0 && a.M <= 4096 && a.iters > 0;
}
size_t sinkhorn_counter_ints(int B, int M) { return (size_t)B * cdiv(M, 128); }

int sinkhorn_run(const SinkhornArgs& a, cudaStream_t st) {
  const int B = a.B, N = a.N, M = a.M;
  if (B == 0) return SFM_OK;
  SFM_REQUIRE(N > 0 && M > 0, "sinkhorn: empty side");
  // constants as the reference computes them in float32 (superglue.py:165-167)
  const float norm = -logf((float)N + (float)M);
  const float log_mu_bin = logf((float)M) + norm;
  const float log_nu_bin = logf((float)N) + norm;
  long long nv = (long long)B * (M + 1), nu = (long long)B * (N + 1);
  fill_kernel<<<cdiv(nv, 256), 256, 0, st>>>(a.v, nv, 0.f);
  fill_kernel<<<cdiv(nu, 256), 256, 0, st>>>(a.u, nu, 0.f);
  SFM_LAUNCH_CHECK_N(2);
  if (sinkhorn_use_scaling(a)) {
    const int nblk = cdiv(N, FUSED_RB);
    const float ealpha = expf(a.alpha);
    const float mu = 1.0f / ((float)N + (float)M), mu_bin = (float)M * mu, nu_bin = (float)N * mu;
    float* K = const_cast<float*>(a.S);   // the score matrix is overwritten by K = exp(S - rowmax)
    sks_prologue_kernel<<<dim3(cdiv(N, 8), B), 256, 0, st>>>(K, N, M, a.alpha, a.kb, a.k16,
                                                             a.k16 != nullptr ? a.rowmax : nullptr);
    fill_kernel<<<cdiv(nv, 256), 256, 0, st>>>(a.v, nv, 1.f);   // B = exp(v) = 1
    dim3 fgrid(nblk, B);
    dim3 cg(cdiv(M, 256) + 1, B);
    const bool h16 = a.k16 != nullptr;
    const void* Kit = h16 ? (const void*)a.k16 : (const void*)K;
    for (int it = 0; it < a.iters; it++) {
      const bool last = a.col_idx != nullptr && it + 1 == a.iters;
#define SKS_LAUNCH(G, H, L) sks_fused_kernel<G, H, L><<<fgrid, 256, 0, st>>>(Kit, N, M, ealpha, a.kb, a.v, a.u, a.fused_part, mu, mu_bin)
#define SKS_PICK(G) { if (h16) { if (last) SKS_LAUNCH(G, true, true); else SKS_LAUNCH(G, true, false); } \
                      else { if (last) SKS_LAUNCH(G, false, true); else SKS_LAUNCH(G, false, false); } }
      if (M <= 1024) SKS_PICK(1) else if (M <= 2048) SKS_PICK(2) else SKS_PICK(4)
#undef SKS_PICK
#undef SKS_LAUNCH
      sks_combine_kernel<<<cg, 256, 0, st>>>(a.fused_part, nblk, N, M, ealpha, a.kb, a.u, a.v, mu, nu_bin,
                                             last ? a.col_idx : nullptr);
    }
    count_launches(2 + 2 * a.iters);
    SFM_CUDA(cudaGetLastError());
    return SFM_OK;
  }
  const int sub = sinkhorn_sub_batch(B, N, M);
  for (int b0 = 0; b0 < B; b0 += sub) {
    const int nb = (B - b0) < sub ? (B - b0) : sub;
    const float* S = a.S + (long long)b0 * N * M;
    float* u = a.u + (long long)b0 * (N + 1);
    float* v = a.v + (long long)b0 * (M + 1);
    float* part = a.part + (long long)b0 * a.R * (M + 1) * 2;
    int* counters = a.counters + b0 * cdiv(M, 128);
    dim3 rgrid(cdiv(N + 1, 8), nb);
    dim3 cgrid(cdiv(M, 128), a.R, nb);
    for (int it = 0; it < a.iters; it++) {
      if (a.fast) {
        sk_row_kernel<true><<<rgrid, 256, 0, st>>>(S, N, M, a.alpha, v, u, norm, log_mu_bin);
        sk_col_kernel<0, true><<<cgrid, 256, 0, st>>>(S, N, M, a.alpha, u, nullptr, v, part, counters, a.R, norm,
                                                      log_nu_bin, nullptr, nullptr);
      } else {
        sk_row_kernel<false><<<rgrid, 256, 0, st>>>(S, N, M, a.alpha, v, u, norm, log_mu_bin);
        sk_col_kernel<0, false><<<cgrid, 256, 0, st>>>(S, N, M, a.alpha, u, nullptr, v, part, counters, a.R, norm,
                                                       log_nu_bin, nullptr, nullptr);
      }
    }
    count_launches(2 * a.iters);
  }
  SFM_CUDA(cudaGetLastError());
  return SFM_OK;
}

int argmax_rows_cols(const SinkhornArgs& a, float norm, float* max0, int* idx0, float* max1, int* idx1,
                     cudaStream_t st) {
  const int B = a.B, N = a.N, M = a.M;
  if (B == 0) return SFM_OK;
  dim3 rgrid(cdiv(N, 8), B);
  sk_row_argmax_kernel<<<rgrid, 256, 0, st>>>(a.S, N, M, a.u, a.v, norm, max0, idx0);
  dim3 cgrid(cdiv(M, 128), a.R, B);
  sk_col_kernel<1, false><<<cgrid, 256, 0, st>>>(a.S, N, M, a.alpha, a.u, a.v, nullptr, a.part, a.counters, a.R, norm, 0.f,
                                          max1, idx1);
  SFM_LAUNCH_CHECK_N(2);
  return SFM_OK;
}

int sinkhorn_match(const SinkhornArgs& a, float thr, float* max0, int* idx0, float* max1, int* idx1,
                   long long* matches0, long long* matches1, float* mscores0, float* mscores1, cudaStream_t st) {
  const int B = a.B, N = a.N, M = a.M;
  if (B == 0) return SFM_OK;
  const float norm = -logf((float)N + (float)M);
  dim3 mgrid(cdiv(N > M ? N : M, 256), B);
  if (sinkhorn_use_scaling(a)) {
    // a.S now holds K, a.u = A, a.v = B: assignment probabilities are K A B (N + M)
    if (a.k16 != nullptr && a.rowmax != nullptr && a.rowmax == max0) {
      // the score matrix was left untouched (the iterations ran on the fp16 copy): S + log B, c_i from max0
      const long long nv = (long long)B * (M + 1);
      log_kernel<<<cdiv(nv, 256), 256, 0, st>>>(a.v, a.part, nv);
      sks_row_argmax_s_kernel<<<dim3(cdiv(N, 8), B), 256, 0, st>>>(a.S, N, M, a.u, a.part, (float)N + (float)M, max0,
                                                                   idx0);
      count_launches(1);
    } else {
      sks_row_argmax_kernel<<<dim3(cdiv(N, 8), B), 256, 0, st>>>(a.S, N, M, a.u, a.v, (float)N + (float)M, max0, idx0);
    }
    if (a.col_idx != idx1 || a.iters == 0)   // column arg-max not produced by the last sweep
      sk_col_kernel<2, false><<<dim3(cdiv(M, 128), a.R, B), 256, 0, st>>>(a.S, N, M, a.alpha, a.u, a.v, nullptr, a.part,
                                                                          a.counters, a.R, 0.f, 0.f, max1, idx1);
    mutual_kernel<<<mgrid, 256, 0, st>>>(max0, idx0, idx1, N, M, thr, 1, matches0, matches1, mscores0, mscores1);
    SFM_LAUNCH_CHECK_N(3);
    return SFM_OK;
  }
  SFM_TRY(argmax_rows_cols(a, norm, max0, idx0, max1, idx1, st));
  mutual_kernel<<<mgrid, 256, 0, st>>>(max0, idx0, idx1, N, M, thr, 0, matches0, matches1, mscores0, mscores1);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int sinkhorn_write_Z(const SinkhornArgs& a, float* Z, cudaStream_t st) {
  if (a.B == 0) return SFM_OK;
  const float norm = -logf((float)a.N + (float)a.M);
  dim3 grid(cdiv(a.M + 1, 256), a.N + 1, a.B);
  write_Z_kernel<<<grid, 256, 0, st>>>(a.S, a.N, a.M, a.alpha, a.u, a.v, norm, Z);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int gather_tokens(const float* store, int store_k, const int* idx, int B, int K, float* X, cudaStream_t st) {
  const long long rows = (long long)B * K;
  if (rows == 0) return SFM_OK;
  gather_tokens_kernel<<<cdiv(rows, 8), 256, 0, st>>>(reinterpret_cast<const float4*>(store), store_k, idx, K,
                                                      reinterpret_cast<float4*>(X), rows);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int pack_tables(const unsigned* tables, int n_pairs, int kcap, const int* lengths, const long long* offsets,
                unsigned* out, cudaStream_t st) {
  if (n_pairs == 0) return SFM_OK;
  pack_tables_kernel<<<n_pairs, 256, 0, st>>>(reinterpret_cast<const uint2*>(tables), kcap, lengths, offsets,
                                              reinterpret_cast<uint2*>(out));
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int pair_list_range(int n, long long first, long long count, int* idx0, int* idx1, cudaStream_t st) {
  if (count == 0) return SFM_OK;
  pair_list_kernel<<<cdiv(count, 256), 256, 0, st>>>(n, first, count, idx0, idx1);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int compact_matches(const long long* matches0, int B, int N, unsigned* tables, int* lengths, cudaStream_t st) {
  if (B == 0) return SFM_OK;
  compact_kernel<<<B, 256, 0, st>>>(matches0, N, tables, lengths);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

}  // namespace sfm
