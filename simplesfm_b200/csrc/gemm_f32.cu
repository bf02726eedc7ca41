#include "gemm_f32.cuh"
#include <algorithm>
#include <cstdint>

namespace sfm {

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NTHREADS = 256;
constexpr int APAD = 4;

// zero the lanes of a float4 read at column k that lie at or beyond K (K not a multiple of 4)
__device__ __forceinline__ float4 mask_k(float4 v, int k, int K) {
  if (k + 3 >= K) {
    if (k + 1 >= K) v.y = 0.f;
    if (k + 2 >= K) v.z = 0.f;
    if (k + 3 >= K) v.w = 0.f;
    if (k >= K) v.x = 0.f;
  }
  return v;
}

// MODE: 0 = NT (B [N,K]), 1 = NN (B [K,N]), 2 = implicit 3x3 conv (B [N,K])
template <int BN, int MODE>
__global__ void __launch_bounds__(NTHREADS) gemm_f32_kernel(GemmF32 p) {
  constexpr int TN = BN / 16;  // columns per thread (8 or 4)
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + APAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int z = blockIdx.z;
  const int z0 = z / p.nb1, z1 = z % p.nb1;
  const float* __restrict__ A = p.A + z0 * p.sA0 + z1 * p.sA1;
  const float* __restrict__ A2 = p.A2 ? p.A2 + z0 * p.sA0 + z1 * p.sA1 : nullptr;
  const float* __restrict__ B = p.B + z0 * p.sB0 + z1 * p.sB1;
  float* __restrict__ C = p.C + z0 * p.sC0 + z1 * p.sC1;

  // ---- A loader geometry: rows ar, ar+64; k offset akq
  const int ar = tid >> 2, akq = (tid & 3) * 4;
  int cy[2], cx[2];
  long long cbase[2];
  bool rvalid[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    int m = m0 + ar + i * 64;
    rvalid[i] = m < p.M;
    if (MODE == 2) {
      int hw = p.convH * p.convW;
      int img = m / hw, rem = m - img * hw;
      cy[i] = rem / p.convW;
      cx[i] = rem - cy[i] * p.convW;
      cbase[i] = (long long)img * hw;
    } else {
      cy[i] = cx[i] = 0;
      cbase[i] = 0;
    }
  }

  float4 ra[2], rb[2];
  // implicit 3x3 convolution: the filter tap changes every Cin / BK k-tiles, so the tap's source pixel (pointer and
  // validity per row) is kept across k-tiles and only the channel offset advances -- the kernel is issue-bound and
  // the per-tile address arithmetic (a division by Cin, bounds tests, 64-bit multiplies) cost a quarter of its slots
  const float* tap_ptr[2] = {nullptr, nullptr};
  int tap_next = 0, ci_next = 0;   // position of the NEXT load_tiles call (k0 advances by BK each call)
  auto load_tiles = [&](int k0) {
    // A
    if (MODE == 2) {
      (void)k0;
      if (ci_next == 0) {   // new tap (block-uniform)
        const int ky = tap_next / 3 - 1, kx = tap_next - (tap_next / 3) * 3 - 1;
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int yy = cy[i] + ky, xx = cx[i] + kx;
          const bool ok = rvalid[i] && yy >= 0 && yy < p.convH && xx >= 0 && xx < p.convW;
          tap_ptr[i] = ok ? A + (cbase[i] + (long long)yy * p.convW + xx) * p.convCin + akq : nullptr;
        }
      }
#pragma unroll
      for (int i = 0; i < 2; i++)
        ra[i] = tap_ptr[i] != nullptr ? *reinterpret_cast<const float4*>(tap_ptr[i] + ci_next)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      ci_next += BK;
      if (ci_next >= p.convCin) {
        ci_next = 0;
        tap_next++;
      }
    } else {
      int k = k0 + akq;
      const float* src = A;
      int ld = p.lda, kk = k;
      if (A2 != nullptr && k >= p.K1) {
        src = A2;
        ld = p.lda2;
        kk = k - p.K1;
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int m = m0 + ar + i * 64;
        ra[i] = (rvalid[i] && k < p.K) ? mask_k(*reinterpret_cast<const float4*>(src + (long long)m * ld + kk), k, p.K)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // B
    if (MODE == 1) {
      constexpr int F4_PER_ROW = BN / 4;
#pragma unroll
      for (int i = 0; i < (BK * F4_PER_ROW) / NTHREADS; i++) {
        int f = tid + i * NTHREADS;
        int kr = f / F4_PER_ROW, nq = (f % F4_PER_ROW) * 4;
        int k = k0 + kr, n = n0 + nq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < p.K) {
          const float* s = B + (long long)k * p.ldb + n;
          if (n + 3 < p.N) {
            v = *reinterpret_cast<const float4*>(s);
          } else {
            if (n < p.N) v.x = s[0];
            if (n + 1 < p.N) v.y = s[1];
            if (n + 2 < p.N) v.z = s[2];
          }
        }
        rb[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < BN / 64; i++) {
        int n = n0 + ar + i * 64;
        int k = k0 + akq;
        rb[i] = (n < p.N && k < p.K) ? mask_k(*reinterpret_cast<const float4*>(B + (long long)n * p.ldb + k), k, p.K)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      int r = ar + i * 64;
      As[buf][akq + 0][r] = ra[i].x;
      As[buf][akq + 1][r] = ra[i].y;
      As[buf][akq + 2][r] = ra[i].z;
      As[buf][akq + 3][r] = ra[i].w;
    }
    if (MODE == 1) {
      constexpr int F4_PER_ROW = BN / 4;
#pragma unroll
      for (int i = 0; i < (BK * F4_PER_ROW) / NTHREADS; i++) {
        int f = tid + i * NTHREADS;
        int kr = f / F4_PER_ROW, nq = (f % F4_PER_ROW) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][kr][nq]) = rb[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < BN / 64; i++) {
        int r = ar + i * 64;
        Bs[buf][akq + 0][r] = rb[i].x;
        Bs[buf][akq + 1][r] = rb[i].y;
        Bs[buf][akq + 2][r] = rb[i].z;
        Bs[buf][akq + 3][r] = rb[i].w;
      }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  const int nk = (p.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      if (TN == 8) {
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < TN / 4; jh++) {
      int n = n0 + jh * 64 + tx * 4;
      if (n >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float x = p.alpha * acc[i][jh * 4 + j];
        if (p.bias != nullptr && n + j < p.N) x += p.bias[n + j];
        if (p.flags & GEMM_RELU) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      float* dst = C + (long long)m * p.ldc + n;
      if (vec_ok && n + 3 < p.N) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (p.flags & GEMM_ACCUM) {
          float4 c = *reinterpret_cast<const float4*>(dst);
          o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
        }
        *reinterpret_cast<float4*>(dst) = o;
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (n + j < p.N) dst[j] = (p.flags & GEMM_ACCUM) ? dst[j] + v[j] : v[j];
      }
    }
  }
}

template <int MODE>
int launch_gemm(const GemmF32& p, cudaStream_t st) {
  SFM_REQUIRE(p.M >= 0 && p.N > 0 && p.K > 0, "gemm_f32: bad dims M=%d N=%d K=%d", p.M, p.N, p.K);
  if (p.M == 0 || p.batch == 0) return SFM_OK;
  if (MODE != 2) SFM_REQUIRE(p.lda % 4 == 0, "gemm_f32: lda %% 4 != 0");
  if (MODE == 0 || MODE == 2) SFM_REQUIRE(p.ldb % 4 == 0, "gemm_f32: ldb %% 4 != 0");
  if (MODE == 1) SFM_REQUIRE(p.ldb % 4 == 0 && p.sB0 % 4 == 0 && p.sB1 % 4 == 0, "gemm_f32(nn): ldb/strides %% 4 != 0");
  if (p.A2) SFM_REQUIRE(p.K1 % BK == 0 && p.lda2 % 4 == 0, "gemm_f32: split point must be a multiple of %d", BK);
  if (MODE == 2) SFM_REQUIRE(p.convCin % BK == 0 && p.K == 9 * p.convCin, "conv3x3_f32: Cin must be a multiple of 16");
  // 64-wide column tiles also when the 128-wide grid would leave most of the machine idle (SuperPoint's 60 x 80
  // layers: 150 CTAs of 8 warps on 148 SMs); the accumulation order per output element does not change
  const long long ctas128 = (long long)cdiv(p.N, 128) * cdiv(p.M, BM) * p.batch;
  if (p.N <= 64 || ctas128 < 2 * 148) {
    dim3 grid(cdiv(p.N, 64), cdiv(p.M, BM), p.batch);
    gemm_f32_kernel<64, MODE><<<grid, NTHREADS, 0, st>>>(p);
  } else {
    dim3 grid(cdiv(p.N, 128), cdiv(p.M, BM), p.batch);
    gemm_f32_kernel<128, MODE><<<grid, NTHREADS, 0, st>>>(p);
  }
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

// conv1a: one input channel, 64 output channels.  thread = (8-channel group, pixel): the group's 72 weights live
// in registers for the whole thread (no shared-memory operand per FMA), a block walks 8 x 32 consecutive pixels
// of the flattened image batch, and every warp store covers 4 adjacent pixels x 64 channels (NHWC, fully coalesced).
// Same accumulation order as before (taps in (ky, kx) order), so fp32 results are unchanged.
constexpr int C1A_PIX_PER_BLOCK = 256;
__global__ void __launch_bounds__(256) conv1a_kernel(const float* __restrict__ img, const float* __restrict__ w9,
                                                     const float* __restrict__ bias, float* __restrict__ out,
                                                     __nv_bfloat16* __restrict__ out_h, int n_img, int H, int W) {
  const int cg = threadIdx.x & 7, lp = threadIdx.x >> 3;   // channel group, pixel lane (0..31)
  float w[8][9], b[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    b[c] = bias[cg * 8 + c];
#pragma unroll
    for (int t = 0; t < 9; t++) w[c][t] = w9[(cg * 8 + c) * 9 + t];
  }
  const int hw = H * W;
  const long long total = (long long)n_img * hw;
  const long long p0 = (long long)blockIdx.x * C1A_PIX_PER_BLOCK;
#pragma unroll 2
  for (int it = 0; it < C1A_PIX_PER_BLOCK / 32; it++) {
    const long long pix = p0 + it * 32 + lp;
    if (pix >= total) break;
    const int n = (int)(pix / hw), rem = (int)(pix - (long long)n * hw);
    const int y = rem / W, x = rem - y * W;
    const float* im = img + (long long)n * hw;
    float v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        v[ky * 3 + kx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? im[yy * W + xx] : 0.f;
      }
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < 9; t++) a = fmaf(v[t], w[c][t], a);
      o[c] = fmaxf(a + b[c], 0.f);
    }
    if (out_h != nullptr) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]),
                     h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
      *reinterpret_cast<uint4*>(out_h + pix * 64 + cg * 8) =
          make_uint4(*reinterpret_cast<unsigned*>(&h0), *reinterpret_cast<unsigned*>(&h1),
                     *reinterpret_cast<unsigned*>(&h2), *reinterpret_cast<unsigned*>(&h3));
    } else {
      float4* dst = reinterpret_cast<float4*>(out + pix * 64 + cg * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// conv1a, four pixels per thread (W % 4 == 0): thread = (8-channel group, 4 consecutive pixels of one row).  The
// 3 x 6 input window comes from three 16-byte loads plus the two edge columns, the index arithmetic (two divisions,
// 36 bounds tests) is paid once per four pixels, and 288 independent FMAs hide the load latency.  Per output the
// taps are accumulated in the same (ky, kx) order as conv1a_kernel: fp32 results are bit-identical.
__global__ void __launch_bounds__(128) conv1a_x4_kernel(const float* __restrict__ img, const float* __restrict__ w9,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        __nv_bfloat16* __restrict__ out_h, int n_img, int H, int W) {
  const int cg = threadIdx.x & 7, lq = threadIdx.x >> 3;   // channel group, quad lane (0..15)
  float w[8][9], b[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    b[c] = bias[cg * 8 + c];
#pragma unroll
    for (int t = 0; t < 9; t++) w[c][t] = w9[(cg * 8 + c) * 9 + t];
  }
  const int wq = W >> 2;                      // quads per row
  const long long quads = (long long)n_img * H * wq;
  for (long long q = (long long)blockIdx.x * 16 + lq; q < quads; q += (long long)gridDim.x * 16) {
    const long long row = q / wq;             // image row index over the whole batch
    const int x0 = (int)(q - row * wq) * 4;
    const int y = (int)(row % H);
    const float* im = img + row * W;          // start of this image row
    float v[3][6];
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      const int yy = y + ky - 1;
      if (yy >= 0 && yy < H) {
        const float* r = im + (ky - 1) * W + x0;
        const float4 m = *reinterpret_cast<const float4*>(r);
        v[ky][0] = x0 > 0 ? r[-1] : 0.f;
        v[ky][1] = m.x; v[ky][2] = m.y; v[ky][3] = m.z; v[ky][4] = m.w;
        v[ky][5] = x0 + 4 < W ? r[4] : 0.f;
      } else {
#pragma unroll
        for (int k = 0; k < 6; k++) v[ky][k] = 0.f;
      }
    }
    const long long pix0 = row * W + x0;
#pragma unroll
    for (int px = 0; px < 4; px++) {
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        float a = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
          for (int kx = 0; kx < 3; kx++) a = fmaf(v[ky][px + kx], w[c][ky * 3 + kx], a);
        o[c] = fmaxf(a + b[c], 0.f);
      }
      if (out_h != nullptr) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]),
                       h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
        *reinterpret_cast<uint4*>(out_h + (pix0 + px) * 64 + cg * 8) =
            make_uint4(*reinterpret_cast<unsigned*>(&h0), *reinterpret_cast<unsigned*>(&h1),
                       *reinterpret_cast<unsigned*>(&h2), *reinterpret_cast<unsigned*>(&h3));
      } else {
        float4* dst = reinterpret_cast<float4*>(out + (pix0 + px) * 64 + cg * 8);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}

__global__ void maxpool2_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n_img, int H, int W,
                                int C4) {
  int Ho = H / 2, Wo = W / 2;
  long long total = (long long)n_img * Ho * Wo * C4;
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  int c = gid % C4;
  long long t = gid / C4;
  int xo = t % Wo;
  t /= Wo;
  int yo = t % Ho;
  int n = t / Ho;
  const float4* src = in + (((long long)n * H + 2 * yo) * W + 2 * xo) * C4 + c;
  float4 a = src[0], b = src[C4], d = src[(long long)W * C4], e = src[(long long)W * C4 + C4];
  float4 r;
  r.x = fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x));
  r.y = fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y));
  r.z = fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z));
  r.w = fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w));
  out[gid] = r;
}

}  // namespace

int gemm_f32_nt(const GemmF32& p, cudaStream_t st) { return launch_gemm<0>(p, st); }
int gemm_f32_nn(const GemmF32& p, cudaStream_t st) { return launch_gemm<1>(p, st); }
int conv3x3_f32(const GemmF32& p, cudaStream_t st) { return launch_gemm<2>(p, st); }

int conv1a_f32(const float* img, const float* w9, const float* bias, float* out, __nv_bfloat16* out_h, int n_img,
               int H, int W, cudaStream_t st) {
  long long total = (long long)n_img * H * W;
  if (total == 0) return SFM_OK;
  if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0) {
    const long long quads = total / 4;
    const int grid = (int)std::min<long long>(cdiv(quads, 16), 148 * 16);
    conv1a_x4_kernel<<<grid, 128, 0, st>>>(img, w9, bias, out, out_h, n_img, H, W);
  } else {
    conv1a_kernel<<<cdiv(total, C1A_PIX_PER_BLOCK), 256, 0, st>>>(img, w9, bias, out, out_h, n_img, H, W);
  }
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int maxpool2_nhwc_f32(const float* in, float* out, int n_img, int H, int W, int C, cudaStream_t st) {
  SFM_REQUIRE(C % 4 == 0, "maxpool2: C%%4 required (C=%d)", C);   // odd H / W: the last row / column is dropped (floor)
  long long total = (long long)n_img * (H / 2) * (W / 2) * (C / 4);
  if (total == 0) return SFM_OK;
  maxpool2_kernel<<<cdiv(total, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(in),
                                                    reinterpret_cast<float4*>(out), n_img, H, W, C / 4);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

}  // namespace sfm
