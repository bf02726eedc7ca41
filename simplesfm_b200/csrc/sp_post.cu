// SuperPoint post-processing chain (HBM/latency-bound, exact integer outputs):
//   softmax(65) + 8x8 pixel shuffle  (reference superpoint.py:127-130)
//   simple_nms                        (superpoint.py:9-24)
//   mask / threshold / border / ordered compaction (superpoint.py:133-145)
//   top-k + descending sort with a fixed tie rule  (superpoint.py:35-39,148-151,165-171)
//   bilinear descriptor sampling + double L2 normalisation (superpoint.py:42-54,157-163)
#include "sp_post.cuh"
#include <algorithm>

namespace sfm {

namespace {

// ------------------------------------------------------------------------------------------
// softmax over 65 channels + drop dustbin + depth-to-space.  One warp per 8x8 cell.
// logits: NHWC [n, h, w, ldl] (65 used); out: [n, 8h, 8w]
__global__ void softmax_shuffle_kernel(const float* __restrict__ logits, int ldl, float* __restrict__ out, int n_img,
                                       int h, int w) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int cells = n_img * h * w;
  if (warp >= cells) return;
  const float* src = logits + (long long)warp * ldl;
  float v0 = src[lane], v1 = src[lane + 32];
  float v2 = lane == 0 ? src[64] : -INFINITY;
  float m = warp_max(fmaxf(fmaxf(v0, v1), v2));
  float e0 = expf(v0 - m), e1 = expf(v1 - m), e2 = lane == 0 ? expf(v2 - m) : 0.f;
  float s = warp_sum(e0 + e1 + e2);
  int img = warp / (h * w), rem = warp - img * h * w;
  int cy = rem / w, cx = rem - cy * w;
  int W = 8 * w;
  float* dst = out + (long long)img * (8 * h) * W + (long long)(cy * 8) * W + cx * 8;
  // channel c -> (row c / 8, col c % 8) inside the cell
  dst[(lane >> 3) * W + (lane & 7)] = e0 / s;
  dst[((lane >> 3) + 4) * W + (lane & 7)] = e1 / s;
}

// ------------------------------------------------------------------------------------------
// simple_nms: 32x32 output tile, halo 5*r, separable (2r+1)-max in shared memory.
constexpr int NMS_TILE = 32;

// Separable (2r+1)-max restricted to where it is needed: `in` is meaningful on [mg, T-mg)^2, the result on
// [mg+r, T-mg-r)^2 (every pool of the chain shrinks the valid region by r, so later pools touch fewer pixels).
__device__ __forceinline__ void pool_sep(const float* __restrict__ in, float* __restrict__ tmp,
                                         float* __restrict__ out, int T, int r, int mg) {
  const int lo = mg, hi = T - mg;            // input extent
  const int olo = mg + r, ohi = T - mg - r;  // output extent
  const int wout = ohi - olo, hin = hi - lo;
  if (wout <= 0) return;
  for (int i = threadIdx.x; i < hin * wout; i += blockDim.x) {
    const int y = lo + i / wout, x = olo + i % wout;
    const float* p = in + y * T + x - r;
    float m = p[0];
    for (int k = 1; k <= 2 * r; k++) m = fmaxf(m, p[k]);
    tmp[y * T + x] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < wout * wout; i += blockDim.x) {
    const int y = olo + i / wout, x = olo + i % wout;
    const float* p = tmp + (y - r) * T + x;
    float m = p[0];
    for (int k = 1; k <= 2 * r; k++) m = fmaxf(m, p[k * T]);
    out[y * T + x] = m;
  }
  __syncthreads();
}

// radius-4 fast path (the reference default): every thread produces FOUR neighbouring outputs from 16-byte shared
// loads (3 per 4 outputs in the row pass, 9 in the column pass, instead of 9 scalar loads per output) and
// three-input maxima.  max is exact and order-independent, so the result equals the scalar chain bit for bit.
__device__ __forceinline__ float fmax3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void pool_sep_r4(const float* __restrict__ in, float* __restrict__ tmp,
                                            float* __restrict__ out, int T, int mg) {
  const int lo = mg, hi = T - mg;
  const int olo = mg + 4, ohi = T - mg - 4;
  const int wout = ohi - olo, hin = hi - lo;     // T % 4 == 0 and mg % 4 == 0: wout % 4 == 0, 16-byte aligned x
  if (wout <= 0) return;
  const int wq = wout >> 2;
  for (int i = threadIdx.x; i < hin * wq; i += blockDim.x) {
    const int y = lo + i / wq, x = olo + 4 * (i % wq);
    const float4 A = *reinterpret_cast<const float4*>(in + y * T + x - 4);
    const float4 B = *reinterpret_cast<const float4*>(in + y * T + x);
    const float4 C = *reinterpret_cast<const float4*>(in + y * T + x + 4);
    // inputs x-4 .. x+7 = A.xyzw B.xyzw C.xyzw ; output k covers inputs k .. k+8
    const float core = fmaxf(fmax3f(A.w, B.x, B.y), fmax3f(B.z, B.w, C.x));   // inputs 3 .. 8
    float4 o;
    o.x = fmaxf(core, fmax3f(A.x, A.y, A.z));
    o.y = fmax3f(core, fmaxf(A.y, A.z), C.y);
    o.z = fmax3f(core, A.z, fmaxf(C.y, C.z));
    o.w = fmaxf(core, fmax3f(C.y, C.z, C.w));
    *reinterpret_cast<float4*>(tmp + y * T + x) = o;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < wout * wq; i += blockDim.x) {
    const int y = olo + i / wq, x = olo + 4 * (i % wq);
    float4 v[9];
#pragma unroll
    for (int k = 0; k < 9; k++) v[k] = *reinterpret_cast<const float4*>(tmp + (y - 4 + k) * T + x);
    float4 o;
    o.x = fmax3f(fmax3f(v[0].x, v[1].x, v[2].x), fmax3f(v[3].x, v[4].x, v[5].x), fmax3f(v[6].x, v[7].x, v[8].x));
    o.y = fmax3f(fmax3f(v[0].y, v[1].y, v[2].y), fmax3f(v[3].y, v[4].y, v[5].y), fmax3f(v[6].y, v[7].y, v[8].y));
    o.z = fmax3f(fmax3f(v[0].z, v[1].z, v[2].z), fmax3f(v[3].z, v[4].z, v[5].z), fmax3f(v[6].z, v[7].z, v[8].z));
    o.w = fmax3f(fmax3f(v[0].w, v[1].w, v[2].w), fmax3f(v[3].w, v[4].w, v[5].w), fmax3f(v[6].w, v[7].w, v[8].w));
    *reinterpret_cast<float4*>(out + y * T + x) = o;
  }
  __syncthreads();
}

// TILE x TILE outputs per CTA with a halo of 5 r on every side.  RC > 0 fixes the radius at compile time (tile side,
// halo and every index division become constants); the reference default r = 4 runs with 64 x 64 tiles (104 x 104
// with the halo: 2.6 x redundant work instead of 5.1 x for 32 x 32 tiles).
template <int TILE, int RC>
__global__ void __launch_bounds__(RC > 0 ? 1024 : 512) nms_kernel(const float* __restrict__ scores, float* __restrict__ out, int H, int W, int r_arg) {
  const int r = RC > 0 ? RC : r_arg;
  extern __shared__ __align__(16) float sm[];
  const int halo = 5 * r;
  const int T = TILE + 2 * halo;
  float* s = sm;              // scores (-inf outside the image)
  float* a = s + T * T;       // work input
  float* tmp = a + T * T;     // separable temp
  float* pl = tmp + T * T;    // pooled
  unsigned char* mx = reinterpret_cast<unsigned char*>(pl + T * T);

  const int img = blockIdx.z;
  const int y0 = blockIdx.y * TILE - halo, x0 = blockIdx.x * TILE - halo;
  const float* src = scores + (long long)img * H * W;
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    int y = y0 + i / T, x = x0 + i % T;
    bool in = y >= 0 && y < H && x >= 0 && x < W;
    s[i] = in ? src[(long long)y * W + x] : -INFINITY;
  }
  __syncthreads();
  if (r == 4) pool_sep_r4(s, tmp, pl, T, 0); else pool_sep(s, tmp, pl, T, r, 0);   // valid on margin r
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    int ty = i / T, tx = i % T;
    int y = y0 + ty, x = x0 + tx;
    bool in = y >= 0 && y < H && x >= 0 && x < W;
    bool val = ty >= r && ty < T - r && tx >= r && tx < T - r;   // pl is only defined there
    mx[i] = (in && val && s[i] == pl[i]) ? 1 : 0;
  }
  __syncthreads();
  for (int round = 0; round < 2; round++) {
    for (int i = threadIdx.x; i < T * T; i += blockDim.x) a[i] = mx[i] ? 1.f : 0.f;
    __syncthreads();
    const int m1 = r * (1 + 2 * round);
    if (r == 4) pool_sep_r4(a, tmp, pl, T, m1); else pool_sep(a, tmp, pl, T, r, m1);   // pl > 0 <=> suppressed neighbourhood (valid on margin m1 + r)
    for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
      int y = y0 + i / T, x = x0 + i % T;
      bool in = y >= 0 && y < H && x >= 0 && x < W;
      int ty = i / T, tx = i % T;
      bool val = ty >= m1 + r && ty < T - m1 - r && tx >= m1 + r && tx < T - m1 - r;
      bool supp = val && pl[i] > 0.f;
      a[i] = in ? (supp ? 0.f : s[i]) : -INFINITY;
      // stash supp in bit 1 of mx
      mx[i] = (mx[i] & 1) | (supp ? 2 : 0);
    }
    __syncthreads();
    if (r == 4) pool_sep_r4(a, tmp, pl, T, m1 + r); else pool_sep(a, tmp, pl, T, r, m1 + r);   // valid on margin m1 + 2r
    for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
      int ty = i / T, tx = i % T;
      int y = y0 + ty, x = x0 + tx;
      bool in = y >= 0 && y < H && x >= 0 && x < W;
      bool val = ty >= m1 + 2 * r && ty < T - m1 - 2 * r && tx >= m1 + 2 * r && tx < T - m1 - 2 * r;
      bool supp = mx[i] & 2;
      bool nm = in && val && (a[i] == pl[i]) && !supp;
      mx[i] = ((mx[i] & 1) || nm) ? 1 : 0;
    }
    __syncthreads();
  }
  float* dst = out + (long long)img * H * W;
  for (int i = threadIdx.x; i < TILE * TILE; i += blockDim.x) {
    int ty = i / TILE, tx = i - ty * TILE;
    int y = blockIdx.y * TILE + ty, x = blockIdx.x * TILE + tx;
    if (y < H && x < W) {
      int j = (ty + halo) * T + tx + halo;
      dst[(long long)y * W + x] = mx[j] ? s[j] : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------
// ordered compaction of candidates: value = mask ? nms : 0; keep value > thr inside the border.
constexpr int SEL_PER_BLOCK = 1024;

__device__ __forceinline__ bool sel_pred(const float* nms, const unsigned char* mask, int p, int H, int W, float thr,
                                         int border, float* val) {
  float v = nms[p];
  if (mask != nullptr && !mask[p]) v = 0.f;
  int y = p / W, x = p - y * W;
  *val = v;
  return (v > thr) && y >= border && y < H - border && x >= border && x < W - border;
}

__global__ void sel_count_kernel(const float* __restrict__ nms, const unsigned char* __restrict__ mask, int H, int W,
                                 float thr, int border, int* __restrict__ block_counts) {
  int img = blockIdx.y;
  const float* s = nms + (long long)img * H * W;
  const unsigned char* mk = mask ? mask + (long long)img * H * W : nullptr;
  int base = blockIdx.x * SEL_PER_BLOCK;
  int cnt = 0;
  for (int i = threadIdx.x; i < SEL_PER_BLOCK; i += blockDim.x) {
    int p = base + i;
    float v;
    if (p < H * W && sel_pred(s, mk, p, H, W, thr, border, &v)) cnt++;
  }
  __shared__ int tot;
  if (threadIdx.x == 0) tot = 0;
  __syncthreads();
  int wsum = cnt;
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&tot, wsum);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[img * gridDim.x + blockIdx.x] = tot;
}

__global__ void sel_scan_kernel(int* __restrict__ block_counts, int nblocks, int* __restrict__ counts) {
  // one block per image; exclusive scan of block counts in place (nblocks is small)
  int img = blockIdx.x;
  int* bc = block_counts + img * nblocks;
  __shared__ int carry;
  __shared__ int wsums[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < nblocks ? bc[i] : 0;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = lane < (blockDim.x >> 5) ? wsums[lane] : 0;
      int winc = w;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      wsums[lane] = winc - w;
    }
    __syncthreads();
    int excl = carry + wsums[wid] + inc - v;
    if (i < nblocks) bc[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[img] = carry;
}

__global__ void sel_write_kernel(const float* __restrict__ nms, const unsigned char* __restrict__ mask, int H, int W,
                                 float thr, int border, const int* __restrict__ block_offsets, int cap,
                                 int* __restrict__ cand_pix, float* __restrict__ cand_score) {
  int img = blockIdx.y;
  const float* s = nms + (long long)img * H * W;
  const unsigned char* mk = mask ? mask + (long long)img * H * W : nullptr;
  int base = blockIdx.x * SEL_PER_BLOCK;
  __shared__ int running;
  __shared__ int wsums[8];
  if (threadIdx.x == 0) running = block_offsets[img * gridDim.x + blockIdx.x];
  __syncthreads();
  // 256 threads, 4 strips of 256 consecutive pixels: order = strip, then thread
  for (int strip = 0; strip < SEL_PER_BLOCK / 256; strip++) {
    int p = base + strip * 256 + threadIdx.x;
    float v = 0.f;
    bool keep = p < H * W && sel_pred(s, mk, p, H, W, thr, border, &v);
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wsums[wid] = __popc(bal);
    __syncthreads();
    int off = running;
    for (int k = 0; k < wid; k++) off += wsums[k];
    off += __popc(bal & ((1u << lane) - 1));
    if (keep && off < cap) {
      cand_pix[(long long)img * cap + off] = p;
      cand_score[(long long)img * cap + off] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < 8; k++) t += wsums[k];
      running += t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// rank sort: position = #{j : s_j > s_i  or (s_j == s_i and j < i)}; keep rank < k.
// Candidates arrive in ascending pixel order, so the tie rule is (score desc, pixel index asc).
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// 64 candidates per block (a few thousand candidates per image must still fill 148 SMs); the keys of all
// candidates stream through a 256-entry shared tile that every thread scans four at a time.
constexpr int RANK_THREADS = 64;
constexpr int SORT_MAX = 8192;   // candidates one CTA sorts in shared memory (64 KB of keys)
__global__ void __launch_bounds__(RANK_THREADS) rank_sort_kernel(const int* __restrict__ cand_pix,
                                                                const float* __restrict__ cand_score,
                                                                const int* __restrict__ counts, int cap, int max_kp,
                                                                int W, int out_cap, float* __restrict__ kpts_xy,
                                                                float* __restrict__ scores_out,
                                                                int* __restrict__ out_counts) {
  int img = blockIdx.y;
  int C = min(counts[img], cap);
  if (C <= SORT_MAX) return;   // sorted by bitonic_rank_kernel
  int keep = (max_kp >= 0 && max_kp < C) ? max_kp : C;
  keep = min(keep, out_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) out_counts[img] = keep;
  const float* sc = cand_score + (long long)img * cap;
  const int* px = cand_pix + (long long)img * cap;
  __shared__ __align__(16) unsigned tile[256];
  for (int blk = blockIdx.x; blk * RANK_THREADS < C; blk += gridDim.x) {   // block-uniform
  int i = blk * RANK_THREADS + threadIdx.x;
  unsigned ki = i < C ? f2ord(sc[i]) : 0u;
  int rank = 0;
  for (int base = 0; base < C; base += 256) {
#pragma unroll
    for (int q = 0; q < 256 / RANK_THREADS; q++) {
      int j = base + q * RANK_THREADS + threadIdx.x;
      tile[q * RANK_THREADS + threadIdx.x] = j < C ? f2ord(sc[j]) : 0u;   // 0 never outranks a real key
    }
    __syncthreads();
    // entries before candidate i win ties (pixel-ascending order): count `>=` there and `>` after it
    const int nge = min(max(i - base, 0), 256);   // tile entries [0, nge) precede i
    for (int t = 0; t < 256; t += 4) {
      const uint4 k4 = *reinterpret_cast<const uint4*>(&tile[t]);
      rank += (k4.x > ki) + (k4.y > ki) + (k4.z > ki) + (k4.w > ki);
      if (t < nge)
        rank += (t < nge && k4.x == ki) + (t + 1 < nge && k4.y == ki) + (t + 2 < nge && k4.z == ki) +
                (t + 3 < nge && k4.w == ki);
    }
    __syncthreads();
  }
  if (i < C && rank < keep) {
    int p = px[i];
    int y = p / W, x = p - y * W;
    kpts_xy[((long long)img * out_cap + rank) * 2 + 0] = (float)x;
    kpts_xy[((long long)img * out_cap + rank) * 2 + 1] = (float)y;
    scores_out[(long long)img * out_cap + rank] = sc[i];
  }
  }
}

// The usual case (a few thousand candidates per image): one CTA per image sorts 64-bit keys
// (inverted ordered score | candidate index) with a bitonic network in shared memory -- ascending keys are
// (score descending, pixel ascending), the same total order as the rank counting above -- and writes the first
// `keep` entries.  O(C log^2 C) instead of O(C^2): 4 700 candidates take ~5 us instead of ~25.
__global__ void __launch_bounds__(1024) bitonic_rank_kernel(const int* __restrict__ cand_pix,
                                                            const float* __restrict__ cand_score,
                                                            const int* __restrict__ counts, int cap, int max_kp, int W,
                                                            int out_cap, float* __restrict__ kpts_xy,
                                                            float* __restrict__ scores_out,
                                                            int* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned long long keys[];
  const int img = blockIdx.x;
  const int C = min(counts[img], cap);
  if (C > SORT_MAX) return;    // rank_sort_kernel takes this image
  int keep = (max_kp >= 0 && max_kp < C) ? max_kp : C;
  keep = min(keep, out_cap);
  if (threadIdx.x == 0) out_counts[img] = keep;
  if (keep == 0) return;
  const float* sc = cand_score + (long long)img * cap;
  const int* px = cand_pix + (long long)img * cap;
  int n2 = 64;
  while (n2 < C) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += 1024)
    keys[i] = i < C ? ((unsigned long long)(~f2ord(sc[i])) << 32) | (unsigned)i : ~0ull;
  __syncthreads();
  // The network is bound by shared-memory traffic (every stage moves all keys in and out), so up to three
  // consecutive stages (strides 4 low, 2 low, low) run on eight keys held in registers per round trip:
  // 35 passes instead of 91 for 8 192 keys.
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0;) {
      const int r = j >= 4 ? 3 : (j == 2 ? 2 : 1);   // stages in this pass
      const int low = j >> (r - 1);                  // smallest stride of the pass
      const int m_cnt = 1 << r;
      for (int gidx = threadIdx.x; gidx < (n2 >> r); gidx += 1024) {
        const int base = ((gidx & ~(low - 1)) << r) | (gidx & (low - 1));
        const bool up = (base & k) == 0;             // the bits of this pass lie below k: one direction per group
        unsigned long long v[8];
#pragma unroll
        for (int m = 0; m < 8; m++)
          if (m < m_cnt) v[m] = keys[base + m * low];
#pragma unroll
        for (int sbit = 2; sbit >= 0; sbit--) {
          if (sbit < r) {
#pragma unroll
            for (int m = 0; m < 8; m++) {
              if (m < m_cnt && (m & (1 << sbit)) == 0) {
                const int o = m | (1 << sbit);
                const unsigned long long x = v[m], y = v[o];
                if ((x > y) == up) {
                  v[m] = y;
                  v[o] = x;
                }
              }
            }
          }
        }
#pragma unroll
        for (int m = 0; m < 8; m++)
          if (m < m_cnt) keys[base + m * low] = v[m];
      }
      __syncthreads();
      j >>= r;
    }
  }
  for (int r = threadIdx.x; r < keep; r += 1024) {
    const int i = (int)(keys[r] & 0xffffffffu);
    const int p = px[i];
    const int y = p / W, x = p - y * W;
    kpts_xy[((long long)img * out_cap + r) * 2 + 0] = (float)x;
    kpts_xy[((long long)img * out_cap + r) * 2 + 1] = (float)y;
    scores_out[(long long)img * out_cap + r] = sc[i];
  }
}

// ------------------------------------------------------------------------------------------
// descriptor sampling.  dmap: raw convDb output NHWC [n, h, w, 256] (normalised per cell on the
// fly = F.normalize(dim=1)); out: channel-major (256, ld_out) per image (reference layout).
__global__ void __launch_bounds__(256) sample_desc_kernel(const float* __restrict__ dmap, int h, int w,
                                                          const float* __restrict__ kpts_xy,
                                                          const int* __restrict__ counts, int kp_cap,
                                                          int align_corners, float* __restrict__ out, int ld_out) {
  __shared__ float tile[256][33];
  const int img = blockIdx.y;
  const int K = min(counts[img], kp_cap);
  const int k0 = blockIdx.x * 32;
  if (k0 >= K) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* dm = dmap + (long long)img * h * w * 256;
  const float sx = (float)(w * 8) - 4.0f - 0.5f, sy = (float)(h * 8) - 4.0f - 0.5f;
  for (int q = 0; q < 4; q++) {
    int kl = warp * 4 + q;
    int k = k0 + kl;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = 0.f;
    if (k < K) {
      float kx = kpts_xy[((long long)img * kp_cap + k) * 2 + 0];
      float ky = kpts_xy[((long long)img * kp_cap + k) * 2 + 1];
      float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(kx, 4.0f), 0.5f), sx), 2.0f), 1.0f);
      float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(ky, 4.0f), 0.5f), sy), 2.0f), 1.0f);
      float px, py;
      if (align_corners) {
        px = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(w - 1));
        py = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(h - 1));
      } else {
        px = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)w), 1.0f), 2.0f);
        py = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)h), 1.0f), 2.0f);
      }
      float fx0 = floorf(px), fy0 = floorf(py);
      float fx = px - fx0, fy = py - fy0;
      int x0 = (int)fx0, y0 = (int)fy0;
#pragma unroll
      for (int t = 0; t < 4; t++) {
        int xi = x0 + (t & 1), yi = y0 + (t >> 1);
        float wt = ((t & 1) ? fx : 1.f - fx) * ((t >> 1) ? fy : 1.f - fy);
        if (xi < 0 || xi >= w || yi < 0 || yi >= h) continue;  // warp-uniform
        const float4* cell = reinterpret_cast<const float4*>(dm + ((long long)yi * w + xi) * 256);
        float4 a = cell[lane], b = cell[32 + lane];
        float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
        ss = warp_sum(ss);
        float nrm = fmaxf(sqrtf(ss), 1e-12f);
        acc[0] += wt * (a.x / nrm); acc[1] += wt * (a.y / nrm); acc[2] += wt * (a.z / nrm); acc[3] += wt * (a.w / nrm);
        acc[4] += wt * (b.x / nrm); acc[5] += wt * (b.y / nrm); acc[6] += wt * (b.z / nrm); acc[7] += wt * (b.w / nrm);
      }
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < 8; c++) ss += acc[c] * acc[c];
      ss = warp_sum(ss);
      float nrm = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int c = 0; c < 8; c++) acc[c] = acc[c] / nrm;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
      tile[lane * 4 + c][kl] = acc[c];
      tile[128 + lane * 4 + c][kl] = acc[4 + c];
    }
  }
  __syncthreads();
  float* o = out + (long long)img * 256 * ld_out;
  for (int c = warp; c < 256; c += 8) {
    int k = k0 + lane;
    if (k < K) o[(long long)c * ld_out + k] = tile[c][lane];
  }
}

}  // namespace

int sp_softmax_shuffle(const float* logits, int ldl, float* out, int n_img, int h, int w, cudaStream_t st) {
  long long threads = (long long)n_img * h * w * 32;
  if (threads == 0) return SFM_OK;
  softmax_shuffle_kernel<<<cdiv(threads, 256), 256, 0, st>>>(logits, ldl, out, n_img, h, w);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int sp_nms(const float* scores, float* out, int n_img, int H, int W, int radius, cudaStream_t st) {
  SFM_REQUIRE(radius >= 0, "simple_nms: nms_radius must be >= 0");
  SFM_REQUIRE(radius <= 8, "simple_nms: nms_radius > 8 is not supported by the tiled kernel (got %d)", radius);
  if (n_img == 0 || H == 0 || W == 0) return SFM_OK;
  constexpr size_t SMEM_MAX = 227 * 1024;   // one opt-in per device to the largest dynamic shared memory of a CTA
  if (radius == 4) {
    constexpr int TILE = 64, T = TILE + 40;
    constexpr size_t smem = (size_t)T * T * (4 * sizeof(float) + 1) + 16;
    static_assert(smem <= SMEM_MAX, "nms tile does not fit");
    SFM_SMEM_OPT_IN((nms_kernel<TILE, 4>), SMEM_MAX);
    dim3 grid(cdiv(W, TILE), cdiv(H, TILE), n_img);
    nms_kernel<TILE, 4><<<grid, 1024, smem, st>>>(scores, out, H, W, radius);
  } else {
    int T = NMS_TILE + 10 * radius;
    size_t smem = (size_t)T * T * (4 * sizeof(float) + 1) + 16;
    SFM_REQUIRE(smem <= SMEM_MAX, "simple_nms: nms_radius %d needs %zu bytes of shared memory per tile (limit %zu)", radius,
                smem, SMEM_MAX);
    SFM_SMEM_OPT_IN((nms_kernel<NMS_TILE, 0>), SMEM_MAX);
    dim3 grid(cdiv(W, NMS_TILE), cdiv(H, NMS_TILE), n_img);
    nms_kernel<NMS_TILE, 0><<<grid, 512, smem, st>>>(scores, out, H, W, radius);
  }
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

size_t sp_select_workspace_ints(int n_img, int H, int W) { return (size_t)n_img * cdiv((long long)H * W, SEL_PER_BLOCK); }

int sp_select(const float* nms, const unsigned char* mask, int n_img, int H, int W, float thr, int border,
              int* block_tmp, int cap, int* cand_pix, float* cand_score, int* counts, cudaStream_t st) {
  if (n_img == 0) return SFM_OK;
  int nblocks = cdiv((long long)H * W, SEL_PER_BLOCK);
  dim3 grid(nblocks, n_img);
  sel_count_kernel<<<grid, 256, 0, st>>>(nms, mask, H, W, thr, border, block_tmp);
  SFM_LAUNCH_CHECK();
  sel_scan_kernel<<<n_img, 256, 0, st>>>(block_tmp, nblocks, counts);
  SFM_LAUNCH_CHECK();
  sel_write_kernel<<<grid, 256, 0, st>>>(nms, mask, H, W, thr, border, block_tmp, cap, cand_pix, cand_score);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

int sp_rank_sort(const int* cand_pix, const float* cand_score, const int* counts, int n_img, int cap, int max_kp,
                 int W, int out_cap, float* kpts_xy, float* scores_out, int* out_counts, cudaStream_t st) {
  if (n_img == 0) return SFM_OK;
  const size_t smem = sizeof(unsigned long long) * SORT_MAX;
  SFM_SMEM_OPT_IN(bitonic_rank_kernel, smem);
  bitonic_rank_kernel<<<n_img, 1024, smem, st>>>(cand_pix, cand_score, counts, cap, max_kp, W, out_cap, kpts_xy,
                                                 scores_out, out_counts);
  // images with more than SORT_MAX candidates (every block returns at once otherwise)
  dim3 grid(std::min(cdiv(cap, RANK_THREADS), 592), n_img);
  rank_sort_kernel<<<grid, RANK_THREADS, 0, st>>>(cand_pix, cand_score, counts, cap, max_kp, W, out_cap, kpts_xy, scores_out,
                                         out_counts);
  SFM_LAUNCH_CHECK_N(2);
  return SFM_OK;
}

int sp_sample_descriptors(const float* dmap, int n_img, int h, int w, const float* kpts_xy, const int* counts,
                          int kp_cap, int align_corners, float* out, int ld_out, cudaStream_t st) {
  if (n_img == 0 || kp_cap == 0) return SFM_OK;
  dim3 grid(cdiv(kp_cap, 32), n_img);
  sample_desc_kernel<<<grid, 256, 0, st>>>(dmap, h, w, kpts_xy, counts, kp_cap, align_corners, out, ld_out);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

// ---- image ingestion: 8-bit grey -> float in [0,1], exactly numpy's `astype(float32) / 255`
// (matcher.py:360: IEEE single-precision division), 16 pixels per thread
namespace {
__global__ void u8_to_unit_kernel(const uint4* __restrict__ in16, const unsigned char* __restrict__ in,
                                  float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long base = i * 16;
  if (base >= n) return;
  if (in16 != nullptr && base + 16 <= n) {
    const uint4 v = in16[i];
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float4 o;
      o.x = __fdiv_rn((float)(w[k] & 0xffu), 255.f);
      o.y = __fdiv_rn((float)((w[k] >> 8) & 0xffu), 255.f);
      o.z = __fdiv_rn((float)((w[k] >> 16) & 0xffu), 255.f);
      o.w = __fdiv_rn((float)(w[k] >> 24), 255.f);
      *reinterpret_cast<float4*>(out + base + 4 * k) = o;
    }
  } else {
    for (long long j = base; j < n && j < base + 16; j++) out[j] = __fdiv_rn((float)in[j], 255.f);
  }
}
}  // namespace

int sp_u8_to_unit_f32(const unsigned char* in, float* out, long long n, cudaStream_t st) {
  if (n == 0) return SFM_OK;
  const bool vec = (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  u8_to_unit_kernel<<<cdiv(cdiv(n, 16), 256), 256, 0, st>>>(vec ? reinterpret_cast<const uint4*>(in) : nullptr, in, out, n);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

}  // namespace sfm
