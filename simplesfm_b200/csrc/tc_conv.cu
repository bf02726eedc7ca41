// 3x3 / pad 1 convolution as an implicit GEMM on tcgen05 (SuperPoint VGG encoder and heads,
// reference simple_sfm/models/superpoint.py:84-99,112-126,157-158).  NHWC bf16 activations.
//
// GEMM view: M = pixels, N = Cout, K = 9 * Cin ordered (tap, ci).  One tile = 8 rows x 16 columns
// of pixels (128 = UMMA M).  The im2col matrix is never built: for every tap (ky, kx) the A operand
// is ONE 4-D TMA box {64 ch, 16 x, 8 y, 1 img} of the input at pixel offset (kx-1, ky-1); the TMA unit
// zero-fills whatever falls outside the image, which is exactly the padding.  The [Cout x 64] weight
// slice of the same k-block streams through the same pipeline stage (weights are L2-resident).
//   warp 0   : TMA producer (A box + weight slice per k-block, 4 stages)
//   warp 1   : tcgen05.mma issuer, accumulators double-buffered in TMEM (2 x Cout columns)
//   warps 2-5: epilogue: tcgen05.ld -> bias + ReLU -> bf16 slab (swizzled smem) -> 4-D TMA store
#include "tc_common.cuh"
#include "tc_path.cuh"

namespace sfm {
namespace tc {

namespace {

constexpr int CONV_STAGES = 4;
constexpr int A_BYTES = 128 * 128;     // 128 pixels x 64 channels bf16
constexpr int SLAB = 128 * 128;
constexpr int CONV_THREADS = 192;

struct ConvBars {
  uint64_t full[CONV_STAGES], empty[CONV_STAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c, int x, int y,
                                            int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(x), "r"(y), "r"(n)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c, int x, int y, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c), "r"(x), "r"(y), "r"(n)
               : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ uint4 lds128q(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162*>(&a), *pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int k = 0; k < 4; k++) pr[k] = __hmax2(pa[k], pb[k]);
  return r;
}

// POOL: the 2 x 2 max-pool that follows conv1b / conv2b / conv3b (superpoint.py:114-119) runs in the epilogue -- a
// tile is 8 x 16 pixels, so every pooling window lies inside it -- and only the pooled map (a quarter of the bytes)
// is written: the full-resolution activation (157 MB for four 480 x 640 images after conv1b) never reaches HBM.
template <int COUT, int POOL>
__global__ void __launch_bounds__(CONV_THREADS, 1)
tc_conv3x3_kernel(const __grid_constant__ CUtensorMap mapIn, const __grid_constant__ CUtensorMap mapW,
                  const __grid_constant__ CUtensorMap mapOut, TcConv p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int W_BYTES = COUT * 128;
  constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  uint8_t* sStage = smem;
  uint8_t* sO = smem + CONV_STAGES * STAGE_BYTES;
  float* sBias = reinterpret_cast<float*>(sO + 2 * SLAB);
  ConvBars* bars = reinterpret_cast<ConvBars*>(sBias + COUT);
  constexpr int TMEM_COLS = 2 * COUT < 32 ? 32 : 2 * COUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapIn);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapOut);
    for (int s = 0; s < CONV_STAGES; s++) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  for (int i = threadIdx.x; i < COUT; i += CONV_THREADS) sBias[i] = p.bias ? p.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int tiles_x = (p.W + 15) / 16, tiles_y = (p.H + 7) / 8;
  const int tiles = p.n * tiles_y * tiles_x;
  const int cblocks = p.Cin / 64;
  const int kblocks = 9 * cblocks;

  // Producer and issuer warps run CONVERGED and one elected lane issues: coordinates, descriptors and barrier
  // addresses then live in uniform registers.  Under `if (lane == 0)` the compiler wraps every TMA / tcgen05.mma in an
  // ELECT / R2UR.BROADCAST / branch sequence of ~85 clocks -- more than the 32 (Cout = 64) or 64 (Cout = 128) tensor
  // clocks of one MMA, which made these kernels issue-bound.
  if (warp == 0) {
    {
      const bool leader = elect_one_sync();
      int stage = 0, phase = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int img = t / (tiles_y * tiles_x), rem = t % (tiles_y * tiles_x);
        const int y0 = (rem / tiles_x) * 8, x0 = (rem % tiles_x) * 16;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* st = sStage + stage * STAGE_BYTES;
          if (leader) {
            mbar_arrive_expect_tx(&bars->full[stage], STAGE_BYTES);
            tma_load_4d(st, &mapIn, &bars->full[stage], cb * 64, x0 + tap % 3 - 1, y0 + tap / 3 - 1, img);
            tma_load_2d(st + A_BYTES, &mapW, &bars->full[stage], tap * p.Cin + cb * 64, 0);
          }
          __syncwarp();
          if (++cb == cblocks) {
            cb = 0;
            tap++;
          }
          if (++stage == CONV_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = elect_one_sync();
      constexpr uint32_t idesc = umma_idesc_bf16(128, COUT, 0);
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sStage), 16, 1024);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sStage + A_BYTES), 16, 1024);
      int stage = 0, phase = 0, tile = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, tile++) {
        const int buf = tile & 1;
        mbar_wait(&bars->acc_empty[buf], ((tile >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * COUT;
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint64_t ad = umma_desc_advance(a_desc0, stage * STAGE_BYTES);
          const uint64_t bd = umma_desc_advance(w_desc0, stage * STAGE_BYTES);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              umma_ss(d_tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc, (kb | k) != 0);
            umma_commit(&bars->empty[stage]);
          }
          __syncwarp();
          if (++stage == CONV_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) umma_commit(&bars->acc_full[buf]);
        __syncwarp();
      }
    }
  } else {
    const int quad = warp & 3;
    const int etid = (warp - 2) * 32 + lane;          // 0..127 (not the pixel index!)
    const int pix = quad * 32 + lane;                 // TMEM lane = pixel inside the tile (y*16 + x)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    int tile = 0, slab_i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, tile++) {
      const int img = t / (tiles_y * tiles_x), rem = t % (tiles_y * tiles_x);
      const int y0 = (rem / tiles_x) * 8, x0 = (rem % tiles_x) * 16;
      const int buf = tile & 1;
      mbar_wait(&bars->acc_full[buf], (tile >> 1) & 1);
      tc_fence_after();
      constexpr int NCH = COUT / 32;
      uint32_t r[2][32];
      tmem_ld_32x32(lane_addr + buf * COUT, r[0]);
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        tmem_wait_ld();
        if (c + 1 < NCH) tmem_ld_32x32(lane_addr + buf * COUT + (c + 1) * 32, r[(c + 1) & 1]);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float x = __uint_as_float(r[c & 1][j]) + sBias[c * 32 + j];
          v[j] = p.relu ? fmaxf(x, 0.f) : x;
        }
        if (!POOL) {
        uint8_t* slab = sO + (slab_i & 1) * SLAB;
        if ((c & 1) == 0) {
          if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // slab used 2 stores ago is free
          bar_sync(1, 128);
        }
        const uint32_t rowa = smem_u32(slab) + pix * 128;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int chunk = (c & 1) * 4 + j;
          sts128(rowa + ((chunk ^ (pix & 7)) << 4), pack_bf16(v[8 * j], v[8 * j + 1]),
                 pack_bf16(v[8 * j + 2], v[8 * j + 3]), pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                 pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        }
        if ((c & 1) == 1) {
          fence_proxy_async();
          bar_sync(1, 128);
          if (etid == 0) {
            tma_store_4d(&mapOut, slab, (c >> 1) * 64, x0, y0, img);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          slab_i++;
        }
        } else {
        // full-resolution slab (64 channels of the 128 pixels) in sO[0 .. 16 KB), pooled slabs (32 pixels x 128 B,
        // double-buffered against the TMA store) behind it
        if ((c & 1) == 0) bar_sync(1, 128);            // the pooling reads of the previous half are done
        const uint32_t rowa = smem_u32(sO) + pix * 128;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int chunk = (c & 1) * 4 + j;
          sts128(rowa + ((chunk ^ (pix & 7)) << 4), pack_bf16(v[8 * j], v[8 * j + 1]),
                 pack_bf16(v[8 * j + 2], v[8 * j + 3]), pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                 pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        }
        if ((c & 1) == 1) {
          if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // pooled slab of 2 stores ago is free
          bar_sync(1, 128);
          uint8_t* pslab = sO + SLAB + (slab_i & 1) * 4096;
          const int pp = etid >> 2, cq = etid & 3;     // pooled pixel (4 rows x 8 columns), 16-channel quarter
          const int sp = (pp >> 3) * 32 + (pp & 7) * 2;   // top-left source pixel: y = 2 py, x = 2 px, 16 pixels per row
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int chunk = cq * 2 + h;
            uint4 m = lds128q(smem_u32(sO) + sp * 128 + ((chunk ^ (sp & 7)) << 4));
            m = max_bf16x8(m, lds128q(smem_u32(sO) + (sp + 1) * 128 + ((chunk ^ ((sp + 1) & 7)) << 4)));
            m = max_bf16x8(m, lds128q(smem_u32(sO) + (sp + 16) * 128 + ((chunk ^ ((sp + 16) & 7)) << 4)));
            m = max_bf16x8(m, lds128q(smem_u32(sO) + (sp + 17) * 128 + ((chunk ^ ((sp + 17) & 7)) << 4)));
            sts128(smem_u32(pslab) + pp * 128 + ((chunk ^ (pp & 7)) << 4), m.x, m.y, m.z, m.w);
          }
          fence_proxy_async();
          bar_sync(1, 128);
          if (etid == 0) {
            tma_store_4d(&mapOut, pslab, (c >> 1) * 64, x0 >> 1, y0 >> 1, img);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          slab_i++;
        }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
    }
    if (etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_nhwc(CUtensorMap* out, const void* base, int n, int H, int W, int C, int box_x = 16, int box_y = 8) {
  static EncodeTiledFn enc = nullptr;
  if (enc == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SFM_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d): n=%d H=%d W=%d C=%d", (int)r, n, H, W, C);
    return SFM_ERR_CUDA;
  }
  return SFM_OK;
}

template <int COUT, int POOL>
int launch_conv(const TcConv& p, int num_sms, cudaStream_t st) {
  CUtensorMap mIn, mW, mOut;
  SFM_TRY(make_tmap_nhwc(&mIn, p.in, p.n, p.H, p.W, p.Cin));
  if (POOL) SFM_TRY(make_tmap_nhwc(&mOut, p.out, p.n, p.H / 2, p.W / 2, COUT, 8, 4));
  else SFM_TRY(make_tmap_nhwc(&mOut, p.out, p.n, p.H, p.W, COUT));
  SFM_TRY(make_tmap_bf16(&mW, p.w, COUT, 9ull * p.Cin, 9ull * p.Cin, COUT));
  const int tiles = p.n * cdiv(p.H, 8) * cdiv(p.W, 16);
  const size_t smem = (size_t)CONV_STAGES * (A_BYTES + COUT * 128) + 2 * SLAB + COUT * sizeof(float) + sizeof(ConvBars) +
                      1024;
  SFM_SMEM_OPT_IN((tc_conv3x3_kernel<COUT, POOL>), smem);
  const int grid = tiles < num_sms ? tiles : num_sms;
  tc_conv3x3_kernel<COUT, POOL><<<grid, CONV_THREADS, smem, st>>>(mIn, mW, mOut, p);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

__global__ void maxpool2_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int n, int H, int W,
                                     int C8) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)n * Ho * Wo * C8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c = gid % C8;
  long long t = gid / C8;
  const int xo = t % Wo;
  t /= Wo;
  const int yo = t % Ho;
  const int img = t / Ho;
  const uint4* src = in + (((long long)img * H + 2 * yo) * W + 2 * xo) * C8 + c;
  uint4 a = src[0], b = src[C8], d = src[(long long)W * C8], e = src[(long long)W * C8 + C8];
  uint4 r;
  const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162*>(&a), *pb = reinterpret_cast<const __nv_bfloat162*>(&b),
                       *pd = reinterpret_cast<const __nv_bfloat162*>(&d), *pe = reinterpret_cast<const __nv_bfloat162*>(&e);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int k = 0; k < 4; k++) pr[k] = __hmax2(__hmax2(pa[k], pb[k]), __hmax2(pd[k], pe[k]));
  out[gid] = r;
}

}  // namespace

int tc_conv3x3(const TcConv& p, int num_sms, cudaStream_t st) {
  SFM_REQUIRE(p.Cin == 64 || p.Cin == 128, "tc_conv3x3: Cin must be 64 or 128 (got %d)", p.Cin);
  SFM_REQUIRE(p.n > 0 && p.H > 0 && p.W > 0, "tc_conv3x3: bad image dims");
  SFM_REQUIRE(!p.pool || (p.H % 2 == 0 && p.W % 2 == 0 && p.Cout <= 128), "tc_conv3x3: pooled output needs even H, W and Cout <= 128");
  switch (p.Cout) {
    case 64: return p.pool ? launch_conv<64, 1>(p, num_sms, st) : launch_conv<64, 0>(p, num_sms, st);
    case 128: return p.pool ? launch_conv<128, 1>(p, num_sms, st) : launch_conv<128, 0>(p, num_sms, st);
    case 256: return launch_conv<256, 0>(p, num_sms, st);
    default: set_error("tc_conv3x3: Cout must be 64, 128 or 256 (got %d)", p.Cout); return SFM_ERR_UNSUPPORTED;
  }
}

int maxpool2_nhwc_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int n, int H, int W, int C, cudaStream_t st) {
  SFM_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_bf16: C%%8, H%%2, W%%2 required");
  const long long total = (long long)n * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return SFM_OK;
  maxpool2_bf16_kernel<<<cdiv(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(in),
                                                         reinterpret_cast<uint4*>(out), n, H, W, C / 8);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

}  // namespace tc
}  // namespace sfm
