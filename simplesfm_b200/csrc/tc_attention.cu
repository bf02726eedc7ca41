// Fused multi-head attention on tcgen05 (reference: attention / MultiHeadedAttention,
// simple_sfm/models/superglue.py:85-108; head_dim 64, 4 heads, the (N x M) probability matrix never
// leaves the SM).
//
// One CTA = (pair, head, 256 query rows) as two 128-row tiles that ping-pong on the tensor pipe:
//   warp 0    : TMA producer (Q tiles once, then K/V tiles through a 3-stage ring)
//   warp 1    : tcgen05.mma issuer:  S_t = Q_t K_j^T  (SS, fp32 in TMEM),  O_t = P_t V_j  (A = P from TMEM)
//   warps 2-5 : softmax warpgroup of tile 0 (thread = query row), warps 6-9: tile 1
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)
// O accumulates in TMEM across K/V tiles.  The softmax threads keep a (possibly stale) reference
// maximum per row and only rescale O / the running sum when the true maximum has grown by more
// than 2^8 (lazy rescale: probabilities stay <= 256, exact after the final division by the sum).
#include "tc_common.cuh"
#include "tc_path.cuh"
#include <stdlib.h>

namespace sfm {
namespace tc {

namespace {

constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, 128-byte swizzled rows
template <int NT> struct AttCfg { static constexpr int THREADS = 64 + NT * 128; };

struct AttBars {
  uint64_t q_full;
  uint64_t kv_full[KV_STAGES], kv_empty[KV_STAGES];
  uint64_t s_full[2], s_free[2], p_full[2], o_full[2];
  uint32_t tmem_base;
};

template <int NT>
__global__ void __launch_bounds__(64 + NT * 128, NT == 1 ? 2 : 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap map, TcAttn p, int units0, int qt0, int qt1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                         // NT tiles
  uint8_t* sK = smem + NT * TILE_BYTES;       // KV_STAGES tiles
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;  // KV_STAGES tiles
  AttBars* bars = reinterpret_cast<AttBars*>(sV + KV_STAGES * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- decode the work unit: side, pair, head, 256-row query block
  int u = blockIdx.x, side = 0;
  if (u >= units0) {
    u -= units0;
    side = 1;
  }
  const int qts = side == 0 ? qt0 : qt1;
  const int head = u & 3;
  const int qblk = (u >> 2) % qts;
  const int b = (u >> 2) / qts;
  const int Nq = p.Nq[side], Nk = p.Nk[side];
  const long long qrow = p.q_row0[side] + (long long)b * Nq + qblk * (NT * 128);
  const long long krow = p.k_row0[side] + (long long)b * Nk;
  const int n_kv = (Nk + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < KV_STAGES; s++) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 1);
    }
    for (int t = 0; t < NT; t++) {
      mbar_init(&bars->s_full[t], 1);
      mbar_init(&bars->p_full[t], 4);
      mbar_init(&bars->o_full[t], 1);
      mbar_init(&bars->s_free[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<NT * 256>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  constexpr uint32_t COL_S = 0, COL_O = NT * 128, COL_P = NT * 128 + NT * 64;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->q_full, NT * TILE_BYTES);
      for (int t = 0; t < NT; t++) tma_load_2d(sQ + t * TILE_BYTES, &map, &bars->q_full, head * 64, (int)qrow + t * 128);
      for (int j = 0; j < n_kv; j++) {
        const int s = j % KV_STAGES;
        mbar_wait(&bars->kv_empty[s], ((j / KV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->kv_full[s], 2 * TILE_BYTES);
        tma_load_2d(sK + s * TILE_BYTES, &map, &bars->kv_full[s], 256 + head * 64, (int)krow + j * 128);
        tma_load_2d(sV + s * TILE_BYTES, &map, &bars->kv_full[s], 512 + head * 64, (int)krow + j * 128);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);  // S = Q K^T : B = K tile, K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);   // O = P V   : B = V tile, MN-major
      uint64_t q_desc[NT], k_desc[KV_STAGES], v_desc[KV_STAGES];
#pragma unroll
      for (int t = 0; t < NT; t++) q_desc[t] = umma_desc_sw128(smem_u32(sQ + t * TILE_BYTES), 16, 1024);
#pragma unroll
      for (int s = 0; s < KV_STAGES; s++) {
        k_desc[s] = umma_desc_sw128(smem_u32(sK + s * TILE_BYTES), 16, 1024);
        v_desc[s] = umma_desc_sw128(smem_u32(sV + s * TILE_BYTES), 1024, 1024);
      }
      auto issue_s = [&](int t, int stage) {
        const uint64_t ad = (NT == 1 || t == 0) ? q_desc[0] : q_desc[NT - 1], bd = stage == 0 ? k_desc[0] : (stage == 1 ? k_desc[1] : k_desc[2]);
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_ss(tmem + COL_S + t * 128, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc_s,
                  k != 0);
        umma_commit(&bars->s_full[t]);
      };
      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->kv_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < NT; t++) issue_s(t, 0);
      for (int j = 0; j < n_kv; j++) {
        const int s = j % KV_STAGES;
        // (1) as soon as a warpgroup has pulled S_t(j) into registers its TMEM region is free again: issue
        //     S_t(j+1) right away so that it is ready when the group finishes the exponentials of tile j
        if (j + 1 < n_kv) {
          const int s1 = (j + 1) % KV_STAGES;
          mbar_wait(&bars->kv_full[s1], ((j + 1) / KV_STAGES) & 1);
          for (int t = 0; t < NT; t++) {
            mbar_wait(&bars->s_free[t], j & 1);
            tc_fence_after();
            issue_s(t, s1);
          }
        }
        // (2) P_t(j) written (much later): accumulate O_t += P_t V_j
        for (int t = 0; t < NT; t++) {
          mbar_wait(&bars->p_full[t], j & 1);
          tc_fence_after();
          const uint64_t vd = s == 0 ? v_desc[0] : (s == 1 ? v_desc[1] : v_desc[2]);
#pragma unroll
          for (int k = 0; k < 8; k++)  // 16 keys per step: 16 V rows = 2048 bytes, P advances 8 columns
            umma_ts(tmem + COL_O + t * 64, tmem + COL_P + t * 64 + k * 8, umma_desc_advance(vd, k * 2048), idesc_o,
                    (j | k) != 0);
          umma_commit(&bars->o_full[t]);
        }
        umma_commit(&bars->kv_empty[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = (warp - 2) >> 2;  // query tile of this warpgroup
    const int quad = warp & 3;      // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    const int qi = qblk * (NT * 128) + t * 128 + row_in_tile;  // query index inside the pair
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const float c = 0.125f * 1.4426950408889634f;  // log2(e) / sqrt(64)
    float m_used = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_kv; j++) {
      mbar_wait(&bars->s_full[t], j & 1);
      tc_fence_after();
      const int valid = min(128, Nk - j * 128);  // columns of this tile that are real keys
      uint32_t sv[128];
#pragma unroll
      for (int ch = 0; ch < 4; ch++)
        tmem_ld_32x32(lane_addr + COL_S + t * 128 + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[ch * 32]));
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free[t]);   // S_t is in registers: the tensor pipe may overwrite it
      if (valid < 128) {
#pragma unroll
        for (int i = 0; i < 128; i++)
          if (i >= valid) sv[i] = 0xff800000u;  // -inf
      }
      // 8 independent chains (a single 128-long FMNMX dependency chain is latency-bound with 2 warps/SMSP)
      float mxa[8];
#pragma unroll
      for (int q = 0; q < 8; q++) mxa[q] = __uint_as_float(sv[q]);
#pragma unroll
      for (int i = 8; i < 128; i += 8)
#pragma unroll
        for (int q = 0; q < 8; q++) mxa[q] = fmaxf(mxa[q], __uint_as_float(sv[i + q]));
      const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])),
                             fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
      // lazy rescale decision (the O rescale itself is deferred until S has been consumed, so that
      // the 128 S registers and the O staging registers are never live together)
      float alpha = 1.f;
      bool grow = false;
      if (j == 0) {
        m_used = mx;
      } else {
        grow = (mx - m_used) * c > 8.0f;
        if (grow) {
          alpha = fast_exp2((m_used - mx) * c);
          m_used = mx;
          l_run *= alpha;
        }
      }
      const bool any_grow = __any_sync(0xffffffffu, grow);
      const float mc = m_used * c;
      if (j > 0) {  // P V (j-1) must have consumed P_t before it is overwritten (and O_t must be quiescent)
        mbar_wait(&bars->o_full[t], (j - 1) & 1);
        tc_fence_after();
      }
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int half = 0; half < 2; half++) {   // two 64-key halves keep the packed-P registers short-lived
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const int e = half * 64 + i;
          const float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), c, -mc));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), c, -mc));
          const float p2 = fast_exp2(fmaf(__uint_as_float(sv[e + 2]), c, -mc));
          const float p3 = fast_exp2(fmaf(__uint_as_float(sv[e + 3]), c, -mc));
          ls[0] += p0; ls[1] += p1; ls[2] += p2; ls[3] += p3;
          pk[i >> 1] = pack_bf16(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16(p2, p3);
        }
        tmem_st_32x32(lane_addr + COL_P + t * 64 + half * 32, pk);
      }
      l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      if (any_grow) {   // warp-uniform: rescale the running output of rows whose reference maximum moved
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + COL_O + t * 64 + ch * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_32x32(lane_addr + COL_O + t * 64 + ch * 32, r);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[t]);
    }
    mbar_wait(&bars->o_full[t], (n_kv - 1) & 1);
    tc_fence_after();
    {
      uint32_t o[64];
      tmem_ld_32x32(lane_addr + COL_O + t * 64, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
      tmem_ld_32x32(lane_addr + COL_O + t * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
      tmem_wait_ld();
      if (qi < Nq) {
        const float inv = 1.f / l_run;
        __nv_bfloat16* dst = p.out + (qrow + t * 128 + row_in_tile) * 256 + head * 64;
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int i = 0; i < 8; i++)
          d4[i] = make_uint4(pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                             pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                             pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                             pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<NT * 256>(tmem);
}

template <int NT>
int launch_att(const TcAttn& p, const CUtensorMap& map, cudaStream_t st) {
  const int qt0 = cdiv(p.Nq[0], NT * 128), qt1 = p.Nq[1] > 0 ? cdiv(p.Nq[1], NT * 128) : 0;
  const int units0 = p.B * 4 * qt0, units1 = p.B * 4 * qt1;
  const size_t smem = (size_t)(NT + 2 * KV_STAGES) * TILE_BYTES + sizeof(AttBars) + 1024;
  static bool configured = false;
  if (!configured) {
    SFM_CUDA(cudaFuncSetAttribute(tc_attention_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  tc_attention_kernel<NT><<<units0 + units1, AttCfg<NT>::THREADS, smem, st>>>(map, p, units0, qt0, qt1 > 0 ? qt1 : 1);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}


}  // namespace

int tc_attention(const TcAttn& p, cudaStream_t st) {
  SFM_REQUIRE(p.B > 0 && p.Nq[0] > 0 && p.Nk[0] > 0 && p.Nq[1] >= 0, "tc_attention: bad dims");
  CUtensorMap map;
  SFM_TRY(make_tmap_bf16(&map, p.qkv, p.rows_total, 768, 768, 128));
  static int nt = getenv("SFM_ATT_NT") ? atoi(getenv("SFM_ATT_NT")) : 2;
  if (nt == 1) return launch_att<1>(p, map, st);
  return launch_att<2>(p, map, st);
}

}  // namespace tc
}  // namespace sfm
