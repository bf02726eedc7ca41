// Fused multi-head attention on tcgen05 (reference: attention / MultiHeadedAttention,
// simple_sfm/models/superglue.py:85-108; head_dim 64, 4 heads, the (N x M) probability matrix never
// leaves the SM).
//
// Work unit = (pair, head, 256 query rows) as two 128-row tiles that ping-pong on the tensor pipe.  CTAs are
// persistent (one per SM, units strided by the grid): the Q tiles of the next unit are loaded into a second
// buffer and its first S = Q K^T is issued while the softmax groups still normalise / store the previous
// unit's output, so barrier setup, TMEM allocation and the load / store latencies are paid once per CTA.
//   warp 0    : TMA producer (Q tiles per unit, K/V tiles through a 3-stage ring that runs across units)
//   warp 1    : tcgen05.mma issuer:  S_t = Q_t K_j^T  (SS, fp32 in TMEM),  O_t = P_t V_j  (A = P from TMEM)
//   warps 2-5 : softmax warpgroup of tile 0 (thread = query row), warps 6-9: tile 1
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)
// O accumulates in TMEM across K/V tiles.  The softmax threads keep a (possibly stale) reference
// maximum per row and only rescale O / the running sum when the true maximum has grown by more
// than 2^8 (lazy rescale: probabilities stay <= 256, exact after the final division by the sum).
#include "tc_common.cuh"
#include "tc_path.cuh"
#include <stdlib.h>
#include <algorithm>

namespace sfm {
namespace tc {

namespace {

constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, 128-byte swizzled rows
template <int NT> struct AttCfg { static constexpr int THREADS = 64 + NT * 128; };

struct AttBars {
  uint64_t q_full[2], q_empty[2];
  uint64_t kv_full[KV_STAGES], kv_empty[KV_STAGES];
  uint64_t s_full[2], s_free[2], p_full[2], o_full[2];
  uint32_t tmem_base;
};

struct Unit {
  int side, head, qblk, b, Nq, Nk, n_kv;
  long long qrow, krow;
};
template <int NT>
__device__ __forceinline__ Unit decode_unit(const TcAttn& p, int u, int units0, int qt0, int qt1) {
  Unit w;
  w.side = 0;
  if (u >= units0) {
    u -= units0;
    w.side = 1;
  }
  const int qts = w.side == 0 ? qt0 : qt1;
  w.head = u & 3;
  w.qblk = (u >> 2) % qts;
  w.b = (u >> 2) / qts;
  w.Nq = p.Nq[w.side];
  w.Nk = p.Nk[w.side];
  w.qrow = p.q_row0[w.side] + (long long)w.b * w.Nq + w.qblk * (NT * 128);
  w.krow = p.k_row0[w.side] + (long long)w.b * w.Nk;
  w.n_kv = (w.Nk + 127) / 128;
  return w;
}

template <int NT>
__global__ void __launch_bounds__(64 + NT * 128, NT == 1 ? 2 : 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap map, TcAttn p, int units0, int qt0, int qt1, int n_units) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                             // 2 buffers x NT tiles
  uint8_t* sK = smem + 2 * NT * TILE_BYTES;       // KV_STAGES tiles
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;      // KV_STAGES tiles
  AttBars* bars = reinterpret_cast<AttBars*>(sV + KV_STAGES * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map);
    for (int q = 0; q < 2; q++) {
      mbar_init(&bars->q_full[q], 1);
      mbar_init(&bars->q_empty[q], 1);
    }
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

  // Every barrier of the S / P / O hand-shake completes exactly once per K/V step, so one CTA-wide step
  // counter g (running across units) gives all parities; the K/V ring stage is g % KV_STAGES.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int g = 0, it = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, it++) {
        const Unit w = decode_unit<NT>(p, u, units0, qt0, qt1);
        const int qb = it & 1;
        mbar_wait(&bars->q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->q_full[qb], NT * TILE_BYTES);
        for (int t = 0; t < NT; t++)
          tma_load_2d(sQ + (qb * NT + t) * TILE_BYTES, &map, &bars->q_full[qb], w.head * 64, (int)w.qrow + t * 128);
        for (int j = 0; j < w.n_kv; j++, g++) {
          const int s = g % KV_STAGES;
          mbar_wait(&bars->kv_empty[s], ((g / KV_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->kv_full[s], 2 * TILE_BYTES);
          tma_load_2d(sK + s * TILE_BYTES, &map, &bars->kv_full[s], 256 + w.head * 64, (int)w.krow + j * 128);
          tma_load_2d(sV + s * TILE_BYTES, &map, &bars->kv_full[s], 512 + w.head * 64, (int)w.krow + j * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && blockIdx.x < n_units) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);  // S = Q K^T : B = K tile, K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);   // O = P V   : B = V tile, MN-major
      const uint64_t q_desc0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV), 1024, 1024);
      // S_t of K/V step `stage`, Q buffer qb
      auto issue_s = [&](int t, int qb, int stage) {
        const uint64_t ad = umma_desc_advance(q_desc0, (qb * NT + t) * TILE_BYTES);
        const uint64_t bd = umma_desc_advance(k_desc0, stage * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_ss(tmem + COL_S + t * 128, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc_s,
                  k != 0);
        umma_commit(&bars->s_full[t]);
      };
      // Issue order follows a STAGGERED schedule: tile 1 runs half a K/V step behind tile 0 (its softmax group
      // starts late, see below), so that one group's exponentials overlap the other group's TMEM load / maximum /
      // waits.  Per global step g:  S0(g+1) | PV1(g-1) | S1(g+1) | PV0(g)  -- each behind the barrier that fires
      // next in that schedule, so the issuer never couples the two groups back into lock-step.
      struct Cursor { int u, it, j, n_kv; bool valid; };
      auto advance = [&](Cursor& c) {
        if (++c.j == c.n_kv) {
          c.u += gridDim.x;
          c.it++;
          c.j = 0;
          c.valid = c.u < n_units;
          if (c.valid) c.n_kv = decode_unit<NT>(p, c.u, units0, qt0, qt1).n_kv;
        }
      };
      auto issue_pv = [&](int t, int g, bool first) {
        mbar_wait(&bars->p_full[t], g & 1);
        tc_fence_after();
        const uint64_t vd = umma_desc_advance(v_desc0, (g % KV_STAGES) * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 8; k++)  // 16 keys per step: 16 V rows = 2048 bytes, P advances 8 columns
          umma_ts(tmem + COL_O + t * 64, tmem + COL_P + t * 64 + k * 8, umma_desc_advance(vd, k * 2048), idesc_o,
                  (first ? 0 : 1) | k);
        umma_commit(&bars->o_full[t]);
      };
      Cursor cur{(int)blockIdx.x, 0, 0, decode_unit<NT>(p, blockIdx.x, units0, qt0, qt1).n_kv, true};
      // very first S of this CTA
      mbar_wait(&bars->q_full[0], 0);
      mbar_wait(&bars->kv_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < NT; t++) issue_s(t, 0, 0);
      bool prev_first = false;   // step g-1 was the first of its unit (its P V overwrites O)
      int g = 0;
      for (; cur.valid; g++) {
        Cursor nxt = cur;
        advance(nxt);
        const bool cur_first = cur.j == 0, cur_last = cur.j + 1 == cur.n_kv;
        const int s1 = (g + 1) % KV_STAGES;
        // S of step g+1 (next K/V tile of this unit, or tile 0 of the next unit with its own Q buffer)
        auto issue_next_s = [&](int t) {
          if (!nxt.valid) return;
          if (t == 0) {
            if (nxt.j == 0) mbar_wait(&bars->q_full[nxt.it & 1], (nxt.it >> 1) & 1);
            mbar_wait(&bars->kv_full[s1], ((g + 1) / KV_STAGES) & 1);
          }
          mbar_wait(&bars->s_free[t], g & 1);
          tc_fence_after();
          issue_s(t, nxt.it & 1, s1);
        };
        issue_next_s(0);
        if (NT > 1 && g > 0) {
          issue_pv(NT - 1, g - 1, prev_first);
          umma_commit(&bars->kv_empty[(g - 1) % KV_STAGES]);   // both tiles are done with K/V stage g-1
        }
        for (int t = 1; t < NT; t++) issue_next_s(t);
        // all S MMAs that read this unit's Q buffer have been issued (the next step belongs to another unit)
        if (cur_last) umma_commit(&bars->q_empty[cur.it & 1]);
        issue_pv(0, g, cur_first);
        if (NT == 1) umma_commit(&bars->kv_empty[g % KV_STAGES]);
        prev_first = cur_first;
        cur = nxt;
      }
      if (NT > 1 && g > 0) {
        issue_pv(NT - 1, g - 1, prev_first);
        umma_commit(&bars->kv_empty[(g - 1) % KV_STAGES]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = (warp - 2) >> 2;  // query tile of this warpgroup
    const int quad = warp & 3;      // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const float c = 0.125f * 1.4426950408889634f;  // log2(e) / sqrt(64)
    int g = 0;
    // tile 1 starts after tile 0 has produced its first probability tile: half a step of lag (see the issuer)
    if (NT > 1 && t == 1 && p.stagger) mbar_wait(&bars->p_full[0], 0);
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
    const Unit w = decode_unit<NT>(p, u, units0, qt0, qt1);
    const int Nk = w.Nk, n_kv = w.n_kv;
    const int qi = w.qblk * (NT * 128) + t * 128 + row_in_tile;  // query index inside the pair
    float m_used = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_kv; j++, g++) {
      mbar_wait(&bars->s_full[t], g & 1);
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
      // 8 independent chains of three-input maxima (a single long dependency chain is latency-bound with
      // 2 warps per scheduler): 128 values in 64 FMNMX3
      float mxa[8];
#pragma unroll
      for (int q = 0; q < 8; q++) mxa[q] = fmaxf(__uint_as_float(sv[q]), __uint_as_float(sv[8 + q]));
#pragma unroll
      for (int i = 16; i < 128; i += 16)
#pragma unroll
        for (int q = 0; q < 8; q++) mxa[q] = fmax3(mxa[q], __uint_as_float(sv[i + q]), __uint_as_float(sv[i + 8 + q]));
      const float mx = fmaxf(fmax3(mxa[0], mxa[1], mxa[2]), fmax3(fmax3(mxa[3], mxa[4], mxa[5]), mxa[6], mxa[7]));
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
        mbar_wait(&bars->o_full[t], (g - 1) & 1);
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
    mbar_wait(&bars->o_full[t], (g - 1) & 1);
    tc_fence_after();
    {
      uint32_t o[64];
      tmem_ld_32x32(lane_addr + COL_O + t * 64, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
      tmem_ld_32x32(lane_addr + COL_O + t * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
      tmem_wait_ld();
      tc_fence_before();   // O_t is in registers: the next unit's first P V may overwrite it
      if (qi < w.Nq) {
        const float inv = 1.f / l_run;
        __nv_bfloat16* dst = p.out + (w.qrow + t * 128 + row_in_tile) * 256 + w.head * 64;
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int i = 0; i < 8; i++)
          d4[i] = make_uint4(pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                             pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                             pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                             pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
      }
    }
    }  // units
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<NT * 256>(tmem);
}

template <int NT>
int launch_att(const TcAttn& p, const CUtensorMap& map, int num_sms, cudaStream_t st) {
  const int qt0 = cdiv(p.Nq[0], NT * 128), qt1 = p.Nq[1] > 0 ? cdiv(p.Nq[1], NT * 128) : 0;
  const int units0 = p.B * 4 * qt0, units1 = p.B * 4 * qt1;
  const int n_units = units0 + units1;
  const size_t smem = (size_t)(2 * NT + 2 * KV_STAGES) * TILE_BYTES + sizeof(AttBars) + 1024;
  static bool configured = false;
  if (!configured) {
    SFM_CUDA(cudaFuncSetAttribute(tc_attention_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int ctas_per_sm = NT == 1 ? 2 : 1;
  const int grid = std::min(n_units, num_sms * ctas_per_sm);
  tc_attention_kernel<NT><<<grid, AttCfg<NT>::THREADS, smem, st>>>(map, p, units0, qt0, qt1 > 0 ? qt1 : 1, n_units);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

}  // namespace

int tc_attention(const TcAttn& p, cudaStream_t st) {
  SFM_REQUIRE(p.B > 0 && p.Nq[0] > 0 && p.Nk[0] > 0 && p.Nq[1] >= 0, "tc_attention: bad dims");
  CUtensorMap map;
  SFM_TRY(make_tmap_bf16(&map, p.qkv, p.rows_total, 768, 768, 128));
  static int nt = getenv("SFM_ATT_NT") ? atoi(getenv("SFM_ATT_NT")) : 2;
  static int stagger = getenv("SFM_ATT_STAGGER") ? atoi(getenv("SFM_ATT_STAGGER")) : 1;
  TcAttn q = p;
  q.stagger = stagger;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SFM_CUDA(cudaGetDevice(&dev));
    SFM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (nt == 1) return launch_att<1>(q, map, num_sms, st);
  return launch_att<2>(q, map, num_sms, st);
}

}  // namespace tc
}  // namespace sfm
