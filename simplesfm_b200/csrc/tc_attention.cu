// Fused multi-head attention on tcgen05 (reference: attention / MultiHeadedAttention,
// simple_sfm/models/superglue.py:85-108; head_dim 64, 4 heads, the (N x M) probability matrix never
// leaves the SM).
//
// Work unit = (pair, head, 256 query rows) as two 128-row tiles that ping-pong on the tensor pipe.  CTAs are
// persistent (one per SM, units strided by the grid): the Q tiles of the next unit are loaded into a second
// buffer and its first S = Q K^T is issued while the softmax groups still normalise / store the previous
// unit's output, so barrier setup, TMEM allocation and the load / store latencies are paid once per CTA.
//   warps 0-3  : softmax warpgroup of tile 0 (thread = query row), warps 4-7: tile 1
//   warp 8     : TMA producer (Q tiles per unit, K/V tiles through a 3-stage ring that runs across units)
//   warps 9-10 : tcgen05.mma issuers, one per query tile (the whole warp runs converged, one elected lane issues):
//                S_t = Q_t K_j^T  (SS, fp32 in TMEM),  O_t = P_t V_j  (A = P from TMEM)
// The unit's output leaves through a swizzled shared-memory staging tile and one TMA store (att_epilogue).
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)
// O accumulates in TMEM across K/V tiles.  The softmax threads exponentiate a step relative to the row maximum
// of the steps BEFORE it (one step of lag) and move that reference only when the running maximum has grown by
// more than 2^8 -- probabilities stay far inside the exponent range of bf16 / fp32 and the result is exact after
// the final division by the sum.  Because nothing about a step's scores has to be known before its first
// exponential, S streams from TMEM in 32-column chunks and loads, maxima and MUFU work overlap.
//
// What bounds the kernel (profiles/README.md, round 2): the arithmetic of a step -- 128 x (FFMA, MUFU.EX2, FADD),
// 64 F2FP, 64 FMNMX3 per row -- runs at the MUFU rate (16 ex2 / clk / SM; tools/micro/exp_mix_bench.cu), i.e.
// 2 x 1 045 clocks per K/V step for the two softmax warps of a scheduler.  The kernel needs about 2 500 per step
// plus 1 400 per unit boundary (clock64 trace: -DSFM_ATT_TRACE, tools/att_trace.py): the remainder is per-warp serial
// latency (first TMEM chunk, P hand-off) that two warps per scheduler cannot fully hide from each other.  Measured and NOT helpful: a start offset between the
// tiles (any offset random-walks away), strict MUFU turn-taking between the paired warps (their bubbles end up
// serialised), PRMT instead of F2FP packing, separate issue threads alone.
#include "tc_common.cuh"
#include "tc_path.cuh"
#include <stdlib.h>
#include <algorithm>

namespace sfm {
namespace tc {

namespace {

constexpr int KV_STAGES = 3;

#ifdef SFM_ATT_TRACE
// debug build only (tools/att_trace.py): clock64 time stamps of CTA 0 for K/V steps [TRACE_G0, TRACE_G0 + TRACE_NG)
constexpr int TRACE_G0 = 64, TRACE_NG = 24;
__device__ long long g_att_trace[TRACE_NG * 64];
#define ATT_TRACE(g, ev)                                                                              \
  do {                                                                                                \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (g) >= TRACE_G0 && (g) < TRACE_G0 + TRACE_NG)   \
      g_att_trace[((g) - TRACE_G0) * 64 + (ev)] = clock64();                                          \
  } while (0)
#else
#define ATT_TRACE(g, ev) do { } while (0)
#endif
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, 128-byte swizzled rows
// warp 0: TMA producer; warps 1..NT: one MMA issuer per query tile; then NT softmax warpgroups
// Softmax warpgroups first (warps 0 .. 4 NT - 1: warpgroup t = query tile t, TMEM quadrant = warp % 4), then the TMA
// producer and the NT issuers.
template <int NT> struct AttCfg {
  static constexpr int SW0 = 0;
  static constexpr int PRODUCER = 4 * NT, ISSUER0 = 4 * NT + 1;
  static constexpr int THREADS = (1 + NT) * 32 + NT * 128;
};
// (`setmaxnreg` cannot move registers to the softmax warpgroups in this layout: the producer / issuer warps form a
// partial warpgroup of three and the instruction never returns there; a 12th warp to complete it costs 8 %.)

__device__ __forceinline__ void att_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Unit epilogue of one softmax thread (out of line: its 64 + 32 live registers must not weigh on the allocation of
// the K/V loop): O row from TMEM -> 1 / l -> bf16.  A whole tile leaves as ONE coalesced TMA store from a swizzled
// staging tile: a thread owns a row, so direct stores touch 32 different 128-byte lines per instruction and the LSU
// back-pressure keeps the tile's warps away from the MUFU for ~1 700 clocks per unit.
__device__ __noinline__ void att_epilogue(uint32_t o_addr, float inv, bool whole, bool row_ok, int t, int wg_tid,
                                          int row_in_tile, uint32_t stage, const CUtensorMap* map_out, int col0,
                                          long long row0, __nv_bfloat16* out) {
  uint32_t o[64];
  tmem_ld_32x32(o_addr, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
  tmem_ld_32x32(o_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
  tmem_wait_ld();
  tc_fence_before();   // O_t is in registers: the next unit's first P V may overwrite it
  auto chunk = [&](int i) {   // 8 output channels of this row as bf16
    return make_uint4(pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                      pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                      pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                      pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
  };
  if (whole) {
    if (wg_tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous unit's store has left smem
    att_bar_sync(1 + t, 128);
    const uint32_t row_base = stage + row_in_tile * 128;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint4 v = chunk(i);
      st_shared_v4(row_base + ((i ^ (row_in_tile & 7)) << 4), v.x, v.y, v.z, v.w);
    }
    fence_proxy_async();
    att_bar_sync(1 + t, 128);
    if (wg_tid == 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(map_out)),
                   "r"(stage), "r"(col0), "r"((int)row0)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  } else if (row_ok) {   // ragged tile: the rows past Nq belong to another pair
    uint4* d4 = reinterpret_cast<uint4*>(out + (row0 + row_in_tile) * 256 + col0);
#pragma unroll
    for (int i = 0; i < 8; i++) d4[i] = chunk(i);
  }
}

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
  if ((qts & (qts - 1)) == 0) {   // the usual case (2048 keypoints: 8 query blocks); a division costs ~250 clocks per unit
    w.qblk = (u >> 2) & (qts - 1);
    w.b = (u >> 2) >> (31 - __clz(qts));
  } else {
    w.qblk = (u >> 2) % qts;
    w.b = (u >> 2) / qts;
  }
  w.Nq = p.Nq[w.side];
  w.Nk = p.Nk[w.side];
  w.qrow = p.q_row0[w.side] + (long long)w.b * w.Nq + w.qblk * (NT * 128);
  w.krow = p.k_row0[w.side] + (long long)w.b * w.Nk;
  w.n_kv = (w.Nk + 127) / 128;
  return w;
}

template <int NT>
__global__ void __launch_bounds__(AttCfg<NT>::THREADS, 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map_out, TcAttn p,
                    int units0, int qt0, int qt1, int n_units) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                             // 2 buffers x NT tiles
  uint8_t* sK = smem + 2 * NT * TILE_BYTES;       // KV_STAGES tiles
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;      // KV_STAGES tiles
  uint8_t* sO = sV + KV_STAGES * TILE_BYTES;      // NT output staging tiles (128 rows x 128 B, 128-byte swizzle)
  AttBars* bars = reinterpret_cast<AttBars*>(sO + NT * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == AttCfg<NT>::PRODUCER && lane == 0) {
    tma_prefetch_desc(&map);
    tma_prefetch_desc(&map_out);
    for (int q = 0; q < 2; q++) {
      mbar_init(&bars->q_full[q], 1);
      mbar_init(&bars->q_empty[q], NT);
    }
    for (int s = 0; s < KV_STAGES; s++) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], NT);
    }
    for (int t = 0; t < NT; t++) {
      mbar_init(&bars->s_full[t], 1);
      mbar_init(&bars->p_full[t], 4);
      mbar_init(&bars->o_full[t], 1);
      mbar_init(&bars->s_free[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == AttCfg<NT>::ISSUER0) tmem_alloc<NT * 256>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  constexpr uint32_t COL_S = 0, COL_O = NT * 128, COL_P = NT * 128 + NT * 64;

  // Every barrier of the S / P / O hand-shake completes exactly once per K/V step, so one CTA-wide step
  // counter g (running across units) gives all parities; the K/V ring stage is g % KV_STAGES.
  if (warp == AttCfg<NT>::PRODUCER) {
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
  } else if (warp > AttCfg<NT>::PRODUCER) {
    // ------------------------------------------------------------------ MMA issuers: ONE PER QUERY TILE
    // Each tile's chain  S_t(g+1) <- s_free_t(g),  P V_t(g) <- p_full_t(g)  is driven by its own thread, so a tile
    // never waits behind a barrier of the other tile (a single in-order issuer couples the two softmax groups:
    // its four waits per step form one cycle, and whichever group runs late stalls the other's P V).  The two
    // tiles share only the K/V ring and the Q buffers, released when BOTH issuers' MMAs on them have completed
    // (kv_empty / q_empty count NT; tcgen05.commit tracks the issuing thread's own MMAs).
    const int t = warp - AttCfg<NT>::ISSUER0;
    // The WHOLE warp runs this loop and one elected lane issues: descriptors, TMEM addresses and barrier addresses then
    // live in uniform registers.  Under `if (lane == 0)` they are per-thread values and every tcgen05.mma is wrapped
    // in an ELECT / R2UR.BROADCAST / branch sequence (~85 clocks per MMA: the eight P V MMAs of a step took 700
    // clocks to issue, which put S (g+1) at 70 % of step g).
    if (blockIdx.x < n_units) {
      const bool leader = elect_one_sync();
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);  // S = Q K^T : B = K tile, K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);   // O = P V   : B = V tile, MN-major
      const uint64_t q_desc0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV), 1024, 1024);
      // S_t of the K/V tile in ring slot `stage`, Q buffer qb
      auto issue_s = [&](int qb, int stage) {
        const uint64_t ad = umma_desc_advance(q_desc0, (qb * NT + t) * TILE_BYTES);
        const uint64_t bd = umma_desc_advance(k_desc0, stage * TILE_BYTES);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_ss(tmem + COL_S + t * 128, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc_s,
                    k != 0);
          umma_commit(&bars->s_full[t]);
        }
      };
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
      Cursor cur{(int)blockIdx.x, 0, 0, decode_unit<NT>(p, blockIdx.x, units0, qt0, qt1).n_kv, true};
      // very first S of this tile
      mbar_wait(&bars->q_full[0], 0);
      mbar_wait(&bars->kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      for (int g = 0; cur.valid; g++) {
        Cursor nxt = cur;
        advance(nxt);
        const bool cur_first = cur.j == 0, cur_last = cur.j + 1 == cur.n_kv;
        ATT_TRACE(g, 48 + t * 6 + 0);
        if (nxt.valid) {
          // S of step g+1 (next K/V tile of this unit, or tile 0 of the next unit with its own Q buffer): issued as
          // soon as the softmax group holds S_t(g) in registers, so it overlaps the exponentials of step g
          const int s1 = (g + 1) % KV_STAGES;
          if (nxt.j == 0) mbar_wait(&bars->q_full[nxt.it & 1], (nxt.it >> 1) & 1);
          mbar_wait(&bars->kv_full[s1], ((g + 1) / KV_STAGES) & 1);
          ATT_TRACE(g, 48 + t * 6 + 1);
          mbar_wait(&bars->s_free[t], g & 1);
          ATT_TRACE(g, 48 + t * 6 + 2);
          tc_fence_after();
          issue_s(nxt.it & 1, s1);
        }
        // every S MMA of this tile that reads this unit's Q buffer has been issued
        if (cur_last && leader) umma_commit(&bars->q_empty[cur.it & 1]);
        mbar_wait(&bars->p_full[t], g & 1);
        ATT_TRACE(g, 48 + t * 6 + 3);
        tc_fence_after();
        const uint64_t vd = umma_desc_advance(v_desc0, (g % KV_STAGES) * TILE_BYTES);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; k++)  // 16 keys per step: 16 V rows = 2048 bytes, P advances 8 columns
            umma_ts(tmem + COL_O + t * 64, tmem + COL_P + t * 64 + k * 8, umma_desc_advance(vd, k * 2048), idesc_o,
                    (cur_first ? 0 : 1) | k);
          ATT_TRACE(g, 48 + t * 6 + 4);
          umma_commit(&bars->o_full[t]);
          umma_commit(&bars->kv_empty[g % KV_STAGES]);   // this tile is done with K/V slot g (S_t(g) and P V_t(g))
          ATT_TRACE(g, 48 + t * 6 + 5);
        }
        __syncwarp();
        cur = nxt;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = (warp - AttCfg<NT>::SW0) >> 2;  // query tile of this warpgroup
    const int quad = warp & 3;      // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const float c = 0.125f * 1.4426950408889634f;  // log2(e) / sqrt(64)
    int g = 0;
    // tile 1 starts after tile 0 has produced its first probability tile: half a step of lag
    if (NT > 1 && t == 1 && p.stagger) mbar_wait(&bars->p_full[0], 0);
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
    const Unit w = decode_unit<NT>(p, u, units0, qt0, qt1);
    const int Nk = w.Nk, n_kv = w.n_kv;
    const int qi = w.qblk * (NT * 128) + t * 128 + row_in_tile;  // query index inside the pair
    // Reference maximum m_ref: the exponentials of step j are taken relative to the maximum of the steps BEFORE
    // it (one step of lag; step 0 uses the maximum of its first 32 keys).  The row maximum of step j is collected
    // in passing, off the critical path, and only moves m_ref for the next step (when it grew by more than 2^8);
    // the matching rescale of O / the running sum happens at the next step, once P V (j) has landed.  Nothing has
    // to be known about a step's scores before its first exponential, so S streams in 32-column chunks and the
    // loads, the maxima and the MUFU work overlap.  bf16 P and the fp32 accumulators have the exponent range for
    // any realistic growth inside one step; a row whose scores jump by more than 2^90 within a step (no finite
    // softmax input does that in practice) is flagged and recomputed exactly by att_redo_kernel.
    float m_ref = 0.f, l_run = 0.f, alpha_p = 1.f;
    bool pend = false, any_pend = false, bad = false;

    for (int j = 0; j < n_kv; j++, g++) {
      ATT_TRACE(g, (warp - AttCfg<NT>::SW0) * 6 + 0);
      mbar_wait(&bars->s_full[t], g & 1);
      ATT_TRACE(g, (warp - AttCfg<NT>::SW0) * 6 + 1);
      tc_fence_after();
      const int valid = min(128, Nk - j * 128);  // columns of this tile that are real keys
      const bool waited = any_pend;
      if (any_pend) {   // warp-uniform, rare: the reference maximum moved after the previous step -> rescale O now
        mbar_wait(&bars->o_full[t], (g - 1) & 1);     // P V (j-1) has landed
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + COL_O + t * 64 + ch * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha_p);
          tmem_st_32x32(lane_addr + COL_O + t * 64 + ch * 32, r);
        }
        tmem_wait_st();
      }
      uint32_t sv[128];
      tmem_ld_32x32(lane_addr + COL_S + t * 128, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
      tmem_wait_ld();
#pragma unroll
      for (int ch = 1; ch < 4; ch++)     // in flight while chunk 0 is exponentiated
        tmem_ld_32x32(lane_addr + COL_S + t * 128 + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[ch * 32]));
      if (valid < 128) {                 // ragged last K/V tile (warp-uniform, rare): mask the missing keys
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 128; i++)
          if (i >= valid) sv[i] = 0xff800000u;  // -inf
      }
      // maximum of 32 scores: 4 chains of three-input maxima
      auto max32 = [&](int base) {
        float a0 = fmax3(__uint_as_float(sv[base]), __uint_as_float(sv[base + 1]), __uint_as_float(sv[base + 2]));
        float a1 = fmax3(__uint_as_float(sv[base + 3]), __uint_as_float(sv[base + 4]), __uint_as_float(sv[base + 5]));
        float a2 = fmax3(__uint_as_float(sv[base + 6]), __uint_as_float(sv[base + 7]), __uint_as_float(sv[base + 8]));
        float a3 = fmax3(__uint_as_float(sv[base + 9]), __uint_as_float(sv[base + 10]), __uint_as_float(sv[base + 11]));
#pragma unroll
        for (int i = 12; i < 28; i += 8) {
          a0 = fmax3(a0, __uint_as_float(sv[base + i]), __uint_as_float(sv[base + i + 1]));
          a1 = fmax3(a1, __uint_as_float(sv[base + i + 2]), __uint_as_float(sv[base + i + 3]));
          a2 = fmax3(a2, __uint_as_float(sv[base + i + 4]), __uint_as_float(sv[base + i + 5]));
          a3 = fmax3(a3, __uint_as_float(sv[base + i + 6]), __uint_as_float(sv[base + i + 7]));
        }
        a0 = fmax3(a0, __uint_as_float(sv[base + 28]), __uint_as_float(sv[base + 29]));
        a1 = fmax3(a1, __uint_as_float(sv[base + 30]), __uint_as_float(sv[base + 31]));
        return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
      };
      float mx = max32(0);
      if (j == 0) m_ref = mx;
      const float mc = m_ref * c;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
      // 32 exponentials of chunk `ch` -> 16 packed bf16 pairs at pk[(ch & 1) * 16 ...]
      auto exp32 = [&](int ch) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int e = ch * 32 + i;
          const float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), c, -mc));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), c, -mc));
          const float p2 = fast_exp2(fmaf(__uint_as_float(sv[e + 2]), c, -mc));
          const float p3 = fast_exp2(fmaf(__uint_as_float(sv[e + 3]), c, -mc));
          ls[0] += p0; ls[1] += p1; ls[2] += p2; ls[3] += p3;
          pk[(ch & 1) * 16 + (i >> 1)] = pack_bf16(p0, p1);
          pk[(ch & 1) * 16 + (i >> 1) + 1] = pack_bf16(p2, p3);
        }
      };
      exp32(0);
      tmem_wait_ld();                  // chunks 1-3 have landed behind the first 32 exponentials
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free[t]);   // S_t is in registers: the tensor pipe may overwrite it
      ATT_TRACE(g, (warp - AttCfg<NT>::SW0) * 6 + 2);
      exp32(1);
      mx = fmax3(mx, max32(32), max32(64));
      mx = fmaxf(mx, max32(96));
      if (j > 0 && !waited) {
        // P V (j-1) must have consumed P_t before it is overwritten; by now that MMA (issue + 256 tensor clocks +
        // commit) has had 64 exponentials of time
        ATT_TRACE(g, (warp - AttCfg<NT>::SW0) * 6 + 3);
        mbar_wait(&bars->o_full[t], (g - 1) & 1);
        ATT_TRACE(g, (warp - AttCfg<NT>::SW0) * 6 + 4);
        tc_fence_after();
      }
      tmem_st_32x32(lane_addr + COL_P + t * 64, pk);
      exp32(2);
      exp32(3);
      tmem_st_32x32(lane_addr + COL_P + t * 64 + 32, pk);
      l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      // the reference for the NEXT step (not after the last one: the final division wants O and the sum in one frame)
      const float growth = (mx - m_ref) * c;
      bad = bad || growth > 90.f;
      pend = (j + 1 < n_kv) && growth > 8.0f;
      alpha_p = 1.f;
      if (pend) {
        alpha_p = fast_exp2(-growth);
        m_ref = mx;
        l_run *= alpha_p;
      }
      any_pend = __any_sync(0xffffffffu, pend);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[t]);
      ATT_TRACE(g, (warp - AttCfg<NT>::SW0) * 6 + 5);
    }
    mbar_wait(&bars->o_full[t], (g - 1) & 1);
    ATT_TRACE(g, (warp - AttCfg<NT>::SW0) * 6 + 3);   // (unit epilogue: the next step has no P V wait, its slots are free)
    tc_fence_after();
    {
      const bool whole = p.tma_out && w.qblk * (NT * 128) + t * 128 + 128 <= w.Nq;
      att_epilogue(lane_addr + COL_O + t * 64, 1.f / l_run, whole, qi < w.Nq, t, (warp - AttCfg<NT>::SW0 - 4 * t) * 32 + lane,
                   row_in_tile, smem_u32(sO + t * TILE_BYTES), &map_out, w.head * 64, w.qrow + t * 128, p.out);
      ATT_TRACE(g, (warp - AttCfg<NT>::SW0) * 6 + 4);
      if (qi < w.Nq && bad && p.redo != nullptr) {   // (practically never) hand the row to the exact recomputation
        const int slot = atomicAdd(p.redo, 1);
        if (slot < p.redo_cap)
          reinterpret_cast<int4*>(p.redo + 4)[slot] =
              make_int4((int)(w.qrow + t * 128 + row_in_tile), w.head, (int)w.krow, Nk);
      }
    }
    }  // units
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // (only the storing threads have groups in flight)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AttCfg<NT>::ISSUER0) tmem_dealloc<NT * 256>(tmem);
}

// Exact recomputation of the rows the main kernel flagged (a score jump of more than 2^90 inside one K/V step; see
// the softmax warpgroups): plain online softmax in fp32, one warp per (query row, head), lane = two head dims.
// One block; with an empty list -- the normal case -- it returns at once.  Leaves the list empty.
__global__ void __launch_bounds__(256) att_redo_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                       __nv_bfloat16* __restrict__ out, int* __restrict__ redo, int cap) {
  const int n = min(redo[0], cap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float c = 0.125f * 1.4426950408889634f;
  for (int e = warp; e < n; e += 8) {
    const int4 ent = reinterpret_cast<const int4*>(redo + 4)[e];
    const long long qrow = ent.x, krow = ent.z;
    const int head = ent.y, Nk = ent.w;
    const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(qkv + qrow * 768 + head * 64 + 2 * lane);
    const float qx = __bfloat162float(q2.x), qy = __bfloat162float(q2.y);
    float m = -INFINITY, l = 0.f, ox = 0.f, oy = 0.f;
    for (int j = 0; j < Nk; j++) {
      const __nv_bfloat16* kr = qkv + (krow + j) * 768 + 256 + head * 64 + 2 * lane;
      const __nv_bfloat162 k2 = *reinterpret_cast<const __nv_bfloat162*>(kr);
      const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(kr + 256);
      const float s = warp_sum(qx * __bfloat162float(k2.x) + qy * __bfloat162float(k2.y)) * c;
      const float mn = fmaxf(m, s);
      const float scale = exp2f(m - mn), pj = exp2f(s - mn);
      l = l * scale + pj;
      ox = ox * scale + pj * __bfloat162float(v2.x);
      oy = oy * scale + pj * __bfloat162float(v2.y);
      m = mn;
    }
    *reinterpret_cast<__nv_bfloat162*>(out + qrow * 256 + head * 64 + 2 * lane) = __floats2bfloat162_rn(ox / l, oy / l);
  }
  __syncthreads();
  if (threadIdx.x == 0) redo[0] = 0;
}

template <int NT>
int launch_att(const TcAttn& p, const CUtensorMap& map, const CUtensorMap& map_out, int num_sms, cudaStream_t st) {
  const int qt0 = cdiv(p.Nq[0], NT * 128), qt1 = p.Nq[1] > 0 ? cdiv(p.Nq[1], NT * 128) : 0;
  const int units0 = p.B * 4 * qt0, units1 = p.B * 4 * qt1;
  const int n_units = units0 + units1;
  const size_t smem = (size_t)(3 * NT + 2 * KV_STAGES) * TILE_BYTES + sizeof(AttBars) + 1024;
  SFM_SMEM_OPT_IN(tc_attention_kernel<NT>, smem);
  const int grid = std::min(n_units, num_sms);
  tc_attention_kernel<NT><<<grid, AttCfg<NT>::THREADS, smem, st>>>(map, map_out, p, units0, qt0, qt1 > 0 ? qt1 : 1, n_units);
  SFM_LAUNCH_CHECK();
  if (p.redo != nullptr) {
    att_redo_kernel<<<1, 256, 0, st>>>(p.qkv, p.out, p.redo, p.redo_cap);
    SFM_LAUNCH_CHECK();
  }
  return SFM_OK;
}

}  // namespace

int tc_attention(const TcAttn& p, int num_sms, cudaStream_t st) {
  SFM_REQUIRE(p.B > 0 && p.Nq[0] > 0 && p.Nk[0] > 0 && p.Nq[1] >= 0, "tc_attention: bad dims");
  CUtensorMap map;
  SFM_TRY(get_tmap_bf16(p.tmaps, &map, p.qkv, p.rows_total, 768, 768, 128));
  CUtensorMap map_out;
  SFM_TRY(get_tmap_bf16(p.tmaps, &map_out, p.out, p.rows_total, 256, 256, 128));
  // tuning knobs are read from the environment once per process (read-only afterwards)
  static const int nt = getenv("SFM_ATT_NT") ? atoi(getenv("SFM_ATT_NT")) : 2;
  static const int stagger = getenv("SFM_ATT_STAGGER") ? atoi(getenv("SFM_ATT_STAGGER")) : 1;
  static const int tma_out = getenv("SFM_ATT_TMA_OUT") ? atoi(getenv("SFM_ATT_TMA_OUT")) : 1;
  TcAttn q = p;
  q.stagger = stagger;
  q.tma_out = tma_out;
  SFM_REQUIRE(num_sms > 0, "tc_attention: bad SM count");
  if (nt == 1) return launch_att<1>(q, map, map_out, num_sms, st);
  return launch_att<2>(q, map, map_out, num_sms, st);
}

}  // namespace tc
}  // namespace sfm

#ifdef SFM_ATT_TRACE
extern "C" __attribute__((visibility("default"))) int sfm_debug_att_trace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, sfm::tc::g_att_trace, sizeof(long long) * n);
}
#endif
