// SuperGlue.forward orchestration (reference simple_sfm/models/superglue.py:235-290).
// Token-major activations: X [T0 + T1, 256], side-0 tokens (b*K0 + k) first, then side 1.
#include <algorithm>
#include "context.cuh"
#include "gemm_f32.cuh"
#include "sg_kernels.cuh"
#include "bn_kernels.cuh"
#include "tc_path.cuh"
#include <stdlib.h>

namespace sfm {

static bool sk_use_k16() {
  static const bool on = getenv("SFM_SK_FP32") == nullptr;
  return on;
}


namespace {

struct SgBuffers {
  float *X, *xk, *kin, *t0, *t1, *qkv, *msgp, *msg, *hid, *att, *bn_part, *bn_scale, *bn_shift, *md, *sk_fused, *sk_kb, *sk_k16f, *scores, *u, *v, *part, *max0, *max1;
  int *counters, *idx0, *idx1;
  int R, ldS;
};

int carve(Arena& a, const SgCall& c, SgBuffers& b) {
  const long long T0 = (long long)c.B * c.K0, T1 = (long long)c.B * c.K1, T = T0 + T1;
  const int Kmax = c.K0 > c.K1 ? c.K0 : c.K1;
  b.ldS = (Kmax + 3) / 4 * 4;
  b.X = a.get<float>(T * 256);
  b.xk = a.get<float>(T * 256);
  b.kin = a.get<float>(T * 4);
  b.t0 = a.get<float>(T * 256);   // keypoint-encoder ping-pong (fp32 SIMT on both precision paths)
  b.t1 = a.get<float>(T * 256);
  const int G = c.groups > 0 ? c.groups : 1;
  // BatchNorm partial moments: 64-row blocks of the stand-alone kernels, or 32-row blocks written by the
  // tcgen05 GEMM epilogue (tc_path.cu)
  b.bn_part = a.get<float>(std::max((size_t)bn_stat_blocks(BnGroups{T0, T1, G}), (size_t)((T0 + T1) / 32 + 1)) * 512 * 2);
  b.bn_scale = a.get<float>((size_t)2 * G * 512);
  b.bn_shift = a.get<float>((size_t)2 * G * 512);
  if (c.precision == 0) {   // fp32-path activations
    b.qkv = a.get<float>(T * 768);
    b.msgp = a.get<float>(T * 256);
    b.msg = a.get<float>(T * 256);
    b.hid = a.get<float>(T * 512);
    b.att = a.get<float>((size_t)c.B * 4 * Kmax * b.ldS);
    b.md = a.get<float>(T * 256);
  }
  b.scores = a.get<float>((size_t)c.B * c.K0 * c.K1);
  b.u = a.get<float>((size_t)c.B * (c.K0 + 1));
  b.v = a.get<float>((size_t)c.B * (c.K1 + 1));
  b.R = 1;
  b.part = c.K1 > 0 ? a.get<float>(sinkhorn_partial_floats(c.B, c.K0, c.K1, &b.R)) : nullptr;   // (K1 == 0: hoisting pass)
  b.counters = a.get<int>(sinkhorn_counter_ints(c.B, c.K1));
  b.sk_fused = c.precision == 1 ? a.get<float>(sinkhorn_fused_floats(c.B, c.K0, c.K1)) : nullptr;
  b.sk_kb = c.precision == 1 ? a.get<float>((size_t)c.B * c.K0) : nullptr;
  // fp16 copy of the scaling-form Sinkhorn matrix (bf16 precision mode): [B, K0, K1] halves
  b.sk_k16f = (c.precision == 1 && sk_use_k16()) ? a.get<float>(((size_t)c.B * c.K0 * c.K1 + 1) / 2) : nullptr;
  b.max0 = a.get<float>((size_t)c.B * c.K0);
  b.max1 = a.get<float>((size_t)c.B * c.K1);
  b.idx0 = a.get<int>((size_t)c.B * c.K0);
  b.idx1 = a.get<int>((size_t)c.B * c.K1);
  return SFM_OK;
}

}  // namespace

// y = [relu(bn(]x W^T + b[))]  over all tokens of both sides
int linear_f32_raw(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int K, int N,
               const float* bias, float* C, int ldc, long long T, int flags, cudaStream_t st) {
  GemmF32 g;
  g.A = A; g.lda = lda; g.A2 = A2; g.lda2 = lda2; g.K1 = K1;
  g.B = W; g.ldb = K; g.C = C; g.ldc = ldc; g.bias = bias;
  g.M = (int)T; g.N = N; g.K = K; g.flags = flags;
  return gemm_f32_nt(g, st);
}

// message = softmax(q k^T / 8) v for the queries of one side against the keys/values of `src`
int attention_f32_raw(const SgBuffers& b, int B, long long q_row0, int Nq, long long s_row0, int Nk, cudaStream_t st) {
  GemmF32 g;
  g.A = b.qkv + q_row0 * 768; g.lda = 768;
  g.B = b.qkv + s_row0 * 768 + 256; g.ldb = 768;
  g.C = b.att; g.ldc = b.ldS;
  g.M = Nq; g.N = Nk; g.K = 64; g.alpha = 0.125f;
  g.batch = B * 4; g.nb1 = 4;
  g.sA0 = (long long)Nq * 768; g.sA1 = 64;
  g.sB0 = (long long)Nk * 768; g.sB1 = 64;
  g.sC0 = 4ll * Nq * b.ldS; g.sC1 = (long long)Nq * b.ldS;
  SFM_TRY(gemm_f32_nt(g, st));
  SFM_TRY(softmax_rows_f32(b.att, (long long)B * 4 * Nq, Nk, b.ldS, st));
  GemmF32 p;
  p.A = b.att; p.lda = b.ldS;
  p.B = b.qkv + s_row0 * 768 + 512; p.ldb = 768;
  p.C = b.msgp + q_row0 * 256; p.ldc = 256;
  p.M = Nq; p.N = 64; p.K = Nk;
  p.batch = B * 4; p.nb1 = 4;
  p.sA0 = 4ll * Nq * b.ldS; p.sA1 = (long long)Nq * b.ldS;
  p.sB0 = (long long)Nk * 768; p.sB1 = 64;
  p.sC0 = (long long)Nq * 256; p.sC1 = 64;
  return gemm_f32_nn(p, st);
}

// keypoint encoder (superglue.py:72-82, 254-255): X += kenc(kin); fp32 SIMT on both precision paths
int kenc_forward_f32(sfm_ctx* h, const SgCall& c, float* X, const float* kin, float* t0, float* t1, float* part,
                     float* scale, float* shift, cudaStream_t st) {
  const SgWeights& w = h->sg;
  const long long T0 = (long long)c.B * c.K0, T1 = (long long)c.B * c.K1, T = T0 + T1;
  const bool train = c.bn_mode == 0;
  const float* cur = kin;
  int ld = 4;
  float* pp[2] = {t0, t1};
  for (int i = 0; i < 5; i++) {
    const int cin = w.kch[i], cout = w.kch[i + 1];
    const bool last = i == 4;
    if (last) {
      SFM_TRY(linear_f32_raw(cur, ld, nullptr, 0, 0, w.kW[i], cin, cout, w.kb[i], X, 256, T, GEMM_ACCUM, st));
    } else if (train) {
      float* out = pp[i & 1];
      SFM_TRY(linear_f32_raw(cur, ld, nullptr, 0, 0, w.kW[i], cin, cout, w.kb[i], out, cout, T, 0, st));
      const BnGroups gr{T0, T1, c.groups > 0 ? c.groups : 1};
      SFM_TRY(bn_train_stats<float>(out, cout, cout, gr, w.kbn[i].gamma, w.kbn[i].beta, 1e-5f, part, scale, shift, st));
      SFM_TRY(bn_apply_relu<float>(out, cout, cout, gr, scale, shift, st));
      cur = out;
      ld = cout;
    } else {
      float* out = pp[i & 1];
      SFM_TRY(linear_f32_raw(cur, ld, nullptr, 0, 0, w.kWf[i], cin, cout, w.kbf[i], out, cout, T, GEMM_RELU, st));
      cur = out;
      ld = cout;
    }
  }
  return SFM_OK;
}

namespace {

int forward_f32(sfm_ctx* h, const SgCall& c, const SgBuffers& b, cudaStream_t st) {
  const SgWeights& w = h->sg;
  auto linear_f32 = [&](auto... args) {
    ProfScope ps(h, SFM_PROF_LINEAR, st);
    return sfm::linear_f32_raw(args...);
  };
  auto attention_f32 = [&](auto... args) {
    ProfScope ps(h, SFM_PROF_ATTENTION, st);
    return sfm::attention_f32_raw(args...);
  };
  const long long T0 = (long long)c.B * c.K0, T1 = (long long)c.B * c.K1, T = T0 + T1;
  const bool train = c.bn_mode == 0;
  {
    ProfScope ps(h, SFM_PROF_OTHER, st);
    SFM_TRY(kenc_forward_f32(h, c, b.X, b.kin, b.t0, b.t1, b.bn_part, b.bn_scale, b.bn_shift, st));
  }
  SFM_CUDA(cudaMemcpyAsync(b.xk, b.X, T * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));  // parity tap
  // ---- attentional GNN (superglue.py:123-139)
  for (int l = 0; l < 18; l++) {
    const SgLayer& L = w.L[l];
    const bool cross = l & 1;
    SFM_TRY(linear_f32(b.X, 256, nullptr, 0, 0, L.Wqkv, 256, 768, L.bqkv, b.qkv, 768, T, 0, st));
    SFM_TRY(attention_f32(b, c.B, 0, c.K0, cross ? T0 : 0, cross ? c.K1 : c.K0, st));
    SFM_TRY(attention_f32(b, c.B, T0, c.K1, cross ? 0 : T0, cross ? c.K0 : c.K1, st));
    SFM_TRY(linear_f32(b.msgp, 256, nullptr, 0, 0, L.Wm, 256, 256, L.bm, b.msg, 256, T, 0, st));
    if (train) {
      SFM_TRY(linear_f32(b.X, 256, b.msg, 256, 256, L.W1, 512, 512, L.b1, b.hid, 512, T, 0, st));
      ProfScope ps(h, SFM_PROF_OTHER, st);
      const BnGroups gr{T0, T1, c.groups > 0 ? c.groups : 1};
      SFM_TRY(bn_train_stats<float>(b.hid, 512, 512, gr, L.bn1.gamma, L.bn1.beta, 1e-5f, b.bn_part, b.bn_scale,
                                    b.bn_shift, st));
      SFM_TRY(bn_apply_relu<float>(b.hid, 512, 512, gr, b.bn_scale, b.bn_shift, st));
    } else {
      SFM_TRY(linear_f32(b.X, 256, b.msg, 256, 256, L.W1f, 512, 512, L.b1f, b.hid, 512, T, GEMM_RELU, st));
    }
    SFM_TRY(linear_f32(b.hid, 512, nullptr, 0, 0, L.W2, 512, 256, L.b2, b.X, 256, T, GEMM_ACCUM, st));
  }
  // ---- final projection + score matrix (superglue.py:261-265)
  SFM_TRY(linear_f32(b.X, 256, nullptr, 0, 0, w.Wf, 256, 256, w.bf, b.md, 256, T, 0, st));
  GemmF32 g;
  g.A = b.md; g.lda = 256;
  g.B = b.md + T0 * 256; g.ldb = 256;
  g.C = b.scores; g.ldc = c.K1;
  g.M = c.K0; g.N = c.K1; g.K = 256; g.alpha = 1.0f / 16.0f;
  g.batch = c.B; g.nb1 = 1;
  g.sA0 = (long long)c.K0 * 256; g.sB0 = (long long)c.K1 * 256; g.sC0 = (long long)c.K0 * c.K1;
  return gemm_f32_nt(g, st);
}

}  // namespace

int sg_forward(sfm_ctx* h, const SgCall& c, void* ws, size_t ws_bytes, bool dry, size_t* need, int tap_which,
               float** tap_ptr, size_t* tap_floats, cudaStream_t st) {
  const bool hoist_pass = c.last_layer < 18;
  SFM_REQUIRE(c.B >= 0 && c.K0 > 0 && (c.K1 > 0 || (hoist_pass && c.K1 == 0)),
              "superglue: B >= 0 and K0, K1 > 0 required (B=%d K0=%d K1=%d)", c.B, c.K0, c.K1);
  SFM_REQUIRE((c.first_layer == 0 && c.last_layer == 18) || (c.precision == 1 && c.bn_mode == 1),
              "superglue: layer-0 hoisting exists for the bf16 path with eval-mode BatchNorm only");
  SFM_REQUIRE(c.first_layer >= 0 && c.first_layer <= 1 && c.last_layer >= 1 && c.last_layer <= 18 &&
                  (c.first_layer == 0 || (c.xs0 && c.xs1 && c.xs_k0 >= c.K0 && c.xs_k1 >= c.K1)) &&
                  (!hoist_pass || c.x_out), "superglue: bad hoisting arguments");
  SFM_REQUIRE(c.precision == 0 || c.precision == 1, "superglue: unknown precision %d", c.precision);
  SFM_REQUIRE(c.bn_mode == 0 || c.bn_mode == 1, "superglue: unknown bn_mode %d", c.bn_mode);
  SFM_REQUIRE(c.groups >= 0 && (c.groups <= 1 || c.B % c.groups == 0), "superglue: B=%d is not divisible into %d chunk groups", c.B, c.groups);
  Arena a(dry ? nullptr : ws, dry ? 0 : ws_bytes);
  SgBuffers b;
  carve(a, c, b);
  TcSgBuffers tb;
  if (c.precision == 1) tc_sg_carve(a, c, tb);
  if (need) *need = a.off;
  const long long T = (long long)c.B * (c.K0 + c.K1);
  if (tap_ptr) {
    // offsets are valid for any base pointer: recompute with the caller's workspace
    Arena a2(ws, ws_bytes);
    SgBuffers b2;
    carve(a2, c, b2);
    if (a2.overflow) {
      set_error("superglue tap: workspace too small");
      return SFM_ERR_WORKSPACE;
    }
    switch (tap_which) {
      case 0: *tap_ptr = b2.xk; *tap_floats = T * 256; break;   // snapshot after kenc (fp32 path)
      case 1: *tap_ptr = b2.X; *tap_floats = T * 256; break;
      case 2: *tap_ptr = b2.scores; *tap_floats = (size_t)c.B * c.K0 * c.K1; break;
      case 3: *tap_ptr = b2.u; *tap_floats = (size_t)c.B * (c.K0 + 1); break;
      case 4: *tap_ptr = b2.v; *tap_floats = (size_t)c.B * (c.K1 + 1); break;
      default: set_error("superglue tap: unknown tap %d", tap_which); return SFM_ERR_INVALID;
    }
    return SFM_OK;
  }
  if (dry) return SFM_OK;
  if (a.overflow) {
    set_error("superglue: workspace too small (%zu bytes given, %zu needed)", ws_bytes, a.off);
    return SFM_ERR_WORKSPACE;
  }
  SFM_REQUIRE(h != nullptr && h->sg.loaded, "superglue: weights not loaded (call sfm_load_superglue)");
  if (c.B == 0) return SFM_OK;
  const long long T0 = (long long)c.B * c.K0;
  if (c.K1 > 0) SFM_CUDA(cudaMemsetAsync(b.counters, 0, sinkhorn_counter_ints(c.B, c.K1) * sizeof(int), st));
  if (c.first_layer > 0) {
    SFM_TRY(gather_tokens(c.xs0, c.xs_k0, c.idx0, c.B, c.K0, b.X, st));
    SFM_TRY(gather_tokens(c.xs1, c.xs_k1, c.idx1, c.B, c.K1, b.X + T0 * 256, st));
  } else {
    SFM_TRY(sg_frontend(c.desc0, 256ll * c.ldk0, c.ldk0, c.kpts0, c.scores0, c.kcap0, c.idx0, c.B, c.K0, c.H0, c.W0,
                        b.X, b.kin, st));
    SFM_TRY(sg_frontend(c.desc1, 256ll * c.ldk1, c.ldk1, c.kpts1, c.scores1, c.kcap1, c.idx1, c.B, c.K1, c.H1, c.W1,
                        b.X + T0 * 256, b.kin + T0 * 4, st));
  }
  if (c.precision == 0) {
    SFM_TRY(forward_f32(h, c, b, st));
  } else {
    tb.t0 = b.t0; tb.t1 = b.t1; tb.bn_part = b.bn_part; tb.bn_scale = b.bn_scale; tb.bn_shift = b.bn_shift;
    SFM_TRY(tc_sg_forward(h, c, b.X, b.kin, b.scores, tb, st));
  }
  if (hoist_pass) {   // the residual stream after layer `last_layer - 1`, token-major, one block per image
    SFM_CUDA(cudaMemcpyAsync(c.x_out, b.X, sizeof(float) * 256 * (size_t)c.B * (c.K0 + c.K1), cudaMemcpyDeviceToDevice, st));
    return SFM_OK;
  }
  SinkhornArgs s;
  s.S = b.scores; s.B = c.B; s.N = c.K0; s.M = c.K1; s.alpha = h->sg.bin_score; s.iters = c.iters;
  s.u = b.u; s.v = b.v; s.part = b.part; s.counters = b.counters; s.R = b.R;
  s.fast = c.precision == 1;
  s.fused_part = b.sk_fused;
  s.kb = b.sk_kb;
  s.col_idx = b.idx1;
  s.rowmax = b.max0;
  s.k16 = reinterpret_cast<__half*>(b.sk_k16f);
  ProfScope ps(h, SFM_PROF_SINKHORN, st);
  SFM_TRY(sinkhorn_run(s, st));
  return sinkhorn_match(s, c.thr, b.max0, b.idx0, b.max1, b.idx1, c.m0, c.m1, c.ms0, c.ms1, st);
}

}  // namespace sfm
