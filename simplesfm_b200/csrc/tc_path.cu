// SuperGlue GNN on the bf16 tcgen05 path (reference simple_sfm/models/superglue.py:254-265).
// fp32 master copy of the residual stream X, bf16 shadow for the tensor-core operands.
#include "tc_path.cuh"
#include "bn_kernels.cuh"
#include "gemm_f32.cuh"

namespace sfm {

void tc_sg_carve(Arena& a, const SgCall& c, TcSgBuffers& b) {
  const long long T0 = (long long)c.B * c.K0, T1 = (long long)c.B * c.K1, T = T0 + T1;
  b.xh = a.get<__nv_bfloat16>(T * 256);
  b.qkv = a.get<__nv_bfloat16>(T * 768);
  b.msgp = a.get<__nv_bfloat16>(T * 256);
  b.msg = a.get<__nv_bfloat16>(T * 256);
  b.hid = a.get<__nv_bfloat16>(T * 512);
  b.mdh = a.get<__nv_bfloat16>(T * 256);
  (void)T0; (void)T1;
}

int tc_sg_forward(sfm_ctx* h, const SgCall& c, float* X, const float* kin, float* scores, const TcSgBuffers& b,
                  cudaStream_t st) {
  using namespace tc;
  const SgWeights& w = h->sg;
  const long long T0 = (long long)c.B * c.K0, T1 = (long long)c.B * c.K1, T = T0 + T1;
  const bool train = c.bn_mode == 0;
  const int sms = h->sm_count;
  {
    ProfScope ps(h, SFM_PROF_OTHER, st);
    SFM_TRY(kenc_forward_f32(h, c, X, kin, b.t0, b.t1, b.bn_part, b.bn_scale, b.bn_shift, st));
    SFM_TRY(f32_to_bf16(X, b.xh, T * 256, st));
  }
  auto linear = [&](const __nv_bfloat16* A, int lda, const __nv_bfloat16* A2, int lda2, int K, int K1,
                    const __nv_bfloat16* W, int N, const float* bias, int relu, __nv_bfloat16* out_h, int ldo,
                    float* out_f, int accumulate) {
    ProfScope ps(h, SFM_PROF_LINEAR, st);
    TcGemm g;
    g.A = A; g.lda = lda; g.A2 = A2; g.lda2 = lda2; g.a_rows = T;
    g.W = W; g.ldw = K; g.w_rows = N;
    g.M = (int)T; g.N = N; g.K = K; g.K1 = K1;
    g.bias = bias; g.relu = relu;
    g.out_h = out_h; g.out_h_ld = ldo;
    g.out_f = out_f; g.out_f_ld = 256; g.accumulate_f = accumulate;
    return tc_gemm(g, sms, st);
  };
  for (int l = 0; l < 18; l++) {
    const SgLayer& L = w.L[l];
    const bool cross = l & 1;
    SFM_TRY(linear(b.xh, 256, nullptr, 0, 256, 256, L.Wqkv_h, 768, L.bqkv, 0, b.qkv, 768, nullptr, 0));
    {
      ProfScope ps(h, SFM_PROF_ATTENTION, st);
      TcAttn a;
      a.qkv = b.qkv; a.rows_total = T; a.out = b.msgp; a.B = c.B;
      a.Nq[0] = c.K0; a.Nq[1] = c.K1;
      a.q_row0[0] = 0; a.q_row0[1] = T0;
      a.Nk[0] = cross ? c.K1 : c.K0; a.Nk[1] = cross ? c.K0 : c.K1;
      a.k_row0[0] = cross ? T0 : 0; a.k_row0[1] = cross ? 0 : T0;
      SFM_TRY(tc_attention(a, st));
    }
    // the merge conv is folded into the MLP input weights (W1c), so the attention output feeds the MLP directly
    if (train) {
      SFM_TRY(linear(b.xh, 256, b.msgp, 256, 512, 256, L.W1c_h, 512, L.b1c, 0, b.hid, 512, nullptr, 0));
      ProfScope ps(h, SFM_PROF_OTHER, st);
      const BnGroups gr{T0, T1, c.groups > 0 ? c.groups : 1};
      SFM_TRY(bn_train_stats<__nv_bfloat16>(b.hid, 512, 512, gr, L.bn1.gamma, L.bn1.beta, 1e-5f, b.bn_part, b.bn_scale,
                                            b.bn_shift, st));
      SFM_TRY(bn_apply_relu<__nv_bfloat16>(b.hid, 512, 512, gr, b.bn_scale, b.bn_shift, st));
    } else {
      SFM_TRY(linear(b.xh, 256, b.msgp, 256, 512, 256, L.W1cf_h, 512, L.b1cf, 1, b.hid, 512, nullptr, 0));
    }
    SFM_TRY(linear(b.hid, 512, nullptr, 0, 512, 512, L.W2_h, 256, L.b2, 0, b.xh, 256, X, 1));
  }
  SFM_TRY(linear(b.xh, 256, nullptr, 0, 256, 256, w.Wf_h, 256, w.bf, 0, b.mdh, 256, nullptr, 0));
  // scores_b = md0_b md1_b^T / 16  (superglue.py:264-265): batched, "weights" = the pair's side-1 descriptors
  ProfScope ps(h, SFM_PROF_LINEAR, st);
  TcGemm g;
  g.A = b.mdh; g.lda = 256; g.a_rows = T; g.a_row_stride = c.K0;
  g.W = b.mdh; g.ldw = 256; g.w_rows = T; g.w_row_base = (int)T0; g.w_row_stride = c.K1;
  g.M = c.K0; g.N = c.K1; g.K = 256; g.K1 = 256; g.batch = c.B;
  g.alpha = 1.0f / 16.0f;
  g.out_f = scores; g.out_f_ld = c.K1; g.out_f_bstride = (long long)c.K0 * c.K1;
  return tc_gemm(g, sms, st);
}

}  // namespace sfm
