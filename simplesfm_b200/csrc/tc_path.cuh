// bf16 tcgen05/TMEM/TMA path of the SuperGlue GNN (throughput mode).  Internal.
#pragma once
#include "context.cuh"

namespace sfm {

struct TcSgBuffers {
  __nv_bfloat16 *xh, *qkv, *msgp, *msg, *hid, *mdh;
  float *t0, *t1, *bn_part, *bn_scale, *bn_shift;   // shared with the fp32 buffers (set by sg_forward)
};

void tc_sg_carve(Arena& a, const SgCall& c, TcSgBuffers& b);
// X: fp32 master descriptors [T,256] (frontend output), kin [T,4]; writes scores [B,K0,K1] f32
int tc_sg_forward(sfm_ctx* h, const SgCall& c, float* X, const float* kin, float* scores, const TcSgBuffers& b,
                  cudaStream_t st);
// keypoint encoder on the fp32 SIMT path (shared by both precisions): X += kenc(kin)
int kenc_forward_f32(sfm_ctx* h, const SgCall& c, float* X, const float* kin, float* t0, float* t1, float* part,
                     float* scale, float* shift, cudaStream_t st);

// C = A W^T + b on the fp32 FFMA GEMM (sg_forward.cu)
int linear_f32_raw(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int K, int N,
                   const float* bias, float* C, int ldc, long long T, int flags, cudaStream_t st);

namespace tc {

struct TmapCache;   // per-handle cache of encoded TMA tensor maps (tc_common.cuh)
TmapCache* tmap_cache_create();
void tmap_cache_destroy(TmapCache* c);

// D_b[M,N] = alpha * A_b[M,K] * W_b[N,K]^T + bias   (bf16 operands, fp32 accumulation in TMEM)
struct TcGemm {
  const __nv_bfloat16* A = nullptr;   // [a_rows, K1] row-major (lda)
  const __nv_bfloat16* A2 = nullptr;  // [a_rows, K-K1] (lda2): concat along K (MLP input [x | message])
  const __nv_bfloat16* W = nullptr;   // [w_rows, K] (ldw)
  long long a_rows = 0, w_rows = 0;
  int lda = 0, lda2 = 0, ldw = 0;
  int M = 0, N = 0, K = 0, K1 = 0;
  int batch = 1;
  int a_row_stride = 0;                  // A row offset per batch
  int w_row_base = 0, w_row_stride = 0;  // W row offset = base + b * stride
  const float* bias = nullptr;
  float alpha = 1.f;
  int relu = 0;
  __nv_bfloat16* out_h = nullptr;
  long long out_h_ld = 0, out_h_bstride = 0;
  int out_row_base = 0, out_row_stride = 0;  // bf16 output row offset (TMA store coordinates)
  float* out_f = nullptr;
  long long out_f_ld = 0, out_f_bstride = 0;
  int accumulate_f = 0;  // out_f += result (residual stream); out_h then holds bf16(out_f)
  // fused train-mode BatchNorm (both optional):
  //  stat_partial: the epilogue also writes, per 32-row block and output column, (mean, centred M2) of the
  //                bf16 output -> [M/32, N, 2] floats, block index = global row / 32 (feeds bn_finalize)
  //  xf_*:         A is normalised on the fly, a <- relu(a * scale[g] + shift[g]) with g the statistic group of
  //                the tile's rows (side 0 rows [0,xf_T0) in groups of xf_rows0, side 1 in groups of xf_rows1)
  float* stat_partial = nullptr;
  const float* xf_scale = nullptr;   // [2 * xf_G, K]
  const float* xf_shift = nullptr;
  long long xf_T0 = 0, xf_rows0 = 1, xf_rows1 = 1;
  int xf_G = 1;
  // filled by the launcher
  int vec_f = 0, coalesced_rmw = 0, num_panels = 0, num_m_tiles = 0, cpp = 1;
  int dbg_flags = 0;
  int l2_prefetch = 2;   // bit 0: TMA-prefetch the next A tile into L2, bit 1: prefetch the next residual rows
  long long* dbg = nullptr;  // optional per-role wait-cycle counters (debug builds, SFM_TC_DEBUG)
  TmapCache* tmaps = nullptr;   // the handle's tensor-map cache (optional)
};
int tc_gemm(const TcGemm& p, int num_sms, cudaStream_t st);

// fused multi-head attention for both sides of one GNN layer (superglue.py:85-108):
// message[q] = softmax(Q K^T / 8) V per head; Q/K/V read from the fused projection buffer
// qkv [T, 768] (q | k | v, heads contiguous), output msgp [T, 256] bf16.
struct TcAttn {
  const __nv_bfloat16* qkv = nullptr;
  long long rows_total = 0;
  __nv_bfloat16* out = nullptr;
  int B = 0;
  int Nq[2] = {0, 0}, Nk[2] = {0, 0};
  long long q_row0[2] = {0, 0}, k_row0[2] = {0, 0};
  int tma_out = 1;   // whole output tiles leave through a swizzled staging tile and one TMA store
  int stagger = 1;   // tile 1 of a CTA runs half a K/V step behind tile 0 (ping-pong on the MUFU pipe)
  TmapCache* tmaps = nullptr;   // the handle's tensor-map cache (optional)
  // exact-recomputation list (handle-owned, zero on entry, left zero): redo[0] = count, entries int4 from redo + 4
  int* redo = nullptr;
  int redo_cap = 0;
};
constexpr int ATT_REDO_CAP = 4096;
inline size_t att_redo_bytes() { return (size_t)(4 + 4 * ATT_REDO_CAP) * sizeof(int); }
int tc_attention(const TcAttn& p, int num_sms, cudaStream_t st);

// 3x3 / pad 1 convolution, NHWC bf16, implicit GEMM on tcgen05 with 4-D TMA im2col (tc_conv.cu)
struct TcConv {
  const __nv_bfloat16* in = nullptr;  // [n, H, W, Cin]
  const __nv_bfloat16* w = nullptr;   // [Cout, 9 * Cin]  (tap-major, channel-minor)
  const float* bias = nullptr;
  __nv_bfloat16* out = nullptr;       // [n, H, W, Cout]; with pool: [n, H / 2, W / 2, Cout]
  int n = 0, H = 0, W = 0, Cin = 0, Cout = 0, relu = 1;
  int pool = 0;                       // fuse nn.MaxPool2d(2, 2) into the epilogue (H, W even): only the pooled map is written
};
int tc_conv3x3(const TcConv& p, int num_sms, cudaStream_t st);
int maxpool2_nhwc_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int n, int H, int W, int C, cudaStream_t st);

// elementwise / reduction helpers of the bf16 path
int f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t st);

}  // namespace tc
}  // namespace sfm
