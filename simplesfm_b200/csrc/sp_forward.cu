// SuperPoint.forward orchestration (reference simple_sfm/models/superpoint.py:109-171).
// Activations are NHWC fp32; convolutions are implicit GEMMs over (ky, kx, ci).
#include "context.cuh"
#include "gemm_f32.cuh"
#include "sp_post.cuh"
#include "tc_path.cuh"

namespace sfm {

namespace {

struct SpBuffers {
  __nv_bfloat16 *A16, *B16, *feat16;   // bf16 path activations
  float *A, *Bf, *feat, *logits, *smap, *nms, *dmap, *cand_score;
  int *cand_pix, *block_tmp, *cand_counts;
  int cap;
};

void carve(Arena& a, int n, int H, int W, int precision, SpBuffers& b) {
  const size_t full = (size_t)n * H * W;
  const size_t cells = (size_t)n * (H / 8) * (W / 8);
  if (precision == 1) {
    b.A16 = a.get<__nv_bfloat16>(full * 64);
    b.B16 = a.get<__nv_bfloat16>(full * 64);
    b.feat16 = a.get<__nv_bfloat16>(cells * 128);
    b.A = b.Bf = nullptr;
  } else {
    b.A16 = b.B16 = b.feat16 = nullptr;
    b.A = a.get<float>(full * 64);
    b.Bf = a.get<float>(full * 64);
  }
  b.feat = a.get<float>(cells * 128);
  b.logits = a.get<float>(cells * 65);
  b.smap = a.get<float>(full);
  b.nms = a.get<float>(full);
  b.dmap = a.get<float>(cells * 256);
  b.cap = H * W;
  b.cand_pix = a.get<int>(full);
  b.cand_score = a.get<float>(full);
  b.block_tmp = a.get<int>(sp_select_workspace_ints(n, H, W));
  b.cand_counts = a.get<int>(n);
}

int conv(const SpConv& cv, const float* in, float* out, int n, int H, int W, bool relu, cudaStream_t st) {
  GemmF32 g;
  g.A = in;
  g.B = cv.w;
  g.ldb = cv.k * cv.k * cv.cin;
  g.C = out;
  g.ldc = cv.cout;
  g.bias = cv.b;
  g.M = n * H * W;
  g.N = cv.cout;
  g.K = cv.k * cv.k * cv.cin;
  g.flags = relu ? GEMM_RELU : 0;
  if (cv.k == 3) {
    g.convH = H;
    g.convW = W;
    g.convCin = cv.cin;
    return conv3x3_f32(g, st);
  }
  g.lda = cv.cin;
  return gemm_f32_nt(g, st);
}

// bf16 throughput path: the eight 3x3 encoder convs and the two head convs run as tcgen05 implicit GEMMs
// (tc_conv.cu), the two 1x1 head convs on the tcgen05 GEMM with fp32 outputs, so that the exact
// post-processing chain (softmax, NMS, selection, sampling) is shared with the fp32 path.
int tconv(sfm_ctx* h, const SpConv& cv, const __nv_bfloat16* in, __nv_bfloat16* out, int n, int H, int W,
          cudaStream_t st, bool pool = false) {
  tc::TcConv p;
  p.in = in; p.w = cv.w_h; p.bias = cv.b; p.out = out;
  p.n = n; p.H = H; p.W = W; p.Cin = cv.cin; p.Cout = cv.cout; p.relu = 1; p.pool = pool ? 1 : 0;
  return tc::tc_conv3x3(p, h->sm_count, st);
}

int head1x1(sfm_ctx* h, const SpConv& cv, const __nv_bfloat16* in, float* out, long long rows, cudaStream_t st) {
  tc::TcGemm g;
  g.A = in; g.lda = cv.cin; g.a_rows = rows; g.W = cv.w_h; g.ldw = cv.cin; g.w_rows = cv.cout;
  g.M = (int)rows; g.N = cv.cout; g.K = cv.cin; g.K1 = cv.cin; g.bias = cv.b;
  g.out_f = out; g.out_f_ld = cv.cout;
  g.tmaps = h->tmaps;
  return tc::tc_gemm(g, h->sm_count, st);
}

int detect_bf16(sfm_ctx* h, const float* images, int n, int H, int W, const unsigned char* mask, float thr,
                int nms_radius, int border, int* cand_counts, const SpBuffers& b, cudaStream_t st) {
  const SpConv* c = h->sp.conv;
  // the 2 x 2 max-pools run in the epilogues of conv1b / conv2b / conv3b: the full-resolution outputs of those layers
  // are never written
  SFM_TRY(conv1a_f32(images, c[0].w, c[0].b, nullptr, b.A16, n, H, W, st));
  SFM_TRY(tconv(h, c[1], b.A16, b.B16, n, H, W, st, true));
  SFM_TRY(tconv(h, c[2], b.B16, b.A16, n, H / 2, W / 2, st));
  SFM_TRY(tconv(h, c[3], b.A16, b.B16, n, H / 2, W / 2, st, true));
  SFM_TRY(tconv(h, c[4], b.B16, b.A16, n, H / 4, W / 4, st));
  SFM_TRY(tconv(h, c[5], b.A16, b.B16, n, H / 4, W / 4, st, true));
  const int hh = H / 8, ww = W / 8;
  const long long cells = (long long)n * hh * ww;
  SFM_TRY(tconv(h, c[6], b.B16, b.A16, n, hh, ww, st));
  SFM_TRY(tconv(h, c[7], b.A16, b.feat16, n, hh, ww, st));
  SFM_TRY(tconv(h, c[8], b.feat16, b.A16, n, hh, ww, st));                 // convPa
  SFM_TRY(head1x1(h, c[9], b.A16, b.logits, cells, st));                    // convPb -> fp32 logits [cells, 65]
  SFM_TRY(sp_softmax_shuffle(b.logits, 65, b.smap, n, hh, ww, st));
  SFM_TRY(sp_nms(b.smap, b.nms, n, H, W, nms_radius, st));
  SFM_TRY(tconv(h, c[10], b.feat16, b.A16, n, hh, ww, st));                // convDa
  SFM_TRY(head1x1(h, c[11], b.A16, b.dmap, cells, st));                     // convDb -> fp32 descriptor map
  SFM_TRY(sp_select(b.nms, mask, n, H, W, thr, border, b.block_tmp, b.cap, b.cand_pix, b.cand_score, b.cand_counts,
                    st));
  if (cand_counts)
    SFM_CUDA(cudaMemcpyAsync(cand_counts, b.cand_counts, n * sizeof(int), cudaMemcpyDeviceToDevice, st));
  return SFM_OK;
}

}  // namespace

int sp_forward_detect(sfm_ctx* h, const float* images, int n, int H, int W, const unsigned char* mask, float thr,
                      int nms_radius, int border, int precision, int* cand_counts, void* ws, size_t ws_bytes,
                      bool dry, size_t* need, cudaStream_t st) {
  SFM_REQUIRE(precision == 0 || precision == 1, "superpoint: unknown precision %d", precision);
  SFM_REQUIRE(n >= 0 && H >= 8 && W >= 8, "superpoint: H and W must be at least 8 (H=%d W=%d)", H, W);
  // Sides that are not multiples of 8 (fp32 path): like the reference, the convolutions run at the full size, every
  // pool floors, and everything after the detector head lives on the (H/8*8, W/8*8) score map (superpoint.py:112-130).
  // A mask has the image's shape and cannot broadcast against that map (the reference raises, superpoint.py:134).
  SFM_REQUIRE((H % 8 == 0 && W % 8 == 0) || precision == 0,
              "superpoint: the bf16 path needs H and W to be multiples of 8 (H=%d W=%d)", H, W);
  SFM_REQUIRE((H % 8 == 0 && W % 8 == 0) || mask == nullptr,
              "superpoint: a mask of size (%d, %d) does not match the (%d, %d) score map", H, W, H / 8 * 8, W / 8 * 8);
  Arena a(dry ? nullptr : ws, dry ? 0 : ws_bytes);
  SpBuffers b;
  carve(a, n, H, W, precision, b);
  if (need) *need = a.off;
  if (dry) return SFM_OK;
  if (a.overflow) {
    set_error("superpoint: workspace too small (%zu bytes given, %zu needed)", ws_bytes, a.off);
    return SFM_ERR_WORKSPACE;
  }
  SFM_REQUIRE(h != nullptr && h->sp.loaded, "superpoint: weights not loaded (call sfm_load_superpoint)");
  if (n == 0) return SFM_OK;
  const SpConv* c = h->sp.conv;
  if (precision == 1) return detect_bf16(h, images, n, H, W, mask, thr, nms_radius, border, cand_counts, b, st);
  // encoder (superpoint.py:112-122); H2 = H / 2 ... are the floored sizes nn.MaxPool2d(2, 2) produces
  const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2, hh = H4 / 2, ww = W4 / 2, Hs = hh * 8, Ws = ww * 8;
  SFM_TRY(conv1a_f32(images, c[0].w, c[0].b, b.A, nullptr, n, H, W, st));
  SFM_TRY(conv(c[1], b.A, b.Bf, n, H, W, true, st));
  SFM_TRY(maxpool2_nhwc_f32(b.Bf, b.A, n, H, W, 64, st));
  SFM_TRY(conv(c[2], b.A, b.Bf, n, H2, W2, true, st));
  SFM_TRY(conv(c[3], b.Bf, b.A, n, H2, W2, true, st));
  SFM_TRY(maxpool2_nhwc_f32(b.A, b.Bf, n, H2, W2, 64, st));
  SFM_TRY(conv(c[4], b.Bf, b.A, n, H4, W4, true, st));
  SFM_TRY(conv(c[5], b.A, b.Bf, n, H4, W4, true, st));
  SFM_TRY(maxpool2_nhwc_f32(b.Bf, b.A, n, H4, W4, 128, st));
  SFM_TRY(conv(c[6], b.A, b.Bf, n, hh, ww, true, st));
  SFM_TRY(conv(c[7], b.Bf, b.feat, n, hh, ww, true, st));
  // detector head (superpoint.py:125-131)
  SFM_TRY(conv(c[8], b.feat, b.A, n, hh, ww, true, st));
  SFM_TRY(conv(c[9], b.A, b.logits, n, hh, ww, false, st));
  SFM_TRY(sp_softmax_shuffle(b.logits, 65, b.smap, n, hh, ww, st));
  SFM_TRY(sp_nms(b.smap, b.nms, n, Hs, Ws, nms_radius, st));
  // descriptor head (superpoint.py:157-158); per-cell L2 normalisation is fused into sampling
  SFM_TRY(conv(c[10], b.feat, b.A, n, hh, ww, true, st));
  SFM_TRY(conv(c[11], b.A, b.dmap, n, hh, ww, false, st));
  // mask / threshold / border (superpoint.py:133-145)
  SFM_TRY(sp_select(b.nms, mask, n, Hs, Ws, thr, border, b.block_tmp, b.cap, b.cand_pix, b.cand_score, b.cand_counts,
                    st));
  if (cand_counts)
    SFM_CUDA(cudaMemcpyAsync(cand_counts, b.cand_counts, n * sizeof(int), cudaMemcpyDeviceToDevice, st));
  return SFM_OK;
}

int sp_forward_describe(sfm_ctx* h, int n, int H, int W, int max_kp, int align_corners, int precision, int kp_cap,
                        float* kpts_xy, float* scores, float* desc, int* counts, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  (void)h;
  SFM_REQUIRE(precision == 0 || precision == 1, "superpoint: unknown precision %d", precision);
  SFM_REQUIRE(max_kp == -1 || max_kp > 0, "\"max_keypoints\" must be positive or \"-1\"");
  SFM_REQUIRE(kp_cap >= 0, "superpoint: kp_cap < 0");
  Arena a(ws, ws_bytes);
  SpBuffers b;
  carve(a, n, H, W, precision, b);
  if (a.overflow) {
    set_error("superpoint: workspace too small (%zu bytes given, %zu needed)", ws_bytes, a.off);
    return SFM_ERR_WORKSPACE;
  }
  if (n == 0) return SFM_OK;
  SFM_TRY(sp_rank_sort(b.cand_pix, b.cand_score, b.cand_counts, n, b.cap, max_kp, W / 8 * 8, kp_cap, kpts_xy, scores,
                       counts, st));   // pixel indices are positions on the (H/8*8, W/8*8) score map
  return sp_sample_descriptors(b.dmap, n, H / 8, W / 8, kpts_xy, counts, kp_cap, align_corners, desc, kp_cap, st);
}

int sp_tap(int n, int H, int W, int precision, int which, void* ws, size_t ws_bytes, float** out, size_t* cnt) {
  Arena a(ws, ws_bytes);
  SpBuffers b;
  carve(a, n, H, W, precision, b);
  if (a.overflow) {
    set_error("superpoint tap: workspace too small");
    return SFM_ERR_WORKSPACE;
  }
  const size_t full = (size_t)n * (H / 8 * 8) * (W / 8 * 8), cells = (size_t)n * (H / 8) * (W / 8);
  switch (which) {
    case 0: *out = b.feat; *cnt = cells * 128; break;
    case 1: *out = b.logits; *cnt = cells * 65; break;
    case 2: *out = b.smap; *cnt = full; break;
    case 3: *out = b.nms; *cnt = full; break;
    case 4: *out = b.dmap; *cnt = cells * 256; break;
    default: set_error("superpoint tap: unknown tap %d", which); return SFM_ERR_INVALID;
  }
  return SFM_OK;
}

}  // namespace sfm
