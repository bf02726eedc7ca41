// extern "C" surface of libsfm_b200 (declared in include/sfm_b200.h).
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <vector>

#include "../../include/sfm_b200.h"
#include "context.cuh"
#include "gemm_f32.cuh"
#include "sg_kernels.cuh"
#include "sp_post.cuh"
#include "tc_path.cuh"

namespace sfm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return SFM_ERR_CUDA;
}

static std::atomic<long long> g_launches{0};
void count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static cudaEvent_t take_event(sfm_ctx* h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
ProfScope::ProfScope(sfm_ctx* h_, int kind_, cudaStream_t st_) : h(h_), kind(kind_), st(st_) {
  if (h && h->profiling) {
    e0 = take_event(h);
    e1 = take_event(h);
    cudaEventRecord(e0, st);
  }
}
ProfScope::~ProfScope() {
  if (e0) {
    cudaEventRecord(e1, st);
    h->prof[kind].push_back({e0, e1});
  }
}

namespace {

template <typename T>
int upload(sfm_ctx* h, const std::vector<T>& host, T** dev) {
  void* p = nullptr;
  SFM_CUDA(cudaMalloc(&p, host.size() * sizeof(T) + 16));
  h->allocs.push_back(p);
  SFM_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dev = (T*)p;
  return SFM_OK;
}

int upload_bf16(sfm_ctx* h, const std::vector<float>& host, __nv_bfloat16** dev) {
  std::vector<__nv_bfloat16> t(host.size());
  for (size_t i = 0; i < host.size(); i++) t[i] = __float2bfloat16(host[i]);
  return upload(h, t, dev);
}

constexpr int SP_DIMS[12][3] = {{1, 64, 3},    {64, 64, 3},   {64, 64, 3},   {64, 64, 3},
                                {64, 128, 3},  {128, 128, 3}, {128, 128, 3}, {128, 128, 3},
                                {128, 256, 3}, {256, 65, 1},  {128, 256, 3}, {256, 256, 1}};

size_t sp_packed_floats() {
  size_t n = 0;
  for (auto& d : SP_DIMS) n += (size_t)d[1] * d[0] * d[2] * d[2] + d[1];
  return n;
}

constexpr int KCH[6] = {3, 32, 64, 128, 256, 256};

size_t sg_packed_floats() {
  size_t n = 1;
  for (int i = 0; i < 5; i++) {
    n += (size_t)KCH[i + 1] * KCH[i] + KCH[i + 1];
    if (i < 4) n += 4 * (size_t)KCH[i + 1];
  }
  size_t layer = 4 * (256 * 256 + 256) + (512 * 512 + 512) + 4 * 512 + (256 * 512 + 256);
  n += 18 * layer;
  n += 256 * 256 + 256;
  return n;
}

struct Reader {
  const float* p;
  size_t left;
  std::vector<float> take(size_t n) {
    std::vector<float> v(p, p + n);
    p += n;
    left -= n;
    return v;
  }
};

void fold_bn(const std::vector<float>& W, const std::vector<float>& b, const std::vector<float>& g,
             const std::vector<float>& be, const std::vector<float>& rm, const std::vector<float>& rv, int cout,
             int cin, std::vector<float>& Wf, std::vector<float>& bf) {
  Wf.resize(W.size());
  bf.resize(cout);
  for (int o = 0; o < cout; o++) {
    float s = g[o] / sqrtf(rv[o] + 1e-5f);
    for (int i = 0; i < cin; i++) Wf[(size_t)o * cin + i] = W[(size_t)o * cin + i] * s;
    bf[o] = (b[o] - rm[o]) * s + be[o];
  }
}

}  // namespace
}  // namespace sfm

using namespace sfm;

extern "C" {

int sfm_abi_version(void) { return SFM_ABI_VERSION; }
const char* sfm_last_error(void) { return sfm::g_err; }
size_t sfm_superpoint_packed_floats(void) { return sp_packed_floats(); }
size_t sfm_superglue_packed_floats(void) { return sg_packed_floats(); }

int sfm_create(int device, sfm_handle_t* out) {
  SFM_REQUIRE(out != nullptr, "sfm_create: out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("sfm_create: no CUDA device available (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return SFM_ERR_CUDA;
  }
  SFM_REQUIRE(device >= 0 && device < count, "sfm_create: device %d out of range [0,%d)", device, count);
  SFM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SFM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("sfm_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
              prop.minor);
    return SFM_ERR_UNSUPPORTED;
  }
  sfm_ctx* h = new sfm_ctx();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->tmaps = tc::tmap_cache_create();
  {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, tc::att_redo_bytes());
    if (e == cudaSuccess) e = cudaMemset(p, 0, tc::att_redo_bytes());
    if (e != cudaSuccess) {
      delete h;
      return cuda_fail(e, "cudaMalloc(attention redo list)", __FILE__, __LINE__);
    }
    h->att_redo = (int*)p;
    h->allocs.push_back(p);
  }
  *out = h;
  return SFM_OK;
}

long long sfm_launch_count(void) { return g_launches.load(); }

int sfm_set_profiling(sfm_handle_t h, int enable) {
  SFM_REQUIRE(h, "sfm_set_profiling: NULL handle");
  h->profiling = enable != 0;
  return SFM_OK;
}

int sfm_get_profile(sfm_handle_t h, float* ms, long long* groups) {
  SFM_REQUIRE(h && ms && groups, "sfm_get_profile: NULL argument");
  SFM_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < SFM_PROF_KINDS; k++) {
    double total = 0.0;
    for (auto& pr : h->prof[k]) {
      float t = 0.f;
      SFM_CUDA(cudaEventElapsedTime(&t, pr.first, pr.second));
      total += t;
      h->event_pool.push_back(pr.first);
      h->event_pool.push_back(pr.second);
    }
    ms[k] = (float)total;
    groups[k] = (long long)h->prof[k].size();
    h->prof[k].clear();
  }
  return SFM_OK;
}

int sfm_destroy(sfm_handle_t h) {
  if (!h) return SFM_OK;
  cudaSetDevice(h->device);
  for (int k = 0; k < SFM_PROF_KINDS; k++)
    for (auto& pr : h->prof[k]) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
  for (void* p : h->allocs) cudaFree(p);
  tc::tmap_cache_destroy(h->tmaps);
  delete h;
  return SFM_OK;
}

int sfm_load_superpoint(sfm_handle_t h, const float* packed, size_t n) {
  SFM_REQUIRE(h && packed, "sfm_load_superpoint: NULL argument");
  SFM_REQUIRE(n == sp_packed_floats(), "sfm_load_superpoint: expected %zu floats, got %zu", sp_packed_floats(), n);
  SFM_CUDA(cudaSetDevice(h->device));
  Reader r{packed, n};
  for (int i = 0; i < 12; i++) {
    const int cin = SP_DIMS[i][0], cout = SP_DIMS[i][1], k = SP_DIMS[i][2];
    std::vector<float> w = r.take((size_t)cout * cin * k * k), b = r.take(cout);
    std::vector<float> wk(w.size());
    for (int o = 0; o < cout; o++)
      for (int ci = 0; ci < cin; ci++)
        for (int t = 0; t < k * k; t++)
          wk[((size_t)o * k * k + t) * cin + ci] = w[((size_t)o * cin + ci) * k * k + t];
    SpConv& c = h->sp.conv[i];
    c.cin = cin; c.cout = cout; c.k = k;
    SFM_TRY(upload(h, wk, &c.w));
    SFM_TRY(upload(h, b, &c.b));
    SFM_TRY(upload_bf16(h, wk, &c.w_h));
  }
  h->sp.loaded = true;
  return SFM_OK;
}

int sfm_load_superglue(sfm_handle_t h, const float* packed, size_t n) {
  SFM_REQUIRE(h && packed, "sfm_load_superglue: NULL argument");
  SFM_REQUIRE(n == sg_packed_floats(), "sfm_load_superglue: expected %zu floats, got %zu", sg_packed_floats(), n);
  SFM_CUDA(cudaSetDevice(h->device));
  SgWeights& w = h->sg;
  Reader r{packed, n};
  w.bin_score = r.take(1)[0];
  for (int i = 0; i < 5; i++) {
    const int cin = KCH[i], cout = KCH[i + 1];
    std::vector<float> W = r.take((size_t)cout * cin), b = r.take(cout);
    int cin_p = cin;
    if (i == 0) {  // pad K 3 -> 4
      std::vector<float> Wp((size_t)cout * 4, 0.f);
      for (int o = 0; o < cout; o++)
        for (int c = 0; c < 3; c++) Wp[o * 4 + c] = W[o * 3 + c];
      W.swap(Wp);
      cin_p = 4;
    }
    SFM_TRY(upload(h, W, &w.kW[i]));
    SFM_TRY(upload(h, b, &w.kb[i]));
    if (i < 4) {
      std::vector<float> g = r.take(cout), be = r.take(cout), rm = r.take(cout), rv = r.take(cout);
      SFM_TRY(upload(h, g, &w.kbn[i].gamma));
      SFM_TRY(upload(h, be, &w.kbn[i].beta));
      SFM_TRY(upload(h, rm, &w.kbn[i].rmean));
      SFM_TRY(upload(h, rv, &w.kbn[i].rvar));
      std::vector<float> Wf, bf;
      fold_bn(W, b, g, be, rm, rv, cout, cin_p, Wf, bf);
      SFM_TRY(upload(h, Wf, &w.kWf[i]));
      SFM_TRY(upload(h, bf, &w.kbf[i]));
      if (i == 3) {
        SFM_TRY(upload_bf16(h, W, &w.kW3_h));
        SFM_TRY(upload_bf16(h, Wf, &w.kW3f_h));
      }
    }
    if (i == 4) SFM_TRY(upload_bf16(h, W, &w.kW4_h));
  }
  int perm[256];
  for (int c = 0; c < 256; c++) perm[c] = (c % 64) * 4 + c / 64;  // head-contiguous c' -> reference channel d*4+h
  for (int l = 0; l < 18; l++) {
    SgLayer& L = w.L[l];
    std::vector<float> Wqkv((size_t)768 * 256), bqkv(768);
    for (int j = 0; j < 3; j++) {
      std::vector<float> W = r.take(256 * 256), b = r.take(256);
      for (int c = 0; c < 256; c++) {
        memcpy(&Wqkv[((size_t)j * 256 + c) * 256], &W[(size_t)perm[c] * 256], 256 * sizeof(float));
        bqkv[j * 256 + c] = b[perm[c]];
      }
    }
    std::vector<float> Wm = r.take(256 * 256), bm = r.take(256), Wmp(256 * 256);
    for (int o = 0; o < 256; o++)
      for (int c = 0; c < 256; c++) Wmp[o * 256 + c] = Wm[o * 256 + perm[c]];
    std::vector<float> W1 = r.take(512 * 512), b1 = r.take(512);
    std::vector<float> g = r.take(512), be = r.take(512), rm = r.take(512), rv = r.take(512);
    std::vector<float> W2 = r.take(256 * 512), b2 = r.take(256);
    std::vector<float> W1f, b1f;
    fold_bn(W1, b1, g, be, rm, rv, 512, 512, W1f, b1f);
    SFM_TRY(upload(h, Wqkv, &L.Wqkv)); SFM_TRY(upload(h, bqkv, &L.bqkv));
    SFM_TRY(upload(h, Wmp, &L.Wm));    SFM_TRY(upload(h, bm, &L.bm));
    SFM_TRY(upload(h, W1, &L.W1));     SFM_TRY(upload(h, b1, &L.b1));
    SFM_TRY(upload(h, g, &L.bn1.gamma)); SFM_TRY(upload(h, be, &L.bn1.beta));
    SFM_TRY(upload(h, rm, &L.bn1.rmean)); SFM_TRY(upload(h, rv, &L.bn1.rvar));
    SFM_TRY(upload(h, W1f, &L.W1f));   SFM_TRY(upload(h, b1f, &L.b1f));
    SFM_TRY(upload(h, W2, &L.W2));     SFM_TRY(upload(h, b2, &L.b2));
    SFM_TRY(upload_bf16(h, Wqkv, &L.Wqkv_h));
    SFM_TRY(upload_bf16(h, Wmp, &L.Wm_h));
    SFM_TRY(upload_bf16(h, W1, &L.W1_h));
    SFM_TRY(upload_bf16(h, W1f, &L.W1f_h));
    SFM_TRY(upload_bf16(h, W2, &L.W2_h));
    // fold the attention merge conv into the first MLP conv (exact algebra, done in double)
    for (int variant = 0; variant < 2; variant++) {
      const std::vector<float>& Wa = variant ? W1f : W1;
      const std::vector<float>& ba = variant ? b1f : b1;
      std::vector<float> Wc((size_t)512 * 512), bc(512);
      for (int o = 0; o < 512; o++) {
        for (int c = 0; c < 256; c++) Wc[(size_t)o * 512 + c] = Wa[(size_t)o * 512 + c];
        double bacc = ba[o];
        for (int k = 0; k < 256; k++) bacc += (double)Wa[(size_t)o * 512 + 256 + k] * bm[k];
        bc[o] = (float)bacc;
        for (int c = 0; c < 256; c++) {
          double acc = 0.0;
          for (int k = 0; k < 256; k++) acc += (double)Wa[(size_t)o * 512 + 256 + k] * Wmp[(size_t)k * 256 + c];
          Wc[(size_t)o * 512 + 256 + c] = (float)acc;
        }
      }
      SFM_TRY(upload_bf16(h, Wc, variant ? &L.W1cf_h : &L.W1c_h));
      SFM_TRY(upload(h, bc, variant ? &L.b1cf : &L.b1c));
    }
  }
  std::vector<float> Wf = r.take(256 * 256), bf = r.take(256);
  SFM_TRY(upload(h, Wf, &w.Wf));
  SFM_TRY(upload(h, bf, &w.bf));
  SFM_TRY(upload_bf16(h, Wf, &w.Wf_h));
  w.loaded = true;
  return SFM_OK;
}

// ---- SuperPoint -------------------------------------------------------------------------------
size_t sfm_superpoint_workspace_bytes(int n_img, int H, int W, int precision) {
  size_t need = 0;
  if (sp_forward_detect(nullptr, nullptr, n_img, H, W, nullptr, 0.f, 0, 0, precision, nullptr, nullptr, 0, true, &need,
                        0) != SFM_OK)
    return 0;
  return need;
}

int sfm_superpoint_detect(sfm_handle_t h, const float* images, int n_img, int H, int W, const uint8_t* mask,
                          float thr, int nms_radius, int border, int precision, int* cand_counts, void* ws,
                          size_t ws_bytes, void* stream) {
  SFM_REQUIRE(h, "sfm_superpoint_detect: NULL handle");
  SFM_REQUIRE(n_img == 0 || images, "sfm_superpoint_detect: NULL images");
  return sp_forward_detect(h, images, n_img, H, W, mask, thr, nms_radius, border, precision, cand_counts, ws, ws_bytes,
                           false, nullptr, (cudaStream_t)stream);
}

int sfm_superpoint_describe(sfm_handle_t h, int n_img, int H, int W, int max_keypoints, int align_corners,
                            int precision, int kp_cap, float* kpts_xy, float* scores, float* desc, int* counts,
                            void* ws, size_t ws_bytes, void* stream) {
  SFM_REQUIRE(h, "sfm_superpoint_describe: NULL handle");
  return sp_forward_describe(h, n_img, H, W, max_keypoints, align_corners, precision, kp_cap, kpts_xy, scores, desc,
                             counts, ws, ws_bytes, (cudaStream_t)stream);
}

int sfm_superpoint_tap(int n_img, int H, int W, int precision, int which, void* ws, size_t ws_bytes, float** out_ptr,
                       size_t* out_floats) {
  SFM_REQUIRE(out_ptr && out_floats, "sfm_superpoint_tap: NULL output");
  return sp_tap(n_img, H, W, precision, which, ws, ws_bytes, out_ptr, out_floats);
}

int sfm_u8_to_unit_f32(const uint8_t* pixels, float* out, size_t n, void* stream) {
  SFM_REQUIRE(n == 0 || (pixels && out), "sfm_u8_to_unit_f32: NULL argument");
  return sp_u8_to_unit_f32(pixels, out, (long long)n, (cudaStream_t)stream);
}

int sfm_simple_nms(const float* scores, float* out, int n_img, int H, int W, int nms_radius, void* stream) {
  SFM_REQUIRE(scores && out, "sfm_simple_nms: NULL argument");
  return sp_nms(scores, out, n_img, H, W, nms_radius, (cudaStream_t)stream);
}

int sfm_sample_descriptors(const float* dmap_nhwc, int n_img, int h, int w, const float* kpts_xy, const int* counts,
                           int kp_cap, int align_corners, float* desc, void* stream) {
  SFM_REQUIRE(dmap_nhwc && kpts_xy && counts && desc, "sfm_sample_descriptors: NULL argument");
  return sp_sample_descriptors(dmap_nhwc, n_img, h, w, kpts_xy, counts, kp_cap, align_corners, desc, kp_cap,
                               (cudaStream_t)stream);
}

// ---- SuperGlue --------------------------------------------------------------------------------
static SgCall shape_call(int B, int K0, int K1, int precision, int groups) {
  SgCall c;
  memset(&c, 0, sizeof(c));
  c.B = B; c.K0 = K0; c.K1 = K1; c.precision = precision; c.groups = groups;
  c.last_layer = 18;
  return c;
}

size_t sfm_superglue_workspace_bytes(int B, int K0, int K1, int precision, int chunk_groups) {
  size_t need = 0;
  SgCall c = shape_call(B, K0, K1, precision, chunk_groups);
  if (sg_forward(nullptr, c, nullptr, 0, true, &need, 0, nullptr, nullptr, 0) != SFM_OK) return 0;
  return need;
}

int sfm_superglue_forward(sfm_handle_t h, const float* desc0, const float* kpts0, const float* scores0, int ldk0,
                          int kcap0, const int* idx0, const float* desc1, const float* kpts1, const float* scores1,
                          int ldk1, int kcap1, const int* idx1, int B, int K0, int K1, int H0, int W0, int H1, int W1,
                          int sinkhorn_iterations, float match_threshold, int bn_mode, int precision,
                          int chunk_groups, int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                          void* ws, size_t ws_bytes, void* stream) {
  SFM_REQUIRE(h, "sfm_superglue_forward: NULL handle");
  SFM_REQUIRE(B == 0 || (desc0 && kpts0 && scores0 && desc1 && kpts1 && scores1), "sfm_superglue_forward: NULL input");
  SFM_REQUIRE(B == 0 || (matches0 && matches1 && mscores0 && mscores1), "sfm_superglue_forward: NULL output");
  SFM_REQUIRE(K0 <= ldk0 && K0 <= kcap0 && K1 <= ldk1 && K1 <= kcap1, "sfm_superglue_forward: K exceeds record size");
  SFM_REQUIRE(sinkhorn_iterations >= 0, "sfm_superglue_forward: negative sinkhorn_iterations");
  SgCall c;
  c.desc0 = desc0; c.kpts0 = kpts0; c.scores0 = scores0; c.ldk0 = ldk0; c.kcap0 = kcap0; c.idx0 = idx0;
  c.desc1 = desc1; c.kpts1 = kpts1; c.scores1 = scores1; c.ldk1 = ldk1; c.kcap1 = kcap1; c.idx1 = idx1;
  c.B = B; c.K0 = K0; c.K1 = K1; c.H0 = H0; c.W0 = W0; c.H1 = H1; c.W1 = W1;
  c.iters = sinkhorn_iterations; c.thr = match_threshold; c.bn_mode = bn_mode; c.precision = precision;
  c.groups = chunk_groups;
  c.m0 = (long long*)matches0; c.m1 = (long long*)matches1; c.ms0 = mscores0; c.ms1 = mscores1;
  return sg_forward(h, c, ws, ws_bytes, false, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int sfm_superglue_hoist_layer0(sfm_handle_t h, const float* desc, const float* kpts, const float* scores, int ldk,
                               int kcap, int n_img, int K, int H, int W, int precision, float* x_store, void* ws,
                               size_t ws_bytes, void* stream) {
  SFM_REQUIRE(h, "sfm_superglue_hoist_layer0: NULL handle");
  SFM_REQUIRE(n_img == 0 || (desc && kpts && scores && x_store), "sfm_superglue_hoist_layer0: NULL argument");
  SFM_REQUIRE(K > 0 && K <= ldk && K <= kcap, "sfm_superglue_hoist_layer0: K exceeds record size");
  SgCall c = shape_call(n_img, K, 0, precision, 1);
  c.desc0 = desc; c.kpts0 = kpts; c.scores0 = scores; c.ldk0 = ldk; c.kcap0 = kcap;
  c.H0 = H; c.W0 = W; c.H1 = H; c.W1 = W;
  c.bn_mode = 1;            // running statistics: the only mode in which layer 0 is independent of the pair
  c.last_layer = 1;
  c.x_out = x_store;
  return sg_forward(h, c, ws, ws_bytes, false, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int sfm_superglue_forward_hoisted(sfm_handle_t h, const float* x_store0, int store_k0, const int* idx0,
                                  const float* x_store1, int store_k1, const int* idx1, int B, int K0, int K1,
                                  int sinkhorn_iterations, float match_threshold, int precision, int64_t* matches0,
                                  int64_t* matches1, float* mscores0, float* mscores1, void* ws, size_t ws_bytes,
                                  void* stream) {
  SFM_REQUIRE(h, "sfm_superglue_forward_hoisted: NULL handle");
  SFM_REQUIRE(B == 0 || (x_store0 && x_store1), "sfm_superglue_forward_hoisted: NULL input");
  SFM_REQUIRE(B == 0 || (matches0 && matches1 && mscores0 && mscores1), "sfm_superglue_forward_hoisted: NULL output");
  SFM_REQUIRE(sinkhorn_iterations >= 0, "sfm_superglue_forward_hoisted: negative sinkhorn_iterations");
  SgCall c = shape_call(B, K0, K1, precision, 1);
  c.xs0 = x_store0; c.xs_k0 = store_k0; c.idx0 = idx0;
  c.xs1 = x_store1; c.xs_k1 = store_k1; c.idx1 = idx1;
  c.first_layer = 1;
  c.bn_mode = 1;
  c.iters = sinkhorn_iterations; c.thr = match_threshold;
  c.m0 = (long long*)matches0; c.m1 = (long long*)matches1; c.ms0 = mscores0; c.ms1 = mscores1;
  return sg_forward(h, c, ws, ws_bytes, false, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int sfm_superglue_tap(int B, int K0, int K1, int precision, int chunk_groups, int which, void* ws, size_t ws_bytes,
                      float** out_ptr, size_t* out_floats) {
  SFM_REQUIRE(out_ptr && out_floats, "sfm_superglue_tap: NULL output");
  SgCall c = shape_call(B, K0, K1, precision, chunk_groups);
  return sg_forward(nullptr, c, ws, ws_bytes, false, nullptr, which, out_ptr, out_floats, 0);
}

// ---- optimal transport ------------------------------------------------------------------------
namespace {
struct OtBuffers {
  float *u, *v, *part, *max0, *max1, *fused, *kb;
  int *counters, *idx0, *idx1;
  int R;
};
void ot_carve(Arena& a, int B, int N, int M, OtBuffers& b) {
  b.u = a.get<float>((size_t)B * (N + 1));
  b.v = a.get<float>((size_t)B * (M + 1));
  b.part = a.get<float>(sinkhorn_partial_floats(B, N, M, &b.R));
  b.counters = a.get<int>(sinkhorn_counter_ints(B, M));
  b.max0 = a.get<float>((size_t)B * N);
  b.max1 = a.get<float>((size_t)B * M);
  b.idx0 = a.get<int>((size_t)B * N);
  b.idx1 = a.get<int>((size_t)B * M);
  b.fused = a.get<float>(sinkhorn_fused_floats(B, N, M));
  b.kb = a.get<float>((size_t)B * N);
}
int ot_prepare(const float* scores, int B, int N, int M, float alpha, int iters, void* ws, size_t ws_bytes,
               OtBuffers& b, SinkhornArgs& s, cudaStream_t st, bool scaled = false) {
  SFM_REQUIRE(scores && B >= 0 && N > 0 && M > 0 && iters >= 0, "sinkhorn: bad arguments (B=%d N=%d M=%d iters=%d)", B, N,
              M, iters);
  Arena a(ws, ws_bytes);
  ot_carve(a, B, N, M, b);
  if (a.overflow) {
    set_error("sinkhorn: workspace too small (%zu bytes given, %zu needed)", ws_bytes, a.off);
    return SFM_ERR_WORKSPACE;
  }
  SFM_CUDA(cudaMemsetAsync(b.counters, 0, sinkhorn_counter_ints(B, M) * sizeof(int), st));
  s.S = scores; s.B = B; s.N = N; s.M = M; s.alpha = alpha; s.iters = iters;
  s.u = b.u; s.v = b.v; s.part = b.part; s.counters = b.counters; s.R = b.R;
  if (scaled) {
    s.fast = true;
    s.fused_part = b.fused;
    s.kb = b.kb;
    s.col_idx = b.idx1;
  }
  return sinkhorn_run(s, st);
}
}  // namespace

size_t sfm_sinkhorn_workspace_bytes(int B, int N, int M) {
  Arena a(nullptr, 0);
  OtBuffers b;
  ot_carve(a, B, N, M, b);
  return a.off;
}

int sfm_log_optimal_transport(const float* scores, int B, int N, int M, float alpha, int iters, float* Z, void* ws,
                              size_t ws_bytes, void* stream) {
  SFM_REQUIRE(Z, "sfm_log_optimal_transport: NULL output");
  OtBuffers b;
  SinkhornArgs s;
  SFM_TRY(ot_prepare(scores, B, N, M, alpha, iters, ws, ws_bytes, b, s, (cudaStream_t)stream));
  return sinkhorn_write_Z(s, Z, (cudaStream_t)stream);
}

int sfm_log_sinkhorn_iterations(const float* Z, const float* log_mu, const float* log_nu, int B, int m, int n, int iters,
                                float* out, void* ws, size_t ws_bytes, void* stream) {
  SFM_REQUIRE(B == 0 || (Z && log_mu && log_nu && out && ws), "sfm_log_sinkhorn_iterations: NULL argument");
  SFM_REQUIRE(B >= 0 && m > 0 && n > 0, "sfm_log_sinkhorn_iterations: bad dims");
  const size_t need = ((size_t)B * m + (size_t)B * n) * sizeof(float);
  if (ws_bytes < need) {
    set_error("sfm_log_sinkhorn_iterations: workspace too small (%zu bytes given, %zu needed)", ws_bytes, need);
    return SFM_ERR_WORKSPACE;
  }
  float* u = (float*)ws;
  return log_sinkhorn_general(Z, log_mu, log_nu, B, m, n, iters, out, u, u + (size_t)B * m, (cudaStream_t)stream);
}

int sfm_sinkhorn_match(const float* scores, int B, int N, int M, float alpha, int iters, float match_threshold,
                       int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1, void* ws,
                       size_t ws_bytes, void* stream) {
  SFM_REQUIRE(matches0 && matches1 && mscores0 && mscores1, "sfm_sinkhorn_match: NULL output");
  OtBuffers b;
  SinkhornArgs s;
  SFM_TRY(ot_prepare(scores, B, N, M, alpha, iters, ws, ws_bytes, b, s, (cudaStream_t)stream));
  return sinkhorn_match(s, match_threshold, b.max0, b.idx0, b.max1, b.idx1, (long long*)matches0, (long long*)matches1,
                        mscores0, mscores1, (cudaStream_t)stream);
}

size_t sfm_sinkhorn_half_workspace_bytes(int B, int N, int M) {
  if (B < 0 || N <= 0 || M <= 0) return 0;
  return sfm_sinkhorn_workspace_bytes(B, N, M) + align_up((size_t)B * N * M * 2, 256);
}

int sfm_sinkhorn_match_scaled_half(const float* scores, int B, int N, int M, float alpha, int iters,
                                   float match_threshold, int64_t* matches0, int64_t* matches1, float* mscores0,
                                   float* mscores1, void* ws, size_t ws_bytes, void* stream) {
  SFM_REQUIRE(matches0 && matches1 && mscores0 && mscores1 && scores, "sfm_sinkhorn_match_scaled_half: NULL argument");
  SFM_REQUIRE(M % 4 == 0 && M <= 4096 && iters > 0 && B >= 0 && N > 0,
              "sfm_sinkhorn_match_scaled_half: needs M %% 4 == 0, M <= 4096, iters > 0");
  const size_t base = sfm_sinkhorn_workspace_bytes(B, N, M);
  SFM_REQUIRE(ws_bytes >= sfm_sinkhorn_half_workspace_bytes(B, N, M), "sfm_sinkhorn_match_scaled_half: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Arena a(ws, base);
  OtBuffers b;
  ot_carve(a, B, N, M, b);
  SFM_CUDA(cudaMemsetAsync(b.counters, 0, sinkhorn_counter_ints(B, M) * sizeof(int), st));
  SinkhornArgs s;
  s.S = scores; s.B = B; s.N = N; s.M = M; s.alpha = alpha; s.iters = iters;
  s.u = b.u; s.v = b.v; s.part = b.part; s.counters = b.counters; s.R = b.R;
  s.fast = true;
  s.fused_part = b.fused;
  s.kb = b.kb;
  s.col_idx = b.idx1;
  s.rowmax = b.max0;                                                   // scores stay intact
  s.k16 = reinterpret_cast<__half*>(static_cast<char*>(ws) + base);    // the iterations stream this copy
  SFM_TRY(sinkhorn_run(s, st));
  return sinkhorn_match(s, match_threshold, b.max0, b.idx0, b.max1, b.idx1, (long long*)matches0, (long long*)matches1,
                        mscores0, mscores1, st);
}

int sfm_sinkhorn_match_scaled(float* scores, int B, int N, int M, float alpha, int iters, float match_threshold,
                              int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1, void* ws,
                              size_t ws_bytes, void* stream) {
  SFM_REQUIRE(matches0 && matches1 && mscores0 && mscores1, "sfm_sinkhorn_match_scaled: NULL output");
  SFM_REQUIRE(M % 4 == 0 && M <= 4096 && iters > 0, "sfm_sinkhorn_match_scaled: needs M %% 4 == 0, M <= 4096, iters > 0");
  OtBuffers b;
  SinkhornArgs s;
  SFM_TRY(ot_prepare(scores, B, N, M, alpha, iters, ws, ws_bytes, b, s, (cudaStream_t)stream, true));
  return sinkhorn_match(s, match_threshold, b.max0, b.idx0, b.max1, b.idx1, (long long*)matches0, (long long*)matches1,
                        mscores0, mscores1, (cudaStream_t)stream);
}

// ---- stand-alone attention in the reference layout (superglue.py:85-89), fp32 FFMA path
namespace {
struct AttBuffers { float *qT, *kT, *vT, *S, *oT; int ldS; };
size_t att_carve(int B, int N, int M, void* ws, size_t ws_bytes, AttBuffers& b) {
  Arena a(ws, ws_bytes);
  b.ldS = (M + 3) / 4 * 4;
  b.qT = a.get<float>((size_t)B * 4 * N * 64);
  b.kT = a.get<float>((size_t)B * 4 * M * 64);
  b.vT = a.get<float>((size_t)B * 4 * M * 64);
  b.oT = a.get<float>((size_t)B * 4 * N * 64);
  b.S = a.get<float>((size_t)B * 4 * N * b.ldS);
  return a.off;
}
}  // namespace

size_t sfm_attention_workspace_bytes(int B, int N, int M) {
  if (B < 0 || N < 0 || M < 0) return 0;
  AttBuffers b;
  return att_carve(B, N, M, nullptr, 0, b);
}

int sfm_attention(const float* query, const float* key, const float* value, int B, int N, int M, float* out,
                  float* prob, void* ws, size_t ws_bytes, void* stream) {
  SFM_REQUIRE(B >= 0 && N >= 0 && M > 0, "sfm_attention: B, N >= 0 and M > 0 required (B=%d N=%d M=%d)", B, N, M);
  if (B == 0 || N == 0) return SFM_OK;
  SFM_REQUIRE(query && key && value && out && ws, "sfm_attention: NULL argument");
  SFM_REQUIRE(B * 4 <= 65535, "sfm_attention: B too large");
  AttBuffers b;
  size_t need = att_carve(B, N, M, ws, ws_bytes, b);
  SFM_REQUIRE(ws_bytes >= need, "sfm_attention: workspace too small (%zu < %zu)", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  // (B,64,4,L): element (b,d,h,l) at b*256L + d*4L + h*L + l  ->  [b][h][l][d]
  SFM_TRY(transpose_batched_f32(query, b.qT, B * 4, 4, 64, N, 4ll * N, 64, 256ll * N, N, 256ll * N, 64ll * N, st));
  SFM_TRY(transpose_batched_f32(key, b.kT, B * 4, 4, 64, M, 4ll * M, 64, 256ll * M, M, 256ll * M, 64ll * M, st));
  SFM_TRY(transpose_batched_f32(value, b.vT, B * 4, 4, 64, M, 4ll * M, 64, 256ll * M, M, 256ll * M, 64ll * M, st));
  GemmF32 g;
  g.A = b.qT; g.lda = 64;
  g.B = b.kT; g.ldb = 64;
  g.C = b.S; g.ldc = b.ldS;
  g.M = N; g.N = M; g.K = 64; g.alpha = 0.125f;
  g.batch = B * 4; g.nb1 = 1;
  g.sA0 = 64ll * N; g.sB0 = 64ll * M; g.sC0 = (long long)N * b.ldS;
  SFM_TRY(gemm_f32_nt(g, st));
  SFM_TRY(softmax_rows_f32(b.S, (long long)B * 4 * N, M, b.ldS, st));
  if (prob) SFM_TRY(copy_rows_f32(b.S, prob, (long long)B * 4 * N, M, b.ldS, M, st));
  GemmF32 p;
  p.A = b.S; p.lda = b.ldS;
  p.B = b.vT; p.ldb = 64;
  p.C = b.oT; p.ldc = 64;
  p.M = N; p.N = 64; p.K = M;
  p.batch = B * 4; p.nb1 = 1;
  p.sA0 = (long long)N * b.ldS; p.sB0 = 64ll * M; p.sC0 = 64ll * N;
  SFM_TRY(gemm_f32_nn(p, st));
  return transpose_batched_f32(b.oT, out, B * 4, 4, N, 64, 64, 4ll * N, 256ll * N, 64ll * N, 256ll * N, N, st);
}

int sfm_compact_matches(const int64_t* matches0, int B, int K0, uint32_t* tables, int* lengths, void* stream) {
  SFM_REQUIRE(B == 0 || (matches0 && tables && lengths), "sfm_compact_matches: NULL argument");
  return compact_matches((const long long*)matches0, B, K0, tables, lengths, (cudaStream_t)stream);
}

int sfm_pack_tables(const uint32_t* tables, int n_pairs, int kcap, const int* lengths, const int64_t* offsets,
                    uint32_t* out, void* stream) {
  SFM_REQUIRE(n_pairs == 0 || (tables && lengths && offsets && out), "sfm_pack_tables: NULL argument");
  SFM_REQUIRE(n_pairs >= 0 && kcap > 0, "sfm_pack_tables: bad dims");
  return pack_tables(tables, n_pairs, kcap, lengths, (const long long*)offsets, out, (cudaStream_t)stream);
}

int sfm_pair_list(int n_images, long long first, long long count, int* idx0, int* idx1, void* stream) {
  const long long total = (long long)n_images * (n_images - 1) / 2;
  SFM_REQUIRE(n_images >= 0 && first >= 0 && count >= 0 && first + count <= total, "sfm_pair_list: range outside the %lld pairs of %d images", total, n_images);
  SFM_REQUIRE(count == 0 || (idx0 && idx1), "sfm_pair_list: NULL output");
  return pair_list_range(n_images, first, count, idx0, idx1, (cudaStream_t)stream);
}

int sfm_sparse_depth(const double* xyz, int n_points, const int* entry_cam, const int* entry_point, long long n_entries,
                     const double* extrinsics, const double* intrinsics, int n_cams, int W, int H, double* depth,
                     double* bounds, long long* winner_ws, int* error_flag, void* stream) {
  SFM_REQUIRE(n_cams >= 0 && W > 0 && H > 0 && n_points >= 0 && n_entries >= 0, "sfm_sparse_depth: bad dims");
  SFM_REQUIRE(n_cams == 0 || (extrinsics && intrinsics && depth && bounds && winner_ws && error_flag), "sfm_sparse_depth: NULL argument");
  SFM_REQUIRE(n_entries == 0 || (xyz && entry_cam && entry_point), "sfm_sparse_depth: NULL entries");
  return sparse_depth_project(xyz, entry_cam, entry_point, n_entries, extrinsics, intrinsics, n_cams, W, H, depth,
                              bounds, winner_ws, error_flag, (cudaStream_t)stream);
}

// ---- kernel-level test hooks of the tcgen05 path ------------------------------------------------
int sfm_tc_gemm(sfm_handle_t h, const void* A, const void* A2, const void* W, int M, int N, int K, int K1,
                const float* bias, float alpha, int relu, void* out_bf16, float* out_f32, int accumulate,
                void* stream) {
  SFM_REQUIRE(h && A && W, "sfm_tc_gemm: NULL argument");
  tc::TcGemm g;
  g.A = (const __nv_bfloat16*)A; g.lda = K1; g.a_rows = M;
  g.A2 = (const __nv_bfloat16*)A2; g.lda2 = K - K1;
  g.W = (const __nv_bfloat16*)W; g.ldw = K; g.w_rows = N;
  g.M = M; g.N = N; g.K = K; g.K1 = K1;
  g.bias = bias; g.alpha = alpha; g.relu = relu;
  g.out_h = (__nv_bfloat16*)out_bf16; g.out_h_ld = N;
  g.out_f = out_f32; g.out_f_ld = N; g.accumulate_f = accumulate;
  g.tmaps = h->tmaps;
  return tc::tc_gemm(g, h->sm_count, (cudaStream_t)stream);
}

int sfm_tc_attention(sfm_handle_t h, const void* qkv, long long rows_total, int B, int Nq0, int Nk0, long long q_row0_0,
                     long long k_row0_0, int Nq1, int Nk1, long long q_row0_1, long long k_row0_1, void* out_bf16,
                     void* stream) {
  SFM_REQUIRE(h && qkv && out_bf16, "sfm_tc_attention: NULL argument");
  tc::TcAttn a;
  a.qkv = (const __nv_bfloat16*)qkv; a.rows_total = rows_total; a.out = (__nv_bfloat16*)out_bf16; a.B = B;
  a.Nq[0] = Nq0; a.Nk[0] = Nk0; a.q_row0[0] = q_row0_0; a.k_row0[0] = k_row0_0;
  a.Nq[1] = Nq1; a.Nk[1] = Nk1; a.q_row0[1] = q_row0_1; a.k_row0[1] = k_row0_1;
  a.tmaps = h->tmaps;
  a.redo = h->att_redo; a.redo_cap = tc::ATT_REDO_CAP;
  return tc::tc_attention(a, h->sm_count, (cudaStream_t)stream);
}

int sfm_tc_conv3x3(sfm_handle_t h, const void* in_nhwc, const void* w, const float* bias, void* out_nhwc, int n, int H,
                   int W, int Cin, int Cout, int relu, void* stream) {
  SFM_REQUIRE(h && in_nhwc && w && out_nhwc, "sfm_tc_conv3x3: NULL argument");
  tc::TcConv p;
  p.in = (const __nv_bfloat16*)in_nhwc; p.w = (const __nv_bfloat16*)w; p.bias = bias; p.out = (__nv_bfloat16*)out_nhwc;
  p.n = n; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.relu = relu & 1; p.pool = (relu >> 1) & 1;
  return tc::tc_conv3x3(p, h->sm_count, (cudaStream_t)stream);
}

// ---- SimpleMatcher ----------------------------------------------------------------------------
namespace {
struct SmBuffers {
  float *x1, *x2, *dmat, *u, *v, *part, *max0, *max1;
  int *counters, *idx0, *idx1;
  int R;
};
void sm_carve(Arena& a, int N, int M, SmBuffers& b) {
  b.x1 = a.get<float>((size_t)N * 256);
  b.x2 = a.get<float>((size_t)M * 256);
  b.dmat = a.get<float>((size_t)N * M);
  b.u = a.get<float>(N + 1);
  b.v = a.get<float>(M + 1);
  b.part = a.get<float>(sinkhorn_partial_floats(1, N, M, &b.R));
  b.counters = a.get<int>(sinkhorn_counter_ints(1, M));
  b.max0 = a.get<float>(N);
  b.max1 = a.get<float>(M);
  b.idx0 = a.get<int>(N);
  b.idx1 = a.get<int>(M);
}
}  // namespace

size_t sfm_simple_match_workspace_bytes(int N, int M) {
  Arena a(nullptr, 0);
  SmBuffers b;
  sm_carve(a, N, M, b);
  return a.off;
}

int sfm_simple_match(const float* desc1, int ld1, int N, const float* desc2, int ld2, int M, float nn_thresh,
                     float* out, int out_ld, int* count, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SFM_REQUIRE(desc1 && desc2 && out && count, "sfm_simple_match: NULL argument");
  SFM_REQUIRE(N > 0 && M > 0, "sfm_simple_match: empty descriptor set");
  if (nn_thresh < 0.f) {
    set_error("'nn_thresh' should be non-negative");
    return SFM_ERR_INVALID;
  }
  Arena a(ws, ws_bytes);
  SmBuffers b;
  sm_carve(a, N, M, b);
  if (a.overflow) {
    set_error("simple_match: workspace too small (%zu bytes given, %zu needed)", ws_bytes, a.off);
    return SFM_ERR_WORKSPACE;
  }
  SFM_CUDA(cudaMemsetAsync(b.counters, 0, sinkhorn_counter_ints(1, M) * sizeof(int), st));
  SFM_CUDA(cudaMemsetAsync(b.u, 0, (N + 1) * sizeof(float), st));
  SFM_CUDA(cudaMemsetAsync(b.v, 0, (M + 1) * sizeof(float), st));
  SFM_TRY(transpose_to_tokens(desc1, ld1, N, b.x1, st));
  SFM_TRY(transpose_to_tokens(desc2, ld2, M, b.x2, st));
  GemmF32 g;
  g.A = b.x1; g.lda = 256; g.B = b.x2; g.ldb = 256; g.C = b.dmat; g.ldc = M;
  g.M = N; g.N = M; g.K = 256;
  SFM_TRY(gemm_f32_nt(g, st));
  SFM_TRY(neg_l2_from_dot(b.dmat, (long long)N * M, st));
  SinkhornArgs s;
  s.S = b.dmat; s.B = 1; s.N = N; s.M = M; s.alpha = 0.f; s.iters = 0;
  s.u = b.u; s.v = b.v; s.part = b.part; s.counters = b.counters; s.R = b.R;
  SFM_TRY(argmax_rows_cols(s, 0.f, b.max0, b.idx0, b.max1, b.idx1, st));
  return simple_mutual(b.max0, b.idx0, b.idx1, N, nn_thresh, out, out_ld, count, st);
}

}  // extern "C"
