// Handle: packed weights resident on one device.  Internal.
#pragma once
#include <vector>
#include "common.cuh"

namespace sfm {

struct SpConv {
  int cin, cout, k;
  float* w;  // [cout][k*k*cin]  (ky, kx, ci) K-major  -- implicit-GEMM B operand
  float* b;  // [cout]
  __nv_bfloat16* w_h;  // bf16 copy of w (tcgen05 path)
};

struct SgBn {
  float *gamma, *beta, *rmean, *rvar;
};

struct SgLayer {
  float* Wqkv;   // [768, 256] rows: q (head-contiguous), k, v
  float* bqkv;   // [768]
  float* Wm;     // [256, 256] input columns head-contiguous
  float* bm;
  float* W1;     // [512, 512]
  float* b1;
  SgBn bn1;
  float* W1f;    // eval-mode BN folded into W1/b1
  float* b1f;
  float* W2;     // [256, 512]
  float* b2;
  // bf16 copies for the tcgen05 path
  __nv_bfloat16 *Wqkv_h, *Wm_h, *W1_h, *W1f_h, *W2_h;
  // merge conv folded into the MLP input layer: mlp0([x | merge(a)]) = [W1x | W1m Wm] [x | a] + (b1 + W1m bm)
  __nv_bfloat16 *W1c_h, *W1cf_h;   // [512, 512] train / eval-folded
  float *b1c, *b1cf;
};

struct SgWeights {
  bool loaded = false;
  float bin_score = 1.f;
  int kch[6] = {4, 32, 64, 128, 256, 256};  // first layer K padded 3 -> 4
  float* kW[5];
  float* kb[5];
  SgBn kbn[4];
  __nv_bfloat16 *kW3_h, *kW3f_h, *kW4_h;  // bf16 copies of the two wide encoder layers (tensor-core path)
  float* kWf[4];  // eval-folded
  float* kbf[4];
  SgLayer L[18];
  float* Wf;
  float* bf;
  __nv_bfloat16* Wf_h;
};

struct SpWeights {
  bool loaded = false;
  SpConv conv[12];
};

}  // namespace sfm

enum { SFM_PROF_LINEAR = 0, SFM_PROF_ATTENTION = 1, SFM_PROF_SINKHORN = 2, SFM_PROF_OTHER = 3, SFM_PROF_KINDS = 4 };

namespace sfm { namespace tc { struct TmapCache; } }

struct sfm_ctx {
  int device = 0;
  sfm::tc::TmapCache* tmaps = nullptr;   // encoded TMA tensor maps of this handle's launches
  int* att_redo = nullptr;               // attention rows handed to the exact recomputation (tc_attention.cu)
  int sm_count = 148;
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[SFM_PROF_KINDS];  // recorded this window
  std::vector<cudaEvent_t> event_pool;
  sfm::SpWeights sp;
  sfm::SgWeights sg;
  std::vector<void*> allocs;  // everything cudaMalloc'ed by the handle
};

namespace sfm {
// CUDA-event bracket around a group of launches on `st` (only when profiling is enabled)
struct ProfScope {
  sfm_ctx* h;
  int kind;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(sfm_ctx* h_, int kind_, cudaStream_t st_);
  ~ProfScope();
};
int sp_forward_detect(sfm_ctx* h, const float* images, int n_img, int H, int W, const unsigned char* mask, float thr,
                      int nms_radius, int border, int precision, int* cand_counts, void* ws, size_t ws_bytes,
                      bool dry, size_t* need, cudaStream_t st);
int sp_forward_describe(sfm_ctx* h, int n_img, int H, int W, int max_kp, int align_corners, int precision, int kp_cap,
                        float* kpts_xy, float* scores, float* desc, int* counts, void* ws, size_t ws_bytes,
                        cudaStream_t st);
int sp_tap(int n_img, int H, int W, int precision, int which, void* ws, size_t ws_bytes, float** out, size_t* n);

struct SgCall {
  const float *desc0, *kpts0, *scores0;
  int ldk0, kcap0;
  const int* idx0;
  const float *desc1, *kpts1, *scores1;
  int ldk1, kcap1;
  const int* idx1;
  int B, K0, K1, H0, W0, H1, W1;
  int iters;
  float thr;
  int bn_mode, precision;
  int groups;  // number of BatchNorm statistic groups (fused Matcher chunks); B % groups == 0
  long long *m0, *m1;
  float *ms0, *ms1;
  // Layer-0 hoisting (eval-mode BatchNorm only): GNN layer 0 is a SELF layer, so desc + kenc and the
  // layer-0 update depend on one image alone and can be computed once per image instead of once per pair.
  //   last_layer < 18 : stop after GNN layer `last_layer - 1` and copy the residual stream to x_out (K1 may be 0)
  //   xs0 / xs1 != 0  : per-image stores [n_img, xs_k, 256] f32 of the residual stream entering layer `first_layer`;
  //                     the pair batch is gathered from them (idx0 / idx1) instead of running the front end
  int first_layer = 0, last_layer = 18;
  const float *xs0 = nullptr, *xs1 = nullptr;
  int xs_k0 = 0, xs_k1 = 0;
  float* x_out = nullptr;
};
// dry = true: only compute the workspace size / tap offsets (no launches, h may be null)
int sg_forward(sfm_ctx* h, const SgCall& c, void* ws, size_t ws_bytes, bool dry, size_t* need, int tap_which,
               float** tap_ptr, size_t* tap_floats, cudaStream_t st);
}  // namespace sfm
