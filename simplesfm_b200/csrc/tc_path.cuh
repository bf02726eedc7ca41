// bf16 tcgen05/TMEM/TMA path of the SuperGlue GNN (throughput mode).  Internal.
#pragma once
#include "context.cuh"

namespace sfm {

struct TcSgBuffers {
  __nv_bfloat16 *xh, *qkv, *msgp, *msg, *hid, *mdh;
  float *hid32, *stats, *mean, *var, *t0, *t1;
};

void tc_sg_carve(Arena& a, const SgCall& c, TcSgBuffers& b);
// X: fp32 master descriptors [T,256] (frontend output), kin [T,4]; writes scores [B,K0,K1] f32
int tc_sg_forward(sfm_ctx* h, const SgCall& c, float* X, const float* kin, float* scores, const TcSgBuffers& b,
                  cudaStream_t st);

}  // namespace sfm
