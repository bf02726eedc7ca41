// Train-mode BatchNorm1d statistics + fused normalise/ReLU (reference superglue.py:56-57 as run:
// batch statistics over all (B x N) columns of one side, biased variance, eps 1e-5).
// Two token groups (side 0 rows [0,T0), side 1 rows [T0,T0+T1)) get separate statistics, exactly
// like the reference's two separate module calls (superglue.py:254-255, 137).
#pragma once
#include "common.cuh"

namespace sfm {

// Rows [0,T0) belong to side 0 and [T0,T0+T1) to side 1; each side is cut into G equal statistic
// groups (G > 1 when several Matcher chunks are fused into one launch: every chunk keeps its own
// batch statistics, exactly as if it had been run alone).
struct BnGroups {
  long long T0, T1;
  int G;
};
int bn_stat_blocks(const BnGroups& g);
// partial: [bn_stat_blocks, C, 2] floats; scale/shift: [2 * G, C]  (y = relu(x * scale + shift))
template <typename T>
int bn_train_stats(const T* X, int ld, int C, const BnGroups& g, const float* gamma, const float* beta, float eps,
                   float* partial, float* scale, float* shift, cudaStream_t st);
// scale/shift from partials produced elsewhere (the tcgen05 GEMM epilogue): [rows / rb, C, 2] (mean, centred M2)
// per rb-row block, block index = global row / rb
int bn_finalize(const float* partial, int rb, int C, const BnGroups& g, const float* gamma, const float* beta, float eps,
                float* scale, float* shift, cudaStream_t st);
template <typename T>
int bn_apply_relu(T* X, int ld, int C, const BnGroups& g, const float* scale, const float* shift, cudaStream_t st);

// out (bf16, ld_out) = relu(X (f32, ld) * scale + shift): hands a fp32 layer to the tensor-core path
int bn_apply_relu_f32_to_bf16(const float* X, int ld, __nv_bfloat16* out, int ld_out, int C, const BnGroups& g,
                              const float* scale, const float* shift, cudaStream_t st);

}  // namespace sfm
