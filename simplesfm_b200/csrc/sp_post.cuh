// SuperPoint post-processing kernels (see sp_post.cu for the reference citations).
#pragma once
#include "common.cuh"

namespace sfm {

int sp_softmax_shuffle(const float* logits, int ldl, float* out, int n_img, int h, int w, cudaStream_t st);
int sp_nms(const float* scores, float* out, int n_img, int H, int W, int radius, cudaStream_t st);
size_t sp_select_workspace_ints(int n_img, int H, int W);
int sp_select(const float* nms, const unsigned char* mask, int n_img, int H, int W, float thr, int border,
              int* block_tmp, int cap, int* cand_pix, float* cand_score, int* counts, cudaStream_t st);
int sp_rank_sort(const int* cand_pix, const float* cand_score, const int* counts, int n_img, int cap, int max_kp,
                 int W, int out_cap, float* kpts_xy, float* scores_out, int* out_counts, cudaStream_t st);
int sp_sample_descriptors(const float* dmap, int n_img, int h, int w, const float* kpts_xy, const int* counts,
                          int kp_cap, int align_corners, float* out, int ld_out, cudaStream_t st);

// 8-bit grey image -> float32 / 255 (bit-identical to the reference's host conversion, matcher.py:360)
int sp_u8_to_unit_f32(const unsigned char* in, float* out, long long n, cudaStream_t st);

}  // namespace sfm
