// SuperGlue element-wise / reduction kernels shared by the fp32 and bf16 paths.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace sfm {

// Gather one batch of pairs from per-image feature records into token-major activations.
//   desc  : per image channel-major (256, ldk) f32 (reference Frame layout), image stride img_stride
//   kpts  : per image (kcap, 2) f32; scores: per image (kcap,)
//   idx   : device int32 [B] image index of each pair's side, or nullptr for identity
//   X     : [B*K, 256] token-major descriptors (f32)
//   kin   : [B*K, 4]   (x_norm, y_norm, score, 0)  -- keypoint-encoder input (superglue.py:62-82)
int sg_frontend(const float* desc, long long desc_img_stride, int ldk, const float* kpts, const float* scores,
                int kcap, const int* idx, int B, int K, int imgH, int imgW, float* X, float* kin,
                cudaStream_t st);

// pair batch gathered from a per-image token-major store [n_img, store_k, 256] (layer-0 hoisting)
int gather_tokens(const float* store, int store_k, const int* idx, int B, int K, float* X, cudaStream_t st);

// in-place softmax over the last dim of [rows, M] (ld)
int softmax_rows_f32(float* S, long long rows, int M, int ld, cudaStream_t st);
// batched 2-D transpose, batch index z -> offsets bs?0 * (z / nb1) + bs?1 * (z % nb1)
int transpose_batched_f32(const float* in, float* out, int batch, int nb1, int R, int C, long long ldi,
                          long long ldo, long long bsi0, long long bsi1, long long bso0, long long bso1,
                          cudaStream_t st);
int copy_rows_f32(const float* in, float* out, long long rows, int M, int ldi, int ldo, cudaStream_t st);

// ---- optimal transport + matching (superglue.py:142-171, 273-283)
struct SinkhornArgs {
  const float* S;   // [B, N, M] scores (already divided by sqrt(d))
  int B, N, M;
  float alpha;      // bin_score
  int iters;
  float* u;         // [B, N+1]
  float* v;         // [B, M+1]
  float* part;      // [B, R, M+1, 2] column partials
  int* counters;    // [B, colblocks] zero-initialised, self-resetting
  int R;            // row splits of the column pass
  float* fused_part = nullptr;  // 3 x [B, ceil(N/64), M]: column partial sums (+ maxima, rows in the last sweep)
  float* rowmax = nullptr;      // optional [B, N]: with k16, keeps S intact and stores the row maxima here (pass max0)
  int* col_idx = nullptr;       // optional [B, M]: column arg-max of the assignment, written by the last sweep
  float* kb = nullptr;          // [B, N] dustbin-column factors exp(alpha - rowmax) of the scaling form
  __half* k16 = nullptr;        // optional [B, N, M] fp16 copy of the scaling-form matrix the iterations stream
  int sub_batch = 0;  // pairs per L2-resident sweep group (0 = auto)
  bool fast = false;  // ex2.approx exponentials (bf16 precision mode) instead of IEEE expf
};
size_t sinkhorn_partial_floats(int B, int N, int M, int* R_out);
size_t sinkhorn_counter_ints(int B, int M);
size_t sinkhorn_fused_floats(int B, int N, int M);
bool sinkhorn_use_scaling(const SinkhornArgs& a);
int sinkhorn_run(const SinkhornArgs& a, cudaStream_t st);
// mutual nearest neighbour + threshold from S, u, v  (never materialises the (N+1)x(M+1) matrix)
int sinkhorn_match(const SinkhornArgs& a, float match_threshold, float* max0, int* idx0, float* max1, int* idx1,
                   long long* matches0, long long* matches1, float* mscores0, float* mscores1, cudaStream_t st);
int argmax_rows_cols(const SinkhornArgs& a, float norm, float* max0, int* idx0, float* max1, int* idx1,
                     cudaStream_t st);
// SimpleMatcher pieces (simple_matcher.py:66-114)
int transpose_to_tokens(const float* desc, int ld, int N, float* X, cudaStream_t st);
int neg_l2_from_dot(float* S, long long n, cudaStream_t st);
int simple_mutual(const float* max0, const int* idx0, const int* idx1, int N, float thr, float* out, int ld,
                  int* count, cudaStream_t st);
// optional: write the full (B, N+1, M+1) log-assignment (what log_optimal_transport returns)
int sinkhorn_write_Z(const SinkhornArgs& a, float* Z, cudaStream_t st);

// log_sinkhorn_iterations on arbitrary couplings / marginals (superglue.py:142-148); u [B,m], v [B,n] scratch
int log_sinkhorn_general(const float* Z, const float* log_mu, const float* log_nu, int B, int m, int n, int iters,
                         float* out, float* u, float* v, cudaStream_t st);

// sparse depth maps from a COLMAP sparse model (sparse_depth.cu): depth [n_cams, W, H] f64 (-1 = empty)
int sparse_depth_project(const double* xyz, const int* entry_cam, const int* entry_point, long long n_entries,
                         const double* extrinsics, const double* intrinsics, int n_cams, int W, int H, double* depth,
                         double* bounds, long long* winner_ws, int* error_flag, cudaStream_t st);

// lexicographic (i < j) pairs [first, first + count) of n images as two int32 index arrays
int pair_list_range(int n, long long first, long long count, int* idx0, int* idx1, cudaStream_t st);
// (L,2) uint32 match-table compaction per pair: rows [i, matches0[i]] for matches0[i] > -1
int compact_matches(const long long* matches0, int B, int N, unsigned* tables, int* lengths, cudaStream_t st);
// flatten per-pair (kcap, 2) tables into one (sum lengths, 2) array; offsets = exclusive prefix sum of lengths
int pack_tables(const unsigned* tables, int n_pairs, int kcap, const int* lengths, const long long* offsets,
                unsigned* out, cudaStream_t st);

}  // namespace sfm
