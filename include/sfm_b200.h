/* libsfm_b200 -- C ABI of the B200-native SuperPoint -> SuperGlue matching path.
 *
 * The reference (palsol/SimpleSfm) has no FFI layer: its boundary for this path is the Python
 * API of `simple_sfm.matchers` / `simple_sfm.models.{superpoint,superglue}`.  Each entry point
 * below names the reference interface it replaces (paths relative to the reference root).  The
 * host-side mirror of those Python classes lives in `simplesfm_b200/` and binds this header
 * with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative SFM_ERR_* code; sfm_last_error() gives
 *     a thread-local message.  No C++ exception crosses the boundary.
 *   - all tensor arguments are DEVICE pointers unless the name ends in `_host`; the caller
 *     (torch) owns inputs, outputs and workspaces.  The handle owns only the packed weights.
 *   - every compute entry point is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - there is no CPU fallback: without a CUDA device sfm_create fails.
 */
#ifndef SFM_B200_H_
#define SFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFM_ABI_VERSION 1

#if defined(__GNUC__)
#define SFM_API __attribute__((visibility("default")))
#else
#define SFM_API
#endif

#define SFM_OK 0
#define SFM_ERR_INVALID (-1)
#define SFM_ERR_CUDA (-2)
#define SFM_ERR_WORKSPACE (-3)
#define SFM_ERR_STATE (-4)
#define SFM_ERR_UNSUPPORTED (-5)

/* BatchNorm1d behaviour inside SuperGlue's MLPs (superglue.py:56-57).  The reference never calls
 * .eval(), so as run it uses batch statistics (SFM_BN_TRAIN, the default of the mirror). */
#define SFM_BN_TRAIN 0
#define SFM_BN_EVAL 1

/* arithmetic of the contraction kernels */
#define SFM_PREC_FP32 0 /* FFMA, reproduces the reference's fp32 results (bit-exact matches) */
#define SFM_PREC_BF16 1 /* tcgen05 bf16 operands, fp32 accumulate (throughput mode)          */

typedef struct sfm_ctx* sfm_handle_t;

SFM_API int sfm_abi_version(void);
SFM_API const char* sfm_last_error(void);

/* one handle per GPU; holds packed weights only */
SFM_API int sfm_create(int device, sfm_handle_t* out);
SFM_API int sfm_destroy(sfm_handle_t h);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------
 * sfm_launch_count: kernels launched by this library in the process so far.
 * profiling: when enabled, CUDA events bracket every launch group of kind 0 = linear layers (GEMMs),
 * 1 = attention, 2 = Sinkhorn + matching, 3 = everything else, on the caller's stream;
 * sfm_get_profile synchronises, returns the summed milliseconds and group counts per kind
 * (arrays of 4) since the previous call, and clears the window. */
SFM_API long long sfm_launch_count(void);
SFM_API int sfm_set_profiling(sfm_handle_t h, int enable);
SFM_API int sfm_get_profile(sfm_handle_t h, float* ms, long long* groups);

/* ---- weights -------------------------------------------------------------------------------
 * SuperPoint.__init__ -> load_state_dict (superpoint.py:101).  `packed_host`: the 12 convs in the
 * order conv1a,1b,2a,2b,3a,3b,4a,4b,Pa,Pb,Da,Db, each as weight (Cout,Cin,kh,kw) then bias, f32,
 * 1 300 865 floats in total. */
SFM_API int sfm_load_superpoint(sfm_handle_t h, const float* packed_host, size_t n_floats);
SFM_API size_t sfm_superpoint_packed_floats(void);

/* SuperGlue.__init__ -> load_state_dict (superglue.py:233).  `packed_host` order:
 *   bin_score; kenc.encoder.{0,3,6,9,12} as (weight, bias) each followed (except the last) by its
 *   BatchNorm (weight, bias, running_mean, running_var); for L in 0..17: proj.0, proj.1, proj.2,
 *   merge (weight, bias each), mlp.0 (weight, bias), mlp.1 BN (weight, bias, running_mean,
 *   running_var), mlp.3 (weight, bias); final_proj (weight, bias).  Channel order as stored by the
 *   reference (interleaved heads); the library re-orders internally. */
SFM_API int sfm_load_superglue(sfm_handle_t h, const float* packed_host, size_t n_floats);
SFM_API size_t sfm_superglue_packed_floats(void);

/* ---- SuperPoint.forward (superpoint.py:109-171) ----------------------------------------------
 * images [n,H,W] f32 in [0,1]; mask [n,H,W] u8 (0 = masked out) or NULL.  H, W multiples of 8.
 * detect : encoder + detector head + softmax/shuffle + simple_nms + mask/threshold/border
 *          compaction; cand_counts[n] = number of surviving candidates per image.
 * describe: top-k (max_keypoints, -1 = all) + score-descending sort (ties: pixel index
 *          ascending) + descriptor head + bilinear sampling + L2 normalise.
 *          kpts_xy [n,kp_cap,2], scores [n,kp_cap], desc [n,256,kp_cap] (reference layout),
 *          counts[n] = min(candidates, max_keypoints, kp_cap).
 * Both use the same workspace, which must stay untouched between the two calls. */
SFM_API size_t sfm_superpoint_workspace_bytes(int n_img, int H, int W, int precision);
SFM_API int sfm_superpoint_detect(sfm_handle_t h, const float* images, int n_img, int H, int W, const uint8_t* mask,
                          float keypoint_threshold, int nms_radius, int remove_borders, int precision,
                          int* cand_counts, void* workspace, size_t workspace_bytes, void* stream);
SFM_API int sfm_superpoint_describe(sfm_handle_t h, int n_img, int H, int W, int max_keypoints, int align_corners,
                            int precision, int kp_cap, float* kpts_xy, float* scores, float* desc, int* counts,
                            void* workspace, size_t workspace_bytes, void* stream);
/* debug/parity taps into the detect workspace: which = 0 feat NHWC [n,H/8,W/8,128], 1 logits
 * [n,H/8,W/8,65], 2 score map [n,H,W], 3 nms map [n,H,W], 4 raw descriptor map NHWC [n,H/8,W/8,256] */
SFM_API int sfm_superpoint_tap(int n_img, int H, int W, int precision, int which, void* workspace, size_t workspace_bytes,
                       float** out_ptr, size_t* out_floats);

/* image ingestion (matcher.py:360-367: `np.array(Image.open(..).convert('L')).astype('float32') / 255`): the
 * 8-bit grey pixels are uploaded as they are and converted on the device, bit-identically (IEEE division) */
SFM_API int sfm_u8_to_unit_f32(const uint8_t* pixels, float* out, size_t n, void* stream);

/* module-level functions of superpoint.py */
SFM_API int sfm_simple_nms(const float* scores, float* out, int n_img, int H, int W, int nms_radius, void* stream); /* :9-24 */
SFM_API int sfm_sample_descriptors(const float* dmap_nhwc, int n_img, int h, int w, const float* kpts_xy, const int* counts,
                           int kp_cap, int align_corners, float* desc, void* stream); /* :42-54 (+ :159 normalise) */

/* ---- SuperGlue.forward (superglue.py:235-290) ------------------------------------------------
 * Per-image feature records: desc (n_img,256,ldk) f32, kpts (n_img,kcap,2), scores (n_img,kcap).
 * A batch of B pairs is (idx0[b], idx1[b]) into those records (NULL = identity, i.e. the record
 * arrays ARE the (B,256,K) batch tensors of the reference).  K0/K1: keypoints used per side (the
 * caller applies the reference's chunk-minimum truncation, matcher.py:82-91).
 * chunk_groups (>= 1): the B pairs are `chunk_groups` consecutive Matcher chunks of B/chunk_groups
 * pairs fused into one launch; train-mode BatchNorm statistics stay per chunk and side, so the
 * results equal `chunk_groups` separate calls (matcher.py:73-101 semantics) at better GPU occupancy.
 * Outputs: matches0 (B,K0) i64, matches1 (B,K1) i64, matching_scores0/1 f32. */
SFM_API size_t sfm_superglue_workspace_bytes(int B, int K0, int K1, int precision, int chunk_groups);
SFM_API int sfm_superglue_forward(sfm_handle_t h,
                          const float* desc0, const float* kpts0, const float* scores0, int ldk0, int kcap0,
                          const int* idx0,
                          const float* desc1, const float* kpts1, const float* scores1, int ldk1, int kcap1,
                          const int* idx1,
                          int B, int K0, int K1, int H0, int W0, int H1, int W1,
                          int sinkhorn_iterations, float match_threshold, int bn_mode, int precision,
                          int chunk_groups,
                          int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                          void* workspace, size_t workspace_bytes, void* stream);
/* Layer-0 hoisting for eval-mode BatchNorm (bf16 path).  GNN layer 0 is a SELF layer (superglue.py:131-139 with
 * names[0] == 'self'), so `desc + kenc(kpts, scores)` (superglue.py:254-255) and the layer-0 update depend on one image
 * alone when BatchNorm uses running statistics: they are computed ONCE PER IMAGE here instead of once per pair.
 *   hoist_layer0   : records of n_img images (K keypoints each, same layout as sfm_superglue_forward's inputs)
 *                    -> x_store [n_img, K, 256] f32, the residual stream entering layer 1.
 *                    Workspace: sfm_superglue_workspace_bytes(n_img, K, K, precision, 1) is sufficient.
 *   forward_hoisted: the rest of SuperGlue.forward (layers 1..17, final projection, optimal transport, mutual
 *                    matches) for B pairs gathered from the stores by idx0 / idx1 (int32 [B], NULL = identity).
 *                    Results are bit-identical to sfm_superglue_forward with bn_mode = eval on the same keypoints
 *                    (tests/test_gpu_hoist.py). */
SFM_API int sfm_superglue_hoist_layer0(sfm_handle_t h, const float* desc, const float* kpts, const float* scores,
                                       int ldk, int kcap, int n_img, int K, int H, int W, int precision,
                                       float* x_store, void* workspace, size_t workspace_bytes, void* stream);
SFM_API int sfm_superglue_forward_hoisted(sfm_handle_t h, const float* x_store0, int store_k0, const int* idx0,
                                          const float* x_store1, int store_k1, const int* idx1, int B, int K0, int K1,
                                          int sinkhorn_iterations, float match_threshold, int precision,
                                          int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                                          void* workspace, size_t workspace_bytes, void* stream);

/* parity taps into the forward workspace: which = 0 descriptors after kenc (token-major
 * [B*K0 + B*K1, 256]), 1 after the GNN, 2 score matrix [B,K0,K1], 3 u [B,K0+1], 4 v [B,K1+1] */
SFM_API int sfm_superglue_tap(int B, int K0, int K1, int precision, int chunk_groups, int which, void* workspace,
                      size_t workspace_bytes, float** out_ptr, size_t* out_floats);

/* attention (superglue.py:85-89) in the reference's own layout, fp32: query (B,64,4,N), key/value (B,64,4,M)
 * -> out (B,64,4,N) and, when prob != NULL, the materialised softmax prob (B,4,N,M). */
SFM_API size_t sfm_attention_workspace_bytes(int B, int N, int M);
SFM_API int sfm_attention(const float* query, const float* key, const float* value, int B, int N, int M, float* out,
                          float* prob, void* workspace, size_t workspace_bytes, void* stream);

/* log_optimal_transport (superglue.py:151-171): scores (B,N,M) -> Z (B,N+1,M+1) */
SFM_API size_t sfm_sinkhorn_workspace_bytes(int B, int N, int M);
SFM_API int sfm_log_optimal_transport(const float* scores, int B, int N, int M, float alpha, int iters, float* Z,
                              void* workspace, size_t workspace_bytes, void* stream);
/* log_sinkhorn_iterations (superglue.py:142-148) on ARBITRARY couplings Z (B,m,n) and marginals log_mu (B,m),
 * log_nu (B,n): out = Z + u + v after `iters` iterations from u = v = 0.  workspace: (B*m + B*n) floats. */
SFM_API int sfm_log_sinkhorn_iterations(const float* Z, const float* log_mu, const float* log_nu, int B, int m, int n,
                                        int iters, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Sinkhorn + mutual-NN + threshold without materialising Z (superglue.py:268-283; BASELINE config 5) */
SFM_API int sfm_sinkhorn_match(const float* scores, int B, int N, int M, float alpha, int iters, float match_threshold,
                       int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Same contract in the throughput formulation (matrix-scaling Sinkhorn, see csrc/sg_kernels.cu): one exp per
 * element once, then two FMAs per element per iteration and a single HBM read of the matrix per iteration.
 * `scores` is OVERWRITTEN (it becomes exp(S - rowmax)).  Requires M % 4 == 0, M <= 4096, iters > 0. */
SFM_API int sfm_sinkhorn_match_scaled(float* scores, int B, int N, int M, float alpha, int iters, float match_threshold,
                                      int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                                      void* workspace, size_t workspace_bytes, void* stream);
/* Same, with the iterations running on a packed fp16 copy of the kernel matrix (BASELINE config 5 "16-bit
 * storage"): half the HBM bytes per sweep, `scores` is NOT modified, the final arg-max passes use the fp32 scores. */
SFM_API size_t sfm_sinkhorn_half_workspace_bytes(int B, int N, int M);
SFM_API int sfm_sinkhorn_match_scaled_half(const float* scores, int B, int N, int M, float alpha, int iters,
                                           float match_threshold, int64_t* matches0, int64_t* matches1,
                                           float* mscores0, float* mscores1, void* workspace, size_t workspace_bytes,
                                           void* stream);

/* ---- kernel-level hooks of the bf16 tcgen05 path (tests/, bench) ---------------------------------
 * sfm_tc_gemm: out[M,N] = alpha * [A | A2][M,K] * W[N,K]^T + bias (bf16 operands, K in {256,512},
 *   K1 = columns taken from A); optional ReLU; writes bf16 and/or fp32 (accumulate: out_f32 += result,
 *   out_bf16 = bf16(out_f32)) -- the linear layers of superglue.py:48-59,103-108.
 * sfm_tc_attention: fused softmax(Q K^T / 8) V for both sides of one layer from the fused projection
 *   buffer qkv [rows_total, 768] bf16 (q | k | v, head-contiguous); side s uses queries at rows
 *   q_row0_s + b*Nq_s and keys/values at rows k_row0_s + b*Nk_s; out [rows_total, 256] bf16
 *   (superglue.py:85-108).  Nq1 = 0 disables the second side. */
SFM_API int sfm_tc_gemm(sfm_handle_t h, const void* A, const void* A2, const void* W, int M, int N, int K, int K1,
                        const float* bias, float alpha, int relu, void* out_bf16, float* out_f32, int accumulate,
                        void* stream);
SFM_API int sfm_tc_attention(sfm_handle_t h, const void* qkv, long long rows_total, int B, int Nq0, int Nk0,
                             long long q_row0_0, long long k_row0_0, int Nq1, int Nk1, long long q_row0_1,
                             long long k_row0_1, void* out_bf16, void* stream);

/* sfm_tc_conv3x3: 3x3 / pad 1 convolution + bias (+ ReLU) as a tcgen05 implicit GEMM (SuperPoint encoder and heads,
 *   superpoint.py:84-99): in [n,H,W,Cin] bf16 NHWC, w [Cout, 9*Cin] bf16 ordered (ky,kx,ci), out [n,H,W,Cout] bf16;
 *   Cin in {64,128}, Cout in {64,128,256}.  relu: bit 0 = ReLU, bit 1 = fuse the nn.MaxPool2d(2, 2) that follows
 *   (superpoint.py:114-119) into the epilogue: out is then [n,H/2,W/2,Cout] (H, W even, Cout <= 128). */
SFM_API int sfm_tc_conv3x3(sfm_handle_t h, const void* in_nhwc, const void* w, const float* bias, void* out_nhwc,
                           int n, int H, int W, int Cin, int Cout, int relu, void* stream);

/* Matcher.__super_glue_matcher tail (matcher.py:103-109): per pair (L,2) uint32 rows
 * [i, matches0[i]] for matches0[i] > -1, in ascending i.  tables (B,K0,2) u32, lengths (B) i32. */
SFM_API int sfm_compact_matches(const int64_t* matches0, int B, int K0, uint32_t* tables, int* lengths, void* stream);
/* The same tables flattened for one device->host copy (and for the multi-GPU gather to the database writer's
 * rank): rows [0, lengths[p]) of every pair's (kcap,2) table -> out (sum lengths, 2) u32 at row offsets[p]
 * (exclusive prefix sum of lengths, int64).  The concatenation is what matcher.py:103-109 hands to
 * colmap_bd_utils.py:354-367 pair by pair. */
SFM_API int sfm_pack_tables(const uint32_t* tables, int n_pairs, int kcap, const int* lengths, const int64_t* offsets,
                            uint32_t* out, void* stream);

/* Pair schedule of Matcher.match_frames without a filter (matcher.py:143-144, itertools.combinations(ids, 2)):
 * entries [first, first + count) of the lexicographic list of (i < j), 0 <= i < j < n_images, as two device int32
 * arrays of image positions -- the index arrays sfm_superglue_forward takes; a rank of the multi-GPU partition
 * generates exactly its own range. */
SFM_API int sfm_pair_list(int n_images, long long first, long long count, int* idx0, int* idx1, void* stream);

/* Post-match consumer generate_sparse_depth_from_colmap (dataset/simple_sfm_dataset_utils.py:195-231, with
 * coords_world_to_cam / coords_cam_to_film / coords_film_to_pixel, utils/coord_conversion.py:61-97,138-159), all
 * cameras of one image size in one call, float64 like the reference.
 *   xyz (n_points,3); entries e = (entry_cam[e], entry_point[e]): point visible in camera, a camera's points in model
 *   order; extrinsics (n_cams,12) rows of world-to-camera [R|t]; intrinsics (n_cams,4) = f_x, f_y, c_x, c_y relative
 *   to the image size; depth out (n_cams,W,H), -1 where empty, on a pixel hit by several points the LAST entry's
 *   depth (the reference's indexed write); bounds out (n_cams,2) = min / max depth over every in-frustum point
 *   (+inf / 0 when none); winner_ws n_cams*W*H int64 scratch; error_flag = 1 when a point lands on x == W or
 *   y == H (the reference raises IndexError there). */
SFM_API int sfm_sparse_depth(const double* xyz, int n_points, const int* entry_cam, const int* entry_point,
                             long long n_entries, const double* extrinsics, const double* intrinsics, int n_cams,
                             int W, int H, double* depth, double* bounds, long long* winner_ws, int* error_flag,
                             void* stream);

/* SimpleMatcher.match_two_way(_cuda) (simple_matcher.py:66-114): desc (256,N) / (256,M) unit f32.
 * out (3, min(N,M)) f32 rows [i, j, dist], count = L. */
SFM_API size_t sfm_simple_match_workspace_bytes(int N, int M);
SFM_API int sfm_simple_match(const float* desc1, int ld1, int N, const float* desc2, int ld2, int M, float nn_thresh,
                     float* out, int out_ld, int* count, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFM_B200_H_ */
