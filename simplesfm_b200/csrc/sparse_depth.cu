// Sparse depth maps from a COLMAP sparse model (reference: generate_sparse_depth_from_colmap,
// simple_sfm/dataset/simple_sfm_dataset_utils.py:147-261 -- a Python loop over every 3-D point for every camera,
// O(cameras x points) dict look-ups, then per camera world_to_cam / cam_to_pixel / an indexed write).
// Here: one thread per (camera, visible point) entry, all cameras of one image size in one launch, float64 like
// the reference (torch.tensor of numpy float64).  The reference's indexed write keeps the LAST point that lands
// on a pixel (points in model order); the same rule is made deterministic with an atomicMax on the entry index.
#include "common.cuh"

namespace sfm {
namespace {

struct Proj {
  long long pix;   // linear index into [C, W, H], -1 = not written, -2 = lands on the x == W / y == H edge
  double depth;
};

__device__ __forceinline__ Proj project(const double* __restrict__ xyz, const double* __restrict__ extr,
                                        const double* __restrict__ intr, int cam, int pt, int W, int H) {
  const double* p = xyz + 3ll * pt;
  const double* E = extr + 12ll * cam;      // rows of [R | t]
  // coords_world_to_cam (utils/coord_conversion.py:138-159): R p + t
  const double x = (E[0] * p[0] + E[1] * p[1] + E[2] * p[2]) + E[3];
  const double y = (E[4] * p[0] + E[5] * p[1] + E[6] * p[2]) + E[7];
  const double z = (E[8] * p[0] + E[9] * p[1] + E[10] * p[2]) + E[11];
  // coords_cam_to_film (:61-78) and coords_film_to_pixel (:81-97), then the scaling to pixels (:211)
  const double den = z >= 0.0 ? fmax(z, 1e-6) : fmin(z, -1e-6);
  const double* K = intr + 4ll * cam;       // f_x, f_y, c_x, c_y as the reference builds them
  const double px = ((x / den) * K[0] + K[2]) * (double)W;
  const double py = ((y / den) * K[1] + K[3]) * (double)H;
  Proj r;
  r.depth = z;
  r.pix = -1;
  if (px >= 0.0 && px <= (double)W && py >= 0.0 && py <= (double)H && z >= 0.0) {   // :212-214
    const long long xi = (long long)px, yi = (long long)py;                          // .long() truncates
    r.pix = (xi >= W || yi >= H) ? -2 : ((long long)cam * W + xi) * H + yi;
  }
  return r;
}

__global__ void __launch_bounds__(256) depth_winner_kernel(const double* __restrict__ xyz, const int* __restrict__ ecam,
                                                           const int* __restrict__ ept, long long E,
                                                           const double* __restrict__ extr,
                                                           const double* __restrict__ intr, int W, int H,
                                                           long long* __restrict__ winner,
                                                           unsigned long long* __restrict__ bounds,
                                                           int* __restrict__ err) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = e < E;
  int cam = -1;
  Proj r;
  r.pix = -1;
  r.depth = 0.0;
  if (live) {
    cam = ecam[e];
    r = project(xyz, extr, intr, cam, ept[e], W, H);
    if (r.pix == -2) atomicExch(err, 1);
    if (r.pix >= 0) atomicMax(reinterpret_cast<unsigned long long*>(winner + r.pix), (unsigned long long)(e + 1));
  }
  // depth bounds per camera over every point inside the frustum, overwritten ones included (:228; the
  // visualisation's normalisation).  Depths are >= 0 there, so their bit patterns order like unsigned integers.
  const bool in = r.pix != -1;
  unsigned long long lo = in ? (unsigned long long)__double_as_longlong(r.depth + 0.0) : ~0ull;
  unsigned long long hi = in ? (unsigned long long)__double_as_longlong(r.depth + 0.0) : 0ull;
  const int cam0 = __shfl_sync(0xffffffffu, cam, 0);
  if (__all_sync(0xffffffffu, cam == cam0 || !in)) {           // entries are sorted by camera: the usual case
    const int cam_any = __reduce_max_sync(0xffffffffu, in ? cam : -1);
    for (int o = 16; o; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && cam_any >= 0) {
      atomicMin(bounds + 2 * cam_any, lo);
      atomicMax(bounds + 2 * cam_any + 1, hi);
    }
  } else if (in) {
    atomicMin(bounds + 2 * cam, lo);
    atomicMax(bounds + 2 * cam + 1, hi);
  }
}

__global__ void __launch_bounds__(256) depth_write_kernel(const double* __restrict__ xyz, const int* __restrict__ ecam,
                                                          const int* __restrict__ ept, long long E,
                                                          const double* __restrict__ extr,
                                                          const double* __restrict__ intr, int W, int H,
                                                          const long long* __restrict__ winner,
                                                          double* __restrict__ depth) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const Proj r = project(xyz, extr, intr, ecam[e], ept[e], W, H);
  if (r.pix >= 0 && winner[r.pix] == e + 1) depth[r.pix] = r.depth;
}

__global__ void fill_f64_kernel(double* p, long long n, double v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void bounds_init_kernel(unsigned long long* b, int n_cams) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_cams) {
    b[2 * i] = 0x7ff0000000000000ull;      // +inf: "no point" (the host substitutes the reference's [-1, 1])
    b[2 * i + 1] = 0ull;
  }
}

}  // namespace

int sparse_depth_project(const double* xyz, const int* entry_cam, const int* entry_point, long long n_entries,
                         const double* extrinsics, const double* intrinsics, int n_cams, int W, int H, double* depth,
                         double* bounds, long long* winner_ws, int* error_flag, cudaStream_t st) {
  const long long n = (long long)n_cams * W * H;
  if (n == 0) return SFM_OK;
  fill_f64_kernel<<<cdiv(n, 256), 256, 0, st>>>(depth, n, -1.0);     // torch.zeros(...) - 1   (:226)
  SFM_CUDA(cudaMemsetAsync(winner_ws, 0, sizeof(long long) * n, st));
  SFM_CUDA(cudaMemsetAsync(error_flag, 0, sizeof(int), st));
  bounds_init_kernel<<<cdiv(n_cams, 256), 256, 0, st>>>(reinterpret_cast<unsigned long long*>(bounds), n_cams);
  SFM_LAUNCH_CHECK_N(2);
  if (n_entries == 0) return SFM_OK;
  depth_winner_kernel<<<cdiv(n_entries, 256), 256, 0, st>>>(xyz, entry_cam, entry_point, n_entries, extrinsics,
                                                            intrinsics, W, H, winner_ws,
                                                            reinterpret_cast<unsigned long long*>(bounds), error_flag);
  depth_write_kernel<<<cdiv(n_entries, 256), 256, 0, st>>>(xyz, entry_cam, entry_point, n_entries, extrinsics,
                                                           intrinsics, W, H, winner_ws, depth);
  SFM_LAUNCH_CHECK_N(2);
  return SFM_OK;
}

}  // namespace sfm
