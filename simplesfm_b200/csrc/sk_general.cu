// log_sinkhorn_iterations on arbitrary couplings and marginals (reference: simple_sfm/models/superglue.py:142-148):
//   u = log_mu - LSE_j(Z + v),  v = log_nu - LSE_i(Z + u)   `iters` times from u = v = 0;   out = Z + u + v
// This is the module-level API function in full generality (any Z, any log_mu / log_nu); the matching path uses
// the dustbin-aware kernels of sg_kernels.cu instead.  Z is streamed twice per iteration (HBM-bound).
#include "sg_kernels.cuh"

namespace sfm {
namespace {

// one warp per row: u_i = log_mu_i - LSE_j(Z_ij + v_j)
__global__ void __launch_bounds__(256) skg_row_kernel(const float* __restrict__ Z, int m, int n,
                                                      const float* __restrict__ v, const float* __restrict__ log_mu,
                                                      float* __restrict__ u) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m) return;
  const float* zr = Z + ((long long)b * m + row) * n;
  const float* vb = v + (long long)b * n;
  Lse acc{-INFINITY, 0.f};
  for (int j = lane; j < n; j += 32) acc = lse_merge(acc, Lse{zr[j] + vb[j], 1.f});
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Lse other{__shfl_xor_sync(0xffffffffu, acc.m, o), __shfl_xor_sync(0xffffffffu, acc.s, o)};
    acc = lse_merge(acc, other);
  }
  if (lane == 0) u[(long long)b * m + row] = log_mu[(long long)b * m + row] - (acc.m + logf(acc.s));
}

// 32 columns per block, 8 row groups: v_j = log_nu_j - LSE_i(Z_ij + u_i)
__global__ void __launch_bounds__(256) skg_col_kernel(const float* __restrict__ Z, int m, int n,
                                                      const float* __restrict__ u, const float* __restrict__ log_nu,
                                                      float* __restrict__ v) {
  __shared__ float sm[8][33], ss[8][33];
  const int b = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  const float* zb = Z + (long long)b * m * n;
  const float* ub = u + (long long)b * m;
  Lse acc{-INFINITY, 0.f};
  if (col < n)
    for (int i = ty; i < m; i += 8) acc = lse_merge(acc, Lse{zb[(long long)i * n + col] + ub[i], 1.f});
  sm[ty][tx] = acc.m;
  ss[ty][tx] = acc.s;
  __syncthreads();
  if (ty == 0 && col < n) {
    for (int k = 1; k < 8; k++) acc = lse_merge(acc, Lse{sm[k][tx], ss[k][tx]});
    v[(long long)b * n + col] = log_nu[(long long)b * n + col] - (acc.m + logf(acc.s));
  }
}

__global__ void skg_apply_kernel(const float* __restrict__ Z, int m, int n, const float* __restrict__ u,
                                 const float* __restrict__ v, float* __restrict__ out, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long r = i / n;          // b * m + row
  const int j = (int)(i - r * n);
  const long long b = r / m;
  out[i] = __fadd_rn(__fadd_rn(Z[i], u[r]), v[b * n + j]);   // the reference's order: (Z + u) + v
}

}  // namespace

int log_sinkhorn_general(const float* Z, const float* log_mu, const float* log_nu, int B, int m, int n, int iters,
                         float* out, float* u, float* v, cudaStream_t st) {
  if (B == 0) return SFM_OK;
  SFM_REQUIRE(m > 0 && n > 0 && iters >= 0, "log_sinkhorn_iterations: bad dims");
  SFM_REQUIRE(B <= 65535, "log_sinkhorn_iterations: batch too large");
  SFM_CUDA(cudaMemsetAsync(u, 0, sizeof(float) * (size_t)B * m, st));
  SFM_CUDA(cudaMemsetAsync(v, 0, sizeof(float) * (size_t)B * n, st));
  for (int it = 0; it < iters; it++) {
    skg_row_kernel<<<dim3(cdiv(m, 8), B), 256, 0, st>>>(Z, m, n, v, log_mu, u);
    skg_col_kernel<<<dim3(cdiv(n, 32), B), 256, 0, st>>>(Z, m, n, u, log_nu, v);
  }
  const long long total = (long long)B * m * n;
  skg_apply_kernel<<<cdiv(total, 256), 256, 0, st>>>(Z, m, n, u, v, out, total);
  SFM_LAUNCH_CHECK_N(2 * iters + 1);
  return SFM_OK;
}

}  // namespace sfm
