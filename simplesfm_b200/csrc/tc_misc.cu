// Element-wise / reduction helpers of the bf16 path: casts, train-mode BatchNorm statistics
// (superglue.py:56-57 as the reference runs it) and the fused normalise + ReLU.
#include "tc_common.cuh"
#include "tc_path.cuh"

namespace sfm {
namespace tc {

namespace {

__global__ void cast_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = in[i];
  out[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}

}  // namespace

int f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t st) {
  SFM_REQUIRE(n % 4 == 0, "f32_to_bf16: n must be a multiple of 4");
  if (n == 0) return SFM_OK;
  cast_kernel<<<cdiv(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out),
                                                n / 4);
  SFM_LAUNCH_CHECK();
  return SFM_OK;
}

}  // namespace tc
}  // namespace sfm
