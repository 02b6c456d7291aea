// Degree-4 real spherical harmonics of the view direction.
// Semantics: SHEncoding.pytorch_fwd -> components_from_spherical_harmonics(4, .) (nerfstudio/field_components/
// encodings.py:797-805, nerfstudio/utils/math.py:31-94), evaluated by NeuRADField on (d+1)/2
// (nerfstudio/fields/base_field.py:136-142).  No gradient flows to the directions (torch.no_grad in the reference).
#include "common.cuh"

namespace nrb {

__device__ __forceinline__ void sh16_eval(float x, float y, float z, float o[16]) {
  const float xx = x * x, yy = y * y, zz = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = 0.4886025119029199f * y;
  o[2] = 0.4886025119029199f * z;
  o[3] = 0.4886025119029199f * x;
  o[4] = 1.0925484305920792f * x * y;
  o[5] = 1.0925484305920792f * y * z;
  o[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
  o[7] = 1.0925484305920792f * x * z;
  o[8] = 0.5462742152960396f * (xx - yy);
  o[9] = 0.5900435899266435f * y * (3.0f * xx - yy);
  o[10] = 2.890611442640554f * x * y * z;
  o[11] = 0.4570457994644658f * y * (5.0f * zz - 1.0f);
  o[12] = 0.3731763325901154f * z * (5.0f * zz - 3.0f);
  o[13] = 0.4570457994644658f * x * (5.0f * zz - 1.0f);
  o[14] = 1.445305721320277f * z * (xx - yy);
  o[15] = 0.5900435899266435f * x * (xx - 3.0f * yy);
}

__global__ void __launch_bounds__(256) sh16_kernel(const float* __restrict__ dirs, float* __restrict__ out, int64_t M,
                                                   int normalize) {
  const int64_t m = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float x = dirs[3 * m], y = dirs[3 * m + 1], z = dirs[3 * m + 2];
  if (normalize) {
    x = (x + 1.0f) * 0.5f;
    y = (y + 1.0f) * 0.5f;
    z = (z + 1.0f) * 0.5f;
  }
  float o[16];
  sh16_eval(x, y, z, o);
  float4* dst = reinterpret_cast<float4*>(out + 16 * m);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

}  // namespace nrb

extern "C" int nrb_sh16(const float* dirs, float* out, int64_t M, int32_t normalize_to_unit_cube,
                        nrb_stream_t stream) {
  NRB_REQUIRE(dirs && out && M >= 0, NRB_ERR_BAD_ARG, "nrb_sh16: null pointer or negative M");
  NRB_REQUIRE(nrb::aligned16(out), NRB_ERR_ALIGNMENT, "nrb_sh16: out must be 16-byte aligned");
  if (M == 0) return NRB_OK;
  nrb::sh16_kernel<<<nrb::blocks_for(M, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dirs, out, M,
                                                                                           normalize_to_unit_cube);
  return nrb::finish_launch("nrb_sh16");
}
