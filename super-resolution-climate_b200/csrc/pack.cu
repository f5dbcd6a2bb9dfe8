// Weight repacking: fp32 OIHW checkpoint layout -> bf16 tap-major UMMA operands.
// (state_dict layout of nn.Conv2d, sres/model/common/cnn.py:8-9.)
#include <cuda_bf16.h>
#include "internal.h"
#include "ptx.cuh"

namespace sres {

// out[t][n][k], 64 k per row.
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int mode,
                                         int n_rows, int cin, int cout_total, int oc_stride, int oc_offset) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 9 * n_rows * 64;
  if (idx >= total) return;
  const int k = idx & 63;
  const int n = (idx >> 6) % n_rows;
  const int t = idx / (64 * n_rows);
  const int ky = t / 3, kx = t % 3;
  float v = 0.f;
  if (mode == 0) {
    // forward: rows = output channels oc(n), k = input channel
    const int oc = n * oc_stride + oc_offset;
    if (oc < cout_total && k < cin) v = w[((oc * cin + k) * 3 + ky) * 3 + kx];
  } else {
    // dgrad: rows = input channels n, k = output channel oc(k), taps flipped
    const int oc = k * oc_stride + oc_offset;
    if (oc < cout_total && n < cin) v = w[((oc * cin + n) * 3 + (2 - ky)) * 3 + (2 - kx)];
  }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  out[idx] = *reinterpret_cast<uint16_t*>(&h);
}

}  // namespace sres

extern "C" int sres_pack_conv_weights(const float* w_oihw, void* out_bf16, int mode, int n_rows, int cin,
                                      int cout_total, int oc_stride, int oc_offset, void* stream) {
  using namespace sres;
  if (!w_oihw || !out_bf16) return set_error(SRES_ERR_INVALID_ARG, "pack: null pointer");
  if (n_rows <= 0 || n_rows > 256 || cin <= 0 || cin > 64 || (mode != 0 && mode != 1))
    return set_error(SRES_ERR_INVALID_ARG, "pack: bad shape");
  const int total = 9 * n_rows * 64;
  pack_conv_weights_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w_oihw, (uint16_t*)out_bf16, mode, n_rows, cin, cout_total, oc_stride, oc_offset);
  SRES_CHECK_LAUNCH("pack: launch");
  return SRES_OK;
}
