// Epilogue reductions shared by the convolution kernels (conv_igemm.cu, conv_pair.cu): per-tile column sums of an
// accumulator block, in the row layout (thread = TMEM lane) and in the accumulator-fragment layout.
#pragma once
#include "ptx.cuh"

namespace sres {

// 16-value butterfly: after the call lane l (even) holds in v[0] the sum over the 32 lanes of
// column (l >> 1).
__device__ __forceinline__ float butterfly16(float (&v)[16], int lane) {
  {
    const bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float send = up ? v[j] : v[j + 8];
      float keep = up ? v[j + 8] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float send = up ? v[j] : v[j + 4];
      float keep = up ? v[j + 4] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float send = up ? v[j] : v[j + 2];
      float keep = up ? v[j + 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
    float send = up ? v[0] : v[1];
    float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

// per-image-segment channel sums of 16 columns -> pool_part[tile][seg][quarter][64]
__device__ __forceinline__ void pool_partials(const float (&v)[16], int seg, int lane, float* dst) {
  const bool any1 = __any_sync(0xffffffffu, seg == 1);
  const bool any0 = __any_sync(0xffffffffu, seg == 0);
  float s0 = 0.f, s1 = 0.f;
  if (any0) {
    float t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = seg == 0 ? v[j] : 0.f;
    s0 = butterfly16(t, lane);
  }
  if (any1) {
    float t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = seg == 1 ? v[j] : 0.f;
    s1 = butterfly16(t, lane);
  }
  if ((lane & 1) == 0) {
    dst[lane >> 1] = s0;
    dst[4 * 64 + (lane >> 1)] = s1;
  }
}

// Column sums in the accumulator-fragment layout: every thread holds 8 partial sums (column 8k + 2(lane%4) + e
// at index 2k + e) over its own rows; three exchange stages over lane bits 4,3,2 leave ONE column total per
// lane: column 16*bit4 + 8*bit3 + 2*(lane%4) + bit2.  7 shuffles per 32 columns.
__device__ __forceinline__ float frag_colsum(float (&w)[8], int lane) {
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? w[i] : w[i + 4];
      const float keep = up ? w[i + 4] : w[i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? w[i] : w[i + 2];
      const float keep = up ? w[i + 2] : w[i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
    const float send = up ? w[0] : w[1];
    const float keep = up ? w[1] : w[0];
    w[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  return w[0];
}

}  // namespace sres
