// The narrow 3x3 convolutions at the two ends of RCAN, on CUDA cores (they are HBM-bound: K or N
// is only 9*C with C = 1..4 image channels, SURVEY.md section 2 op inventory).
//
//   conv3x3_small_in   : planar fp32 NCHW (C_s channels) -> 64-feature PTL rows.
//                        head conv forward (sres/model/rcan/network.py:13,23) and, with transposed +
//                        flipped weights, the input-gradient of the tail conv (network.py:16).
//   small_in_wgrad     : weight/bias gradient of the head conv (reduction over all positions).
//   small_out_wgrad    : weight/bias gradient of the tail conv 64 -> C_s.
#include "internal.h"
#include "ptx.cuh"

namespace sres {

constexpr int kMaxSmallC = 4;

// ---------------------------------------------------------------------------------------------
// planar (B,Cs,H,W) fp32  ->  PTL [rows][64]
//   transposed == 0:  w is (64, Cs, 3, 3):  out[q][n] = b[n] + sum_{c,t} w[n][c][t] * in[c][q+off(t)]
//   transposed == 1:  w is (Cs, 64, 3, 3):  out[q][n] =        sum_{c,t} w[c][n][8-t] * in[c][q+off(t)]
// One thread = one position x 8 output features.  Padding rows are written as zeros.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv3x3_small_in_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                        int B, int Cs, int H, int W, int transposed, int unshuffle, float* __restrict__ out_f32,
                        uint16_t* __restrict__ out_bf16) {
  __shared__ float sw[64 * kMaxSmallC * 9];  // [n][c][t]
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 64 * Cs * 9; i += blockDim.x) {
    const int t = i % 9, c = (i / 9) % Cs, n = i / (9 * Cs);
    sw[i] = transposed ? w[(c * 64 + n) * 9 + (8 - t)] : w[(n * Cs + c) * 9 + t];
  }
  if (threadIdx.x < 64) sb[threadIdx.x] = (bias && !transposed) ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int P = W + 1, RP = (H + 1) * P;
  const long long total = (long long)B * RP * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cg = int(idx & 7);
    const long long q = idx >> 3;
    const int b = int(q / RP), rem = int(q - (long long)b * RP);
    const int y = rem / P, x = rem - y * P;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const bool pad = (x == W) || (y == H);
    if (!pad) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = sb[cg * 8 + j];
      for (int c = 0; c < Cs; ++c) {
        const float* ip = in + ((size_t)b * Cs + c) * H * W;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          const float v = __ldg(ip + (size_t)yy * W + xx);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(sw[((cg * 8 + j) * Cs + c) * 9 + t], v, acc[j]);
        }
      }
    }
    long long oq = q;
    if (unshuffle > 1) {  // PixelUnshuffle(f) store: sub-grid (y%f, x%f), position (y/f, x/f)
      const int f = unshuffle;
      const int Pl = W / f + 1, Rl = H / f + 1;
      const int sub = (y % f) * f + (x % f);
      oq = (long long)sub * B * Rl * Pl + (long long)b * Rl * Pl + (long long)(y / f) * Pl + (x / f);
    }
    if (out_f32) {
      *reinterpret_cast<float4*>(out_f32 + oq * 64 + cg * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(out_f32 + oq * 64 + cg * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (out_bf16)
      *reinterpret_cast<uint4*>(out_bf16 + oq * 64 + cg * 8) =
          make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                     pack_bf16x2(acc[6], acc[7]));
  }
}

// ---------------------------------------------------------------------------------------------
// head-conv weight gradient: dW[n][c][t] = sum_q (g1+g2)[q][n] * in[c][q+off(t)],  db[n] = sum_q g[q][n]
// block = 64 features x 4 position lanes; per-block partials then a fixed-order reduce.
// ---------------------------------------------------------------------------------------------
constexpr int kSwAcc = kMaxSmallC * 9 + 1;

__global__ void __launch_bounds__(256)
small_in_wgrad_kernel(const float* __restrict__ g1, const float* __restrict__ g2, const float* __restrict__ in, int B,
                      int Cs, int H, int W, float* __restrict__ part) {
  __shared__ float sm[4][64][kSwAcc];
  const int n = threadIdx.x & 63, ln = threadIdx.x >> 6;
  const int P = W + 1, RP = (H + 1) * P;
  const long long npos = (long long)B * RP;
  float acc[kSwAcc];
#pragma unroll
  for (int i = 0; i < kSwAcc; ++i) acc[i] = 0.f;
  for (long long q = (long long)blockIdx.x * 4 + ln; q < npos; q += (long long)gridDim.x * 4) {
    const int b = int(q / RP), rem = int(q - (long long)b * RP);
    const int y = rem / P, x = rem - y * P;
    if (x == W || y == H) continue;
    float gv = g1[q * 64 + n];
    if (g2) gv += g2[q * 64 + n];
    acc[kSwAcc - 1] += gv;
    for (int c = 0; c < Cs; ++c) {
      const float* ip = in + ((size_t)b * Cs + c) * H * W;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        acc[c * 9 + t] = fmaf(gv, __ldg(ip + (size_t)yy * W + xx), acc[c * 9 + t]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kSwAcc; ++i) sm[ln][n][i] = acc[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * kSwAcc; i += 256) {
    const int nn = i / kSwAcc, e = i % kSwAcc;
    part[(size_t)blockIdx.x * 64 * kSwAcc + i] = sm[0][nn][e] + sm[1][nn][e] + sm[2][nn][e] + sm[3][nn][e];
  }
}

__global__ void small_in_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int Cs,
                                             float* __restrict__ dw, float* __restrict__ db, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * kSwAcc) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * 64 * kSwAcc + i];
  const int n = i / kSwAcc, e = i % kSwAcc;
  if (e == kSwAcc - 1) {
    if (db) db[n] = accumulate ? db[n] + s : s;
  } else if (e < Cs * 9) {
    float* o = dw + (size_t)n * Cs * 9 + e;
    *o = accumulate ? *o + s : s;
  }
}

// ---------------------------------------------------------------------------------------------
// tail-conv weight gradient: dW[c][k][t] = sum_q dout[c][q] * U[q+off(t)][k],  db[c] = sum_q dout[c][q]
// Looping over the rows p of U and scattering to the 9 taps reads U once:
//   dW[c][k][t] += U[p][k] * dout[c][p - off(t)].
// block = 64 input features x 4 row lanes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
small_out_wgrad_kernel(const float* __restrict__ dout, const uint16_t* __restrict__ u, int B, int Cs, int H, int W,
                       float* __restrict__ part) {
  __shared__ float sm[4][64][kSwAcc];
  const int k = threadIdx.x & 63, ln = threadIdx.x >> 6;
  const int P = W + 1, RP = (H + 1) * P;
  const long long npos = (long long)B * RP;
  float acc[kSwAcc];
#pragma unroll
  for (int i = 0; i < kSwAcc; ++i) acc[i] = 0.f;
  for (long long q = (long long)blockIdx.x * 4 + ln; q < npos; q += (long long)gridDim.x * 4) {
    const int b = int(q / RP), rem = int(q - (long long)b * RP);
    const int y = rem / P, x = rem - y * P;
    if (x == W || y == H) continue;
    const uint16_t raw = u[q * 64 + k];
    const float uv = __uint_as_float(uint32_t(raw) << 16);
    for (int c = 0; c < Cs; ++c) {
      const float* dp = dout + ((size_t)b * Cs + c) * H * W;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        // output pixel that sees row p through tap t:  (y,x) - (t/3-1, t%3-1)
        const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        acc[c * 9 + t] = fmaf(uv, __ldg(dp + (size_t)yy * W + xx), acc[c * 9 + t]);
      }
    }
    // bias gradient: feature lane k < Cs sums its own output channel
    if (k < Cs) acc[kSwAcc - 1] += __ldg(dout + ((size_t)b * Cs + k) * H * W + (size_t)y * W + x);
  }
#pragma unroll
  for (int i = 0; i < kSwAcc; ++i) sm[ln][k][i] = acc[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * kSwAcc; i += 256) {
    const int kk = i / kSwAcc, e = i % kSwAcc;
    part[(size_t)blockIdx.x * 64 * kSwAcc + i] = sm[0][kk][e] + sm[1][kk][e] + sm[2][kk][e] + sm[3][kk][e];
  }
}

__global__ void small_out_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int Cs,
                                              float* __restrict__ dw, float* __restrict__ db, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * kSwAcc) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * 64 * kSwAcc + i];
  const int k = i / kSwAcc, e = i % kSwAcc;
  if (e == kSwAcc - 1) {
    if (db && k < Cs) db[k] = accumulate ? db[k] + s : s;
  } else if (e < Cs * 9) {
    const int c = e / 9, t = e % 9;
    float* o = dw + ((size_t)c * 64 + k) * 9 + t;
    *o = accumulate ? *o + s : s;
  }
}

static int small_grid() {
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  return sms * 4;
}

}  // namespace sres

using namespace sres;

extern "C" int sres_conv3x3_small_in(const float* in_nchw, const float* w, const float* bias, int B, int Cs, int H,
                                     int W, int transposed, int unshuffle, float* out_f32, void* out_bf16,
                                     void* stream) {
  if (!in_nchw || !w || (!out_f32 && !out_bf16)) return set_error(SRES_ERR_INVALID_ARG, "small_in: null pointer");
  if (Cs < 1 || Cs > kMaxSmallC) return set_error(SRES_ERR_UNSUPPORTED, "small_in: 1..4 image channels supported");
  if (B <= 0 || H <= 0 || W <= 0) return set_error(SRES_ERR_INVALID_ARG, "small_in: bad geometry");
  if (unshuffle > 1 && (H % unshuffle || W % unshuffle))
    return set_error(SRES_ERR_INVALID_ARG, "small_in: unshuffle factor must divide H and W");
  const long long total = (long long)B * (H + 1) * (W + 1) * 8;
  long long blocks = (total + 255) / 256;
  const int cap = small_grid() * 4;
  if (blocks > cap) blocks = cap;
  conv3x3_small_in_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in_nchw, w, bias, B, Cs, H, W, transposed,
                                                                         unshuffle, out_f32, (uint16_t*)out_bf16);
  SRES_CHECK_LAUNCH("small_in: launch");
  return SRES_OK;
}

extern "C" size_t sres_small_wgrad_workspace_bytes(void) { return (size_t)small_grid() * 64 * kSwAcc * sizeof(float); }

extern "C" int sres_small_in_wgrad(const float* g1_f32, const float* g2_f32, const float* in_nchw, int B, int Cs,
                                   int H, int W, float* dw, float* db, int accumulate, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!g1_f32 || !in_nchw || !dw || !workspace) return set_error(SRES_ERR_INVALID_ARG, "small_in_wgrad: null pointer");
  if (Cs < 1 || Cs > kMaxSmallC) return set_error(SRES_ERR_UNSUPPORTED, "small_in_wgrad: 1..4 image channels supported");
  const int grid = small_grid();
  if (workspace_bytes < (size_t)grid * 64 * kSwAcc * sizeof(float))
    return set_error(SRES_ERR_INVALID_ARG, "small_in_wgrad: workspace too small");
  small_in_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g1_f32, g2_f32, in_nchw, B, Cs, H, W, (float*)workspace);
  SRES_CHECK_LAUNCH("small_in_wgrad: launch");
  small_in_wgrad_reduce_kernel<<<(64 * kSwAcc + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, grid,
                                                                                            Cs, dw, db, accumulate);
  SRES_CHECK_LAUNCH("small_in_wgrad: reduce launch");
  return SRES_OK;
}

extern "C" int sres_small_out_wgrad(const float* dout_nchw, const void* u_bf16, int B, int Cs, int H, int W, float* dw,
                                    float* db, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dout_nchw || !u_bf16 || !dw || !workspace) return set_error(SRES_ERR_INVALID_ARG, "small_out_wgrad: null pointer");
  if (Cs < 1 || Cs > kMaxSmallC) return set_error(SRES_ERR_UNSUPPORTED, "small_out_wgrad: 1..4 image channels supported");
  const int grid = small_grid();
  if (workspace_bytes < (size_t)grid * 64 * kSwAcc * sizeof(float))
    return set_error(SRES_ERR_INVALID_ARG, "small_out_wgrad: workspace too small");
  small_out_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dout_nchw, (const uint16_t*)u_bf16, B, Cs, H, W,
                                                                 (float*)workspace);
  SRES_CHECK_LAUNCH("small_out_wgrad: launch");
  small_out_wgrad_reduce_kernel<<<(64 * kSwAcc + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, grid,
                                                                                             Cs, dw, db, accumulate);
  SRES_CHECK_LAUNCH("small_out_wgrad: reduce launch");
  return SRES_OK;
}
