// tcgen05.shift throughput: independent column groups, and concurrency with tcgen05.mma on the other accumulator stage.
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "ptx.cuh"
using namespace sres;
__device__ __forceinline__ void tmem_shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" :: "r"(taddr) : "memory");
}
// mode 0: MMAs only (12 x N192 per iteration); 1: shifts only (24 per iteration: 8 groups once + 8 groups twice);
// 2: both, shifts after the MMAs; 3: both, interleaved 2 shifts after every MMA; 4: shifts only, all on one column group
__global__ void __launch_bounds__(128, 1) k(int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 140 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, 192, 0, 0);
    constexpr uint32_t dhi = sdesc_hi_sw128(1024);
    const uint32_t w_lo = sdesc_lo(smem_u32(smem), 16);
    const uint32_t a_lo = sdesc_lo(smem_u32(smem) + 73728, 16);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (leader) {
        for (int it = 0; it < iters; ++it) {
          const uint32_t dst = tmem + uint32_t((it & 1) * 256);         // MMA stage
          const uint32_t sh = tmem + uint32_t(((it & 1) ^ 1) * 256);    // the other stage gets shifted
          int s = 0;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (mode == 0 || mode == 2 || mode == 3)
                umma_bf16_lohi_p(dst, a_lo + uint32_t(ky * 392 + kk * 2), dhi, w_lo + uint32_t(ky * 192 * 8 + kk * 2), dhi, idesc, (ky | kk) != 0);
              if (mode == 3) {   // 2 shifts per MMA: D_1 groups once (s < 8), D_2 groups twice
                for (int e = 0; e < 2; ++e, ++s) tmem_shift_down(sh + uint32_t(s < 8 ? 64 + 8 * s : 128 + 8 * ((s - 8) & 7)));
              }
            }
          }
          if (mode == 1 || mode == 2) {
#pragma unroll
            for (int g = 0; g < 8; ++g) tmem_shift_down(sh + uint32_t(64 + 8 * g));
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int g = 0; g < 8; ++g) tmem_shift_down(sh + uint32_t(128 + 8 * g));
          }
          if (mode == 4) {
            for (int g = 0; g < 24; ++g) tmem_shift_down(sh + 64u);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1, 1);
      t1 = clock64();
    }
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 1; } } while (0)
int main() {
  long long* d; CK(cudaMalloc(&d, 4096 * 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  const char* names[] = {"12 MMAs N192 only", "24 shifts only (independent groups)", "12 MMAs then 24 shifts (other stage)", "12 MMAs with 2 shifts after each", "24 shifts on ONE column group (dependent)"};
  for (int mode = 0; mode < 5; ++mode) {
    const int iters = 20;
    k<<<148, 128, 160 * 1024>>>(mode, iters, d);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    printf("%-45s %8.1f clk per iteration\n", names[mode], double(h[74]) / iters);
  }
  return 0;
}
