// Micro-benchmarks behind the N = 192 convolution design (run on a B200: tools/ubench.sh):
//   A. cycles per tcgen05.mma (M128 x N x K16, SS mode, bf16) for N = 64 / 128 / 192 / 256, all SMs busy --
//      does the shared-memory operand read (4 KB of A + 32 N bytes of B per instruction) pace the N = 64 form?
//   B. tcgen05.ld throughput (32x32b.x32, 4 KB per warp instruction) with 4 / 8 / 12 warps
//   C. warp-shuffle throughput with 8 warps
// Stand-alone: nvcc -gencode arch=compute_100a,code=sm_100a -I super-resolution-climate_b200/csrc tools/ubench_umma.cu
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "ptx.cuh"

using namespace sres;

template <int N>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int n_mma, int a_stride16, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  // zero operands (bf16 zeros): results are irrelevant, the operand traffic is not
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t dhi = sdesc_hi_sw128(1024);
    const uint32_t w_lo = sdesc_lo(smem_u32(smem), 16);                // "weights": first 72 KB
    const uint32_t a_lo = sdesc_lo(smem_u32(smem) + 73728, 16);        // "halo window" behind them
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {   // rep 0 warms up
      __syncwarp();
      t0 = clock64();
      if (leader) {
        constexpr int kTaps = 576 / N;   // 9 / 4 / 3 / 2 operand blocks per "tile"
        for (int rep_t = 0; rep_t < n_mma / (kTaps * 4); ++rep_t) {
#pragma unroll
          for (int t = 0; t < kTaps; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t a = a_lo + uint32_t((N == 64 ? (t / 3) * a_stride16 + (t % 3) * 8 : t * a_stride16) + k * 2) + uint32_t((rep_t & 1) * 29696 / 16);
              const uint32_t b = w_lo + uint32_t(t * N * 8 + k * 2);
              umma_bf16_lohi_p(tmem, a, dhi, b, dhi, idesc, (rep_t | t | k) != 0);
            }
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

__global__ void __launch_bounds__(512, 1) ldtm_rate_kernel(int n_warps, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < n_warps) {
    const uint32_t base = tmem + (uint32_t((warp & 3) * 32) << 16);
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld32(base + uint32_t((i * 32) & 255), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  if (acc == 0x12345u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// three loads in flight before the wait (the N = 192 epilogue's pattern)
__global__ void __launch_bounds__(512, 1) ldtm3_rate_kernel(int n_warps, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < n_warps) {
    const uint32_t base = tmem + (uint32_t((warp & 3) * 32) << 16);
    for (int i = 0; i < iters; ++i) {
      uint32_t a[32], b[32], c[32];
      tmem_ld32(base + uint32_t(((i & 1) * 256)), a);
      tmem_ld32(base + uint32_t(((i & 1) * 256) + 64), b);
      tmem_ld32(base + uint32_t(((i & 1) * 256) + 128), c);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= a[j] + b[j] + c[j];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  if (acc == 0x12345u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(512, 1) shfl_rate_kernel(int iters, long long* out, float* sink) {
  float v[8];
  for (int j = 0; j < 8; ++j) v[j] = threadIdx.x * 0.5f + j;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += __shfl_up_sync(0xffffffffu, v[(j + 1) & 7], 1);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0;
  for (int j = 0; j < 8; ++j) s += v[j];
  if (s == 1.2345f) sink[0] = s;
}

static long long median(std::vector<long long> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 1; } } while (0)

template <int N>
static int run_umma(int grid, int n_mma, int a_stride16, long long* d_out) {
  CK(cudaFuncSetAttribute(umma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  umma_rate_kernel<N><<<grid, 128, 220 * 1024>>>(n_mma, a_stride16, d_out);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  const long long med = median(h);
  const double per = double(med) / n_mma;
  printf("UMMA M128 N%-3d K16  grid %3d  %4d MMAs: %7lld clk  = %6.1f clk/MMA  (tensor floor %3d)  operand B/clk %.0f  a_stride %d B\n",
         N, grid, n_mma, med, per, N / 2, (4096.0 + 32.0 * N) / per, a_stride16 * 16);
  return 0;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d_out; uint32_t* d_sink;
  CK(cudaMalloc(&d_out, 4096 * sizeof(long long)));
  CK(cudaMalloc(&d_sink, 64));
  printf("SMs: %d\n", sms);
  for (int grid : {1, sms}) {
    for (int stride : {49 * 8, 8}) {   // tap shift of one image row (49 rows) / of one PTL row
      if (run_umma<64>(grid, 720, stride, d_out)) return 1;
      if (run_umma<128>(grid, 352, stride, d_out)) return 1;
      if (run_umma<192>(grid, 240, stride, d_out)) return 1;
      if (run_umma<256>(grid, 176, stride, d_out)) return 1;
    }
  }
  for (int nw : {1, 4, 8, 12, 16}) {
    ldtm_rate_kernel<<<sms, 512>>>(nw, 256, d_out, d_sink);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(2 * sms);
    CK(cudaMemcpy(h.data(), d_out, 2 * sms * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<long long> t; for (int i = 0; i < sms; ++i) t.push_back(h[2 * i + 1]);
    const long long med = median(t);
    printf("LDTM 32x32b.x32, %2d warps x 256 loads (4 KB each, wait after every load): %7lld clk -> %.1f B/clk/SM, %.1f clk per load per warp\n", nw, med,
           double(nw) * 256 * 4096 / med, double(med) / 256);
    ldtm3_rate_kernel<<<sms, 512>>>(nw, 128, d_out, d_sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, 2 * sms * sizeof(long long), cudaMemcpyDeviceToHost));
    t.clear(); for (int i = 0; i < sms; ++i) t.push_back(h[2 * i + 1]);
    const long long med3 = median(t);
    printf("LDTM 3 loads in flight, %2d warps x 128 x 3 loads: %7lld clk -> %.1f B/clk/SM\n", nw, med3, double(nw) * 384 * 4096 / med3);
  }
  {
    shfl_rate_kernel<<<sms, 256>>>(256, d_out, (float*)d_sink);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), d_out, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    printf("SHFL.UP + FADD, 8 warps x 2048: %lld clk -> %.2f clk per warp-shuffle per SM\n", median(h), double(median(h)) / (8 * 2048));
    shfl_rate_kernel<<<sms, 512>>>(256, d_out, (float*)d_sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    printf("SHFL.UP + FADD, 16 warps x 2048: %lld clk -> %.2f clk per warp-shuffle per SM\n", median(h), double(median(h)) / (16 * 2048));
  }
  return 0;
}
