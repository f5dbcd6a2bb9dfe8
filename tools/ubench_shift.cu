// What does tcgen05.shift do to a matrix of fp32 accumulators in TMEM?  (PTX: "shifts 32-byte elements down by one row")
// Fills TMEM lanes 0..127 x columns 0..63 with lane*256+col, issues shift(s), reads everything back and prints which
// (lane, column) cells changed; then times a train of shifts.  Stand-alone, run on a B200.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "ptx.cuh"
using namespace sres;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
         "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" :: "r"(taddr) : "memory");
}

// mode: lane offset (0/32/64/96) and column offset of the shift address, number of shifts
__global__ void __launch_bounds__(128, 1) shift_probe(int lane_off, int col_off, int nshift, uint32_t* out, long long* clk) {
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  const uint32_t mine = tmem + (uint32_t(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) v[j] = uint32_t((warp * 32 + lane) * 256 + c0 + j);
    tmem_st16(mine + c0, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int i = 0; i < nshift; ++i) tmem_shift_down(tmem + (uint32_t(lane_off) << 16) + uint32_t(col_off));
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 1);
  if (threadIdx.x == 0) { t1 = clock64(); clk[0] = t1 - t0; }
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(mine + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 1; } } while (0)

static int probe(int lane_off, int col_off, int nshift, uint32_t* d_out, long long* d_clk) {
  shift_probe<<<1, 128>>>(lane_off, col_off, nshift, d_out, d_clk);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> h(128 * 64);
  long long clk = 0;
  CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost));
  printf("shift at lane %d col %d x%d: %lld clk\n", lane_off, col_off, nshift, clk);
  // summarise: for each column, the set of lanes whose value changed and where the new value came from
  int cmin = 64, cmax = -1, changed = 0;
  for (int l = 0; l < 128; ++l)
    for (int c = 0; c < 64; ++c)
      if (h[l * 64 + c] != uint32_t(l * 256 + c)) { ++changed; if (c < cmin) cmin = c; if (c > cmax) cmax = c; }
  printf("  cells changed: %d, columns %d..%d\n", changed, cmin, cmax);
  if (cmax >= 0) {
    const int c = cmin;
    printf("  column %d: lane <- source lane (source col): ", c);
    for (int l = 0; l < 128; ++l) {
      const uint32_t v = h[l * 64 + c];
      if (v != uint32_t(l * 256 + c)) printf("%d<-%u(%u) ", l, v >> 8, v & 255);
    }
    printf("\n");
  }
  return 0;
}

int main() {
  uint32_t* d_out; long long* d_clk;
  CK(cudaMalloc(&d_out, 128 * 64 * 4));
  CK(cudaMalloc(&d_clk, 64));
  if (probe(0, 0, 1, d_out, d_clk)) return 1;
  if (probe(0, 8, 1, d_out, d_clk)) return 1;
  if (probe(32, 16, 1, d_out, d_clk)) return 1;
  if (probe(0, 0, 2, d_out, d_clk)) return 1;
  if (probe(0, 0, 64, d_out, d_clk)) return 1;
  if (probe(0, 0, 256, d_out, d_clk)) return 1;
  return 0;
}
