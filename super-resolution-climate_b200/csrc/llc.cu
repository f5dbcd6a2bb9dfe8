// Raw LLC4320 fields -> region of interest, on the device.
//
// The reference reads one big-endian float32 file per variable and time step that holds only the OCEAN points of the
// 13 nx^2 LLC grid (".shrunk"), scatters it into the full grid through the land mask of a template file (hFacC != 0),
// marks land as NaN, unfolds the LLC faces into one (3 nx, 4 nx) east|west array and crops the region of interest:
//   sres/base/source/swot/raw.py:133-145 (load_file), :38-45 (subset_roi), sres/base/source/swot/util.py:3-55 (mds2d).
// Per file that is several passes over ~1 GB on the host.  Here the template is digested ONCE into a gather index of the
// region (for every ROI pixel: the position of its value in the shrunk file, or -1 for land); after that a file costs one
// host->device copy of its bytes and one gather kernel that byte-swaps on the fly.
//
//   unfolded (Y, X), nx = grid size:   X <  nx       -> d[Y nx + X]                       faces 1-3
//                                      X < 2nx       -> d[3 nx^2 + Y nx + (X - nx)]       faces 4-6
//                                      X >= 2nx      -> d[7 nx^2 + (X - 2nx) 3nx + (3nx - 1 - Y)]   faces 8-13, transposed + flipped
#include "internal.h"

namespace sres {

constexpr int kLlcBlock = 1024;   // grid points per counting block (32 mask words)

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// pass 1: land mask of the template as bits (1 = ocean: the big-endian float is neither +0 nor -0) and ocean points per block
__global__ void __launch_bounds__(kLlcBlock)
llc_mask_kernel(const uint32_t* __restrict__ tmpl_be, long long n, uint32_t* __restrict__ bits, int32_t* __restrict__ block_count) {
  const long long i = (long long)blockIdx.x * kLlcBlock + threadIdx.x;
  const bool ocean = i < n && (bswap32(tmpl_be[i]) & 0x7fffffffu) != 0u;
  const uint32_t word = __ballot_sync(0xffffffffu, ocean);
  __shared__ int s_cnt[kLlcBlock / 32];
  if ((threadIdx.x & 31) == 0) {
    bits[i >> 5] = word;
    s_cnt[threadIdx.x >> 5] = __popc(word);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w = 0; w < kLlcBlock / 32; ++w) c += s_cnt[w];
    block_count[blockIdx.x] = c;
  }
}

// pass 2: exclusive scan of the block counts (one block; the list has 13 nx^2 / 1024 entries); total -> *n_ocean
__global__ void __launch_bounds__(1024)
llc_scan_kernel(const int32_t* __restrict__ block_count, int nblocks, long long* __restrict__ block_prefix, long long* __restrict__ n_ocean) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < nblocks ? block_count[i] : 0;
    long long incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long w = s_warp[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      s_warp[threadIdx.x] = w;
    }
    __syncthreads();
    const long long before = s_carry + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0) + incl - v;
    if (i < nblocks) block_prefix[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry += s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_ocean = s_carry;
}

// pass 3: for every pixel of the region of interest, where its value sits in a shrunk file (-1 = land)
__global__ void __launch_bounds__(256)
llc_roi_index_kernel(const uint32_t* __restrict__ bits, const long long* __restrict__ block_prefix, int nx, int y0, int ys,
                     int x0, int xs, int32_t* __restrict__ roi_index) {
  const long long npix = (long long)ys * xs;
  const long long nx2 = (long long)nx * nx;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const int Y = y0 + int(p / xs), X = x0 + int(p % xs);
    long long src;
    if (X < nx) src = (long long)Y * nx + X;
    else if (X < 2 * nx) src = 3 * nx2 + (long long)Y * nx + (X - nx);
    else src = 7 * nx2 + (long long)(X - 2 * nx) * 3 * nx + (3 * nx - 1 - Y);
    const long long word = src >> 5;
    const uint32_t w = bits[word];
    int32_t out = -1;
    if ((w >> (src & 31)) & 1u) {
      long long rank = block_prefix[src / kLlcBlock];
      for (long long k = (src / kLlcBlock) * (kLlcBlock / 32); k < word; ++k) rank += __popc(bits[k]);
      rank += __popc(w & ((1u << (src & 31)) - 1u));
      out = (int32_t)rank;
    }
    roi_index[p] = out;
  }
}

// per file: out[p] = value of pixel p (byte-swapped from the big-endian file) or NaN for land
__global__ void __launch_bounds__(256)
llc_gather_kernel(const uint32_t* __restrict__ data_be, long long n_data, const int32_t* __restrict__ roi_index, long long npix,
                  float* __restrict__ out) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const int32_t i = roi_index[p];
    out[p] = (i >= 0 && i < n_data) ? __uint_as_float(bswap32(data_be[i])) : __uint_as_float(0x7fc00000u);
  }
}

}  // namespace sres

using namespace sres;

static long long llc_points(int nx) { return 13LL * nx * nx; }

extern "C" size_t sres_llc_index_workspace_bytes(int nx) {
  if (nx <= 0) return 0;
  const long long n = llc_points(nx);
  const long long nblocks = (n + kLlcBlock - 1) / kLlcBlock;
  return (size_t)(nblocks * (kLlcBlock / 32) * 4 + nblocks * 4 + nblocks * 8 + 64);
}

extern "C" int sres_llc_build_roi_index(const void* template_be, int nx, int y0, int ys, int x0, int xs, void* workspace,
                                        size_t workspace_bytes, int32_t* roi_index, int64_t* n_ocean_dev, void* stream) {
  if (!template_be || !workspace || !roi_index || !n_ocean_dev) return set_error(SRES_ERR_INVALID_ARG, "llc: null pointer");
  if (nx <= 0 || y0 < 0 || x0 < 0 || ys <= 0 || xs <= 0 || y0 + (long long)ys > 3LL * nx || x0 + (long long)xs > 4LL * nx)
    return set_error(SRES_ERR_INVALID_ARG, "llc: region of interest outside the unfolded (3 nx, 4 nx) grid");
  const long long n = llc_points(nx);
  if (n / 13 > 0x7fffffffLL / 13) return set_error(SRES_ERR_UNSUPPORTED, "llc: grid too large for 32-bit file positions");
  if (workspace_bytes < sres_llc_index_workspace_bytes(nx)) return set_error(SRES_ERR_INVALID_ARG, "llc: workspace too small");
  const long long nblocks = (n + kLlcBlock - 1) / kLlcBlock;
  uint32_t* bits = (uint32_t*)workspace;
  int32_t* counts = (int32_t*)(bits + nblocks * (kLlcBlock / 32));
  long long* prefix = (long long*)(((uintptr_t)(counts + nblocks) + 15) & ~(uintptr_t)15);
  cudaStream_t st = (cudaStream_t)stream;
  llc_mask_kernel<<<(unsigned)nblocks, kLlcBlock, 0, st>>>((const uint32_t*)template_be, n, bits, counts);
  SRES_CHECK_LAUNCH("llc: mask launch");
  llc_scan_kernel<<<1, 1024, 0, st>>>(counts, (int)nblocks, prefix, (long long*)n_ocean_dev);
  SRES_CHECK_LAUNCH("llc: scan launch");
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const long long npix = (long long)ys * xs;
  long long grid = (npix + 255) / 256;
  if (grid > sms * 32LL) grid = sms * 32LL;
  llc_roi_index_kernel<<<(unsigned)grid, 256, 0, st>>>(bits, prefix, nx, y0, ys, x0, xs, roi_index);
  SRES_CHECK_LAUNCH("llc: index launch");
  return SRES_OK;
}

extern "C" int sres_llc_gather_roi(const void* data_be, int64_t n_data, const int32_t* roi_index, int64_t npix, float* out,
                                   void* stream) {
  if (!data_be || !roi_index || !out || n_data < 0 || npix <= 0) return set_error(SRES_ERR_INVALID_ARG, "llc: bad argument");
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  long long grid = (npix + 255) / 256;
  if (grid > sms * 32LL) grid = sms * 32LL;
  llc_gather_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)data_be, n_data, roi_index, npix, out);
  SRES_CHECK_LAUNCH("llc: gather launch");
  return SRES_OK;
}
