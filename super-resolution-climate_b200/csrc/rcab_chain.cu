// A residual group's whole RCAB chain (forward) in ONE launch, image-resident (sm_100a).
//
// The tile-parallel path runs an RCAB as two launches -- the fused convolution pair (conv_pair.cu) and the channel-attention
// apply kernel (ca.cu) -- and both are grid-wide: every launch boundary is a barrier over all 148 SMs with its own pipeline
// fill and drain, and the squeeze-excite gate needs the pooled mean of a WHOLE image before one row of it can be scaled.
// Per RCAB that is 26 + 17 us at B = 64 for 18 us of tensor-core work and 10 us of streaming.
//
// But nothing in the chain couples two images (no BatchNorm, no inter-tile halo): image b's twenty RCABs are one
// dependent chain that no other image ever has to wait for.  So here a CLUSTER of K CTAs (K = 2 for 48 x 48 tiles) owns
// one image for the whole chain: its T = ceil((H+1)(W+1)/128) image-aligned M tiles are split between the K CTAs, and per
// RCAB each CTA runs
//     conv1 (+bias, ReLU -> T1)  |S1|  conv2 (+bias -> T2, channel sums in registers)  |S2|  gate MLP, x += T2 * s  |S3|
// with the same TMA -> tcgen05 -> TMEM -> TMA-store pipeline as the pair kernel.  S1..S3 are barrier.cluster -- K CTAs, not
// the grid -- so launch gaps and grid-wide fills / drains disappear and the grid may be any size (clusters of later waves
// are independent images).  (The hope that clusters would drift apart, one image's streaming phase under other images'
// tensor-core phases, did not come true: equal work keeps them in step, and the phase runs at the 24 B/clk one SM gets out
// of L2 wherever the others are.  Measured result, DESIGN 3.1c: 642 instead of 1032 launches per training step at the SAME
// step time; opt-in, SRES_RCAB_CHAIN=1.)  Weights of the next convolution are prefetched into the region the previous one has finished with; biases,
// squeeze-excite parameters and the saved tensors are addressed by the block index inside the kernel (4-D tensor maps:
// channel, image row, image, buffer).  Image-local tensor-map coordinates also make the padding work by itself: rows
// outside [0, (H+1)(W+1)) are zero-filled on load and dropped on store.
//
// Replaces, per launch, n_blocks x RCAB of the reference: sres/model/rcan/network.py:50-64 (RCAB: conv, ReLU, conv,
// CALayer, res += x) with CALayer network.py:31-47, as looped by ResidualGroup network.py:66-77.
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
#include "internal.h"
#include "conv_epi.cuh"

namespace sres {

constexpr int kCThreads = 384;
constexpr int kCAcc = 4;
constexpr int kCStages = 4;
constexpr int kCWBytes = 9 * 64 * 128;
constexpr int kCConvW = 64 * 64 * 9;
constexpr int kCTail = 3072;   // barriers (<= 512 B), biases, pooled sums, MLP vectors

struct ChainParams {
  int B, H, W, P, RP, T;       // image geometry; T = M tiles per image
  int K, tpc;                  // CTAs per image (cluster size), tiles per CTA
  int nstage, stage_rows, box_rows;
  int stagger;                 // cycles: cluster i starts (i % 3) * stagger late, so that the streaming phases of different images do not coincide
  int ap_nbuf;                 // >= 3: the streaming phase moves its rows with bulk copies through that many 48-row shared-memory buffers; 0: direct loads / stores
  int lend;                    // 1: the weight region a convolution does not need lends two more slots to the halo ring
  int off_ring, off_s16, off_tail;
  int n_blocks, hid;
  int xb_first, xb_ring;       // XB buffer holding block r's input: xb_ring ? (xb_first + r) % xb_ring : xb_first + r
  int t_first, t_fixed;        // T1 / T2 buffer of block r: t_fixed ? t_first : t_first + r
  const float* params;         // fp32 parameters of the first RCAB (conv1.w, conv1.b, conv2.w, conv2.b, du0.w, du0.b, du2.w, du2.b)
  long long rcab_stride;       // floats between consecutive RCABs
  const float* x_in;           // group input, fp32 PTL: block 0's trunk value
  float* xf;                   // running fp32 trunk (fp32 PTL), written by every block
  uint16_t* xb_base;           // bf16 XB buffers, buffer i at xb_base + i * buf_stride
  const uint16_t* t2_base;     // bf16 T2 buffers (read back by the apply phase)
  long long buf_stride;        // bf16 elements per buffer = B * RP * 64
  float* save_mean;            // [n or 1][B][64]
  float* save_s;
  long long save_stride;       // floats between consecutive blocks' saves (0: keep the last only)
  float* pool_scratch;         // [B][K][64]
  long long* timeline;         // bring-up only
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2,
                                             int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// 1-D bulk copies (no tensor map): global -> shared with mbarrier completion, shared -> global in a bulk group
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" : : "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ float chain_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// conv1 epilogue of one tile, row layout (thread = TMEM lane = output row): bias, ReLU, zero on padding rows, bf16 slab
// (64B-swizzled 32 rows x 32 channels per warp, as the TMA store expects).
__device__ __forceinline__ void chain_epi_rows(const ChainParams& p, const float* s_bias, uint8_t* s16, int r_img, int half,
                                               int lane, uint32_t trow) {
  const int y = r_img / p.P;
  const int x = r_img - y * p.P;
  const bool pad = (x == p.W) || (y >= p.H);
  uint32_t raw[32];
  tmem_ld32(trow, raw);
  tmem_ld_wait();
  const int sw3 = (lane >> 1) & 3;
  uint8_t* r16 = s16 + lane * 64;
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    const int c0 = half * 32 + ch * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * j);
      v[4 * j + 0] = fmaxf(__uint_as_float(raw[ch * 16 + 4 * j + 0]) + b4.x, 0.f);
      v[4 * j + 1] = fmaxf(__uint_as_float(raw[ch * 16 + 4 * j + 1]) + b4.y, 0.f);
      v[4 * j + 2] = fmaxf(__uint_as_float(raw[ch * 16 + 4 * j + 2]) + b4.z, 0.f);
      v[4 * j + 3] = fmaxf(__uint_as_float(raw[ch * 16 + 4 * j + 3]) + b4.w, 0.f);
    }
    if (pad) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
      *reinterpret_cast<uint4*>(r16 + (((ch * 2 + j) ^ sw3) << 4)) =
          make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                     pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  }
}

// conv2 epilogue of one tile, accumulator-fragment layout (thread = rows frq + 8j x column pairs 8k + 2fm): bias, zero on
// padding rows, bf16 slab, and the thread's running column sums (reduced across the warp once per block, not per tile).
__device__ __forceinline__ void chain_epi_frag(const ChainParams& p, const float (&bias8)[8], float (&cs)[8], uint8_t* s16,
                                               int r_img0, int lane, uint32_t trow) {
  const int fm = lane & 3, frq = lane >> 2;
  const int r_img = r_img0 + lane;
  const int y = r_img / p.P;
  const int x = r_img - y * p.P;
  const bool pad = (x == p.W) || (y >= p.H);
  uint32_t fa[16], fb[16];
  tmem_ld_frag16(trow, fa);
  tmem_ld_frag16(trow + (16u << 16), fb);
  const unsigned padmask = __ballot_sync(0xffffffffu, pad);
  tmem_ld_wait();
  const int sw3 = (frq >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = frq + 8 * j;
    const bool rpad = (padmask >> r) & 1u;
    uint8_t* row16 = s16 + r * 64 + 4 * fm;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t* src = (j < 2) ? fa : fb;
      float v0 = __uint_as_float(src[4 * k + 2 * (j & 1)]) + bias8[2 * k];
      float v1 = __uint_as_float(src[4 * k + 2 * (j & 1) + 1]) + bias8[2 * k + 1];
      if (rpad) { v0 = 0.f; v1 = 0.f; }
      cs[2 * k] += v0; cs[2 * k + 1] += v1;
      *reinterpret_cast<uint32_t*>(row16 + ((k ^ sw3) << 4)) = pack_bf16x2(v0, v1);
    }
  }
}

// kLend / kBulk / kDbg are compile-time: the default instance must not carry a single extra branch in the MMA-issuing
// thread's loop (the same kernel with the three switches as run-time flags was 10 % slower end to end).
template <bool kLend, bool kBulk, bool kDbg>
__global__ void __launch_bounds__(kCThreads, 1)
rcab_chain_fwd_kernel(const __grid_constant__ CUtensorMap tmXa, const __grid_constant__ CUtensorMap tmT1s,
                      const __grid_constant__ CUtensorMap tmT1a, const __grid_constant__ CUtensorMap tmT2s,
                      const __grid_constant__ CUtensorMap tmW, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem + p.off_ring;
  const int stage_bytes = p.stage_rows * 128;
  uint8_t* tail = smem + p.off_tail;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(tail);   // [kCStages]
  uint64_t* bar_empty = bar_full + kCStages;                // [kCStages]
  uint64_t* bar_w = bar_empty + kCStages;                   // [2] conv1 / conv2 weights
  uint64_t* bar_tfull = bar_w + 2;                          // [kCAcc]
  uint64_t* bar_tempty = bar_tfull + kCAcc;                 // [kCAcc]
  uint64_t* bar_ap = bar_tempty + kCAcc;                    // [4] streaming-phase buffers
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_ap + 4);
  float* s_bias = reinterpret_cast<float*>(tail + 256);     // [2][64]
  float* s_pool = s_bias + 128;                             // [4][64]
  float* sm_m = s_pool + 256;                               // [64]
  float* sm_h = sm_m + 64;                                  // [64]
  float* sm_s = sm_h + 64;                                  // [64]

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const int k = p.K > 1 ? (int)cluster_ctarank() : 0;
  const int b = blockIdx.x / p.K;
  const int j0 = k * p.tpc;
  const int j1 = min(p.T, j0 + p.tpc);
  const int n_my = max(0, j1 - j0);
  long long* tl = (kDbg && p.timeline) ? p.timeline + (size_t)blockIdx.x * 16 : nullptr;
  if (kDbg) {
    if (tl && tid == 0) { tl[0] = clock64(); for (int i = 1; i < 16; ++i) tl[i] = 0; }
    __syncthreads();
  }

  if (tid == 0) {
    tma_prefetch_desc(&tmXa);
    tma_prefetch_desc(&tmT1s);
    tma_prefetch_desc(&tmT1a);
    tma_prefetch_desc(&tmT2s);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < kCStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(&bar_w[0], 1);
    mbar_init(&bar_w[1], 1);
    for (int i = 0; i < kCAcc; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], 8);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bar_ap[i], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, kCAcc * 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
  pdl_launch_dependents();

  // weights of block 0 (packed by an earlier kernel of the stream, never by the immediate predecessor)
  const bool leader = (warp == 0 || warp == 1) ? elect_one() : false;
  if (warp == 0 && n_my > 0 && leader) {
    for (int c = 0; c < (kLend ? 1 : 2); ++c) {
      mbar_expect_tx(&bar_w[c], kCWBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d((smem + c * kCWBytes) + t * 64 * 128, &tmW, &bar_w[c], 0, c * 576 + t * 64);
    }
  }
  __syncwarp();
  pdl_wait();
  if (p.stagger > 0) {
    // All images cost the same, so left alone every cluster reaches its streaming phase (x += T2 * s: 0.9 MB per CTA
    // through L2) at the same moment and the phase runs at the AGGREGATE L2 rate (24 B/clk per SM) while the tensor cores
    // idle everywhere; a one-time offset of a third of a block period spreads the phases out.
    const long long until = clock64() + (long long)(b % 3) * p.stagger;
    while (clock64() < until) __nanosleep(200);
    __syncthreads();
  }

  int it = 0;                    // tiles this role has handled so far (TMEM stage and, without lending, ring slot / parity)
  // Lending mode: the ring has two slots of its own (0, 1) and two more (2, 3) inside the weight region the running
  // convolution does NOT need -- W[1] during conv1 (conv2's weights land there while conv1's last two tiles run), W[0] during
  // conv2 (the next block's conv1 weights likewise).  Tile jj of an n-tile phase takes slot (jj - n + 2) mod 4, so the phase
  // always ends on slots 0, 1 and the lent slots are idle two tile-times before the phase ends.  `sbits` holds, per slot,
  // the parity of its next use (all three roles walk the same sequence).
  uint32_t sbits = 0;
  int prev_last = -1;            // producer: slot of the previous phase's last tile
  int ap_it = 0;                 // streaming-phase chunks handled so far (buffer / parity), same in every thread
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;   // bring-up: cycles per phase, summed over the blocks (thread 128)
  const bool stamp = kDbg && tl != nullptr && tid == 128;
  if (stamp) tprev = clock64();
#define CSTAMP(i)                                  \
  do {                                             \
    if (kDbg && stamp) {                                 \
      const long long now__ = clock64();           \
      tacc[i] += now__ - tprev;                    \
      tprev = now__;                               \
    }                                              \
  } while (0)
  const uint32_t row_step = uint32_t(p.P) * 8;
  const int wq = warp & 3, ew = warp - 4, half = ew >> 2;
  uint8_t* s16 = smem + p.off_s16 + (ew >= 0 ? ew : 0) * 2048;
  const int fm = lane & 3;
  const int row_lo = j0 * 128, row_hi = min(p.RP, j1 * 128);

#pragma unroll 1
  for (int r = 0; r < p.n_blocks; ++r) {
    const float* prm = p.params + (long long)r * p.rcab_stride;
    const int xb_in = p.xb_ring ? (p.xb_first + r) % p.xb_ring : p.xb_first + r;
    const int xb_out = p.xb_ring ? (p.xb_first + r + 1) % p.xb_ring : p.xb_first + r + 1;
    const int tb = p.t_fixed ? p.t_first : p.t_first + r;
    if (tid < 64) s_bias[tid] = prm[kCConvW + tid];
    else if (tid < 128) s_bias[tid] = prm[2 * kCConvW + 64 + (tid - 64)];
    __syncthreads();

#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      // ============================== convolution c of block r ==============================
      if (warp == 0) {
        // ---- TMA producer ----
        if (leader) fence_proxy_async_all();   // the apply phase's generic stores (XB) precede these async-proxy reads
        if (kLend) {
          // lent slots overlay the previous convolution's weights: wait until its last MMAs have retired
          if (prev_last >= 0) mbar_wait(&bar_empty[prev_last], ((sbits >> prev_last) & 1u) ^ 1u, 6);
          uint8_t* lent = smem + (1 - c) * kCWBytes;
          for (int jj = 0; jj < n_my; ++jj, ++it) {
            const int slot = (jj - n_my + 2) & 3;
            const uint32_t ph = (sbits >> slot) & 1u;
            sbits ^= 1u << slot;
            mbar_wait(&bar_empty[slot], ph ^ 1, 1);
            const int row0 = (j0 + jj) * 128 - (p.P + 1);
            uint8_t* dst = slot < 2 ? smem_a + slot * stage_bytes : lent + (slot - 2) * stage_bytes;
            if (leader) {
              mbar_expect_tx(&bar_full[slot], stage_bytes);
              for (int rr = 0; rr < p.stage_rows; rr += p.box_rows)
                tma_load_4d(dst + rr * 128, c == 0 ? &tmXa : &tmT1a, &bar_full[slot], 0, row0 + rr, b, c == 0 ? xb_in : tb);
            }
            __syncwarp();
          }
          if (n_my > 0) prev_last = (n_my - 1 - n_my + 2) & 3;   // = 1
          // weights of the NEXT convolution (conv2 of this block / conv1 of the next) go where the lent slots were, once those
          // have been released for the last time in this phase -- two tile-times before the phase ends
          if (n_my > 0 && (c == 0 || r + 1 < p.n_blocks)) {
            mbar_wait(&bar_empty[2], ((sbits >> 2) & 1u) ^ 1u, 7);
            mbar_wait(&bar_empty[3], ((sbits >> 3) & 1u) ^ 1u, 7);
            if (leader) {
              const int conv = c == 0 ? 2 * r + 1 : 2 * (r + 1);
              mbar_expect_tx(&bar_w[1 - c], kCWBytes);
              for (int t = 0; t < 9; ++t) tma_load_2d(lent + t * 64 * 128, &tmW, &bar_w[1 - c], 0, conv * 576 + t * 64);
            }
            __syncwarp();
          }
        } else {
          for (int jj = 0; jj < n_my; ++jj, ++it) {
            const int slot = it % p.nstage;
            const uint32_t ph = (it / p.nstage) & 1;
            mbar_wait(&bar_empty[slot], ph ^ 1, 1);
            const int row0 = (j0 + jj) * 128 - (p.P + 1);
            uint8_t* dst = smem_a + slot * stage_bytes;
            if (leader) {
              mbar_expect_tx(&bar_full[slot], stage_bytes);
              for (int rr = 0; rr < p.stage_rows; rr += p.box_rows)
                tma_load_4d(dst + rr * 128, c == 0 ? &tmXa : &tmT1a, &bar_full[slot], 0, row0 + rr, b, c == 0 ? xb_in : tb);
            }
            __syncwarp();
          }
          // the next block's weights for this convolution go into the same region once this phase's MMAs have retired
          // (the last tile's ring slot has been released)
          if (r + 1 < p.n_blocks && n_my > 0) {
            const int last = it - 1;
            mbar_wait(&bar_empty[last % p.nstage], (last / p.nstage) & 1, 7);
            if (leader) {
              mbar_expect_tx(&bar_w[c], kCWBytes);
              for (int t = 0; t < 9; ++t)
                tma_load_2d((smem + c * kCWBytes) + t * 64 * 128, &tmW, &bar_w[c], 0, (2 * (r + 1) + c) * 576 + t * 64);
            }
            __syncwarp();
          }
        }
      } else if (warp == 1) {
        // ---- MMA issuer ----
        if (n_my > 0) {
          constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
          constexpr uint32_t dhi = sdesc_hi_sw128(1024);
          long long ww = 0;
          if (kDbg && tl) ww = clock64();
          mbar_wait(&bar_w[c], r & 1, 2);
          if (kDbg && tl && lane == 0) tl[10] += clock64() - ww;   // MMA warp waiting for the weights
          tc_fence_after();
          const uint32_t w_lo = sdesc_lo(smem_u32((smem + c * kCWBytes)), 16);
          const uint32_t a_lo0 = sdesc_lo(smem_u32(smem_a), 16);
          const uint32_t a_lent = sdesc_lo(smem_u32(smem + (1 - c) * kCWBytes), 16);
          for (int jj = 0; jj < n_my; ++jj, ++it) {
            int slot;
            uint32_t ph;
            if (kLend) {
              slot = (jj - n_my + 2) & 3;
              ph = (sbits >> slot) & 1u;
              sbits ^= 1u << slot;
            } else {
              slot = it % p.nstage;
              ph = (it / p.nstage) & 1;
            }
            const int acc = it % kCAcc;
            const uint32_t aph = (it / kCAcc) & 1;
            long long w0 = 0, w1 = 0;
            if (kDbg && tl) w0 = clock64();
            mbar_wait(&bar_tempty[acc], aph ^ 1, 3);
            if (kDbg && tl) w1 = clock64();
            mbar_wait(&bar_full[slot], ph, 4);
            if (kDbg && tl && lane == 0) { tl[11] += w1 - w0; tl[12] += clock64() - w1; }   // MMA warp waiting for the epilogue / for TMA
            tc_fence_after();
            const uint32_t a_tile = (kLend && slot >= 2) ? a_lent + uint32_t((slot - 2) * stage_bytes) / 16
                                                           : a_lo0 + uint32_t(slot * stage_bytes) / 16;
            const uint32_t d_tmem = tmem_base + uint32_t(acc * 64);
            if (leader) {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                const uint32_t a_tap = a_tile + uint32_t(t / 3) * row_step + uint32_t(t % 3) * 8;
                const uint32_t b_tap = w_lo + uint32_t(t * 64 * 8);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  if (t == 0 && kk == 0) umma_bf16_lohi<false>(d_tmem, a_tap, dhi, b_tap, dhi, idesc);
                  else umma_bf16_lohi<true>(d_tmem, a_tap + kk * 2, dhi, b_tap + kk * 2, dhi, idesc);
                }
              }
              umma_commit(&bar_empty[slot]);
              umma_commit(&bar_tfull[acc]);
            }
            __syncwarp();
          }
        }
      } else if (warp >= 4) {
        // ---- epilogue ----
        float bias8[8], cs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          bias8[i] = s_bias[64 + half * 32 + 8 * (i >> 1) + 2 * fm + (i & 1)];
          cs[i] = 0.f;
        }
        for (int jj = 0; jj < n_my; ++jj, ++it) {
          const int acc = it % kCAcc;
          const uint32_t aph = (it / kCAcc) & 1;
          const int r_img0 = (j0 + jj) * 128 + wq * 32;
          if (lane == 0) bulk_wait_read<0>();   // the slab may be overwritten
          __syncwarp();
          mbar_wait(&bar_tfull[acc], aph, 5);
          tc_fence_after();
          const uint32_t trow = tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc * 64 + half * 32);
          if (c == 0) chain_epi_rows(p, s_bias, s16, r_img0 + lane, half, lane, trow);
          else chain_epi_frag(p, bias8, cs, s16, r_img0, lane, trow);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_tempty[acc]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(c == 0 ? &tmT1s : &tmT2s, s16, half * 32, r_img0, b, tb);
            bulk_commit();
          }
        }
        if (lane == 0) bulk_wait_all<0>();   // this warp's slabs have reached global memory
        __syncwarp();
        if (c == 1) {
          // one column total per lane: column 16*bit4 + 8*bit3 + 2*(lane%4) + bit2 of this warp's 32-column half
          const float tot = frag_colsum(cs, lane);
          const int fcol = ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + 2 * fm + ((lane >> 2) & 1);
          s_pool[wq * 64 + half * 32 + fcol] = tot;
        }
      }
      CSTAMP(c == 0 ? 0 : 2);   // conv1 / conv2 phase of this thread's role
      if (c == 0) {
        // S1: every T1 tile of this image is in global memory before anyone loads a conv2 halo window
        if (p.K > 1) cluster_sync_all(); else __syncthreads();
        CSTAMP(1);
      }
    }

    // ============================== pooled mean -> gate ==============================
    __syncthreads();
    if (tid < 64) {
      const float sum = (s_pool[tid] + s_pool[64 + tid]) + (s_pool[128 + tid] + s_pool[192 + tid]);
      __stcg(p.pool_scratch + ((size_t)b * p.K + k) * 64 + tid, n_my > 0 ? sum : 0.f);
      __threadfence();
    }
    if (p.K > 1) cluster_sync_all(); else __syncthreads();   // S2
    CSTAMP(3);
    const float* w1 = prm + 2 * (kCConvW + 64);
    const float* b1 = w1 + p.hid * 64;
    const float* w2 = b1 + p.hid;
    const float* b2 = w2 + 64 * p.hid;
    if (tid < 64) {
      float tot = 0.f;
      for (int kk = 0; kk < p.K; ++kk) tot += __ldcg(p.pool_scratch + ((size_t)b * p.K + kk) * 64 + tid);
      sm_m[tid] = tot / float(p.H * p.W);
    }
    __syncthreads();
    {
      const float m0 = sm_m[lane], m1 = sm_m[lane + 32];
      for (int j = warp; j < p.hid; j += kCThreads / 32) {
        const float v = chain_warp_sum(fmaf(__ldg(w1 + j * 64 + lane), m0, __ldg(w1 + j * 64 + 32 + lane) * m1));
        if (lane == 0) sm_h[j] = fmaxf(v + __ldg(b1 + j), 0.f);
      }
    }
    __syncthreads();
    if (tid < 64) {
      float z = __ldg(b2 + tid);
      for (int j = 0; j < p.hid; ++j) z = fmaf(__ldg(w2 + tid * p.hid + j), sm_h[j], z);
      const float sg = 1.f / (1.f + expf(-z));
      sm_s[tid] = sg;
      if (k == 0) {
        p.save_mean[(long long)r * p.save_stride + b * 64 + tid] = sm_m[tid];
        p.save_s[(long long)r * p.save_stride + b * 64 + tid] = sg;
      }
    }
    __syncthreads();

    CSTAMP(4);
    // ============================== x <- x + T2 * s, bf16 copy ==============================
    {
      const int cg = tid & 7;
      float s8[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { s8[j] = sm_s[cg * 4 + j]; s8[4 + j] = sm_s[32 + cg * 4 + j]; }
      const float* xin = r == 0 ? p.x_in : p.xf;
      const uint16_t* t2 = p.t2_base + (long long)tb * p.buf_stride;
      uint16_t* xbo = p.xb_base + (long long)xb_out * p.buf_stride;
      if (kBulk) {
        // The rows of this CTA are contiguous in every tensor, so they move as plain 1-D bulk copies (async proxy, no tensor
        // map) through kCh-row buffers in the idle halo-ring / slab space: x chunk (fp32) and T2 chunk (bf16) in, updated in
        // place by one thread per 8 channels of a row, out again as the new fp32 trunk and its bf16 copy.  With ordinary
        // loads / stores the phase ran at 24 B/clk per SM (L1tex wavefronts: four half-used lines per warp instruction) --
        // 39 000 cycles per block against 26 000 for a whole convolution.
        constexpr int kCh = 48, kXB = kCh * 256, kBuf = kCh * 384;
        const int NB = p.ap_nbuf;
        const int n_rows = max(0, row_hi - row_lo);
        const int nch = (n_rows + kCh - 1) / kCh;
        const int lr = tid >> 3;
        auto issue_load = [&](int i, int seq) {
          const int rows_i = min(kCh, n_rows - i * kCh);
          const int bi = seq % NB;
          uint8_t* xs = smem_a + bi * kBuf;
          const size_t q0 = (size_t)b * p.RP + row_lo + i * kCh;
          mbar_expect_tx(&bar_ap[bi], uint32_t(rows_i) * 384u);
          bulk_load_1d(xs, xin + q0 * 64, uint32_t(rows_i) * 256u, &bar_ap[bi]);
          bulk_load_1d(xs + kXB, t2 + q0 * 64, uint32_t(rows_i) * 128u, &bar_ap[bi]);
        };
        if (tid == 0)
          for (int i = 0; i < min(nch, NB - 1); ++i) issue_load(i, ap_it + i);
        for (int i = 0; i < nch; ++i) {
          const int seq = ap_it + i;
          const int bi = seq % NB;
          const int rows_i = min(kCh, n_rows - i * kCh);
          uint8_t* xs = smem_a + bi * kBuf;
          mbar_wait(&bar_ap[bi], uint32_t(seq / NB) & 1u, 8);
          if (lr < rows_i) {
            float4* xa_p = reinterpret_cast<float4*>(xs + lr * 256 + cg * 16);
            float4* xc_p = reinterpret_cast<float4*>(xs + lr * 256 + 128 + cg * 16);
            uint2* ta_p = reinterpret_cast<uint2*>(xs + kXB + lr * 128 + cg * 8);
            uint2* tc_p = reinterpret_cast<uint2*>(xs + kXB + lr * 128 + 64 + cg * 8);
            const float4 xa = *xa_p, xc = *xc_p;
            const uint2 ta = *ta_p, tc = *tc_p;
            float o[8];
            o[0] = fmaf(bf16_lo(ta.x), s8[0], xa.x); o[1] = fmaf(bf16_hi(ta.x), s8[1], xa.y);
            o[2] = fmaf(bf16_lo(ta.y), s8[2], xa.z); o[3] = fmaf(bf16_hi(ta.y), s8[3], xa.w);
            o[4] = fmaf(bf16_lo(tc.x), s8[4], xc.x); o[5] = fmaf(bf16_hi(tc.x), s8[5], xc.y);
            o[6] = fmaf(bf16_lo(tc.y), s8[6], xc.z); o[7] = fmaf(bf16_hi(tc.y), s8[7], xc.w);
            *xa_p = make_float4(o[0], o[1], o[2], o[3]);
            *xc_p = make_float4(o[4], o[5], o[6], o[7]);
            *ta_p = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
            *tc_p = make_uint2(pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
          fence_proxy_async_smem();   // generic writes of the chunk -> the bulk stores' async-proxy reads
          __syncthreads();
          if (tid == 0) {
            const size_t q0 = (size_t)b * p.RP + row_lo + i * kCh;
            bulk_store_1d(p.xf + q0 * 64, xs, uint32_t(rows_i) * 256u);
            bulk_store_1d(xbo + q0 * 64, xs + kXB, uint32_t(rows_i) * 128u);
            bulk_commit();
            if (i + NB - 1 < nch) {
              bulk_wait_read<1>();   // every store but the one just issued has read its buffer: chunk i-1's buffer is free
              issue_load(i + NB - 1, seq + NB - 1);
            }
          }
        }
        ap_it += nch;
        if (tid == 0) bulk_wait_all<0>();   // the new trunk values are in global memory
      } else {
#pragma unroll 2
        for (int rr = row_lo + (tid >> 3); rr < row_hi; rr += kCThreads / 8) {
          const size_t q = (size_t)b * p.RP + rr;
          const uint2 ta = __ldcg(reinterpret_cast<const uint2*>(t2 + q * 64 + cg * 4));
          const uint2 tc = __ldcg(reinterpret_cast<const uint2*>(t2 + q * 64 + 32 + cg * 4));
          const float4 xa = __ldcg(reinterpret_cast<const float4*>(xin + q * 64 + cg * 4));
          const float4 xc = __ldcg(reinterpret_cast<const float4*>(xin + q * 64 + 32 + cg * 4));
          float o[8];
          o[0] = fmaf(bf16_lo(ta.x), s8[0], xa.x); o[1] = fmaf(bf16_hi(ta.x), s8[1], xa.y);
          o[2] = fmaf(bf16_lo(ta.y), s8[2], xa.z); o[3] = fmaf(bf16_hi(ta.y), s8[3], xa.w);
          o[4] = fmaf(bf16_lo(tc.x), s8[4], xc.x); o[5] = fmaf(bf16_hi(tc.x), s8[5], xc.y);
          o[6] = fmaf(bf16_lo(tc.y), s8[6], xc.z); o[7] = fmaf(bf16_hi(tc.y), s8[7], xc.w);
          __stcg(reinterpret_cast<float4*>(p.xf + q * 64 + cg * 4), make_float4(o[0], o[1], o[2], o[3]));
          __stcg(reinterpret_cast<float4*>(p.xf + q * 64 + 32 + cg * 4), make_float4(o[4], o[5], o[6], o[7]));
          __stcg(reinterpret_cast<uint2*>(xbo + q * 64 + cg * 4), make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3])));
          __stcg(reinterpret_cast<uint2*>(xbo + q * 64 + 32 + cg * 4), make_uint2(pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
        }
      }
    }
    __threadfence();
    fence_proxy_async_all();   // generic stores of XB -> the next block's TMA loads (this CTA's and the partner's)
    CSTAMP(5);
    if (p.K > 1) cluster_sync_all(); else __syncthreads();   // S3
    CSTAMP(6);
  }
  if (kDbg && stamp)
    for (int i = 0; i < 7; ++i) tl[1 + i] = tacc[i];
#undef CSTAMP

  tc_fence_before();
  __syncthreads();
  if (kDbg && tl && tid == 0) tl[9] = clock64();
  if (warp == 2) tmem_dealloc(tmem_base, kCAcc * 64);
}

// ---------------------------------------------------------------------------------------------------------------------
// Overlapped variant (SRES_CHAIN_OVERLAP=1): the streaming phase of block r runs UNDER conv1 of block r + 1.
//
// The plain kernel above spends 39 000 of its 100 000 cycles per block in x += T2 * s with the tensor core idle, and that
// phase cannot go faster by itself: it moves 0.92 MB per CTA at 24 B/clk, the rate one SM gets out of L2 for a mixed
// read / write stream (ncu: DRAM 26 %, L2 21 % of peak over the whole kernel -- nothing chip-wide is saturated).  But conv1 of
// the next block needs, for its tile j, only the updated rows of tiles j-1, j, j+1.  So six extra "stream" warps (512 threads
// per CTA) walk the CTA's tiles in a fixed order and publish each finished tile on a shared-memory mbarrier; the TMA producer
// of conv1 waits for tile j+1's barrier before it loads the halo window of tile j.  The two CTAs of a cluster walk their
// tiles AWAY from the boundary between them (rank 0 downwards, rank 1 upwards), so the one tile each needs from the other
// is the first one finished; it is published with a remote arrive on the partner's barrier.  Per block there remain two
// cluster barriers: after conv1 (T1 complete) and in the pooled-mean exchange.
// Measured: correct, and slower than the plain kernel -- six warps stream 0.92 MB in 54 000 cycles where twelve
// need 39 000, and the gated conv1 trails them; kept as a parity-tested opt-in.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kOThreads = 512;
constexpr int kOStream = 6;        // stream warps: 2, 3, 12, 13, 14, 15
constexpr int kOMaxTiles = 16;     // tiles per CTA (one publication barrier each)

__global__ void __launch_bounds__(kOThreads, 1)
rcab_chain_ovl_kernel(const __grid_constant__ CUtensorMap tmXa, const __grid_constant__ CUtensorMap tmT1s,
                      const __grid_constant__ CUtensorMap tmT1a, const __grid_constant__ CUtensorMap tmT2s,
                      const __grid_constant__ CUtensorMap tmW, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem + p.off_ring;
  const int stage_bytes = p.stage_rows * 128;
  uint8_t* tail = smem + p.off_tail;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(tail);   // [kCStages]
  uint64_t* bar_empty = bar_full + kCStages;                // [kCStages]
  uint64_t* bar_w = bar_empty + kCStages;                   // [2]
  uint64_t* bar_tfull = bar_w + 2;                          // [kCAcc]
  uint64_t* bar_tempty = bar_tfull + kCAcc;                 // [kCAcc]
  uint64_t* bar_xb = bar_tempty + kCAcc;                    // [kOMaxTiles] tile (in walking order) updated by the stream warps
  uint64_t* bar_nbr = bar_xb + kOMaxTiles;                  // [1] the partner CTA's boundary tile updated (remote arrives)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_nbr + 1);
  float* s_bias = reinterpret_cast<float*>(tail + 512);     // [2][64]
  float* s_pool = s_bias + 128;                             // [4][64]
  float* sm_m = s_pool + 256;
  float* sm_h = sm_m + 64;
  float* sm_s = sm_h + 64;

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const int k = p.K > 1 ? (int)cluster_ctarank() : 0;
  const int b = blockIdx.x / p.K;
  const int j0 = k * p.tpc;
  const int j1 = min(p.T, j0 + p.tpc);
  const int n_my = max(0, j1 - j0);
  const bool down = p.K > 1 && k == 0;                      // rank 0 walks its tiles downwards, away from the boundary
  auto tile_of = [&](int jj) { return down ? j1 - 1 - jj : j0 + jj; };
  long long* tl = p.timeline ? p.timeline + (size_t)blockIdx.x * 16 : nullptr;   // bring-up: cycles summed over the blocks
  if (tl && tid == 0) { tl[0] = clock64(); for (int i = 1; i < 16; ++i) tl[i] = 0; }

  if (tid == 0) {
    tma_prefetch_desc(&tmXa);
    tma_prefetch_desc(&tmT1s);
    tma_prefetch_desc(&tmT1a);
    tma_prefetch_desc(&tmT2s);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < kCStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(&bar_w[0], 1);
    mbar_init(&bar_w[1], 1);
    for (int i = 0; i < kCAcc; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], 8);
    }
    for (int i = 0; i < kOMaxTiles; ++i) mbar_init(&bar_xb[i], kOStream / 2);   // one group of three stream warps per tile
    mbar_init(bar_nbr, kOStream / 2);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, kCAcc * 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.K > 1) cluster_sync_all();   // the partner's barriers exist before anyone arrives on them remotely
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
  pdl_launch_dependents();

  const bool leader = (warp == 0 || warp == 1) ? elect_one() : false;
  if (warp == 0 && n_my > 0 && leader) {
    for (int c = 0; c < 2; ++c) {
      mbar_expect_tx(&bar_w[c], kCWBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d((smem + c * kCWBytes) + t * 64 * 128, &tmW, &bar_w[c], 0, c * 576 + t * 64);
    }
  }
  __syncwarp();
  pdl_wait();

  int it = 0;
  const uint32_t row_step = uint32_t(p.P) * 8;
  const int wq = warp & 3, ew = warp - 4, half = (ew >> 2) & 1;
  const bool is_epi = warp >= 4 && warp < 12;
  const bool is_stream = warp == 2 || warp == 3 || warp >= 12;
  uint8_t* s16 = smem + p.off_s16 + (is_epi ? ew : 0) * 2048;
  const int fm = lane & 3;
  auto xb_idx = [&](int r) { return p.xb_ring ? (p.xb_first + r) % p.xb_ring : p.xb_first + r; };
  auto t_idx = [&](int r) { return p.t_fixed ? p.t_first : p.t_first + r; };

  // one convolution (c = 0: conv1, 1: conv2) of block rb by the producer / MMA / epilogue warps; `gate`: the producer waits
  // for the stream warps' publication of block rb - 1's update before each halo window
  auto conv_phase = [&](const int c, const int rb, const bool gate) {
    const int tb = t_idx(rb);
    if (warp == 0) {
      if (leader) fence_proxy_async_all();
      const uint32_t xpar = uint32_t(rb - 1) & 1u;
      for (int jj = 0; jj < n_my; ++jj, ++it) {
        const int slot = it % p.nstage;
        const uint32_t ph = (it / p.nstage) & 1;
        if (gate) {
          long long g0 = 0;
          if (tl) g0 = clock64();
          mbar_wait(&bar_xb[jj], xpar, 9);                        // (tile jj - 1 was waited for one iteration ago)
          mbar_wait(&bar_xb[min(jj + 1, n_my - 1)], xpar, 9);
          if (jj == 0 && p.K > 1) mbar_wait(bar_nbr, xpar, 10);
          if (tl && lane == 0) tl[5] += clock64() - g0;   // producer waiting for the stream warps (own tiles + partner's boundary tile)
          if (leader) fence_proxy_async_all();   // the stream warps' generic stores -> this TMA load
        }
        mbar_wait(&bar_empty[slot], ph ^ 1, 1);
        const int row0 = tile_of(jj) * 128 - (p.P + 1);
        uint8_t* dst = smem_a + slot * stage_bytes;
        if (leader) {
          mbar_expect_tx(&bar_full[slot], stage_bytes);
          for (int rr = 0; rr < p.stage_rows; rr += p.box_rows)
            tma_load_4d(dst + rr * 128, c == 0 ? &tmXa : &tmT1a, &bar_full[slot], 0, row0 + rr, b, c == 0 ? xb_idx(rb) : tb);
        }
        __syncwarp();
      }
      if (rb + 1 < p.n_blocks && n_my > 0) {   // the next block's weights of this convolution, once this phase's MMAs have retired
        const int last = it - 1;
        mbar_wait(&bar_empty[last % p.nstage], (last / p.nstage) & 1, 7);
        if (leader) {
          mbar_expect_tx(&bar_w[c], kCWBytes);
          for (int t = 0; t < 9; ++t)
            tma_load_2d((smem + c * kCWBytes) + t * 64 * 128, &tmW, &bar_w[c], 0, (2 * (rb + 1) + c) * 576 + t * 64);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      if (n_my > 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
        constexpr uint32_t dhi = sdesc_hi_sw128(1024);
        mbar_wait(&bar_w[c], rb & 1, 2);
        tc_fence_after();
        const uint32_t w_lo = sdesc_lo(smem_u32(smem + c * kCWBytes), 16);
        const uint32_t a_lo0 = sdesc_lo(smem_u32(smem_a), 16);
        for (int jj = 0; jj < n_my; ++jj, ++it) {
          const int slot = it % p.nstage;
          const uint32_t ph = (it / p.nstage) & 1;
          const int acc = it % kCAcc;
          const uint32_t aph = (it / kCAcc) & 1;
          mbar_wait(&bar_tempty[acc], aph ^ 1, 3);
          mbar_wait(&bar_full[slot], ph, 4);
          tc_fence_after();
          const uint32_t a_tile = a_lo0 + uint32_t(slot * stage_bytes) / 16;
          const uint32_t d_tmem = tmem_base + uint32_t(acc * 64);
          if (leader) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint32_t a_tap = a_tile + uint32_t(t / 3) * row_step + uint32_t(t % 3) * 8;
              const uint32_t b_tap = w_lo + uint32_t(t * 64 * 8);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (t == 0 && kk == 0) umma_bf16_lohi<false>(d_tmem, a_tap, dhi, b_tap, dhi, idesc);
                else umma_bf16_lohi<true>(d_tmem, a_tap + kk * 2, dhi, b_tap + kk * 2, dhi, idesc);
              }
            }
            umma_commit(&bar_empty[slot]);
            umma_commit(&bar_tfull[acc]);
          }
          __syncwarp();
        }
      }
    } else if (is_epi) {
      float bias8[8], cs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        bias8[i] = s_bias[64 + half * 32 + 8 * (i >> 1) + 2 * fm + (i & 1)];
        cs[i] = 0.f;
      }
      for (int jj = 0; jj < n_my; ++jj, ++it) {
        const int acc = it % kCAcc;
        const uint32_t aph = (it / kCAcc) & 1;
        const int r_img0 = tile_of(jj) * 128 + wq * 32;
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        mbar_wait(&bar_tfull[acc], aph, 5);
        tc_fence_after();
        const uint32_t trow = tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc * 64 + half * 32);
        if (c == 0) chain_epi_rows(p, s_bias, s16, r_img0 + lane, half, lane, trow);
        else chain_epi_frag(p, bias8, cs, s16, r_img0, lane, trow);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[acc]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(c == 0 ? &tmT1s : &tmT2s, s16, half * 32, r_img0, b, tb);
          bulk_commit();
        }
      }
      if (lane == 0) bulk_wait_all<0>();
      __syncwarp();
      if (c == 1) {
        const float tot = frag_colsum(cs, lane);
        const int fcol = ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + 2 * fm + ((lane >> 2) & 1);
        s_pool[wq * 64 + half * 32 + fcol] = tot;
      }
    }
  };
  auto barrier_all = [&]() { if (p.K > 1) cluster_sync_all(); else __syncthreads(); };

  {
    const float* prm = p.params;
    if (tid < 64) s_bias[tid] = prm[kCConvW + tid];
    else if (tid < 128) s_bias[tid] = prm[2 * kCConvW + 64 + (tid - 64)];
  }
  __syncthreads();
  conv_phase(0, 0, false);
  barrier_all();   // T1 of block 0 complete

#pragma unroll 1
  for (int r = 0; r < p.n_blocks; ++r) {
    const float* prm = p.params + (long long)r * p.rcab_stride;
    long long c0 = 0;
    if (tl) c0 = clock64();
    conv_phase(1, r, false);
    if (tl && tid == 128) tl[1] += clock64() - c0;      // conv2 phase (epilogue thread)
    // ---- pooled mean -> gate ----
    __syncthreads();
    if (tid < 64) {
      const float sum = (s_pool[tid] + s_pool[64 + tid]) + (s_pool[128 + tid] + s_pool[192 + tid]);
      __stcg(p.pool_scratch + ((size_t)b * p.K + k) * 64 + tid, n_my > 0 ? sum : 0.f);
      __threadfence();
    }
    barrier_all();
    const float* w1 = prm + 2 * (kCConvW + 64);
    const float* b1 = w1 + p.hid * 64;
    const float* w2 = b1 + p.hid;
    const float* b2 = w2 + 64 * p.hid;
    if (tid < 64) {
      float tot = 0.f;
      for (int kk = 0; kk < p.K; ++kk) tot += __ldcg(p.pool_scratch + ((size_t)b * p.K + kk) * 64 + tid);
      sm_m[tid] = tot / float(p.H * p.W);
    }
    if (r + 1 < p.n_blocks) {   // biases of the next block (this block's epilogues are done with theirs)
      const float* pn = prm + p.rcab_stride;
      if (tid >= 128 && tid < 192) s_bias[tid - 128] = pn[kCConvW + (tid - 128)];
      else if (tid >= 192 && tid < 256) s_bias[64 + tid - 192] = pn[2 * kCConvW + 64 + (tid - 192)];
    }
    __syncthreads();
    {
      const float m0 = sm_m[lane], m1 = sm_m[lane + 32];
      for (int j = warp; j < p.hid; j += kOThreads / 32) {
        const float v = chain_warp_sum(fmaf(__ldg(w1 + j * 64 + lane), m0, __ldg(w1 + j * 64 + 32 + lane) * m1));
        if (lane == 0) sm_h[j] = fmaxf(v + __ldg(b1 + j), 0.f);
      }
    }
    __syncthreads();
    if (tid < 64) {
      float z = __ldg(b2 + tid);
      for (int j = 0; j < p.hid; ++j) z = fmaf(__ldg(w2 + tid * p.hid + j), sm_h[j], z);
      const float sg = 1.f / (1.f + expf(-z));
      sm_s[tid] = sg;
      if (k == 0) {
        p.save_mean[(long long)r * p.save_stride + b * 64 + tid] = sm_m[tid];
        p.save_s[(long long)r * p.save_stride + b * 64 + tid] = sg;
      }
    }
    __syncthreads();

    // ---- x <- x + T2 * s by the stream warps, tile by tile, under conv1 of the next block ----
    long long b0 = 0;
    if (tl) { b0 = clock64(); if (tid == 128) tl[2] += b0 - c0; }   // conv2 + pool exchange + MLP
    if (is_stream) {
      // Two groups of three warps take alternate tiles: a tile's publication needs a gpu-scope fence after its stores
      // (MEMBAR.ALL.GPU + the async-proxy fence: about a store round trip), and with all six warps on every tile that fence
      // sat on the critical path of every tile (5 900 cycles per tile).  With two groups one group's fence overlaps the
      // other's loads and arithmetic, and each group issues its next tile's first loads before it fences.
      const int sidx = warp < 4 ? warp - 2 : warp - 10;
      const int grp = sidx & 1, lr = (sidx >> 1) * 4 + (lane >> 3), cg = lane & 7;
      float s8[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { s8[j] = sm_s[cg * 4 + j]; s8[4 + j] = sm_s[32 + cg * 4 + j]; }
      const float* xin = r == 0 ? p.x_in : p.xf;
      const uint16_t* t2 = p.t2_base + (long long)t_idx(r) * p.buf_stride;
      uint16_t* xbo = p.xb_base + (long long)xb_idx(r + 1) * p.buf_stride;
      constexpr int kU = 6, kStep = (kOStream / 2) * 4, kHalf = kU * kStep;   // 2 x 72 rows >= one 128-row tile
      uint2 ta[kU], tc[kU];
      float4 xa[kU], xc[kU];
      auto issue_loads = [&](int jj, int h) {
        const int rbase = tile_of(jj) * 128, rend = min(p.RP, rbase + 128);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int rr = min(rbase + h * kHalf + lr + u * kStep, rend - 1);
          const size_t q = (size_t)b * p.RP + rr;
          ta[u] = __ldcg(reinterpret_cast<const uint2*>(t2 + q * 64 + cg * 4));
          tc[u] = __ldcg(reinterpret_cast<const uint2*>(t2 + q * 64 + 32 + cg * 4));
          xa[u] = __ldcg(reinterpret_cast<const float4*>(xin + q * 64 + cg * 4));
          xc[u] = __ldcg(reinterpret_cast<const float4*>(xin + q * 64 + 32 + cg * 4));
        }
      };
      auto update_rows = [&](int jj, int h) {
        const int rbase = tile_of(jj) * 128, rend = min(p.RP, rbase + 128);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int rr = rbase + h * kHalf + lr + u * kStep;
          if (rr < rend) {
            const size_t q = (size_t)b * p.RP + rr;
            float o[8];
            o[0] = fmaf(bf16_lo(ta[u].x), s8[0], xa[u].x); o[1] = fmaf(bf16_hi(ta[u].x), s8[1], xa[u].y);
            o[2] = fmaf(bf16_lo(ta[u].y), s8[2], xa[u].z); o[3] = fmaf(bf16_hi(ta[u].y), s8[3], xa[u].w);
            o[4] = fmaf(bf16_lo(tc[u].x), s8[4], xc[u].x); o[5] = fmaf(bf16_hi(tc[u].x), s8[5], xc[u].y);
            o[6] = fmaf(bf16_lo(tc[u].y), s8[6], xc[u].z); o[7] = fmaf(bf16_hi(tc[u].y), s8[7], xc[u].w);
            __stcg(reinterpret_cast<float4*>(p.xf + q * 64 + cg * 4), make_float4(o[0], o[1], o[2], o[3]));
            __stcg(reinterpret_cast<float4*>(p.xf + q * 64 + 32 + cg * 4), make_float4(o[4], o[5], o[6], o[7]));
            __stcg(reinterpret_cast<uint2*>(xbo + q * 64 + cg * 4), make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3])));
            __stcg(reinterpret_cast<uint2*>(xbo + q * 64 + 32 + cg * 4), make_uint2(pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
          }
        }
      };
      if (grp < n_my) issue_loads(grp, 0);
      for (int jj = grp; jj < n_my; jj += 2) {
        update_rows(jj, 0);
        issue_loads(jj, 1);
        update_rows(jj, 1);
        if (jj + 2 < n_my) issue_loads(jj + 2, 0);   // in flight while this tile's stores are fenced
        __threadfence();              // this thread's rows are visible gpu-wide ...
        fence_proxy_async_all();      // ... also to the async proxy (TMA loads of this CTA and of the partner)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bar_xb[jj]);
          if (jj == 0 && p.K > 1) mbar_arrive_cluster(mapa_u32(smem_u32(bar_nbr), uint32_t(k ^ 1)));
        }
      }
      if (tl && warp == 2 && lane == 0) tl[3] += clock64() - b0;   // streaming of this block (stream warp 2)
    } else if (r + 1 < p.n_blocks) {
      conv_phase(0, r + 1, true);
      if (tl && tid == 128) tl[4] += clock64() - b0;               // gated conv1 of the next block (epilogue thread)
    }
    barrier_all();   // T1 of block r + 1 complete, block r's update complete everywhere
    if (tl && tid == 128) tl[6] += clock64() - b0;                 // whole overlapped phase incl. the barrier
  }

  tc_fence_before();
  __syncthreads();
  if (tl && tid == 0) tl[9] = clock64();
  if (warp == 2) tmem_dealloc(tmem_base, kCAcc * 64);
}

// shared-memory plan; false when two weight regions leave no room for a two-slot halo ring
static bool chain_plan(int H, int W, ChainParams* p, size_t* smem_bytes) {
  const int smem_max = 232448;
  if (W + 2 > 128) return false;
  const int rows = 128 + 2 * (W + 2);
  const int box = (rows + 7) / 8 * 8 <= 256 ? (rows + 7) / 8 * 8 : 64;
  const int stage_rows = (rows + box - 1) / box * box;
  int ns = (smem_max - 1024 - 2 * kCWBytes - 16384 - kCTail) / (stage_rows * 128);
  if (ns > kCStages) ns = kCStages;
  if (ns < 2) return false;
  static const int lend_env = [] { const char* e = getenv("SRES_CHAIN_LEND"); return e ? atoi(e) : 0; }();
  const int lend = (lend_env && ns < 4 && 2 * stage_rows * 128 <= kCWBytes) ? 1 : 0;
  if (lend) ns = 2;
  (void)H;
  if (p) {
    p->box_rows = box; p->stage_rows = stage_rows; p->nstage = ns; p->lend = lend;
    static const int bulk_env = [] { const char* e = getenv("SRES_CHAIN_BULK"); return e ? atoi(e) : 0; }();
    const int nb = (ns * stage_rows * 128 + 16384) / (48 * 384);   // 48-row buffers that fit the halo ring + epilogue slabs
    p->ap_nbuf = (bulk_env && nb >= 3) ? (nb > 4 ? 4 : nb) : 0;
    int off = 2 * kCWBytes;
    p->off_ring = off; off += ns * stage_rows * 128;
    p->off_s16 = off; off += 16384;
    p->off_tail = off; off += kCTail;
    *smem_bytes = (size_t)off + 1024;
  }
  return true;
}

// 4-D TMA descriptor over bf16 activation buffers [nbuf][B][RP][64]: coordinates (channel, image row, image, buffer);
// box = box_ch channels x box_rows rows; image rows outside [0, RP) are zero-filled on load and dropped on store
static int make_tmap_act4d(CUtensorMap* out, const void* base, int RP, int B, int nbuf, int box_ch, int box_rows) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn enc = [] {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      fp = nullptr;
    }
    return (EncodeTiledFn)fp;
  }();
  if (!enc) return set_error(SRES_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 127) != 0) return set_error(SRES_ERR_INVALID_ARG, "TMA operand must be 128-byte aligned");
  cuuint64_t dims[4] = {64, (cuuint64_t)RP, (cuuint64_t)B, (cuuint64_t)nbuf};
  cuuint64_t strides[3] = {128, (cuuint64_t)RP * 128, (cuuint64_t)B * RP * 128};
  cuuint32_t box[4] = {(cuuint32_t)box_ch, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult rc = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (4-D activations) failed (CUresult %d, RP %d, B %d, nbuf %d, box %d x %d)",
             (int)rc, RP, B, nbuf, box_ch, box_rows);
    return set_error(SRES_ERR_CUDA, msg);
  }
  return SRES_OK;
}

}  // namespace sres

using namespace sres;

extern "C" int sres_rcab_chain_supported(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  if ((long long)B * (H + 1) * (W + 1) > 0x7fffff00LL) return 0;
  return chain_plan(H, W, nullptr, nullptr) ? 1 : 0;
}

extern "C" size_t sres_rcab_chain_scratch_bytes(int B) { return (size_t)(B > 0 ? B : 0) * 2 * 64 * sizeof(float); }

extern "C" int sres_rcab_chain_fwd(const sres_rcab_chain_args* a, void* stream_) {
  if (!a) return set_error(SRES_ERR_INVALID_ARG, "rcab_chain: null arguments");
  if (!sres_rcab_chain_supported(a->B, a->H, a->W)) return set_error(SRES_ERR_UNSUPPORTED, "rcab_chain: geometry not supported");
  if (!a->xb_bf16 || !a->t1_bf16 || !a->t2_bf16 || !a->wpack_bf16 || !a->params || !a->x_in_f32 || !a->x_f32 || !a->save_mean ||
      !a->save_s || !a->scratch)
    return set_error(SRES_ERR_INVALID_ARG, "rcab_chain: null pointer");
  if (a->n_blocks < 1 || a->hidden < 1 || a->hidden > 64 || a->xb_count < 2 || a->t_count < 1)
    return set_error(SRES_ERR_INVALID_ARG, "rcab_chain: bad block / buffer counts");
  if (a->xb_ring ? (a->xb_ring < 2 || a->xb_ring > a->xb_count) : (a->xb_first < 0 || a->xb_first + a->n_blocks >= a->xb_count))
    return set_error(SRES_ERR_INVALID_ARG, "rcab_chain: XB buffer range");
  if (a->t_fixed ? (a->t_first < 0 || a->t_first >= a->t_count) : (a->t_first < 0 || a->t_first + a->n_blocks > a->t_count))
    return set_error(SRES_ERR_INVALID_ARG, "rcab_chain: T1 / T2 buffer range");
  cudaStream_t stream = (cudaStream_t)stream_;
  ChainParams p{};
  p.B = a->B; p.H = a->H; p.W = a->W; p.P = a->W + 1; p.RP = (a->H + 1) * (a->W + 1);
  p.T = (p.RP + 127) / 128;
  static const int k_env = [] { const char* e = getenv("SRES_CHAIN_K"); return e ? atoi(e) : 0; }();
  p.K = (k_env == 1 || p.T < 2) ? 1 : 2;
  p.tpc = (p.T + p.K - 1) / p.K;
  p.n_blocks = a->n_blocks; p.hid = a->hidden;
  p.xb_first = a->xb_first; p.xb_ring = a->xb_ring; p.t_first = a->t_first; p.t_fixed = a->t_fixed;
  p.params = a->params; p.rcab_stride = a->rcab_stride;
  p.x_in = a->x_in_f32; p.xf = a->x_f32;
  p.xb_base = (uint16_t*)a->xb_bf16; p.t2_base = (const uint16_t*)a->t2_bf16;
  p.buf_stride = (long long)a->B * p.RP * 64;
  p.save_mean = a->save_mean; p.save_s = a->save_s; p.save_stride = a->save_stride;
  p.pool_scratch = (float*)a->scratch;
  p.timeline = (long long*)a->debug_timeline;
  {
    static const int stagger_env = [] { const char* e = getenv("SRES_CHAIN_STAGGER"); return e ? atoi(e) : 0; }();
    p.stagger = (a->n_blocks >= 4 && a->B * 2 <= device_sm_count()) ? stagger_env : 0;
  }
  size_t smem = 0;
  if (!chain_plan(a->H, a->W, &p, &smem)) return set_error(SRES_ERR_UNSUPPORTED, "rcab_chain: no room for the halo ring");

  CUtensorMap tmXa, tmT1s, tmT1a, tmT2s, tmW;
  int rc;
  if ((rc = make_tmap_act4d(&tmXa, a->xb_bf16, p.RP, a->B, a->xb_count, 64, p.box_rows))) return rc;
  if ((rc = make_tmap_act4d(&tmT1s, a->t1_bf16, p.RP, a->B, a->t_count, 32, 32))) return rc;
  if ((rc = make_tmap_act4d(&tmT1a, a->t1_bf16, p.RP, a->B, a->t_count, 64, p.box_rows))) return rc;
  if ((rc = make_tmap_act4d(&tmT2s, a->t2_bf16, p.RP, a->B, a->t_count, 32, 32))) return rc;
  if ((rc = make_tmap_rows64(&tmW, a->wpack_bf16, (uint64_t)a->n_blocks * 2 * 576, 64))) return rc;

  cudaError_t e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a->B * p.K); cfg.blockDim = dim3(kCThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (p.K > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = p.K; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  count_launch();
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);   // (cheap; per device and instance)
    if (e2 != cudaSuccess) return e2;
    return cudaLaunchKernelEx(&cfg, kern, tmXa, tmT1s, tmT1a, tmT2s, tmW, p);
  };
  static const int ovl_env = [] { const char* e = getenv("SRES_CHAIN_OVERLAP"); return e ? atoi(e) : 0; }();
  if (ovl_env && p.tpc <= kOMaxTiles) {
    cfg.blockDim = dim3(kOThreads);
    e = launch(rcab_chain_ovl_kernel);
    if (e != cudaSuccess) return set_cuda_error(e, "rcab_chain: launch (overlapped)");
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "rcab_chain: launch (overlapped)");
    return SRES_OK;
  }
  const bool dbg = p.timeline != nullptr, lend = p.lend != 0, bulk = p.ap_nbuf >= 3;
  if (dbg) e = lend ? (bulk ? launch(rcab_chain_fwd_kernel<true, true, true>) : launch(rcab_chain_fwd_kernel<true, false, true>))
                    : (bulk ? launch(rcab_chain_fwd_kernel<false, true, true>) : launch(rcab_chain_fwd_kernel<false, false, true>));
  else e = lend ? (bulk ? launch(rcab_chain_fwd_kernel<true, true, false>) : launch(rcab_chain_fwd_kernel<true, false, false>))
                : (bulk ? launch(rcab_chain_fwd_kernel<false, true, false>) : launch(rcab_chain_fwd_kernel<false, false, false>));
  if (e != cudaSuccess) return set_cuda_error(e, "rcab_chain: launch");
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "rcab_chain: launch");
  return SRES_OK;
}
