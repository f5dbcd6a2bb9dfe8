// 3x3 convolution over the padded tile layout as a tcgen05 implicit GEMM (sm_100a).
//
//   D[q][n] = sum_{tap,k} A[q + off(tap)][k] * Wp[tap][n][k]        q = PTL row, k = 64 in-features
//
// One persistent CTA per SM.  A "stage" is MT consecutive 128-row M tiles.  For a stage the TMA
// producer loads ONE halo window of PTL rows [q0 - (P+1), q0 + 128*MT + (P+1)) into shared memory
// (128B-swizzled rows, out-of-range rows zero-filled by TMA); all 9 taps of all MT tiles are then
// UMMA operands taken from that single window by shifting the descriptor start address by
// (ky*P + kx) rows -- the halo is read from L2 once, not once per tap.  The packed weights of all 9
// taps stay resident in shared memory for the life of the CTA.  Accumulators live in TMEM, double
// buffered so the epilogue warps drain tile i while the tensor core works on tile i+1.
//
// Warp roles (256 threads): 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue.
//
// Replaces nn.Conv2d(64, n, 3, padding=1) forward / input-gradient of the reference
// (sres/model/common/cnn.py:8-9 used at sres/model/rcan/network.py:14-16,55,71, blocks.py:62-64).
#include "ptx.cuh"
#include "internal.h"

namespace sres {

constexpr int kMaxStages = 4;
constexpr int kBoxRows = 64;  // rows per TMA box (8 KB)

struct ConvKParams {
  int B, H, W, P, R;      // input geometry, P = W+1, R = H+1
  int npos;               // rows of the input PTL
  int n_tiles, n_stages;  // 128-row tiles, MT-tile stages
  int mt;                 // tiles per stage
  int nstage;             // smem ring depth
  int stage_rows;         // rows per smem stage (multiple of kBoxRows)
  int n_out, c_real;
  unsigned flags;
  int map_mode, sub_i, sub_j, sf;
  int debug_flags;
  const float* bias;
  const float* resid;
  const float* resid2;
  const uint16_t* mask;
  float* out_f32;
  uint16_t* out_bf16;
  float* pool_part;
  float* out_nchw;
  // TMA-staged epilogue (identity mapping): per-warp 32-row slabs in shared memory
  int tma_epi;          // 1: outputs / addends go through smem slabs + TMA, 0: direct global accesses
  int use_o16, use_msk, use_r32, use_o32;
  int off_s16, off_msk, off_s32, off_tail;  // byte offsets from the aligned smem base
};

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

template <int N_OUT>
__global__ void __launch_bounds__(256, 1)
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmO16, const __grid_constant__ CUtensorMap tmMsk,
                     const __grid_constant__ CUtensorMap tmR32, const __grid_constant__ CUtensorMap tmO32,
                     const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B aligned carve-up (128B swizzle atoms are 1024 B).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kWBytes = 9 * N_OUT * 128;
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + kWBytes;
  const int stage_bytes = p.stage_rows * 128;
  uint8_t* tail = smem + p.off_tail;
  uint8_t* slab16 = smem + p.off_s16;   // [mt][4 warps][32 rows x 128 B]  bf16 output staging
  uint8_t* slabmk = smem + p.off_msk;   // [mt][4 warps][32 rows x 128 B]  bf16 ReLU-mask tile (TMA loaded)
  uint8_t* slab32 = smem + p.off_s32;   // [mt][4 warps][2 halves][32 rows x 128 B]  fp32 addend in -> fp32 output
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(tail);  // [kMaxStages]
  uint64_t* bar_empty = bar_full + kMaxStages;              // [kMaxStages]
  uint64_t* bar_w = bar_empty + kMaxStages;                 // [1]
  uint64_t* bar_tfull = bar_w + 1;                          // [2]
  uint64_t* bar_tempty = bar_tfull + 2;                     // [2]
  uint64_t* bar_in = bar_tempty + 2;                        // [4 warps][2 tiles] epilogue operand loads
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_in + 8);
  float* s_bias = reinterpret_cast<float*>(tmem_holder + 2);  // [N_OUT]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int acc_cols = p.mt * N_OUT;  // TMEM columns per accumulator stage
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(2 * acc_cols)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], 4);  // one arrive per epilogue warp
    }
    for (int i = 0; i < 8; ++i) mbar_init(&bar_in[i], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, tmem_cols);
    tmem_relinquish();
  }
  if (threadIdx.x < N_OUT) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(bar_w, kWBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(smem_w + t * N_OUT * 128, &tmW, bar_w, 0, t * N_OUT);
      int it = 0;
      for (int s = blockIdx.x; s < p.n_stages; s += gridDim.x, ++it) {
        const int slot = it % p.nstage;
        const uint32_t ph = (it / p.nstage) & 1;
        mbar_wait(&bar_empty[slot], ph ^ 1, 1);
        mbar_expect_tx(&bar_full[slot], stage_bytes);
        const int row0 = s * p.mt * 128 - (p.P + 1);
        uint8_t* dst = smem_a + slot * stage_bytes;
        for (int r = 0; r < p.stage_rows; r += kBoxRows)
          tma_load_2d(dst + r * 128, &tmA, &bar_full[slot], 0, row0 + r);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, N_OUT, 0, 0);
      const uint32_t w_addr = smem_u32(smem_w);
      const uint32_t a_addr0 = smem_u32(smem_a);
      const bool use_bo = p.debug_flags & 1;
      mbar_wait(bar_w, 0, 2);
      int it = 0;
      for (int s = blockIdx.x; s < p.n_stages; s += gridDim.x, ++it) {
        const int slot = it % p.nstage;
        const uint32_t ph = (it / p.nstage) & 1;
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&bar_tempty[acc], aph ^ 1, 3);
        mbar_wait(&bar_full[slot], ph, 4);
        tc_fence_after();
        const uint32_t a_stage = a_addr0 + slot * stage_bytes;
        for (int m = 0; m < p.mt; ++m) {
          if ((s * p.mt + m) >= p.n_tiles) break;
          const uint32_t d_tmem = tmem_base + uint32_t(acc * acc_cols + m * N_OUT);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int ky = t / 3, kx = t % 3;
            const uint32_t a_tap = a_stage + uint32_t((m * 128 + ky * p.P + kx) * 128);
            const uint32_t b_tap = w_addr + uint32_t(t * N_OUT * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t aa = a_tap + k * 32, bb = b_tap + k * 32;
              const uint64_t da = make_sdesc_sw128(aa, 16, 1024, use_bo ? (aa >> 7) & 7 : 0);
              const uint64_t db = make_sdesc_sw128(bb, 16, 1024, 0);
              umma_bf16(d_tmem, da, db, idesc, (t | k) ? 1u : 0u);
            }
          }
        }
        umma_commit(&bar_empty[slot]);  // smem slot free once these MMAs retire
        umma_commit(&bar_tfull[acc]);   // accumulators ready
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int RP = p.R * p.P;
    int it = 0;
    for (int s = blockIdx.x; s < p.n_stages; s += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      if (p.tma_epi) {
        // the slabs of the previous stage must have been read out by their TMA stores; then prefetch this
        // stage's fp32 addend / mask tiles while the tensor core is still working on it
        if (lane == 0) {
          bulk_wait_read<0>();
          if (p.use_msk | p.use_r32) {
            for (int m = 0; m < p.mt; ++m) {
              const int tile = s * p.mt + m;
              if (tile >= p.n_tiles) break;
              const int row0 = tile * 128 + wq * 32;
              uint64_t* bi = &bar_in[wq * 2 + m];
              mbar_expect_tx(bi, (p.use_msk ? 4096u : 0u) + (p.use_r32 ? 8192u : 0u));
              if (p.use_msk) tma_load_2d(slabmk + (m * 4 + wq) * 4096, &tmMsk, bi, 0, row0);
              if (p.use_r32) {
                tma_load_2d(slab32 + (m * 4 + wq) * 8192, &tmR32, bi, 0, row0);
                tma_load_2d(slab32 + (m * 4 + wq) * 8192 + 4096, &tmR32, bi, 32, row0);
              }
            }
          }
        }
        __syncwarp();
      }
      mbar_wait(&bar_tfull[acc], aph, 5);
      tc_fence_after();
      for (int m = 0; m < p.mt; ++m) {
        const int tile = s * p.mt + m;
        if (tile >= p.n_tiles) break;
        if (p.tma_epi) {
          // ---------------- TMA-staged epilogue (identity mapping, 64 outputs) ----------------
          const int row0 = tile * 128 + wq * 32;
          const int q = row0 + lane;
          const int b = q / RP;
          const int rem = q - b * RP;
          const int y = rem / p.P;
          const int x = rem - y * p.P;
          const bool pad = (x == p.W) || (y == p.H) || (q >= p.npos);
          const int seg = (b != (tile * 128) / RP) ? 1 : 0;
          const uint32_t trow = tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc * acc_cols + m * N_OUT);
          uint8_t* s16 = slab16 + (m * 4 + wq) * 4096 + lane * 128;
          const uint8_t* smk = slabmk + (m * 4 + wq) * 4096 + lane * 128;
          uint8_t* s32 = slab32 + (m * 4 + wq) * 8192 + lane * 128;
          const int sw = lane & 7;
          if (p.use_msk | p.use_r32) mbar_wait(&bar_in[wq * 2 + m], it & 1, 6);
#pragma unroll 1
          for (int ch = 0; ch < N_OUT / 16; ++ch) {
            uint32_t raw[16];
            tmem_ld16(trow + ch * 16, raw);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) + s_bias[ch * 16 + j];
            uint8_t* h32 = s32 + (ch >> 1) * 4096;   // half row (32 floats) this chunk lives in
            const int c32 = (ch & 1) * 4;            // first 16-byte chunk inside that half
            if (p.use_r32) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 r = *reinterpret_cast<const float4*>(h32 + (((c32 + j) ^ sw) << 4));
                v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
              }
            }
            if (p.resid2 && q < p.npos) {
              const float4* rp = reinterpret_cast<const float4*>(p.resid2 + (long long)q * 64 + ch * 16);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 r = rp[j];
                v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
              }
            }
            if (p.flags & SRES_EPI_RELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.use_msk) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint4 mk = *reinterpret_cast<const uint4*>(smk + (((ch * 2 + j) ^ sw) << 4));
                const uint32_t w4[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (!(bf16_lo(w4[e]) > 0.f)) v[8 * j + 2 * e] = 0.f;
                  if (!(bf16_hi(w4[e]) > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
                }
              }
            }
            if (pad) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (p.use_o32) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(h32 + (((c32 + j) ^ sw) << 4)) =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (p.use_o16) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                *reinterpret_cast<uint4*>(s16 + (((ch * 2 + j) ^ sw) << 4)) =
                    make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                               pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
            }
            if (p.flags & SRES_EPI_POOL) {
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
                float* pp = p.pool_part + ((long long)tile * 2 * 4 + wq) * 64 + ch * 16 + (lane >> 1);
                pp[0] = s0;
                pp[4 * 64] = s1;
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.use_o16) tma_store_2d(&tmO16, slab16 + (m * 4 + wq) * 4096, 0, row0);
            if (p.use_o32) {
              tma_store_2d(&tmO32, slab32 + (m * 4 + wq) * 8192, 0, row0);
              tma_store_2d(&tmO32, slab32 + (m * 4 + wq) * 8192 + 4096, 32, row0);
            }
            bulk_commit();
          }
          continue;
        }
        const int q = tile * 128 + wq * 32 + lane;
        const bool inrange = q < p.npos;
        const int b = q / RP;
        const int rem = q - b * RP;
        const int y = rem / p.P;
        const int x = rem - y * p.P;
        const bool pad = (x == p.W) || (y == p.H) || !inrange;
        // output row
        long long oq = q;
        bool ovalid = inrange;
        if (p.map_mode == SRES_MAP_SHUFFLE) {
          const int P2 = p.sf * p.W + 1, R2 = p.sf * p.H + 1;
          const int oy = p.sf * y + p.sub_i, ox = p.sf * x + p.sub_j;
          ovalid = inrange && oy < R2 && ox < P2;
          oq = (long long)b * R2 * P2 + (long long)oy * P2 + ox;
        } else if (p.map_mode == SRES_MAP_UNSHUFFLE) {
          const int Pl = p.W / p.sf + 1, Rl = p.H / p.sf + 1;
          const int sub = (y % p.sf) * p.sf + (x % p.sf);
          oq = (long long)sub * p.B * Rl * Pl + (long long)b * Rl * Pl + (long long)(y / p.sf) * Pl + (x / p.sf);
        }
        const int seg = (b != (tile * 128) / RP) ? 1 : 0;
        const uint32_t trow = tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc * acc_cols + m * N_OUT);
#pragma unroll 1
        for (int ch = 0; ch < N_OUT / 16; ++ch) {
          uint32_t raw[16];
          tmem_ld16(trow + ch * 16, raw);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) + s_bias[ch * 16 + j];
          if (p.resid && ovalid) {
            const float4* rp = reinterpret_cast<const float4*>(p.resid + oq * 64 + ch * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 r = rp[j];
              v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
            }
          }
          if (p.resid2 && ovalid) {
            const float4* rp = reinterpret_cast<const float4*>(p.resid2 + oq * 64 + ch * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 r = rp[j];
              v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
            }
          }
          if (p.flags & SRES_EPI_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.mask && inrange) {
            const uint4* mp = reinterpret_cast<const uint4*>(p.mask + (long long)q * 64 + ch * 16);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4 mk = mp[j];
              uint32_t w4[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (!(bf16_lo(w4[e]) > 0.f)) v[8 * j + 2 * e] = 0.f;
                if (!(bf16_hi(w4[e]) > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
              }
            }
          }
          if (pad) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
          if (ovalid) {
            if (p.out_f32) {
              float4* op = reinterpret_cast<float4*>(p.out_f32 + oq * 64 + ch * 16);
#pragma unroll
              for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (p.out_bf16) {
              uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + oq * 64 + ch * 16);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                op[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                   pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
            }
            if (p.out_nchw && !pad) {
              for (int c = 0; c < p.c_real - ch * 16 && c < 16; ++c)
                p.out_nchw[(((long long)b * p.c_real + ch * 16 + c) * p.H + y) * p.W + x] = v[c];
            }
          }
          if (p.flags & SRES_EPI_POOL) {
            // v is already 0 at padding rows.
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
              float* pp = p.pool_part + ((long long)tile * 2 * 4 + wq) * 64 + ch * 16 + (lane >> 1);
              pp[0] = s0;
              pp[4 * 64] = s1;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[acc]);
    }
    if (p.tma_epi && lane == 0) bulk_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int launch_conv(const sres_conv_args* a, cudaStream_t stream) {
  if (!a || !a->in_bf16 || !a->wpack_bf16) return set_error(SRES_ERR_INVALID_ARG, "conv: null input");
  if (a->n_out != 64 && a->n_out != 16) return set_error(SRES_ERR_UNSUPPORTED, "conv: n_out must be 64 or 16");
  if (a->B <= 0 || a->H <= 0 || a->W <= 0) return set_error(SRES_ERR_INVALID_ARG, "conv: bad geometry");
  const int sf = a->shuffle_factor > 0 ? a->shuffle_factor : 2;
  if (a->map_mode == SRES_MAP_UNSHUFFLE && (a->H % sf || a->W % sf))
    return set_error(SRES_ERR_INVALID_ARG, "conv: unshuffle factor must divide H and W");
  if ((a->epi_flags & SRES_EPI_POOL) && !a->pool_part) return set_error(SRES_ERR_INVALID_ARG, "conv: pool_part missing");

  ConvKParams p{};
  p.B = a->B; p.H = a->H; p.W = a->W; p.P = a->W + 1; p.R = a->H + 1;
  const long long npos = (long long)p.B * p.R * p.P;
  if (npos > 0x7fffff00LL) return set_error(SRES_ERR_UNSUPPORTED, "conv: batch too large for 32-bit rows");
  p.npos = (int)npos;
  p.n_tiles = (p.npos + 127) / 128;
  p.n_out = a->n_out; p.c_real = a->c_real;
  p.flags = a->epi_flags; p.map_mode = a->map_mode; p.sub_i = a->sub_i; p.sub_j = a->sub_j; p.sf = sf;
  p.debug_flags = a->debug_flags;
  p.bias = a->bias; p.resid = a->resid_f32; p.resid2 = a->resid2_f32; p.mask = (const uint16_t*)a->mask_bf16;
  p.out_f32 = a->out_f32; p.out_bf16 = (uint16_t*)a->out_bf16; p.pool_part = a->pool_part; p.out_nchw = a->out_nchw;
  // TMA-staged epilogue whenever rows map to themselves; scattered (PixelShuffle) and planar stores go direct
  p.tma_epi = (a->map_mode == SRES_MAP_IDENT && a->n_out == 64 && !a->out_nchw && !(a->debug_flags & 2)) ? 1 : 0;
  if (p.tma_epi) {
    p.use_o16 = a->out_bf16 != nullptr; p.use_msk = a->mask_bf16 != nullptr;
    p.use_r32 = a->resid_f32 != nullptr; p.use_o32 = a->out_f32 != nullptr;
  }
  const int wbytes = 9 * a->n_out * 128;
  const int smem_max = 232448;  // 227 KB
  const int tail_bytes = 1024;
  const int per_tile_slab = (p.use_o16 ? 16384 : 0) + (p.use_msk ? 16384 : 0) + ((p.use_r32 | p.use_o32) ? 32768 : 0);
  int chosen = 0;
  for (int mt = 2; mt >= 1 && !chosen; --mt) {
    const int rows = mt * 128 + 2 * (p.P + 1);
    const int stage_rows = (rows + kBoxRows - 1) / kBoxRows * kBoxRows;
    const int avail = smem_max - 1024 - wbytes - tail_bytes - mt * per_tile_slab;
    int ns = avail / (stage_rows * 128);
    if (ns > kMaxStages) ns = kMaxStages;
    if (ns >= 2 || (mt == 1 && ns >= 1)) {
      p.mt = mt; p.nstage = ns; p.stage_rows = stage_rows;
      chosen = 1;
    }
  }
  if (!chosen) return set_error(SRES_ERR_UNSUPPORTED, "conv: image too wide for the flat halo window");
  p.n_stages = (p.n_tiles + p.mt - 1) / p.mt;
  int off = wbytes + p.nstage * p.stage_rows * 128;
  p.off_s16 = off; off += p.use_o16 ? p.mt * 16384 : 0;
  p.off_msk = off; off += p.use_msk ? p.mt * 16384 : 0;
  p.off_s32 = off; off += (p.use_r32 | p.use_o32) ? p.mt * 32768 : 0;
  p.off_tail = off; off += tail_bytes;
  const size_t smem = (size_t)off + 1024;  // + alignment slack

  CUtensorMap tmA, tmW, tmO16, tmMsk, tmR32, tmO32;
  int rc = make_tmap_rows64(&tmA, a->in_bf16, (uint64_t)p.npos, kBoxRows);
  if (rc) return rc;
  rc = make_tmap_rows64(&tmW, a->wpack_bf16, (uint64_t)(9 * a->n_out), a->n_out);
  if (rc) return rc;
  tmO16 = tmA; tmMsk = tmA; tmR32 = tmA; tmO32 = tmA;
  if (p.tma_epi) {
    if (p.use_o16 && (rc = make_tmap_rows64(&tmO16, a->out_bf16, (uint64_t)p.npos, 32))) return rc;
    if (p.use_msk && (rc = make_tmap_rows64(&tmMsk, a->mask_bf16, (uint64_t)p.npos, 32))) return rc;
    if (p.use_r32 && (rc = make_tmap_rows64_f32(&tmR32, a->resid_f32, (uint64_t)p.npos, 32))) return rc;
    if (p.use_o32 && (rc = make_tmap_rows64_f32(&tmO32, a->out_f32, (uint64_t)p.npos, 32))) return rc;
  }

  int sms = device_sm_count();
  if (sms <= 0) return set_error(SRES_ERR_NO_DEVICE, "conv: no CUDA device");
  const int grid = p.n_stages < sms ? p.n_stages : sms;
  cudaError_t e;
  if (a->n_out == 64) {
    e = cudaFuncSetAttribute(conv3x3_igemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) return set_cuda_error(e, "conv: smem attribute");
    conv3x3_igemm_kernel<64><<<grid, 256, smem, stream>>>(tmA, tmW, tmO16, tmMsk, tmR32, tmO32, p);
  } else {
    e = cudaFuncSetAttribute(conv3x3_igemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) return set_cuda_error(e, "conv: smem attribute");
    conv3x3_igemm_kernel<16><<<grid, 256, smem, stream>>>(tmA, tmW, tmO16, tmMsk, tmR32, tmO32, p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "conv: launch");
  return SRES_OK;
}

}  // namespace sres

extern "C" int sres_conv_mtiles(int B, int H, int W) {
  long long npos = (long long)B * (H + 1) * (W + 1);
  return (int)((npos + 127) / 128);
}

extern "C" int sres_conv3x3_igemm(const sres_conv_args* args, void* stream) {
  return sres::launch_conv(args, (cudaStream_t)stream);
}
