/*
 * sres_b200 -- C ABI of the B200-native RCAN hot path for nasa-nccs-hpda/super-resolution-climate.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its
 * hot path is stock torch.nn modules called from
 *     sres/model/rcan/network.py:22-27      RCAN.forward
 *     sres/controller/dual_trainer.py:557-571  ModelTrainer.apply_network  (bicubic down + model)
 *     sres/controller/dual_trainer.py:221-234  ModelTrainer.loss  -> sres/controller/stats.py:5-8
 *     sres/controller/dual_trainer.py:322-323  mloss.backward(); optimizer.step()
 *     sres/base/source/swot/raw.py:216-233   SWOTRawDataLoader.get_tiles
 *     sres/controller/dual_trainer.py:449-480  ModelTrainer.assemble_images
 * Each entry point below names the reference lines it replaces.  Signatures are plain C: raw
 * device pointers, explicit sizes, a cudaStream_t passed as void*.  Every function returns an
 * int status (SRES_OK == 0), never throws, never allocates device memory (the caller owns all
 * buffers, including workspaces) and keeps no mutable global state, so calls on different
 * streams are independent.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Activation layout ("padded tile layout", PTL).  A batch of B feature maps of H x W pixels with
 * 64 channels is stored as   rows = B*(H+1)*(W+1) positions,  64 channels per position,
 * position (b,y,x) at row  b*(H+1)*(W+1) + y*(W+1) + x.  Column x == W and row y == H of every
 * map are zero padding shared between neighbours, so the 3x3 tap (dy,dx) of position q is simply
 * row q + dy*(W+1) + dx -- a constant row shift.  Kernels keep the padding rows at exactly 0.
 * bf16 PTL rows are 128 bytes (one TMA/UMMA 128B-swizzle row); fp32 PTL rows are 256 bytes.
 */
#ifndef SRES_B200_H_
#define SRES_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* exported symbol (the library is built with -fvisibility=hidden) */
#define SRES_API __attribute__((visibility("default")))

#define SRES_OK 0
#define SRES_ERR_INVALID_ARG 1
#define SRES_ERR_UNSUPPORTED 2  /* geometry / channel count outside what the kernels cover */
#define SRES_ERR_CUDA 3         /* a CUDA runtime / driver call failed; see sres_last_error */
#define SRES_ERR_NO_DEVICE 4

#define SRES_NF 64 /* feature channels of the RCAN trunk (nfeatures, config/model/rcan-10-20-64.yaml:4) */

/* ------------------------------------------------------------------------------------------ */
/* library / device                                                                           */
/* ------------------------------------------------------------------------------------------ */
/* ABI version of this header (bumped on incompatible change). */
SRES_API int sres_abi_version(void);
/* Human readable text of the last failure on the calling thread. */
SRES_API const char* sres_last_error(void);
/* Number of SMs of the current device (grid sizing); <0 on error. */
SRES_API int sres_device_sm_count(void);
/* Rows of the padded tile layout for a (B,H,W) batch. */
SRES_API int64_t sres_ptl_rows(int B, int H, int W);

/* ------------------------------------------------------------------------------------------ */
/* 3x3 convolution, 64 input features, tensor cores (tcgen05 implicit GEMM)                   */
/* replaces nn.Conv2d(64, n, 3, padding=1) forward and its input-gradient:                    */
/*   sres/model/common/cnn.py:8-9 (default_conv) as used at sres/model/rcan/network.py:14-16, */
/*   55, 71 and sres/model/rcan/blocks.py:62-64 (Upsampler convs).                            */
/* ------------------------------------------------------------------------------------------ */
#define SRES_EPI_RELU 1u       /* v = max(v, 0)                       (network.py:57, nn.ReLU) */
#define SRES_EPI_POOL 2u       /* emit per-tile channel sums for the CA average pool (network.py:35,45) */

#define SRES_MAP_IDENT 0       /* output position == input position                              */
#define SRES_MAP_SHUFFLE 1     /* PixelShuffle(2) store: (b,y,x) -> (b,2y+sub_i,2x+sub_j) (blocks.py:65) */
#define SRES_MAP_UNSHUFFLE 2   /* inverse: (b,y,x) -> sub-grid (y&1,x&1), position (b,y/2,x/2)     */

typedef struct sres_conv_args {
  const void* in_bf16;     /* [rows_in][64] bf16 PTL, geometry (B,H,W)                          */
  const void* wpack_bf16;  /* [9][n_out][64] bf16, tap-major, from sres_pack_conv_weights        */
  const float* bias;       /* [n_out] fp32 or NULL                                              */
  const float* resid_f32;  /* fp32 PTL indexed like the OUTPUT, added before ReLU; may alias out_f32; or NULL */
  const void* mask_bf16;   /* bf16 PTL indexed like the INPUT: v = mask > 0 ? v : 0 (ReLU backward) or NULL */
  float* out_f32;          /* fp32 PTL output or NULL                                           */
  void* out_bf16;          /* bf16 PTL output or NULL                                           */
  float* pool_part;        /* [n_mtiles][2][4][64] fp32 partial channel sums (SRES_EPI_POOL) or NULL */
  float* out_nchw;         /* [B][c_real][H][W] fp32 planar output (n_out == 16 path) or NULL   */
  int32_t B, H, W;         /* geometry of the input                                             */
  int32_t n_out;           /* 64, or 16 (narrow tail conv, c_real <= 16 real channels)          */
  int32_t c_real;          /* real output channels when out_nchw is used                        */
  uint32_t epi_flags;      /* SRES_EPI_*                                                        */
  int32_t map_mode;        /* SRES_MAP_*                                                        */
  int32_t sub_i, sub_j;    /* sub-pixel of SRES_MAP_SHUFFLE                                     */
  int32_t debug_flags;     /* bit0: put (addr>>7)&7 in the UMMA descriptor base_offset field    */
} sres_conv_args;

/* Number of 128-position M tiles of a (B,H,W) batch (size of pool_part's leading dim). */
SRES_API int sres_conv_mtiles(int B, int H, int W);
SRES_API int sres_conv3x3_igemm(const sres_conv_args* args, void* stream);

/* Repack fp32 OIHW 3x3 weights (the checkpoint layout, state_dict of nn.Conv2d) to the bf16
 * tap-major operand the tensor-core kernels read.
 *   mode 0 (forward):   out[t][n][k] = w[oc(n)][k][ky][kx],              t = ky*3+kx
 *   mode 1 (dgrad):     out[t][n][k] = w[oc(k)][n][2-ky][2-kx]           (transposed, flipped)
 *   oc(c) = c*oc_stride + oc_offset   (PixelShuffle sub-convs: oc_stride 4, oc_offset 2i+j)
 *   n_rows = rows of the packed operand (64, or 16 with zero rows beyond the real ones);
 *   cin = input channels of w (64);  cout_total = output channels of w.                      */
SRES_API int sres_pack_conv_weights(const float* w_oihw, void* out_bf16, int mode, int n_rows, int cin,
                           int cout_total, int oc_stride, int oc_offset, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* 3x3 convolution weight + bias gradient (split-K tcgen05 GEMM + deterministic reduce)       */
/* replaces the grad_weight / grad_bias outputs of aten::convolution_backward for the same    */
/* nn.Conv2d modules, reached from mloss.backward() (sres/controller/dual_trainer.py:322).     */
/*   x_bf16  [rows][64] bf16 PTL, the conv input;  dy_bf16 [rows][64] bf16 PTL, grad of output */
/*   dw_oihw fp32 (cout_total, 64, 3, 3); channel n of dy maps to oc = n*oc_stride+oc_offset   */
/*   dbias   fp32 (cout_total) or NULL;  accumulate != 0 adds into dw/dbias instead of storing */
/*   workspace: at least sres_conv_wgrad_workspace_bytes() bytes of device memory              */
/* ------------------------------------------------------------------------------------------ */
SRES_API size_t sres_conv_wgrad_workspace_bytes(void);
SRES_API int sres_conv3x3_wgrad(const void* x_bf16, const void* dy_bf16, int B, int H, int W, float* dw_oihw,
                                float* dbias, int cout_total, int oc_stride, int oc_offset, int accumulate,
                                void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRES_B200_H_ */
