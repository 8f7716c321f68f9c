/*
 * libvideosd — C ABI of the B200-native per-frame LCM img2img hot path of venetanji/videosd.
 *
 * This is the boundary a binding sits on (ctypes from Python). It replaces, for the hot
 * path only, what the reference reaches through
 *     diffusert/videopipeline.py:75-128   VideoSDPipeline.infer            (frame-transform interface)
 *     diffusert/lcm/lcm_controlnet.py:380-618  LatentConsistencyModelPipeline_controlnet.__call__
 *     diffusert/lcm/lcm_controlnet.py:905-1071 LCMScheduler_X.set_timesteps / step / add_noise
 *     diffusert/server.py:108,117         frame.to_image() / VideoFrame.from_image (libswscale colour conversion)
 * and, below those, the third-party diffusers modules UNet2DConditionModel / AutoencoderTiny / VaeImageProcessor.
 *
 * Conventions: every function returns 0 on success and a negative code on failure; vsd_last_error() returns
 * the message for the calling thread. Plain pointers and sizes only. "dev" pointers are CUDA device pointers
 * owned by the caller; "host" pointers are host memory. No function falls back to the CPU.
 */
#ifndef VIDEOSD_H
#define VIDEOSD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSD_ABI_VERSION 1

const char* vsd_last_error(void);
int vsd_abi_version(void);
/* Non-zero (and an error message) if a tensor-core kernel's internal barrier wait timed out since the last call. */
int vsd_check_pipeline_fault(void);

/* ------------------------------------------------------------------ operator entry points (parity tests) */

/* Implicit-GEMM convolution / linear on tcgen05 tensor cores.
 *   x      dev bf16 NHWC view: nb images of h x w pixels, c channels (multiple of 64), ldx elements per pixel
 *   taps   9: 3x3 stride 1 pad 1 (replaces torch conv2d inside diffusers ResnetBlock2D / TAESD Block)
 *          1: 1x1 conv or nn.Linear over nb*h*w rows
 *   wt     dev bf16 [n][taps*c], tap-major then channel (OHWI)
 *   out    dev bf16 (out_f32=0) or fp32 (out_f32=1) [rows][ldo]
 *   bias   dev fp32 [n] or NULL; rowvec dev fp32 [nb][n] or NULL (time-embedding broadcast);
 *   residual dev bf16 [rows][ldr] or NULL
 *   act    0 none, 1 GEGLU (weight rows interleaved per block_n tile: [values | gates]; out has n/2 columns)
 *   block_n / splits: 0 = choose automatically
 */
int vsd_op_conv_gemm(const void* x, int nb, int h, int w, int c, int ldx, int taps, const void* wt, int n, void* out,
                     int ldo, int out_f32, const float* bias, const float* rowvec, const void* residual, int ldr,
                     int act, int block_n, int splits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOSD_H */
