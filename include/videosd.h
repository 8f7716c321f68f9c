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

/* Bring-up variant: CTA (0,0,0) writes clock64() stamps {start, setup done, first operands landed, last MMA issued,
 * accumulator ready, epilogue stores issued, teardown} to dbg (device int64[8]). */
int vsd_op_conv_gemm_timed(const void* x, int nb, int h, int w, int c, int ldx, int taps, const void* wt, int n, void* out,
                           int ldo, const float* bias, int block_n, int splits, int occ, int kb_per_stage, long long* dbg,
                           void* stream);

/* Fused flash-style attention on tcgen05 (replaces F.scaled_dot_product_attention in diffusers AttnProcessor2_0,
 * reached from lcm_controlnet.py:568-577).
 *   q, k : dev bf16 [batch*rows_per_img][heads*dk_pad], dk_pad = round_up(d, 64), per-head zero padding
 *   vt   : dev bf16 [vt_rows = heads*d][batch*vt_cols_per_img]  (V transposed; keys along the row)
 *   out  : dev bf16 [batch*nq][ldo], head h in columns [h*d, h*d+d)
 * softmax scale is 1/sqrt(d); keys >= nk inside an image's slot are masked. */
int vsd_op_attention(const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* out, int ldo,
                     int batch, int heads, int d, int nq, int nk, int q_rows_per_img, int k_rows_per_img,
                     int vt_cols_per_img, int vt_rows, void* stream);

/* The (block_n, splits, occ, kb_per_stage, mode) configurations the engine's autotuner may pick for a GEMM shape, and a launch
 * with one of them: the operator-level sweep (tests/test_gpu_tuner_sweep.py) checks every one against fp32, so what the tuner can
 * select is what was tested. cand: host int[max_cand][5]; returns the count. stride2 / pad: 3x3 stride-2 taps through TMA element
 * strides. act: 0, 1 (GEGLU), 2 (quick-GELU), | 16 ReLU. */
int vsd_op_gemm_candidates(const void* x, int nb, int h, int w, int c, int ldx, int taps, int stride2, int pad, const void* wt, int n,
                           void* out, int ldo, int out_f32, const float* bias, const float* rowvec, const void* residual, int ldr,
                           int act, int* cand, int max_cand);
int vsd_op_conv_gemm_cfg(const void* x, int nb, int h, int w, int c, int ldx, int taps, int stride2, int pad, const void* wt, int n,
                         void* out, int ldo, int out_f32, const float* bias, const float* rowvec, const void* residual, int ldr, int act,
                         int block_n, int splits, int occ, int kb_per_stage, int mode, void* stream);

/* Linear(LayerNorm(x)) with the LayerNorm folded into the tcgen05 GEMM (BasicTransformerBlock.norm1/2/3 followed by
 * to_q|to_k / to_v / ff.net.0.proj): w_raw bf16 [n][c] is the original weight; swapped = 1 computes out^T (tokens along the
 * columns, the V^T projection); act = 1: GEGLU with value / gate rows interleaved per 128-row tile. stats [rows][nst][2]:
 * partial row sums of x and x^2 left by the GEMM that produced x (vsd_op_linear_stats), NULL = computed here. */
int vsd_op_linear_ln(const void* x, int rows, int c, int ldx, const void* w_raw, int n, const float* gamma, const float* beta,
                     const float* bias, float eps, void* out, int ldo, int swapped, int act, int block_n, const float* stats,
                     int nst, void* stream);
/* Linear (+ bias, + residual) that also leaves those row statistics of its output: stats_out [rows][n_tiles][2]; returns
 * n_tiles. mode 0 plain, 1 CTA pairs; splits > 1 = split-K reduced inside the cluster. */
int vsd_op_linear_stats(const void* x, int rows, int c, int ldx, const void* w, int n, const float* bias, const void* residual,
                        int ldr, void* out, int ldo, float* stats_out, int block_n, int splits, int mode, void* stream);

/* GroupNorm (+ optional SiLU) over NHWC bf16; statistics in fp32 (torch.nn.GroupNorm in ResnetBlock2D /
 * Transformer2DModel). x, y: dev bf16 [nb*hw][ld]. */
int vsd_op_groupnorm(const void* x, int ldx, void* y, int ldy, const float* gamma, const float* beta, int nb, int hw,
                     int c, int groups, float eps, int silu, void* stream);
/* LayerNorm over the last dimension (BasicTransformerBlock.norm1/2/3). */
int vsd_op_layernorm(const void* x, int ldx, void* y, int ldy, const float* gamma, const float* beta, int rows, int c,
                     float eps, void* stream);
/* F.interpolate(mode="nearest") on NHWC bf16 (Upsample2D / TAESD nn.Upsample). */
int vsd_op_upsample_nearest(const void* x, int ldx, void* y, int ldy, int nb, int hi, int wi, int ho, int wo, int c,
                            void* stream);
/* 3x3 stride-2 pad-1 patch gather to [nb*ho*wo][9*c] (Downsample2D and the TAESD encoder strided convs). */
int vsd_op_im2col_s2(const void* x, int ldx, void* y, int nb, int hi, int wi, int c, int ho, int wo, void* stream);
/* 3x3 pad-1 convolution with <=4 input channels. x_kind 0: fp32 NHWC; 1: u8 RGB with the VaeImageProcessor
 * normalisation and TAESD (x+1)/2 prologue; 2: fp32 NHWC with the TAESD decoder tanh(z/3)*3 prologue.
 * wt: dev fp32 [cout][3][3][cin]. y: dev bf16. */
int vsd_op_conv3x3_small_cin(const void* x, int x_kind, int nb, int h, int w, int cin, const float* wt,
                             const float* bias, void* y, int ldy, int cout, int relu, void* stream);
/* Sobel edge map -> ControlNet control image (diffusert/lcm/canny_gpu.py:27-44 + lcm_controlnet.py:218-248).
 * rgb: dev u8 [nb][h][w][3]; mag: dev fp32 scratch [nb][h][w]; maxbits: dev u32 [nb]; control: dev fp32 [nb][h][w][3]. */
int vsd_op_sobel_control(const uint8_t* rgb, float* mag, unsigned int* maxbits, float* control, int nb, int h, int w,
                         float low, float high, void* stream);
/* Center crop + Lanczos resize, bit-identical to PIL Image.crop + resize(LANCZOS) (diffusert/videopipeline.py:92-107).
 * src: dev u8 [nb][in_h][in_w][3]; tmp: dev u8 scratch [nb][ch][w][3]; out: dev u8 [nb][h][w][3]; *_bounds: dev i32 [n][2]
 * (first tap, tap count); *_coeffs: dev i32 [n][ksize], 22-bit fixed point (videosd_b200/resample.py). */
int vsd_op_crop_resize(const uint8_t* src, int in_w, int in_h, int x0, int y0, int cw, int ch, uint8_t* tmp, uint8_t* out, int w,
                       int h, const int* h_bounds, const int* h_coeffs, int h_ksize, const int* v_bounds, const int* v_coeffs,
                       int v_ksize, int nb, void* stream);
/* Direct 3x3 conv (pad 1, stride 1|2) + optional SiLU for narrow layers (ControlNetConditioningEmbedding). bf16 NHWC. */
int vsd_op_conv3x3_direct(const void* x, int ldx, int nb, int hi, int wi, int cin, const void* wt, const float* bias, void* y,
                          int ldy, int cout, int stride, int silu, void* stream);
/* LCMScheduler_X.add_noise (lcm_controlnet.py:1046-1071) and .step (:948-1043) on fp32 latents. */
int vsd_op_add_noise(const float* x0, const float* noise, float* out, float sqrt_alpha, float sqrt_one_minus_alpha,
                     long n, void* stream);
int vsd_op_lcm_step(const float* eps, const float* x, const float* z, float* x_prev, float* denoised, float sqrt_a,
                    float sqrt_1ma, float c_skip, float c_out, float sqrt_ap, float sqrt_1map, int has_noise, long n,
                    void* stream);
/* YUV420P (BT.601 limited) -> RGB24, replacing frame.to_image() (server.py:108). Planar inputs, nb frames. */
int vsd_op_yuv420_to_rgb(const uint8_t* y, const uint8_t* u, const uint8_t* v, uint8_t* rgb, int nb, int h, int w,
                         void* stream);
/* Decoder tail (x*2-1 if taesd_denorm) + VaeImageProcessor.postprocess (lcm_controlnet.py:609-611) + RGB24 ->
 * YUV420P (VideoFrame.from_image + encoder reformat, server.py:117). Any of rgb / y,u,v may be NULL. */
int vsd_op_pack_rgb_yuv420(const float* img, int ldi, uint8_t* rgb, uint8_t* y, uint8_t* u, uint8_t* v, int nb, int h,
                           int w, int taesd_denorm, void* stream);

/* ------------------------------------------------------------------ engine (per-frame path) */

/* One context per GPU (the reference creates one VideoSDPipeline Ray actor per GPU, videopipeline.py:11-32).
 * Calls on a context must be serialised by the caller (server.py keeps <= 1 in-flight infer per GPU, :132-137). */
typedef struct vsd_ctx vsd_ctx;

vsd_ctx* vsd_create(int device);             /* NULL on failure (see vsd_last_error) */
void vsd_destroy(vsd_ctx* ctx);
/* Additional lane on the parent's GPU: own stream / buffers / CUDA graph, shares the parent's weights read-only.
 * Lanes keep several frames in flight on one GPU (co-located sessions, or one stream pipelined); the reference
 * allows at most one in-flight frame per GPU (server.py:132-137). The parent must outlive its lanes. */
vsd_ctx* vsd_create_lane(vsd_ctx* parent);

/* Weights by diffusers state-dict name, prefixed "unet." or "vae.", fp32 host data in
 * PyTorch layout (conv: [Cout][Cin][kh][kw], linear: [out][in]). Replaces from_pretrained in
 * videopipeline.py:49-72. Converted to bf16 and repacked for the tensor-core kernels on load. */
int vsd_load_weight(vsd_ctx* ctx, const char* name, const float* host_f32, const int64_t* shape, int ndim);
int vsd_num_weights(vsd_ctx* ctx);
/* GEMM autotuner (on by default): at plan-build time each distinct GEMM shape is timed over (block_n, split-K,
 * CTAs/SM, k-blocks per stage, halo tiles, CTA pairs) candidates with the L2 flushed. frames_in_flight: 0 = off (shape
 * heuristics), 1 = lowest latency of a single frame, n > 1 = n frames in flight on this GPU (lanes): a candidate's cost is
 * its duration x max(share of the SMs it occupies, 1/n). vsd_tuning_report dumps the choices as text. */
int vsd_set_autotune(vsd_ctx* ctx, int frames_in_flight);
int vsd_tuning_report(vsd_ctx* ctx, char* buf, long cap);
int vsd_tuning_load(vsd_ctx* ctx, const char* text);   /* returns the number of entries loaded */
long vsd_tuning_misses(vsd_ctx* ctx);                   /* shapes timed on the device because no loaded table entry covered them */

/* Working size: `batch` frames of height x width (multiples of 8; infer(height=, width=) at videopipeline.py:75-88).
 * Must be called after the weights are loaded; invalidates schedule, contexts and noise. */
int vsd_configure(vsd_ctx* ctx, int batch, int height, int width);

/* LCM schedule (LCMScheduler_X.set_timesteps, lcm_controlnet.py:905-938, computed by the host):
 *   timesteps[steps]; scalars[steps][6] = sqrt(abar_t), sqrt(1-abar_t), c_skip, c_out, sqrt(abar_prev),
 *   sqrt(1-abar_prev) (:995-1036); add_noise coefficients at timesteps[0] (:1046-1071); the 256-d guidance
 *   embedding (:347-368); has_step_noise = len(timesteps) > 1 (:1032). Builds the launch plan. */
int vsd_set_schedule(vsd_ctx* ctx, int steps, const int* timesteps, const float* scalars, float add_noise_a,
                     float add_noise_b, const float* w_embedding256, int has_step_noise);

/* ControlNet branch (the reference runs it before every UNet pass, lcm_controlnet.py:558-566,
 * on the Sobel edge map of the frame, videopipeline.py:109). Needs "controlnet.*" weights (diffusers ControlNetModel names).
 * scales13 = logspace(-1,0,13) * controlnet_scale (guess mode). Toggling `enabled` requires vsd_set_schedule again. */
int vsd_set_controlnet(vsd_ctx* ctx, int enabled, const float* scales13);

/* Prompt context (CLIP last_hidden_state, 77 x 768 fp32, host) for batch slot `slot`; projects it through every
 * cross-attention to_k / to_v once (the reference recomputes them every step of every frame, lcm_controlnet.py:449). */
int vsd_set_context(vsd_ctx* ctx, int slot, const float* context_77x768);

/* Noise tensors, NHWC fp32 host: init [batch][h/8][w/8][4] (:331) and per-step [steps][batch][h/8][w/8][4] (:1033). */
int vsd_set_noise(vsd_ctx* ctx, const float* init_noise_nhwc, const float* step_noise_nhwc);

/* One frame batch, host planes in / host planes out, synchronous. YUV420P: y [batch][h][w], u, v [batch][h/2][w/2].
 * Replaces frame.to_image() -> infer -> VideoFrame.from_image (server.py:104-117). */
int vsd_infer_yuv420(vsd_ctx* ctx, const uint8_t* y, const uint8_t* u, const uint8_t* v, uint8_t* out_y,
                     uint8_t* out_u, uint8_t* out_v);
/* Packed RGB24 in / out ([batch][h][w][3]); the PIL-compatible path of VideoSDPipeline.infer. */
int vsd_infer_rgb(vsd_ctx* ctx, const uint8_t* rgb_in, uint8_t* rgb_out);

/* AutoencoderKL option (the pipeline's declared VAE, lcm_controlnet.py:69; encode = `latent_dist
 * .sample() * scaling_factor` :298-313, decode = `vae.decode(denoised / scaling_factor)` :594-596). kind 0 = AutoencoderTiny
 * (default, what the reference loads), 1 = AutoencoderKL (diffusers keys under "vae_kl."); switching invalidates the schedule.
 * vsd_set_vae_noise: the sample() noise, fp32 [batch][h/8][w/8][4] host (zeros => the distribution's mean). */
int vsd_set_vae(vsd_ctx* ctx, int kind);
int vsd_set_vae_noise(vsd_ctx* ctx, const float* noise_nhwc);

/* CLIP text encoder (diffusert/lcm/lcm_controlnet.py:175-179 `self.text_encoder(ids)[0]`): the SD1.5
 * text tower (transformers CLIPTextModel keys under the "text_encoder." prefix, loaded with vsd_load_weight). token_ids: 77 ids
 * (tokenizer output padded to max_length, host); context: fp32 [77][768] last_hidden_state (host), ready for vsd_set_context.
 * Runs on the context's stream, once per prompt change. */
int vsd_encode_prompt(vsd_ctx* ctx, const int* token_ids_77, float* context_77x768);

/* GPU center-crop + Lanczos resize (videopipeline.py:92-107 does PIL crop + resize(LANCZOS) on
 * the CPU). The host passes Pillow-compatible windows / 22-bit coefficients (videosd_b200/resample.py); results are
 * bit-identical to Pillow. Geometry: input frames in_w x in_h, crop (x0, y0, cw, ch) -> working size. */
int vsd_set_resize(vsd_ctx* ctx, int in_w, int in_h, int x0, int y0, int cw, int ch, const int* h_bounds, const int* h_coeffs,
                   int h_ksize, const int* v_bounds, const int* v_coeffs, int v_ksize);
int vsd_infer_rgb_resized(vsd_ctx* ctx, const uint8_t* rgb_src, uint8_t* rgb_out);
/* YUV420P planes at the source geometry in (even in_w, in_h), working-size planes out: colour conversion at the source size
 * (frame.to_image(), server.py:104-108), then crop + Lanczos, the frame, and the pack to YUV420P. */
int vsd_infer_yuv420_resized(vsd_ctx* ctx, const uint8_t* y, const uint8_t* u, const uint8_t* v, uint8_t* out_y, uint8_t* out_u,
                             uint8_t* out_v);
int vsd_debug_read_rgb_in(vsd_ctx* ctx, uint8_t* host);

/* Split form of vsd_infer_yuv420 (asynchronous on the context's stream; vsd_sync waits). */
int vsd_upload_yuv420(vsd_ctx* ctx, const uint8_t* y, const uint8_t* u, const uint8_t* v);
int vsd_run_yuv420(vsd_ctx* ctx);
int vsd_download_yuv420(vsd_ctx* ctx, uint8_t* y, uint8_t* u, uint8_t* v);
int vsd_sync(vsd_ctx* ctx);
void* vsd_stream(vsd_ctx* ctx);               /* cudaStream_t of the context */
long vsd_launches_per_frame(vsd_ctx* ctx, int yuv);
long vsd_arena_peak_bytes(vsd_ctx* ctx);

/* Debug taps for the parity tests: fp32 NHWC device buffers copied to the host.
 * what: "init_latents", "noisy", "image" ([batch][h][w][4], 3 used), or with index = step: "eps", "latents", "denoised". */
int vsd_debug_read(vsd_ctx* ctx, const char* what, int index, float* host, long nfloats);
int vsd_debug_unet(vsd_ctx* ctx, const float* latents_nhwc, int step, float* eps_nhwc);
int vsd_debug_run_eager(vsd_ctx* ctx, int yuv);
/* Times each tagged section of the launch plan as its own CUDA graph ("tag kernels microseconds" lines). */
int vsd_debug_profile_sections(vsd_ctx* ctx, int depth, int reps, char* buf, long cap);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOSD_H */
