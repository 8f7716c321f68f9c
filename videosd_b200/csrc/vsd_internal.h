// Internal host-side declarations shared by the translation units of libvideosd.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace vsd {

typedef __nv_bfloat16 bf16;

// ---- error plumbing: every host entry returns 0 or a negative code and records a message.
void set_error(const std::string& msg);
const char* get_error();
#define VSD_CHECK_CUDA(expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            ::vsd::set_error(std::string(#expr) + " -> " + cudaGetErrorString(_e) + " @" + __FILE__ + \
                             ":" + std::to_string(__LINE__));                                        \
            return -2;                                                                               \
        }                                                                                            \
    } while (0)
#define VSD_REQUIRE(cond, msg)                                                             \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            ::vsd::set_error(std::string("requirement failed: ") + #cond + " : " + (msg)); \
            return -1;                                                                     \
        }                                                                                  \
    } while (0)

// ---- kernel launch with programmatic dependent launch (PDL) enabled; VSD_PDL=0 in the environment disables it
int pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// same, as thread-block clusters of (cluster_x, 1, cluster_z) CTAs
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, int cluster_x, int cluster_z,
                                    cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = (unsigned)cluster_z;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- tensor maps (driver entry point resolved at run time; the library does not link libcuda)
int make_tmap_act(CUtensorMap* m, const void* base, int C, int W, int H, int N, int ld, int boxW, int boxH, int boxN,
                  int elem_stride = 1);
int make_tmap_2d(CUtensorMap* m, const void* base, int K, int rows, int ld, int box_rows);

// ---- tcgen05 implicit-GEMM convolution / linear kernel -------------------------------------------
// out[pixel, n] = epilogue( sum_{tap, c} A[pixel + tap offset, c] * Wt[n, tap*cin + c] )
enum { ACT_NONE = 0, ACT_GEGLU = 1, ACT_QUICK_GELU = 2,   // x * sigmoid(1.702 x) after the bias (CLIP MLP)
       ACT_RELU_FLAG = 16,       // ReLU (after the residual add) may be OR-ed in
       ACT_A_STATIC_FLAG = 32,   // the row operand is the constant one (swapped-operand V^T projection)
       ACT_NO_STATIC_FLAG = 64,  // neither operand is constant
       ACT_RES_F32_FLAG = 128,   // the residual operand is fp32 (with an fp32 output: an fp32 residual stream updated in place)
       // LayerNorm folded into this GEMM (see LnFuse): the row statistics come from the GEMM that produced the normalised
       // operand (ACT_ROWSTATS_FLAG there) and are applied to the accumulator in the epilogue
       ACT_LN_A_FLAG = 256,      // the normalised operand is A (rows of the output)
       ACT_LN_B_FLAG = 512,      // the normalised operand is B (columns of the output: swapped-operand V^T projection)
       ACT_ROWSTATS_FLAG = 1024 };  // the epilogue also leaves per-row partial sums of x, x^2 of the tensor it stores

struct GemmParams {
    // A operand traversal (NHWC activation, stride-1 taps; linear layers use H=NB=1, W=rows)
    int taps, cin;        // 1 or 9 ; channels per tap (multiple of 64)
    int H, W, NB;         // output == input spatial extent
    int BW, BH, BN;       // tile rectangle, BW*BH*BN == 128
    int tiles_w, tiles_h, tiles_n;
    // B operand
    int N;                // rows of the weight matrix (GEMM N)
    int block_n;          // multiple of 32, <= 256
    int tmem_cols;        // power of two >= block_n
    int stages, kb_per_stage;
    int cstride, cshift;  // 3x3 taps read input pixel (cstride * w + dx + cshift): stride-2 convolutions, asymmetric padding
    int pair;             // CTA pairs (cta_group::2): see conv_gemm_kernel<.., kPair>
    int persist;          // persistent weight-stationary 3x3 convolution (conv_persist_kernel)
    int halo;             // 3x3 conv halo mode (8x16 tiles, column-shifted 8x18 activation tiles shared by 3 row taps)
    // K split
    int kb_total, kb_per_split, splits;
    // epilogue
    void* out;            // bf16 or fp32 [rows, ldo]
    int ldo, out_f32;
    const float* bias;    // [N] or null
    const float* rowvec;  // [NB, N] or null (per-image broadcast, e.g. time embedding projection)
    const bf16* residual; // [rows, ldr] or null (fp32 rows when res_f32)
    int ldr, res_f32;
    const float* out_scale;  // optional device scalar: out = (acc + bias) * scale + residual (ControlNet conditioning scale)
    int act, relu;
    float* partial;       // [splits, rows, N] fp32 when splits > 1
    int a_static, b_static;  // operand is constant (weights): its first stages may be fetched before pdl_wait()
    // Epilogue through shared memory + bulk tensor stores: each epilogue warp stages its 32 rows x 32 columns in the
    // swizzled layout of mapC's box (sbw x sbh x sbn pixels) and one lane issues the store; TMA clips rows / columns
    // outside the tensor. tma_out: 0 = direct st.global, 1 = bf16 output (4-D map), 2 = fp32 split-K partials (5-D map).
    // tma_res: the residual tile is fetched by TMA (mapR) into the staging region while the main loop runs.
    int tma_out, tma_res;
    int lbw, lbh;         // log2 of BW, BH
    int cluster_k;        // split-K inside a thread-block cluster (1,1,splits): partial tiles are reduced through distributed
                          // shared memory by the cluster itself (fixed order => deterministic); no workspace, no second kernel
    int sbw, sbh, sbn;
    unsigned int stage_off, bar_off;   // byte offsets of the staging region / the mbarrier block in dynamic smem
    long long* dbg;       // optional: CTA (0,0,0) writes clock64() phase stamps here (bring-up only)
    unsigned int* trace;  // optional (VSD_TRACE=1): four device counters -- CTAs entered / CTAs that own their TMEM columns /
                          // CTAs finished / pair CTAs past the first cluster barrier -- read by the host watchdog when the device stops making progress
    // LayerNorm folded into the GEMM: W' = W * gamma (done once at load), out = rstd * (acc - mean * wsum) + (W beta + b).
    // ln_mode 1: mean / rstd per output ROW (A operand rows), wsum per column, the bias pointer carries W beta + b.
    // ln_mode 2: mean / rstd per output COLUMN (B operand rows), ln_wsum and ln_rowbias per output row.
    // The statistics come from the GEMM that produced the normalised tensor: ln_stats [rows][ln_nst][2] = per N tile of that
    // GEMM, sum x and sum x^2 of the row (rowstats_out of the producer; ln_nst = its N tile count).
    int ln_mode;
    const float* ln_wsum;
    const float* ln_rowbias;
    float ln_eps;
    const float* ln_stats;
    int ln_nst;
    float* rowstats_out;  // producer side: [rows][gridDim.y][2]
};

// optional LayerNorm plumbing of one GEMM: consumer side (mode != 0) and / or producer side (stats_out != null)
struct LnFuse { int mode; const float* wsum; const float* rowbias; float eps; const float* stats; int nst; float* stats_out; };

struct GemmOp {
    CUtensorMap mapA, mapB, mapC, mapR;
    GemmParams p;
    dim3 grid;
    int smem_bytes;
};

// Describe a conv3x3(stride 1, pad 1) / conv1x1 / linear as a GemmOp. `act_*` describe the NHWC input view.
struct ActView {
    const void* ptr;
    int NB, H, W, C, ld;  // ld = elements between consecutive pixels (>= C)
    // 3x3 convolutions only: stride 2 reads the input through a TMA tensor map with element strides (2, 2) -- no im2col.
    // pad 1 = symmetric zero padding (UNet / TAESD downsamplers), pad 0 = zeros right / below only (AutoencoderKL).
    int stride = 1, pad = 1;
};
int build_gemm_op(GemmOp* op, const ActView& a, int taps, const bf16* wt, int N, int ldw, void* out, int ldo,
                  int out_f32, const float* bias, const float* rowvec, const bf16* residual, int ldr, int act,
                  float* partial_ws, size_t partial_ws_bytes, int force_block_n, int force_splits,
                  int force_occupancy = 0, int force_kb_per_stage = 0, int force_halo = 0, const LnFuse* ln = nullptr);
int launch_gemm_op(const GemmOp& op, cudaStream_t st);
struct GemmCand { int bn, splits, occ, kbs, mode; };
int enumerate_gemm_candidates(const ActView& a, int taps, const bf16* wt, int N, int ldw, void* outp, int ldo, int out_f32,
                              const float* bias, const float* rowvec, const bf16* res, int ldr, int act, float* ws, size_t ws_bytes,
                              const LnFuse* ln, std::vector<GemmCand>* out);
int gemm_init();  // sets func attributes; call once per process after a device is selected

// ---- tcgen05 attention ---------------------------------------------------------------------------
struct AttnOp {
    CUtensorMap mapQ, mapK, mapVt;
    int heads, d, dk_pad, dv_pad;
    int nq, nk;        // per image
    int batch;
    int q_rows_per_img, k_rows_per_img, vt_cols_per_img;
    bf16* out; int ldo;
    float scale_log2e;
    int stages, tmem_cols, smem_bytes;
    unsigned int* trace; // see GemmParams::trace
    int variant, poly;   // 0: attention_kernel (one 128-query tile per CTA); 2: attention2_kernel (two tiles), every poly-th ex2 on the FMA pipe
    dim3 grid;
};
int build_attn_op(AttnOp* op, const bf16* q, int ldq, const bf16* k, int ldk, const bf16* vt, int ldvt, bf16* out,
                  int ldo, int batch, int heads, int d, int nq, int nk, int q_rows_per_img, int k_rows_per_img,
                  int vt_cols_per_img, int vt_rows);
int launch_attn_op(const AttnOp& op, cudaStream_t st);
int attn_init();
int bw_init();   // per-device function attributes of the bandwidth kernels

// ---- bandwidth kernels ---------------------------------------------------------------------------
int launch_groupnorm(const bf16* x, int ldx, bf16* y, int ldy, const float* gamma, const float* beta, int NB, int HW,
                     int C, int groups, float eps, int silu, float* partial_ws /* two-kernel fallback only */, cudaStream_t st);
int groupnorm_ws_floats(int NB, int HW, int C, int groups);
int launch_layernorm(const bf16* x, int ldx, bf16* y, int ldy, const float* gamma, const float* beta, int rows, int C,
                     float eps, cudaStream_t st);
int launch_upsample_nearest(const bf16* x, int ldx, bf16* y, int ldy, int NB, int Hi, int Wi, int Ho, int Wo, int C,
                            cudaStream_t st);
int launch_im2col_s2(const bf16* x, int ldx, bf16* y, int NB, int Hi, int Wi, int C, int Ho, int Wo, cudaStream_t st, int pad = 1);
int launch_softmax_rows(const float* S, int lds, bf16* P, int ldp, int rows, int cols, float scale, cudaStream_t st);
int launch_kl_sample(const float* enc8, const float* wq, const float* bq, const float* noise, float* z, long px, float scaling,
                     cudaStream_t st);
int launch_kl_post_quant(const float* lat, const float* wp, const float* bp, float* out, long px, float inv_scaling, cudaStream_t st);
int launch_rowstats(const bf16* x, int ldx, int rows, int C, float* out, cudaStream_t st);
int launch_ln_fold_weight(bf16* W, int N, int K, const float* gamma, const float* beta, const float* bias, float* wsum, float* wb,
                          cudaStream_t st);
int launch_chain_weights(const bf16* Wp, const bf16* W2, const float* b2, const float* bp, bf16* Wc, float* bc, int C, int K2,
                         cudaStream_t st);
int launch_fold_v_bias(const bf16* wo, const float* bv, const float* bo, float* out, int C, cudaStream_t st);
int launch_splitk_reduce(const GemmParams& p, long rows, cudaStream_t st);
int launch_conv3x3_small_cin(const void* x, int x_kind, int NB, int H, int W, int Cin, const float* w, const float* bias,
                             bf16* y, int ldy, int Cout, int relu /*0 none, 1 ReLU, 2 SiLU*/, cudaStream_t st,
                             const bf16* res = nullptr, int ldr = 0);
int launch_clip_embed(const int* ids, const bf16* tok, const bf16* pos, float* x, int T, int C, int vocab, cudaStream_t st);
int launch_layernorm_f32in(const float* x, bf16* y, const float* gamma, const float* beta, int rows, int C, float eps, cudaStream_t st);
int launch_clip_attention(const bf16* qkv, bf16* out, int T, int heads, cudaStream_t st);
int launch_crop_resize(const uint8_t* src, int in_w, int in_h, int x0, int y0, int cw, int ch, uint8_t* tmp, uint8_t* out,
                       int W, int H, const int* hb, const int* hk, int hks, const int* vb, const int* vk, int vks, int NB,
                       cudaStream_t st);
int launch_sobel_control(const uint8_t* rgb, float* mag, unsigned int* maxbits, float* control, int NB, int H, int W,
                         float low, float high, cudaStream_t st);
int launch_conv3x3_direct(const bf16* x, int ldx, int NB, int Hi, int Wi, int Cin, const bf16* w, const float* bias, bf16* y,
                          int ldy, int Cout, int stride, int silu, cudaStream_t st);
int launch_add_noise(const float* x0, const float* noise, float* out, float a, float b, long n, cudaStream_t st);
int launch_lcm_step(const float* eps, const float* x, const float* z, float* x_prev, float* denoised, float sqrt_a,
                    float sqrt_1ma, float c_skip, float c_out, float sqrt_ap, float sqrt_1map, int has_noise, long n,
                    cudaStream_t st);
int launch_yuv420_to_rgb(const uint8_t* y, const uint8_t* u, const uint8_t* v, uint8_t* rgb, int NB, int H, int W,
                         cudaStream_t st);
int launch_pack_rgb_yuv420(const float* img, int ldi, uint8_t* rgb, uint8_t* y, uint8_t* u, uint8_t* v, int NB, int H,
                           int W, int taesd_denorm, cudaStream_t st);
// y[out] = act_out( W[out][in] * act_in(x) + b ), fp32, one warp per output (time-embedding path, M = 1)
int launch_gemv_f32(const float* W, const float* x, const float* b, float* y, int out, int in, int silu_in, int silu_out,
                    cudaStream_t st);
int attn_dk_pad(int d);
int attn_dv_pad(int d);
int ensure_init();

unsigned int read_trap_code_gemm();
unsigned int read_trap_code_attn();
// device addresses of the same words on the current device (for asynchronous copies behind a frame)
const unsigned int* trap_code_addr_gemm();
const unsigned int* trap_code_addr_attn();

}  // namespace vsd
