// C-ABI entry points for single operators (declared in include/videosd.h, section "operator entry points").
// These are what the -m gpu parity tests call, one kernel family at a time, with device pointers owned by
// the caller (torch tensors' data_ptr()). The frame-level API lives in vsd_engine.cu.
#include "vsd_internal.h"
#include "../../include/videosd.h"
#include <cstdlib>
#include <mutex>
#include <vector>

namespace vsd {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }

int pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VSD_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}

static std::mutex g_init_mu;
static std::vector<int> g_inited_devices;

// Per-device one-time setup (kernel attributes). Cheap to call repeatedly.
int ensure_init() {
    int dev = -1;
    VSD_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_init_mu);
    for (int d : g_inited_devices)
        if (d == dev) return 0;
    cudaDeviceProp prop;
    VSD_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("libvideosd requires an sm_100 (Blackwell B200) device; found sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor) + ". There is no fallback path.");
        return -4;
    }
    int rc = gemm_init();
    if (rc) return rc;
    rc = attn_init();
    if (rc) return rc;
    rc = bw_init();
    if (rc) return rc;
    g_inited_devices.push_back(dev);
    return 0;
}

static float* g_ws = nullptr;
static size_t g_ws_bytes = 0;
static int ensure_ws(size_t bytes) {
    if (bytes <= g_ws_bytes) return 0;
    if (g_ws) cudaFree(g_ws);
    g_ws = nullptr;
    g_ws_bytes = 0;
    VSD_CHECK_CUDA(cudaMalloc(&g_ws, bytes));
    g_ws_bytes = bytes;
    return 0;
}

// VSD_POISON=1 (operator tests): before a GEMM is launched its split-K workspace and -- when it is a dense tensor of its own --
// its output are filled with 0xFF bytes (NaN as bf16 and as fp32), so that a tile the kernel fails to write shows up as NaN
// instead of as whatever an earlier, correct configuration of the same shape left there (the sweeps run thousands of
// configurations over the same buffers).
static int poison_for_test(const GemmOp& op, void* out, int ldo, int out_f32, int n_out, const void* residual, cudaStream_t st) {
    static const bool on = getenv("VSD_POISON") && atoi(getenv("VSD_POISON")) != 0;
    if (!on) return 0;
    const long rows_out = (long)op.p.NB * op.p.H * op.p.W;   // output geometry (strided convolutions: not the input's)
    if (op.p.splits > 1 && op.p.partial)
        VSD_CHECK_CUDA(cudaMemsetAsync(op.p.partial, 0xFF, (size_t)op.p.splits * rows_out * op.p.N * 4, st));
    if (ldo == n_out && residual != out)
        VSD_CHECK_CUDA(cudaMemsetAsync(out, 0xFF, (size_t)rows_out * ldo * (out_f32 ? 4 : 2), st));
    return 0;
}

}  // namespace vsd

using namespace vsd;

extern "C" {

const char* vsd_last_error(void) { return get_error(); }

int vsd_abi_version(void) { return VSD_ABI_VERSION; }

int vsd_check_pipeline_fault(void) {
    unsigned int a = read_trap_code_gemm();
    unsigned int b = read_trap_code_attn();
    if (a || b) {
        char buf[200];
        snprintf(buf, sizeof(buf), "device-side wait timed out: gemm=0x%08x attn=0x%08x", a, b);
        set_error(buf);
        return -5;
    }
    return 0;
}

int vsd_op_conv_gemm(const void* x, int nb, int h, int w, int c, int ldx, int taps, const void* wt, int n, void* out,
                     int ldo, int out_f32, const float* bias, const float* rowvec, const void* residual, int ldr,
                     int act, int block_n, int splits, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    const long rows = (long)nb * h * w;
    rc = ensure_ws((size_t)16 * rows * n * 4 + 1024);
    if (rc) return rc;
    GemmOp op;
    ActView a{x, nb, h, w, c, ldx};
    // test knobs: act bit 8 requests the 3x3 halo mode, bit 9 CTA pairs (cta_group::2)
    // bit 10: persistent conv; bits 13 / 14: force / forbid the in-cluster split-K reduction
    const int want_halo = ((act & 256) ? 1 : 0) | ((act & 512) ? 2 : 0) | ((act & 1024) ? 4 : 0) | ((act & 8192) ? 8 : 0) | ((act & 16384) ? 16 : 0);
    // bit 11: 3x3 stride 2 (pad 1); bit 12 with it: zeros right / below only
    if (act & 2048) { a.stride = 2; a.pad = (act & 4096) ? 0 : 1; }
    act &= ~(256 | 512 | 1024 | 2048 | 4096 | 8192 | 16384);
    rc = build_gemm_op(&op, a, taps, reinterpret_cast<const bf16*>(wt), n, taps * c, out, ldo, out_f32, bias, rowvec,
                       reinterpret_cast<const bf16*>(residual), ldr, act, g_ws, g_ws_bytes, block_n, splits, 0, 0, want_halo);
    if (rc) return rc;
    rc = poison_for_test(op, out, ldo, out_f32, (act & 0xF) == ACT_GEGLU ? n / 2 : n, residual, reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
    return launch_gemm_op(op, reinterpret_cast<cudaStream_t>(stream));
}

/* The configurations the autotuner may pick for this GEMM (enumerate_gemm_candidates): cand receives up to max_cand rows of
 * (block_n, splits, occ, kb_per_stage, mode). act: ACT_* flags as in vsd_op_conv_gemm (no test knobs); stride2 / pad as there.
 * Returns the number of candidates (>= 0) or a negative error. */
int vsd_op_gemm_candidates(const void* x, int nb, int h, int w, int c, int ldx, int taps, int stride2, int pad, const void* wt, int n,
                           void* out, int ldo, int out_f32, const float* bias, const float* rowvec, const void* residual, int ldr,
                           int act, int* cand, int max_cand) {
    int rc = ensure_init();
    if (rc) return rc < 0 ? rc : -1;
    rc = ensure_ws((size_t)16 * nb * h * w * n * 4 + 1024);
    if (rc) return rc < 0 ? rc : -1;
    ActView a{x, nb, h, w, c, ldx};
    if (stride2) { a.stride = 2; a.pad = pad; }
    std::vector<GemmCand> v;
    enumerate_gemm_candidates(a, taps, reinterpret_cast<const bf16*>(wt), n, taps * c, out, ldo, out_f32, bias, rowvec,
                              reinterpret_cast<const bf16*>(residual), ldr, act, g_ws, g_ws_bytes, nullptr, &v);
    int k = 0;
    for (; k < (int)v.size() && k < max_cand; ++k) {
        cand[k * 5 + 0] = v[k].bn; cand[k * 5 + 1] = v[k].splits; cand[k * 5 + 2] = v[k].occ; cand[k * 5 + 3] = v[k].kbs; cand[k * 5 + 4] = v[k].mode;
    }
    return k;
}

/* One explicit configuration (a row of vsd_op_gemm_candidates). */
int vsd_op_conv_gemm_cfg(const void* x, int nb, int h, int w, int c, int ldx, int taps, int stride2, int pad, const void* wt, int n,
                         void* out, int ldo, int out_f32, const float* bias, const float* rowvec, const void* residual, int ldr, int act,
                         int block_n, int splits, int occ, int kb_per_stage, int mode, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    rc = ensure_ws((size_t)16 * nb * h * w * n * 4 + 1024);
    if (rc) return rc;
    ActView a{x, nb, h, w, c, ldx};
    if (stride2) { a.stride = 2; a.pad = pad; }
    GemmOp op;
    rc = build_gemm_op(&op, a, taps, reinterpret_cast<const bf16*>(wt), n, taps * c, out, ldo, out_f32, bias, rowvec,
                       reinterpret_cast<const bf16*>(residual), ldr, act, g_ws, g_ws_bytes, block_n, splits, occ, kb_per_stage, mode);
    if (rc) return rc;
    rc = poison_for_test(op, out, ldo, out_f32, (act & 0xF) == ACT_GEGLU ? n / 2 : n, residual, reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
    return launch_gemm_op(op, reinterpret_cast<cudaStream_t>(stream));
}

/* bring-up: same as vsd_op_conv_gemm but CTA (0,0,0) records 7 clock64() phase stamps into dbg (device int64[8]) */
int vsd_op_conv_gemm_timed(const void* x, int nb, int h, int w, int c, int ldx, int taps, const void* wt, int n, void* out,
                           int ldo, const float* bias, int block_n, int splits, int occ, int kb_per_stage, long long* dbg,
                           void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    rc = ensure_ws((size_t)16 * nb * h * w * n * 4 + 1024);
    if (rc) return rc;
    GemmOp op;
    ActView a{x, nb, h, w, c, ldx};
    rc = build_gemm_op(&op, a, taps, reinterpret_cast<const bf16*>(wt), n, taps * c, out, ldo, 0, bias, nullptr, nullptr, 0,
                       0, g_ws, g_ws_bytes, block_n, splits, occ & 0xFF, kb_per_stage, occ >> 8 /* mode: 1 halo, 2 pairs */);
    if (rc) return rc;
    op.p.dbg = dbg;
    return launch_gemm_op(op, reinterpret_cast<cudaStream_t>(stream));
}

/* Linear(LayerNorm(x)) with the LayerNorm folded into the GEMM (tests): w_raw is the ORIGINAL bf16 weight [n][c]; a scratch copy
 * is scaled by gamma, then out = rstd * (W' x - mean * wsum) + (W beta + bias). swapped = 0: x [rows][ldx] is the row operand,
 * out [rows][ldo]; swapped = 1: out^T = W LN(x)^T, out [n][ldo] with the tokens along the columns (the V^T projection).
 * act: 0 or 1 (GEGLU; weight / bias rows interleaved per 128-row tile). stats: [rows][nst][2] partial row sums of x, x^2 as a
 * producing GEMM leaves them (vsd_op_linear_stats), or NULL to have them computed here. */
int vsd_op_linear_ln(const void* x, int rows, int c, int ldx, const void* w_raw, int n, const float* gamma, const float* beta,
                     const float* bias, float eps, void* out, int ldo, int swapped, int act, int block_n, const float* stats,
                     int nst, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    bf16* w = nullptr;
    float *wsum = nullptr, *wb = nullptr, *own = nullptr;
    VSD_CHECK_CUDA(cudaMalloc(&w, (size_t)n * c * 2));
    VSD_CHECK_CUDA(cudaMalloc(&wsum, (size_t)n * 4));
    VSD_CHECK_CUDA(cudaMalloc(&wb, (size_t)n * 4));
    VSD_CHECK_CUDA(cudaMemcpyAsync(w, w_raw, (size_t)n * c * 2, cudaMemcpyDeviceToDevice, st));
    rc = launch_ln_fold_weight(w, n, c, gamma, beta, bias, wsum, wb, st);
    if (!rc && !stats) {
        VSD_CHECK_CUDA(cudaMalloc(&own, (size_t)rows * 8));
        rc = launch_rowstats(reinterpret_cast<const bf16*>(x), ldx, rows, c, own, st);
        stats = own;
        nst = 1;
    }
    GemmOp op;
    const int mode = (act & 256) ? 2 : 0;   // test knob: bit 8 = CTA pairs (non-swapped only)
    act &= ~256;
    if (!rc) {
        if (swapped) {
            LnFuse ln{2, wsum, wb, eps, stats, nst, nullptr};
            ActView aw{w, 1, 1, n, c, c};
            rc = build_gemm_op(&op, aw, 1, reinterpret_cast<const bf16*>(x), rows, ldx, out, ldo, 0, nullptr, nullptr, nullptr, 0,
                               ACT_NONE | ACT_A_STATIC_FLAG, nullptr, 0, block_n, 1, 0, 0, 0, &ln);
        } else {
            LnFuse ln{1, wsum, nullptr, eps, stats, nst, nullptr};
            ActView ax{x, 1, 1, rows, c, ldx};
            rc = build_gemm_op(&op, ax, 1, w, n, c, out, ldo, 0, wb, nullptr, nullptr, 0, act, nullptr, 0, block_n, 1, mode ? 1 : 0, 0, mode, &ln);
        }
    }
    if (!rc) rc = launch_gemm_op(op, st);
    cudaStreamSynchronize(st);
    cudaFree(w); cudaFree(wsum); cudaFree(wb);
    if (own) cudaFree(own);
    return rc;
}

/* out = x W^T + bias (+ residual), and per row and per N tile the sums of the stored values and of their squares in
 * stats_out [rows][n_tiles][2] (what a LayerNorm folded into the next GEMM consumes). mode: 0 plain, 1 CTA pairs, 8 in-cluster
 * split-K (splits > 1). Returns the N tile count (> 0) or a negative error. */
int vsd_op_linear_stats(const void* x, int rows, int c, int ldx, const void* w, int n, const float* bias, const void* residual,
                        int ldr, void* out, int ldo, float* stats_out, int block_n, int splits, int mode, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    LnFuse ln{0, nullptr, nullptr, 0.f, nullptr, 0, stats_out};
    rc = ensure_ws((size_t)16 * rows * n * 4 + 1024);   // the common split-K bound (the in-cluster reduction does not touch it)
    if (rc) return rc < 0 ? rc : -1;
    GemmOp op;
    ActView ax{x, 1, 1, rows, c, ldx};
    rc = build_gemm_op(&op, ax, 1, reinterpret_cast<const bf16*>(w), n, c, out, ldo, 0, bias, nullptr,
                       reinterpret_cast<const bf16*>(residual), ldr, ACT_NONE, g_ws, g_ws_bytes, block_n, splits > 1 ? splits : 1, 0, 0,
                       mode == 1 ? 2 : (splits > 1 ? 8 : 16), &ln);
    if (rc) return rc < 0 ? rc : -1;
    rc = launch_gemm_op(op, reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc < 0 ? rc : -1;
    return (int)op.grid.y;
}

int vsd_op_attention(const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* out, int ldo,
                     int batch, int heads, int d, int nq, int nk, int q_rows_per_img, int k_rows_per_img,
                     int vt_cols_per_img, int vt_rows, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    AttnOp op;
    rc = build_attn_op(&op, reinterpret_cast<const bf16*>(q), ldq, reinterpret_cast<const bf16*>(k), ldk,
                       reinterpret_cast<const bf16*>(vt), ldvt, reinterpret_cast<bf16*>(out), ldo, batch, heads, d, nq,
                       nk, q_rows_per_img, k_rows_per_img, vt_cols_per_img, vt_rows);
    if (rc) return rc;
    return launch_attn_op(op, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_groupnorm(const void* x, int ldx, void* y, int ldy, const float* gamma, const float* beta, int nb, int hw,
                     int c, int groups, float eps, int silu, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    rc = ensure_ws((size_t)groupnorm_ws_floats(nb, hw, c, groups) * 4 + 1024);
    if (rc) return rc;
    return launch_groupnorm(reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<bf16*>(y), ldy, gamma, beta, nb, hw,
                            c, groups, eps, silu, g_ws, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_layernorm(const void* x, int ldx, void* y, int ldy, const float* gamma, const float* beta, int rows, int c,
                     float eps, void* stream) {
    return launch_layernorm(reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<bf16*>(y), ldy, gamma, beta, rows, c,
                            eps, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_upsample_nearest(const void* x, int ldx, void* y, int ldy, int nb, int hi, int wi, int ho, int wo, int c,
                            void* stream) {
    return launch_upsample_nearest(reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<bf16*>(y), ldy, nb, hi, wi,
                                   ho, wo, c, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_im2col_s2(const void* x, int ldx, void* y, int nb, int hi, int wi, int c, int ho, int wo, void* stream) {
    return launch_im2col_s2(reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<bf16*>(y), nb, hi, wi, c, ho, wo,
                            reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_conv3x3_small_cin(const void* x, int x_kind, int nb, int h, int w, int cin, const float* wt,
                             const float* bias, void* y, int ldy, int cout, int relu, void* stream) {
    return launch_conv3x3_small_cin(x, x_kind, nb, h, w, cin, wt, bias, reinterpret_cast<bf16*>(y), ldy, cout, relu,
                                    reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_sobel_control(const uint8_t* rgb, float* mag, unsigned int* maxbits, float* control, int nb, int h, int w,
                         float low, float high, void* stream) {
    return launch_sobel_control(rgb, mag, maxbits, control, nb, h, w, low, high, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_crop_resize(const uint8_t* src, int in_w, int in_h, int x0, int y0, int cw, int ch, uint8_t* tmp, uint8_t* out, int w,
                       int h, const int* h_bounds, const int* h_coeffs, int h_ksize, const int* v_bounds, const int* v_coeffs,
                       int v_ksize, int nb, void* stream) {
    if (ensure_init()) return 1;
    return launch_crop_resize(src, in_w, in_h, x0, y0, cw, ch, tmp, out, w, h, h_bounds, h_coeffs, h_ksize, v_bounds, v_coeffs,
                              v_ksize, nb, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_conv3x3_direct(const void* x, int ldx, int nb, int hi, int wi, int cin, const void* wt, const float* bias, void* y,
                          int ldy, int cout, int stride, int silu, void* stream) {
    return launch_conv3x3_direct(reinterpret_cast<const bf16*>(x), ldx, nb, hi, wi, cin, reinterpret_cast<const bf16*>(wt), bias,
                                 reinterpret_cast<bf16*>(y), ldy, cout, stride, silu, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_add_noise(const float* x0, const float* noise, float* out, float sqrt_alpha, float sqrt_one_minus_alpha,
                     long n, void* stream) {
    return launch_add_noise(x0, noise, out, sqrt_alpha, sqrt_one_minus_alpha, n, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_lcm_step(const float* eps, const float* x, const float* z, float* x_prev, float* denoised, float sqrt_a,
                    float sqrt_1ma, float c_skip, float c_out, float sqrt_ap, float sqrt_1map, int has_noise, long n,
                    void* stream) {
    return launch_lcm_step(eps, x, z, x_prev, denoised, sqrt_a, sqrt_1ma, c_skip, c_out, sqrt_ap, sqrt_1map, has_noise, n,
                           reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_yuv420_to_rgb(const uint8_t* y, const uint8_t* u, const uint8_t* v, uint8_t* rgb, int nb, int h, int w,
                         void* stream) {
    return launch_yuv420_to_rgb(y, u, v, rgb, nb, h, w, reinterpret_cast<cudaStream_t>(stream));
}

int vsd_op_pack_rgb_yuv420(const float* img, int ldi, uint8_t* rgb, uint8_t* y, uint8_t* u, uint8_t* v, int nb, int h,
                           int w, int taesd_denorm, void* stream) {
    return launch_pack_rgb_yuv420(img, ldi, rgb, y, u, v, nb, h, w, taesd_denorm, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
