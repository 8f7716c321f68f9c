// Frame-level engine behind the C ABI (include/videosd.h, "engine" section): owns the weights (bf16, repacked
// for the tcgen05 kernels), builds a static launch plan for one (batch, height, width, steps) configuration
//   YUV420/RGB in -> TAESD encode -> add noise -> steps x (UNet, LCM step) -> TAESD decode -> pack RGB/YUV420
// and replays it as one CUDA graph per frame batch (PDL edges; the V^T projections, resnet shortcuts and the ControlNet run
// as parallel branches on side streams). Mirrors the sequencing of diffusert/lcm/lcm_controlnet.py:380-618 with the
// UNet2DConditionModel / AutoencoderTiny topology of SURVEY.md Appendix A; optional: the canny ControlNet + Sobel front end,
// center crop + Lanczos resize of the input, the CLIP text tower (once per prompt), AutoencoderKL instead of AutoencoderTiny.
// Activations are NHWC bf16; latents, scheduler math and normalisation statistics are fp32. Lanes (vsd_create_lane) keep
// several frames in flight on one copy of the weights.
#include "vsd_internal.h"
#include "../../include/videosd.h"
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace vsd {

// ------------------------------------------------------------------------------------------------ helpers
static inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

static void parallel_for(long n, const std::function<void(long, long)>& fn) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw ? (hw > 16 ? 16 : hw) : 4);
    if (n < 1 << 16) nt = 1;
    std::vector<std::thread> th;
    const long chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const long a = t * chunk, b = std::min(n, a + chunk);
        if (a >= b) break;
        th.emplace_back(fn, a, b);
    }
    for (auto& t : th) t.join();
}

struct View {  // NHWC bf16 activation view
    bf16* p = nullptr;
    int nb = 0, h = 0, w = 0, c = 0, ld = 0;
    long rows() const { return (long)nb * h * w; }
    View slice(int c0, int cn) const { View v = *this; v.p = p + c0; v.c = cn; return v; }
    ActView act() const { return ActView{p, nb, h, w, c, ld}; }
    ActView act_rows() const { return ActView{p, 1, 1, (int)rows(), c, ld}; }  // as a [rows, c] matrix
};

struct DevW {
    void* p = nullptr;
    std::vector<int64_t> shape;
    int is_bf16 = 0;
};

class Arena {
  public:
    char* base = nullptr;
    size_t cap = 0, off = 0, peak = 0;
    int init(size_t bytes) {
        VSD_CHECK_CUDA(cudaMalloc(&base, bytes));
        VSD_CHECK_CUDA(cudaMemset(base, 0, bytes));
        cap = bytes; off = 0; peak = 0;
        return 0;
    }
    void destroy() { if (base) cudaFree(base); base = nullptr; }
    void* alloc(size_t bytes) {
        const size_t a = (off + 255) & ~size_t(255);
        if (a + bytes > cap) return nullptr;
        off = a + bytes;
        if (off > peak) peak = off;
        return base + a;
    }
    size_t mark() const { return off; }
    void release(size_t m) { off = m; }
};

struct Launch {
    std::function<int(cudaStream_t)> fn;
    std::string tag;   // dotted scope, e.g. "step0.unet.down0.attn1.ff1" (used by the section profiler)
    // Independent work inside a frame runs on side streams (parallel branches of the captured graph): a launch with side = k
    // (1, 2) forks from the main stream right where it sits in the plan; a later launch whose `join` has bit k-1 set waits
    // for side stream k. Stream 1 carries short branches (V^T projection, resnet shortcut), stream 2 the ControlNet.
    int side = 0, join = 0;
    int operator()(cudaStream_t st) const { return fn(st); }
};
static thread_local std::string g_scope;   // current tag while a plan is being built
struct Scope {
    std::string saved;
    explicit Scope(const std::string& name) : saved(g_scope) { g_scope = g_scope.empty() ? name : g_scope + "." + name; }
    ~Scope() { g_scope = saved; }
};
template <typename F>
static Launch mk(F f, const char* op) { return Launch{std::function<int(cudaStream_t)>(f), g_scope + "." + op}; }

static constexpr int kMaxSteps = 50;
struct StepScalars { float sqrt_a, sqrt_1ma, c_skip, c_out, sqrt_ap, sqrt_1map; int t; };

struct Engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::unordered_map<std::string, DevW> w;
    bool owns_weights = true;   // false for a lane created with vsd_create_lane (weights belong to the parent)
    // configuration
    int NB = 0, H = 0, W = 0, h8 = 0, w8 = 0;
    int steps = 0, has_step_noise = 0;
    float an_a = 0, an_b = 0;  // add_noise coefficients at timesteps[0]
    std::vector<StepScalars> sc;
    std::vector<float> w_emb;  // 256
    bool configured = false, schedule_set = false, context_set = false;
    // buffers
    Arena arena;
    size_t arena_static_mark = 0;
    float* splitk_ws = nullptr; size_t splitk_bytes = 0;
    float* splitk_ws_side[2] = {nullptr, nullptr};   // split-K workspaces of the side-stream branches (they run concurrently)
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
    int lane_index = 0, lanes_created = 1;
    Engine* root = nullptr;   // the engine owning the weights (null for that engine itself)
    uint8_t *d_y = nullptr, *d_u = nullptr, *d_v = nullptr, *d_rgb_in = nullptr;       // inputs
    uint8_t *d_oy = nullptr, *d_ou = nullptr, *d_ov = nullptr, *d_rgb_out = nullptr;   // outputs
    float *init_latents = nullptr, *noisy = nullptr, *init_noise = nullptr, *step_noise = nullptr, *image = nullptr;
    std::vector<float*> eps, lat, den;  // per step
    bf16* ctx_bf16 = nullptr;  // [NB][128][768]
    // per transformer layer cross-attention K / V^T caches (filled by set_context)
    struct XAttn { std::string prefix; int C, d, dkp; bf16* k2; bf16* v2t; };
    std::vector<XAttn> xattn;
    // per step: temb projections per resnet [NB][cout]
    std::unordered_map<std::string, std::vector<float*>> temb;  // resnet prefix -> per-step device pointer
    float *t_emb_in = nullptr, *t_h = nullptr, *t_emb = nullptr;
    // plans
    std::vector<Launch> plan_pre_yuv, plan_pre_rgb, plan_core, plan_post;
    std::vector<std::vector<Launch>> plan_unet;  // per step (debug entry)
    cudaGraphExec_t graph_yuv = nullptr, graph_rgb = nullptr;
    // GEMM autotuner: (block_n, splits, occupancy, mode) per distinct shape, timed with L2 flushed. 0 = off (heuristics);
    // n >= 1 = tune for n frames in flight on this GPU: a launch costs its duration times max(share of the SMs it holds, 1/n),
    // so with one frame in flight the fastest configuration wins and with several, configurations that leave SMs to the
    // other frames are preferred over split-K / many small CTAs.
    int autotune = 1;
    struct Tuned { int bn, splits, occ, kbs, halo; float us; };
    std::unordered_map<std::string, Tuned> tuned;
    long tune_misses = 0;   // shapes that had to be timed on the device (not found in a loaded tuning table)
    void* flush_buf = nullptr; size_t flush_bytes = 0;
    void* tune_w = nullptr; size_t tune_w_bytes = 0;   // autotuner: cold copies of the weight operand being timed
    // ControlNet: canny/Sobel-conditioned residuals added to the UNet skips every step
    // GPU center-crop + Lanczos resize of arbitrary-size input frames
    struct Resize { int in_w = 0, in_h = 0, x0 = 0, y0 = 0, cw = 0, ch = 0, hks = 0, vks = 0;
                    int *hb = nullptr, *hk = nullptr, *vb = nullptr, *vk = nullptr; uint8_t *src = nullptr, *tmp = nullptr, *yuv = nullptr; } rz;
    // CLIP text encoder: built on first use, independent of configure()
    struct Clip { bool built = false; int* ids = nullptr; float* x = nullptr; bf16 *xn = nullptr, *qkv = nullptr, *att = nullptr, *h = nullptr,
                  *out = nullptr; std::vector<Launch> plan; } clip;
    // VAE: 0 = AutoencoderTiny / TAESD (what the reference loads, videopipeline.py:67-69), 1 = AutoencoderKL (the pipeline's
    // declared VAE type; weights under "vae_kl.")
    int vae_kind = 0;
    float* vae_noise = nullptr;   // [NB][h8][w8][4] noise of latent_dist.sample()
    std::unordered_map<std::string, float*> derived;   // small tensors computed from weights at plan-build time (persistent)
    bool cn_enabled = false;
    float* cn_scales = nullptr;        // device [13]: logspace(-1,0,13) * conditioning scale (guess mode)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t done_ev = nullptr;   // blocking-sync completion event (sync_engine)
    double wait_ema_us = 0.0, wait_last_us = 0.0;   // host wait of the last frames (two-phase frame wait, sync_engine)
    // VSD_TRACE=1: per tensor-core launch of the plans three device counters (GemmParams::trace), dumped by the watchdog
    unsigned int* trace_dev = nullptr;
    std::vector<std::string> trace_tags;
    std::vector<int> trace_ctas;
    // device-side wait-timeout words of the two tensor-core translation units, copied behind every frame into pinned host
    // memory so the production entry points can report a fault without an extra synchronisation
    const unsigned int* fault_dev[2] = {nullptr, nullptr};
    unsigned int* fault_host = nullptr;
    long launches_per_frame_yuv = 0;
    std::string err;
};

#define ENG_REQUIRE(cond, msg)                                              \
    do {                                                                    \
        if (!(cond)) {                                                      \
            set_error(std::string("engine: ") + (msg) + " [" #cond "]");    \
            return -1;                                                      \
        }                                                                   \
    } while (0)

// LayerNorm folded into the consuming GEMM (GemmParams::ln_mode): the weight is scaled by gamma IN PLACE, once per device
// pointer for the whole process (lanes and other engines sharing the weights find the derived vectors here).
struct LnFold { float* wsum; float* wb; };
static std::mutex g_ln_mu;
static std::unordered_map<const void*, LnFold> g_ln_folded;
static bool ln_fuse_enabled() {
    static const bool on = !(getenv("VSD_LN_FUSE") && atoi(getenv("VSD_LN_FUSE")) == 0);
    return on;
}
// Transformer tail folded into one GEMM (launch_chain_weights): keyed by the ff.net.2 weight pointer, shared like the LN folds.
struct ChainW { bf16* wc; float* bc; };
static std::unordered_map<const void*, ChainW> g_chain;
static bool ff_out_fuse_enabled() {
    static const bool on = !(getenv("VSD_FF_OUT_FUSE") && atoi(getenv("VSD_FF_OUT_FUSE")) == 0);
    return on;
}
static void ln_unfold_forget(const void* w) {   // the weight buffer is being freed / replaced
    std::lock_guard<std::mutex> lk(g_ln_mu);
    auto it = g_ln_folded.find(w);
    if (it != g_ln_folded.end()) {
        cudaFree(it->second.wsum);
        cudaFree(it->second.wb);
        g_ln_folded.erase(it);
    }
    auto ic = g_chain.find(w);
    if (ic != g_chain.end()) {
        cudaFree(ic->second.wc);
        cudaFree(ic->second.bc);
        g_chain.erase(ic);
    }
}

// ------------------------------------------------------------------------------------------------ watchdog
// VSD_WATCHDOG_S=<seconds> (off by default; bench.py and the GPU tests switch it on): a stream synchronisation that does not
// return within that time means a kernel is stuck on the device (nothing in-process can recover that), so the process reports
// it on stderr and aborts instead of hanging its caller forever.
static std::mutex g_wd_mu;
static char g_tune_now[256] = "";   // the candidate the autotuner is timing right now (watchdog / hang diagnostics)
static std::vector<long long> g_wd_deadlines;   // one entry per synchronisation in progress (milliseconds, steady clock)
static long long now_ms() {
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static int watchdog_seconds() {
    static const int s = getenv("VSD_WATCHDOG_S") ? atoi(getenv("VSD_WATCHDOG_S")) : 0;
    return s;
}
static constexpr int kTraceSlots = 16384;
static std::mutex g_trace_mu;
static std::vector<Engine*> g_trace_engines;
static bool trace_enabled() {
    static const bool on = getenv("VSD_TRACE") && atoi(getenv("VSD_TRACE")) != 0;
    return on;
}
// a counter triple for the tensor-core launch being added to a plan (null when tracing is off / the table is full)
static unsigned int* trace_slot(Engine* e, const std::string& tag, int ctas) {
    if (!trace_enabled()) return nullptr;
    std::lock_guard<std::mutex> lk(g_trace_mu);
    if (!e->trace_dev) {
        if (cudaMalloc(reinterpret_cast<void**>(&e->trace_dev), (size_t)kTraceSlots * 4 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
        cudaMemset(e->trace_dev, 0, (size_t)kTraceSlots * 4 * sizeof(unsigned int));
        g_trace_engines.push_back(e);
    }
    if ((int)e->trace_tags.size() >= kTraceSlots) return nullptr;
    e->trace_tags.push_back(tag);
    e->trace_ctas.push_back(ctas);
    return e->trace_dev + 4 * (e->trace_tags.size() - 1);
}
static void trace_forget(Engine* e) {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    for (size_t i = 0; i < g_trace_engines.size(); ++i)
        if (g_trace_engines[i] == e) { g_trace_engines.erase(g_trace_engines.begin() + (long)i); break; }
    if (e->trace_dev) cudaFree(e->trace_dev);
    e->trace_dev = nullptr;
}
// called by the watchdog thread while a kernel is stuck: copies run on their own stream (the copy engines still work)
static void trace_dump() {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    for (Engine* e : g_trace_engines) {
        if (cudaSetDevice(e->device) != cudaSuccess) continue;
        cudaStream_t st = nullptr;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) continue;
        const size_t n = e->trace_tags.size();
        std::vector<unsigned int> h(n * 4);
        if (cudaMemcpyAsync(h.data(), e->trace_dev, n * 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            fprintf(stderr, "videosd trace: engine %p: copy failed\n", (void*)e);
            continue;
        }
        int shown = 0;
        for (size_t i = 0; i < n; ++i) {
            const unsigned in = h[4 * i], own = h[4 * i + 1], out = h[4 * i + 2], met = h[4 * i + 3];
            if (in == out && in == own) continue;
            fprintf(stderr, "videosd trace: engine %p (lane %d, %dx%dx%d) launch %zu %s: %d CTAs per launch, entered %u, met the peer %u, own TMEM %u, finished %u\n",
                    (void*)e, e->lane_index, e->NB, e->H, e->W, i, e->trace_tags[i].c_str(), e->trace_ctas[i], in, met, own, out);
            if (++shown >= 40) break;
        }
        if (!shown) fprintf(stderr, "videosd trace: engine %p (lane %d, %dx%dx%d): every traced launch finished (%zu traced)\n",
                            (void*)e, e->lane_index, e->NB, e->H, e->W, n);
    }
    fflush(stderr);
}
static void watchdog_start_once() {
    static std::once_flag once;
    std::call_once(once, [] {
        std::thread([] {
            for (;;) {
                std::this_thread::sleep_for(std::chrono::milliseconds(500));
                const long long t = now_ms();
                bool late = false;
                {
                    std::lock_guard<std::mutex> lk(g_wd_mu);
                    for (long long d : g_wd_deadlines) late = late || t > d;
                }
                if (late) {
                    fprintf(stderr, "videosd: the device did not finish a frame within %d s (VSD_WATCHDOG_S): a kernel is stuck; aborting%s%s\n",
                            watchdog_seconds(), g_tune_now[0] ? " -- the autotuner was timing " : "", g_tune_now);
                    fflush(stderr);
                    trace_dump();
                    _exit(86);
                }
            }
        }).detach();
    });
}
struct WatchdogScope {   // registers a deadline for the synchronisation in progress
    long long mine = 0;
    WatchdogScope() {
        const int wd = watchdog_seconds();
        if (wd <= 0) return;
        watchdog_start_once();
        mine = now_ms() + 1000LL * wd;
        std::lock_guard<std::mutex> lk(g_wd_mu);
        g_wd_deadlines.push_back(mine);
    }
    ~WatchdogScope() {
        if (!mine) return;
        std::lock_guard<std::mutex> lk(g_wd_mu);
        for (size_t i = 0; i < g_wd_deadlines.size(); ++i)
            if (g_wd_deadlines[i] == mine) { g_wd_deadlines.erase(g_wd_deadlines.begin() + (long)i); break; }
    }
};
// Wait for the engine's stream. cudaStreamSynchronize spins on a host core; with several frames in flight per GPU (one host
// thread per lane, 6 lanes x 8 ranks on a 32-core box) that is more spinning threads than cores. Blocking on a
// cudaEventBlockingSync event instead costs 0.5 - 1 ms of wake-up latency per frame (measured: e2e 142.5 against 146.1 fps
// with spinning, single stream 70.3 against 73.3). So lanes and lane-pool roots wait for a FRAME by napping: the thread
// sleeps in short naps (a quarter of the time left until the predicted completion, at most 300 us) and looks at the
// completion event between them, polls with yields once the predicted completion (the shorter of the last wait and their
// moving average) is less than 0.6 ms away, and falls back to the blocking wait when the frame is much later than predicted.
// A frame that finishes EARLY (the load dropped: other lanes went idle) is noticed within one nap, not at the predicted
// time. Cost: ~2 % of one core per lane whatever the number of lanes and ranks. Every other synchronisation of a lane
// blocks; a single-lane engine spins. VSD_BLOCKING_SYNC = 0 / 1 / 2 forces spinning / blocking / napping.
static double elapsed_us(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
}
static cudaError_t sync_engine(Engine* e, bool frame = false) {
    static const int mode = getenv("VSD_BLOCKING_SYNC") ? atoi(getenv("VSD_BLOCKING_SYNC")) : -1;
    const bool lanes = e->autotune > 1 || !e->owns_weights;
    WatchdogScope wd;
    if (mode == 0 || (mode < 0 && !lanes)) return cudaStreamSynchronize(e->stream);
    if (!e->done_ev) {
        const cudaError_t r = cudaEventCreateWithFlags(&e->done_ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (r != cudaSuccess) return r;
    }
    const cudaError_t r = cudaEventRecord(e->done_ev, e->stream);
    if (r != cudaSuccess) return r;
    if (mode == 1 || !frame) return cudaEventSynchronize(e->done_ev);
    const auto t0 = std::chrono::steady_clock::now();
    bool done = false;
    if (e->wait_ema_us > 0.0) {
        const double pred = std::min(e->wait_ema_us, e->wait_last_us);
        const double limit_us = pred * 1.25 + 2000.0;
        for (;;) {
            const cudaError_t q = cudaEventQuery(e->done_ev);
            if (q == cudaSuccess) { done = true; break; }
            if (q != cudaErrorNotReady) return q;
            const double t = elapsed_us(t0);
            if (t > limit_us) break;                      // much later than predicted: stop polling
            const double left = pred - t;
            if (left > 600.0) std::this_thread::sleep_for(std::chrono::microseconds((long long)std::min(300.0, left * 0.25)));
            else std::this_thread::yield();
        }
    }
    if (!done) {
        const cudaError_t w = cudaEventSynchronize(e->done_ev);
        if (w != cudaSuccess) return w;
    }
    e->wait_last_us = elapsed_us(t0);
    e->wait_ema_us = e->wait_ema_us > 0.0 ? 0.8 * e->wait_ema_us + 0.2 * e->wait_last_us : e->wait_last_us;
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------ weights
static bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

static int upload(DevW* dw, const void* host, size_t bytes) {
    VSD_CHECK_CUDA(cudaMalloc(&dw->p, bytes));
    VSD_CHECK_CUDA(cudaMemcpy(dw->p, host, bytes, cudaMemcpyHostToDevice));
    return 0;
}

// rows of `src` ([rows][cols] fp32) -> bf16 rows placed at dst_row(r) of a zeroed [dst_rows][cols] matrix
static void scatter_rows_bf16(std::vector<uint16_t>& dst, const float* src, long rows, long cols,
                              const std::function<long(long)>& dst_row) {
    parallel_for(rows, [&](long a, long b) {
        for (long r = a; r < b; ++r) {
            const long dr = dst_row(r);
            uint16_t* d = dst.data() + dr * cols;
            const float* s = src + r * cols;
            for (long c = 0; c < cols; ++c) d[c] = f32_to_bf16_rne(s[c]);
        }
    });
}

static int load_weight(Engine* e, const std::string& name, const float* host, const int64_t* shape, int ndim) {
    ENG_REQUIRE(ndim >= 1 && ndim <= 4, "weight rank must be 1..4");
    long numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= shape[i];
    DevW dw;
    dw.shape.assign(shape, shape + ndim);
    std::string key = name;
    if (ndim == 4 && name.find("quant_conv") != std::string::npos) {   // AutoencoderKL 1x1 (post_)quant_conv: fp32 [out][in]
        int rc = upload(&dw, host, (size_t)numel * 4);
        if (rc) return rc;
    } else if (ndim == 4) {
        const long co = shape[0], ci = shape[1], kh = shape[2], kw = shape[3];
        if (ci <= 4) {  // edge convolutions stay fp32, OHWI
            std::vector<float> t((size_t)numel);
            for (long o = 0; o < co; ++o)
                for (long i = 0; i < ci; ++i)
                    for (long y = 0; y < kh; ++y)
                        for (long x = 0; x < kw; ++x)
                            t[((o * kh + y) * kw + x) * ci + i] = host[((o * ci + i) * kh + y) * kw + x];
            int rc = upload(&dw, t.data(), t.size() * 4);
            if (rc) return rc;
        } else {  // bf16 [Cout][kh*kw*Cin], tap-major
            std::vector<uint16_t> t((size_t)numel);
            parallel_for(co, [&](long a, long b) {
                for (long o = a; o < b; ++o)
                    for (long i = 0; i < ci; ++i)
                        for (long y = 0; y < kh; ++y)
                            for (long x = 0; x < kw; ++x)
                                t[((o * kh + y) * kw + x) * ci + i] = f32_to_bf16_rne(host[((o * ci + i) * kh + y) * kw + x]);
            });
            dw.is_bf16 = 1;
            int rc = upload(&dw, t.data(), t.size() * 2);
            if (rc) return rc;
        }
    } else if (ndim == 2) {
        const long out = shape[0], in = shape[1];
        const bool is_time = name.find("time_embedding.") != std::string::npos || name.find("time_emb_proj") != std::string::npos;
        if (is_time) {
            int rc = upload(&dw, host, (size_t)numel * 4);
            if (rc) return rc;
        } else if (ends_with(name, "attn1.to_q.weight") || ends_with(name, "attn1.to_k.weight")) {
            // fused, head-padded [2*8*dkp][C]: Q rows first, K rows second
            const int heads = 8, d = (int)(out / heads), dkp = attn_dk_pad(d);
            const bool is_k = ends_with(name, "attn1.to_k.weight");
            key = name.substr(0, name.size() - strlen("to_q.weight")) + "qk.weight";
            auto it = e->w.find(key);
            const long rows_total = 2L * heads * dkp;
            if (it == e->w.end()) {
                DevW nw;
                nw.is_bf16 = 1;
                nw.shape = {rows_total, in};
                VSD_CHECK_CUDA(cudaMalloc(&nw.p, (size_t)rows_total * in * 2));
                VSD_CHECK_CUDA(cudaMemset(nw.p, 0, (size_t)rows_total * in * 2));
                it = e->w.emplace(key, nw).first;
            }
            std::vector<uint16_t> t((size_t)heads * dkp * in, 0);
            scatter_rows_bf16(t, host, out, in, [&](long r) { return (r / d) * dkp + (r % d); });
            VSD_CHECK_CUDA(cudaMemcpy(reinterpret_cast<uint16_t*>(it->second.p) + (is_k ? (size_t)heads * dkp * in : 0),
                                      t.data(), t.size() * 2, cudaMemcpyHostToDevice));
            return 0;
        } else if (ends_with(name, "attn2.to_q.weight") || ends_with(name, "attn2.to_k.weight")) {
            const int heads = 8, d = (int)(out / heads), dkp = attn_dk_pad(d);
            std::vector<uint16_t> t((size_t)heads * dkp * in, 0);
            scatter_rows_bf16(t, host, out, in, [&](long r) { return (r / d) * dkp + (r % d); });
            dw.is_bf16 = 1;
            dw.shape = {(int64_t)heads * dkp, in};
            int rc = upload(&dw, t.data(), t.size() * 2);
            if (rc) return rc;
        } else if (ends_with(name, "ff.net.0.proj.weight")) {
            // GEGLU: interleave per 128-row tile [64 value rows | 64 gate rows]
            const long half = out / 2;
            ENG_REQUIRE(half % 64 == 0, "GEGLU width must be a multiple of 64");
            std::vector<uint16_t> t((size_t)numel);
            scatter_rows_bf16(t, host, out, in, [&](long r) {
                const bool gate = r >= half;
                const long j = gate ? r - half : r;
                return (j / 64) * 128 + (gate ? 64 : 0) + (j % 64);
            });
            dw.is_bf16 = 1;
            int rc = upload(&dw, t.data(), t.size() * 2);
            if (rc) return rc;
        } else {
            std::vector<uint16_t> t((size_t)numel);
            scatter_rows_bf16(t, host, out, in, [](long r) { return r; });
            dw.is_bf16 = 1;
            int rc = upload(&dw, t.data(), t.size() * 2);
            if (rc) return rc;
        }
    } else {  // vectors: fp32
        if (ends_with(name, "ff.net.0.proj.bias")) {
            const long out = shape[0], half = out / 2;
            std::vector<float> t((size_t)out);
            for (long r = 0; r < out; ++r) {
                const bool gate = r >= half;
                const long j = gate ? r - half : r;
                t[(j / 64) * 128 + (gate ? 64 : 0) + (j % 64)] = host[r];
            }
            int rc = upload(&dw, t.data(), t.size() * 4);
            if (rc) return rc;
        } else {
            int rc = upload(&dw, host, (size_t)numel * 4);
            if (rc) return rc;
        }
    }
    auto old = e->w.find(key);
    if (old != e->w.end()) {
        ln_unfold_forget(old->second.p);
        cudaFree(old->second.p);
        e->w.erase(old);
    }
    e->w.emplace(key, dw);
    return 0;
}

// ------------------------------------------------------------------------------------------------ GEMM autotuner
// One candidate = kTuneCopies launches back to back (PDL edges between them, like the frame's chain), each reading its own COLD
// copy of the constant operand (in a frame every weight is read once per pass out of 1.7 GB: never L2-resident), the
// activations warm as in the frame; L2 is flushed before the burst. Averaging over the burst resolves differences far below
// the ~2 us granularity of a single event pair, so the choice is stable from run to run.
static constexpr int kTuneCopies = 8;
static int time_gemm(Engine* e, const GemmOp* ops, int nops, float* us) {
    float best = 1e30f;
    static const int reps = getenv("VSD_TUNE_REPS") ? atoi(getenv("VSD_TUNE_REPS")) : 3;
    for (int rep = 0; rep < reps; ++rep) {
        VSD_CHECK_CUDA(cudaMemsetAsync(e->flush_buf, rep, e->flush_bytes, e->stream));   // evict L2: weights come from HBM
        VSD_CHECK_CUDA(cudaEventRecord(e->ev0, e->stream));
        for (int i = 0; i < nops; ++i) {
            int rc = launch_gemm_op(ops[i], e->stream);
            if (rc) return rc;
        }
        VSD_CHECK_CUDA(cudaEventRecord(e->ev1, e->stream));
        {
            WatchdogScope wd;
            VSD_CHECK_CUDA(cudaEventSynchronize(e->ev1));
        }
        float ms = 0.f;
        VSD_CHECK_CUDA(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        const float t = ms * 1e3f / (float)nops;
        if (t < best) best = t;
    }
    *us = best;
    return 0;
}

static int tune_gemm(Engine* e, const ActView& a, int taps, const bf16* wt, int N, int ldw, void* outp, int ldo,
                     int out_f32, const float* bias, const float* rowvec, const bf16* res, int ldr, int act,
                     Engine::Tuned* result, const LnFuse* ln = nullptr) {
    if (!e->flush_buf) {
        e->flush_bytes = (size_t)192 << 20;
        VSD_CHECK_CUDA(cudaMalloc(&e->flush_buf, e->flush_bytes));
        VSD_CHECK_CUDA(cudaEventCreate(&e->ev0));
        VSD_CHECK_CUDA(cudaEventCreate(&e->ev1));
    }
    // cold copies of the constant operand (the weight matrix `wt`; for swapped-operand / two-activation GEMMs the same data)
    const bf16* wcopy[kTuneCopies];
    for (int i = 0; i < kTuneCopies; ++i) wcopy[i] = wt;
    if (!(act & (ACT_A_STATIC_FLAG | ACT_NO_STATIC_FLAG))) {
        const size_t wbytes = ((size_t)N * ldw * 2 + 255) & ~size_t(255);
        if (wbytes * kTuneCopies > e->tune_w_bytes) {
            if (e->tune_w) cudaFree(e->tune_w);
            e->tune_w = nullptr; e->tune_w_bytes = 0;
            VSD_CHECK_CUDA(cudaMalloc(&e->tune_w, wbytes * kTuneCopies));
            e->tune_w_bytes = wbytes * kTuneCopies;
        }
        for (int i = 0; i < kTuneCopies; ++i) {
            VSD_CHECK_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(e->tune_w) + i * wbytes, wt, (size_t)N * ldw * 2, cudaMemcpyDeviceToDevice, e->stream));
            wcopy[i] = reinterpret_cast<const bf16*>(reinterpret_cast<char*>(e->tune_w) + i * wbytes);
        }
    }
    // every configuration the kernel family offers for this shape (enumerate_gemm_candidates, gemm_tc.cu: the same list the
    // operator-level sweep test runs against fp32), timed on the cold weight copies
    std::vector<GemmCand> cands;
    enumerate_gemm_candidates(a, taps, wcopy[0], N, ldw, outp, ldo, out_f32, bias, rowvec, res, ldr, act, e->splitk_ws, e->splitk_bytes, ln,
                              &cands);
    Engine::Tuned best{0, 1, 0, 1, 0, 1e30f};
    float best_cost = 1e30f;
    for (const GemmCand& c : cands) {
        snprintf(g_tune_now, sizeof(g_tune_now), "%dx%dx%dx%d taps %d N %d act %d: bn=%d splits=%d occ=%d kbs=%d mode=%d", a.NB, a.H, a.W, a.C, taps, N,
                 act, c.bn, c.splits, c.occ, c.kbs, c.mode);
        GemmOp ops[kTuneCopies];
        bool ok = true;
        for (int i = 0; i < kTuneCopies && ok; ++i)
            ok = !build_gemm_op(&ops[i], a, taps, wcopy[i], N, ldw, outp, ldo, out_f32, bias, rowvec, res, ldr, act, e->splitk_ws,
                                e->splitk_bytes, c.bn, c.splits, c.occ, c.kbs, c.mode, ln);
        if (!ok) continue;
        float us = 0.f;
        int rc = time_gemm(e, ops, kTuneCopies, &us);
        if (rc) return rc;
        const long ctas = (long)ops[0].grid.x * ops[0].grid.y * ops[0].grid.z;
        const float util = std::min(1.0f, (float)ctas / (148.0f * (float)c.occ));
        const float cost = us * std::max(util, 1.0f / (float)std::max(e->autotune, 1));
        if (cost < best_cost) { best_cost = cost; best = Engine::Tuned{c.bn, c.splits, c.occ, c.kbs, c.mode, us}; }
    }
    g_tune_now[0] = 0;
    if (best.bn == 0) { set_error("autotune found no valid GEMM configuration"); return -1; }
    *result = best;
    return 0;
}

// ------------------------------------------------------------------------------------------------ plan builder
// "unet.down_blocks.0.resnets.1" -> "down_blocks0_resnets1" (one dot-free component for the profiler scopes)
static std::string short_name(const std::string& p) {
    std::string r;
    size_t start = p.find('.') == std::string::npos ? 0 : p.find('.') + 1;
    for (size_t i = start; i < p.size(); ++i) r += (p[i] == '.') ? '_' : p[i];
    return r;
}

struct Builder {
    Engine* e;
    std::vector<Launch>* out;
    int rc = 0;
    std::string fail;
    int cur_stream = 0;    // launches pushed now run on: 0 main stream, 1 / 2 side streams
    int nested = 0;        // begin_side() inside a branch: the inner work simply stays serial on that branch's stream
    size_t side_from = 0;
    int join_pending = 0;  // bitmask: the next launch pushed on the main stream waits for these side streams
    static bool branches_enabled() {
        static const bool on = !(getenv("VSD_BRANCHES") && atoi(getenv("VSD_BRANCHES")) == 0);
        return on;
    }
    void begin_side(int k = 1) {
        if (!branches_enabled()) return;
        if (cur_stream != 0) { ++nested; return; }
        cur_stream = k;
        side_from = out->size();
    }
    void end_side() {
        if (!branches_enabled()) return;
        if (nested > 0) { --nested; return; }
        for (size_t i = side_from; i < out->size(); ++i) (*out)[i].side = cur_stream;
        cur_stream = 0;
    }
    void join_next(int k = 1) {
        if (branches_enabled() && cur_stream == 0) join_pending |= 1 << (k - 1);
    }
    void mark_join(size_t first) {
        if (join_pending && cur_stream == 0 && out->size() > first) { (*out)[first].join |= join_pending; join_pending = 0; }
    }
    float* splitk_workspace() { return cur_stream == 0 ? e->splitk_ws : e->splitk_ws_side[cur_stream - 1]; }

    const bf16* wb(const std::string& n) {
        auto it = e->w.find(n);
        if (it == e->w.end() || !it->second.is_bf16) { miss(n); return nullptr; }
        return reinterpret_cast<const bf16*>(it->second.p);
    }
    const float* wf(const std::string& n) {
        auto it = e->w.find(n);
        if (it == e->w.end() || it->second.is_bf16) { miss(n); return nullptr; }
        return reinterpret_cast<const float*>(it->second.p);
    }
    bool has(const std::string& n) { return e->w.find(n) != e->w.end(); }
    void miss(const std::string& n) {
        if (!rc) { rc = -1; fail = "missing or mistyped weight: " + n; }
    }
    void bad(const std::string& m) {
        if (!rc) { rc = -1; fail = m; }
    }
    View alloc(int nb, int h, int w, int c) {
        View v;
        v.nb = nb; v.h = h; v.w = w; v.c = c; v.ld = c;
        v.p = reinterpret_cast<bf16*>(e->arena.alloc((size_t)nb * h * w * c * 2));
        if (!v.p) bad("activation arena exhausted");
        return v;
    }
    float* alloc_f32(size_t n) {
        float* p = reinterpret_cast<float*>(e->arena.alloc(n * 4));
        if (!p) bad("activation arena exhausted");
        return p;
    }

    // out = epilogue(conv/linear(a)); `a` may be any NHWC view, `o` any view with o.c == N (or N/2 for GEGLU)
    // returns the N tile count of the launch (the row-statistics layout a LayerNorm-folded consumer needs), 0 on failure
    int gemm(const ActView& a, int taps, const bf16* wt, int N, int ldw, void* outp, int ldo, int out_f32,
             const float* bias, const float* rowvec, const bf16* res, int ldr, int act, const float* out_scale = nullptr,
             const LnFuse* ln = nullptr) {
        if (rc) return 0;
        // part of the tuning key: other kernel paths / a restricted set of configurations
        if (ln && ln->mode) act |= (ln->mode == 1 ? ACT_LN_A_FLAG : ACT_LN_B_FLAG);
        if (ln && ln->stats_out) act |= ACT_ROWSTATS_FLAG;
        GemmOp op;
        int fbn = 0, fsp = 0, focc = 0, fkbs = 0, fhalo = 0;
        char key[160];
        if (e->autotune) {
            snprintf(key, sizeof(key), "%dx%dx%dx%d|t%d|n%d|a%d|f%d|r%d", a.NB, a.H, a.W, a.C, taps + 100 * (a.stride - 1) + 1000 * (1 - a.pad),
                     N, act, out_f32, res ? 1 : 0);
            auto it = e->tuned.find(key);
            if (it == e->tuned.end()) {
                Engine::Tuned t;
                int r = tune_gemm(e, a, taps, wt, N, ldw, outp, ldo, out_f32, bias, rowvec, res, ldr, act, &t, ln);
                if (r) { rc = r; fail = get_error(); return 0; }
                it = e->tuned.emplace(key, t).first;
                ++e->tune_misses;
            }
            fbn = it->second.bn; fsp = it->second.splits; focc = it->second.occ; fkbs = it->second.kbs; fhalo = it->second.halo;
        }
        int r = build_gemm_op(&op, a, taps, wt, N, ldw, outp, ldo, out_f32, bias, rowvec, res, ldr, act,
                              splitk_workspace(), e->splitk_bytes, fbn, fsp, focc, fkbs, fhalo, ln);
        if (r && e->autotune) {
            // a table entry written by another build of the kernels may no longer be a valid configuration: tune this shape now
            snprintf(key, sizeof(key), "%dx%dx%dx%d|t%d|n%d|a%d|f%d|r%d", a.NB, a.H, a.W, a.C, taps + 100 * (a.stride - 1) + 1000 * (1 - a.pad),
                     N, act, out_f32, res ? 1 : 0);
            Engine::Tuned t;
            r = tune_gemm(e, a, taps, wt, N, ldw, outp, ldo, out_f32, bias, rowvec, res, ldr, act, &t, ln);
            if (!r) {
                e->tuned[key] = t;
                ++e->tune_misses;
                r = build_gemm_op(&op, a, taps, wt, N, ldw, outp, ldo, out_f32, bias, rowvec, res, ldr, act, splitk_workspace(),
                                  e->splitk_bytes, t.bn, t.splits, t.occ, t.kbs, t.halo, ln);
            }
        }
        if (r) { rc = r; fail = get_error(); return 0; }
        op.p.out_scale = out_scale;
        {
            char what[160];
            snprintf(what, sizeof(what), ".gemm[%dx%dx%dx%d t%d N%d act%d bn%d sp%d pair%d ck%d halo%d persist%d smem%d tmem%d]", a.NB, a.H, a.W, a.C, taps, N, act,
                     op.p.block_n, op.p.splits, op.p.pair, op.p.cluster_k, op.p.halo, op.p.persist, op.smem_bytes, op.p.tmem_cols);
            op.p.trace = trace_slot(e, g_scope + what, (int)(op.grid.x * op.grid.y * op.grid.z));
        }
        const size_t first = out->size();
        out->push_back(mk([op](cudaStream_t st) { return launch_gemm_op(op, st); }, "gemm"));
        mark_join(first);
        return (int)op.grid.y;
    }
    void conv(const View& x, const std::string& name, int taps, const View& o, const float* rowvec, const View* res,
              int act, bool has_bias = true) {
        const bf16* wt = wb(name + ".weight");
        const float* b = has_bias ? wf(name + ".bias") : nullptr;
        if (rc) return;
        gemm(x.act(), taps, wt, o.c, taps * x.c, o.p, o.ld, 0, b, rowvec, res ? res->p : nullptr, res ? res->ld : 0, act);
    }
    void groupnorm(const View& x, const std::string& name, float eps, int silu, const View& o) {
        const float* g = wf(name + ".weight");
        const float* b = wf(name + ".bias");
        if (rc) return;
        float* ws = alloc_f32((size_t)groupnorm_ws_floats(x.nb, x.h * x.w, x.c, 32));
        if (rc) return;
        const View xi = x, oo = o;
        out->push_back(mk([=](cudaStream_t st) {
            return launch_groupnorm(xi.p, xi.ld, oo.p, oo.ld, g, b, xi.nb, xi.h * xi.w, xi.c, 32, eps, silu, ws, st);
        }, "gn"));
    }
    void layernorm(const View& x, const std::string& name, const View& o) {
        const float* g = wf(name + ".weight");
        const float* b = wf(name + ".bias");
        if (rc) return;
        const View xi = x, oo = o;
        out->push_back(mk([=](cudaStream_t st) {
            return launch_layernorm(xi.p, xi.ld, oo.p, oo.ld, g, b, (int)xi.rows(), xi.c, 1e-5f, st);
        }, "ln"));
    }
    void attention(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* vt, int ldvt, const View& o, int heads,
                   int d, int nq, int nk, int q_rows, int k_rows, int vt_cols, int vt_rows) {
        if (rc) return;
        AttnOp op;
        int r = build_attn_op(&op, q, ldq, k, ldk, vt, ldvt, o.p, o.ld, o.nb, heads, d, nq, nk, q_rows, k_rows, vt_cols,
                              vt_rows);
        if (r) { rc = r; fail = get_error(); return; }
        {
            char what[128];
            snprintf(what, sizeof(what), ".attn[b%d h%d d%d nq%d nk%d variant%d smem%d tmem%d]", o.nb, heads, d, nq, nk, op.variant, op.smem_bytes, op.tmem_cols);
            op.trace = trace_slot(e, g_scope + what, (int)(op.grid.x * op.grid.y * op.grid.z));
        }
        const size_t first = out->size();
        out->push_back(mk([op](cudaStream_t st) { return launch_attn_op(op, st); }, "attn"));
        mark_join(first);
    }

    // W' = W * gamma in place (once), wsum / wb derived vectors: see ln_fold_weight_kernel
    LnFold ln_fold(const std::string& wname, const std::string& norm, const float* bias) {
        LnFold f{nullptr, nullptr};
        auto it = e->w.find(wname);
        const float* g = wf(norm + ".weight");
        const float* b = wf(norm + ".bias");
        if (it == e->w.end() || !it->second.is_bf16) { miss(wname); return f; }
        if (rc) return f;
        std::lock_guard<std::mutex> lk(g_ln_mu);
        auto fi = g_ln_folded.find(it->second.p);
        if (fi != g_ln_folded.end()) return fi->second;
        const int N = (int)it->second.shape[0], K = (int)it->second.shape[1];
        if (cudaMalloc(&f.wsum, (size_t)N * 4) != cudaSuccess || cudaMalloc(&f.wb, (size_t)N * 4) != cudaSuccess ||
            launch_ln_fold_weight(reinterpret_cast<bf16*>(it->second.p), N, K, g, b, bias, f.wsum, f.wb, e->stream) ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) {
            bad("LayerNorm fold failed for " + wname);
            return LnFold{nullptr, nullptr};
        }
        g_ln_folded.emplace(it->second.p, f);
        return f;
    }

    // [Wp W2 | Wp], Wp b2 + bp for the transformer tail (computed once per weight set)
    ChainW chain_weights(const std::string& w2name, const std::string& b2name, const std::string& wpname, const std::string& bpname) {
        ChainW c{nullptr, nullptr};
        auto i2 = e->w.find(w2name);
        auto ip = e->w.find(wpname);
        const float* b2 = wf(b2name);
        const float* bp = wf(bpname);
        if (i2 == e->w.end() || !i2->second.is_bf16) { miss(w2name); return c; }
        if (ip == e->w.end() || !ip->second.is_bf16) { miss(wpname); return c; }
        if (rc) return c;
        std::lock_guard<std::mutex> lk(g_ln_mu);
        auto fi = g_chain.find(i2->second.p);
        if (fi != g_chain.end()) return fi->second;
        const int C = (int)i2->second.shape[0], K2 = (int)i2->second.shape[1];
        if (cudaMalloc(&c.wc, (size_t)C * (K2 + C) * 2) != cudaSuccess || cudaMalloc(&c.bc, (size_t)C * 4) != cudaSuccess ||
            launch_chain_weights(reinterpret_cast<const bf16*>(ip->second.p), reinterpret_cast<const bf16*>(i2->second.p), b2, bp, c.wc,
                                 c.bc, C, K2, e->stream) ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) {
            bad("transformer tail fold failed for " + w2name);
            return ChainW{nullptr, nullptr};
        }
        g_chain.emplace(i2->second.p, c);
        return c;
    }

    // ---- diffusers ResnetBlock2D (Appendix A.3)
    void resnet(const View& x, const std::string& p, const float* temb_rowvec, const View& o, float eps = 1e-5f) {
        Scope sc_(short_name(p));
        const size_t m = e->arena.mark();
        View sc;
        if (x.c != o.c) {   // the 1x1 shortcut only needs x: it runs beside norm1 / conv1 / norm2 on the side stream
            sc = alloc(x.nb, x.h, x.w, o.c);
            begin_side();
            conv(x, p + ".conv_shortcut", 1, sc, nullptr, nullptr, ACT_NONE);
            end_side();
        }
        View t1 = alloc(x.nb, x.h, x.w, x.c);
        groupnorm(x, p + ".norm1", eps, 1, t1);
        View h1 = alloc(x.nb, x.h, x.w, o.c);
        conv(t1, p + ".conv1", 9, h1, temb_rowvec, nullptr, ACT_NONE);
        View t2 = alloc(x.nb, x.h, x.w, o.c);
        groupnorm(h1, p + ".norm2", eps, 1, t2);
        if (x.c != o.c) {
            join_next();
            conv(t2, p + ".conv2", 9, o, nullptr, &sc, ACT_NONE);
        } else {
            conv(t2, p + ".conv2", 9, o, nullptr, &x, ACT_NONE);
        }
        e->arena.release(m);
    }

    // ---- diffusers Transformer2DModel with one BasicTransformerBlock (Appendix A.4)
    void transformer(const View& x, const std::string& p, const View& o) {
        Scope sc_(short_name(p));
        const size_t m = e->arena.mark();
        const int C = x.c, heads = 8, d = C / heads, dkp = attn_dk_pad(d);
        const int HW = x.h * x.w, NB = x.nb;
        const long M = x.rows();
        const std::string tb = p + ".transformer_blocks.0";
        View g = alloc(NB, x.h, x.w, C);
        groupnorm(x, p + ".norm", 1e-6f, 0, g);
        View h0 = alloc(NB, x.h, x.w, C);
        // The three LayerNorms are folded into the GEMMs that consume them (no kernel, no normalised copy of the stream). The
        // GEMM that PRODUCES the residual stream (proj_in, attn1.to_out, attn2.to_out) leaves per row and per N tile the sums
        // of x and x^2 of the values it stores (ACT_ROWSTATS_FLAG); the consumers read the raw stream and apply
        //   W LN(x) + b = rstd * (W' x - mean * wsum) + (W beta + b),   W' = W * gamma folded into the weights at plan build.
        // VSD_LN_FUSE=0 keeps the separate layernorm_kernel (decided once per process: the fold rewrites the weights in place).
        const bool fuse = ln_fuse_enabled();
        auto stats_buf = [&]() { return fuse ? alloc_f32((size_t)M * (C / 32) * 2) : nullptr; };   // [rows][<= C/32 tiles][2]
        float* st0 = stats_buf();
        float* st1 = stats_buf();
        float* st2 = stats_buf();
        const LnFuse p0{0, nullptr, nullptr, 0.f, nullptr, 0, st0}, p1{0, nullptr, nullptr, 0.f, nullptr, 0, st1},
            p2{0, nullptr, nullptr, 0.f, nullptr, 0, st2};
        const int nst0 = gemm(g.act(), 1, wb(p + ".proj_in.weight"), C, C, h0.p, h0.ld, 0, wf(p + ".proj_in.bias"), nullptr, nullptr, 0,
                              ACT_NONE, nullptr, fuse ? &p0 : nullptr);
        // self attention
        View n1 = h0;
        LnFold f_qk{nullptr, nullptr}, f_v{nullptr, nullptr}, f_q2{nullptr, nullptr}, f_ff{nullptr, nullptr};
        if (fuse) {
            f_qk = ln_fold(tb + ".attn1.qk.weight", tb + ".norm1", nullptr);
            f_v = ln_fold(tb + ".attn1.to_v.weight", tb + ".norm1", nullptr);
            f_q2 = ln_fold(tb + ".attn2.to_q.weight", tb + ".norm2", nullptr);
            f_ff = ln_fold(tb + ".ff.net.0.proj.weight", tb + ".norm3", wf(tb + ".ff.net.0.proj.bias"));
        } else {
            n1 = alloc(NB, x.h, x.w, C);
            layernorm(h0, tb + ".norm1", n1);
        }
        const LnFuse l_qk{1, f_qk.wsum, nullptr, 1e-5f, st0, nst0, nullptr};
        View qk = alloc(NB, x.h, x.w, 2 * heads * dkp);
        // V^T[C][tokens] = Wv * n1^T : weight as the row operand, activations as the column operand
        const int cols_img = (HW + 7) / 8 * 8;
        const int ldvt = NB * cols_img;
        bf16* vt = reinterpret_cast<bf16*>(e->arena.alloc((size_t)C * ldvt * 2));
        if (!vt) bad("activation arena exhausted");
        const bf16* wv = wb(tb + ".attn1.to_v.weight");
        if (!rc) {
            begin_side();   // the V^T projection only needs norm1's input: it overlaps the Q|K projection on the main stream
            ActView aw{wv, 1, 1, C, C, C};
            if (cols_img == HW) {
                const LnFuse l_v{2, f_v.wsum, f_v.wb, 1e-5f, st0, nst0, nullptr};
                gemm(aw, 1, n1.p, (int)M, n1.ld, vt, ldvt, 0, nullptr, nullptr, nullptr, 0, ACT_NONE | ACT_A_STATIC_FLAG, nullptr,
                     fuse ? &l_v : nullptr);
            } else {
                for (int b = 0; b < NB; ++b) {
                    const LnFuse l_v{2, f_v.wsum, f_v.wb, 1e-5f, fuse ? st0 + (size_t)b * HW * nst0 * 2 : nullptr, nst0, nullptr};
                    gemm(aw, 1, n1.p + (long)b * HW * n1.ld, HW, n1.ld, vt + (long)b * cols_img, ldvt, 0, nullptr, nullptr,
                         nullptr, 0, ACT_NONE | ACT_A_STATIC_FLAG, nullptr, fuse ? &l_v : nullptr);
                }
            }
            end_side();
        }
        gemm(n1.act_rows(), 1, wb(tb + ".attn1.qk.weight"), qk.c, C, qk.p, qk.ld, 0, fuse ? f_qk.wb : nullptr, nullptr, nullptr, 0, ACT_NONE,
             nullptr, fuse ? &l_qk : nullptr);
        join_next();
        View a1 = alloc(NB, x.h, x.w, C);
        attention(qk.p, qk.ld, qk.p + heads * dkp, qk.ld, vt, ldvt, a1, heads, d, HW, HW, HW, HW, cols_img, C);
        View h1 = alloc(NB, x.h, x.w, C);
        const int nst1 = gemm(a1.act_rows(), 1, wb(tb + ".attn1.to_out.0.weight"), C, C, h1.p, h1.ld, 0, wf(tb + ".attn1.to_out.0.bias"),
                              nullptr, h0.p, h0.ld, ACT_NONE, nullptr, fuse ? &p1 : nullptr);
        // cross attention against the cached context projections
        View n2 = h1;
        if (!fuse) {
            n2 = alloc(NB, x.h, x.w, C);
            layernorm(h1, tb + ".norm2", n2);
        }
        const LnFuse l_q2{1, f_q2.wsum, nullptr, 1e-5f, st1, nst1, nullptr};
        View q2 = alloc(NB, x.h, x.w, heads * dkp);
        gemm(n2.act_rows(), 1, wb(tb + ".attn2.to_q.weight"), q2.c, C, q2.p, q2.ld, 0, fuse ? f_q2.wb : nullptr, nullptr, nullptr, 0, ACT_NONE,
             nullptr, fuse ? &l_q2 : nullptr);
        const Engine::XAttn* xa = nullptr;
        for (auto& c : e->xattn)
            if (c.prefix == tb) xa = &c;
        if (!xa) bad("cross-attention cache missing for " + tb);
        View a2 = alloc(NB, x.h, x.w, C);
        if (!rc) attention(q2.p, q2.ld, xa->k2, heads * dkp, xa->v2t, NB * 128, a2, heads, d, HW, 77, HW, 128, 128, C);
        // ff.net.2 (+ h2) and proj_out (+ x) are one GEMM over K = [ff | h2] (chain_weights): ff and h2 live side by side
        const bool fuse_out = ff_out_fuse_enabled();
        View cat = fuse_out ? alloc(NB, x.h, x.w, 5 * C) : View();
        View h2 = fuse_out ? cat.slice(4 * C, C) : alloc(NB, x.h, x.w, C);
        const int nst2 = gemm(a2.act_rows(), 1, wb(tb + ".attn2.to_out.0.weight"), C, C, h2.p, h2.ld, 0, wf(tb + ".attn2.to_out.0.bias"),
                              nullptr, h1.p, h1.ld, ACT_NONE, nullptr, fuse ? &p2 : nullptr);
        // feed-forward (GEGLU fused into the first GEMM's epilogue)
        View n3 = h2;
        if (!fuse) {
            n3 = alloc(NB, x.h, x.w, C);
            layernorm(h2, tb + ".norm3", n3);
        }
        const LnFuse l_ff{1, f_ff.wsum, nullptr, 1e-5f, st2, nst2, nullptr};
        View ff = fuse_out ? cat.slice(0, 4 * C) : alloc(NB, x.h, x.w, 4 * C);
        gemm(n3.act_rows(), 1, wb(tb + ".ff.net.0.proj.weight"), 8 * C, C, ff.p, ff.ld, 0,
             fuse ? f_ff.wb : wf(tb + ".ff.net.0.proj.bias"), nullptr, nullptr, 0, ACT_GEGLU, nullptr, fuse ? &l_ff : nullptr);
        if (fuse_out) {
            const ChainW cw = chain_weights(tb + ".ff.net.2.weight", tb + ".ff.net.2.bias", p + ".proj_out.weight", p + ".proj_out.bias");
            gemm(cat.act(), 1, cw.wc, C, 5 * C, o.p, o.ld, 0, cw.bc, nullptr, x.p, x.ld, ACT_NONE);
            e->arena.release(m);
            return;
        }
        View h3 = alloc(NB, x.h, x.w, C);
        gemm(ff.act_rows(), 1, wb(tb + ".ff.net.2.weight"), C, 4 * C, h3.p, h3.ld, 0, wf(tb + ".ff.net.2.bias"), nullptr,
             h2.p, h2.ld, ACT_NONE);
        conv(h3, p + ".proj_out", 1, o, nullptr, &x, ACT_NONE);
        e->arena.release(m);
    }

    // 3x3 stride-2 convolution: the taps are fetched by TMA with element strides (2, 2) straight from the input (no im2col
    // buffer, one kernel less, 0.2 ms per 512 x 512 frame). pad 1 = diffusers Downsample2D(padding=1) / TAESD; pad 0 = zeros
    // right / below only (AutoencoderKL).
    // History: at the end of round 1 the 360 x 640 frame test produced NaNs with this path (UNet passes 1 and 2 only). In round
    // 2 that no longer reproduces: every autotuner candidate of the odd geometry (45x80 -> 23x40 -> 12x20 -> 6x10, 4 291
    // configurations, NaN-filled outputs) matches fp32 at operator level (tools/gpu_check.py s2_sweep) and three engine runs at
    // 360 x 640 are clean (profiles/r02_summary.md). Since the cause was never named, the strided-TMA path is the default only
    // for even input extents (every size the benchmarks and the reference UI's 64-pixel steps produce); odd extents keep
    // im2col_s2_kernel + GEMM. VSD_TMA_S2=1 / 0 forces it on / off everywhere.
    void conv_s2(const View& x, const std::string& name, const View& o, bool has_bias, int pad = 1) {
        static const int s2_mode = getenv("VSD_TMA_S2") ? (atoi(getenv("VSD_TMA_S2")) != 0 ? 1 : 0) : -1;
        const bool use_im2col = s2_mode == 0 || (s2_mode < 0 && ((x.h | x.w) & 1));
        const bf16* wt = wb(name + ".weight");
        const float* b = has_bias ? wf(name + ".bias") : nullptr;
        if (rc) return;
        if (!use_im2col) {
            ActView av = x.act();
            av.stride = 2;
            av.pad = pad;
            gemm(av, 9, wt, o.c, 9 * x.c, o.p, o.ld, 0, b, nullptr, nullptr, 0, ACT_NONE);
            return;
        }
        const size_t m = e->arena.mark();
        View cols = alloc(o.nb, o.h, o.w, 9 * x.c);
        if (rc) return;
        const View xi = x, ci = cols, oo = o;
        out->push_back(mk([=](cudaStream_t st) {
            return launch_im2col_s2(xi.p, xi.ld, ci.p, xi.nb, xi.h, xi.w, xi.c, oo.h, oo.w, st, pad);
        }, "im2col"));
        gemm(cols.act(), 1, wt, o.c, 9 * x.c, o.p, o.ld, 0, b, nullptr, nullptr, 0, ACT_NONE);
        e->arena.release(m);
    }
    void upsample(const View& x, const View& o) {
        if (rc) return;
        const View xi = x, oo = o;
        out->push_back(mk([=](cudaStream_t st) {
            return launch_upsample_nearest(xi.p, xi.ld, oo.p, oo.ld, xi.nb, xi.h, xi.w, oo.h, oo.w, xi.c, st);
        }, "upsample"));
    }
};

static int ds(int v) { return (v - 1) / 2 + 1; }  // conv k3 s2 p1 output size

// Allocates the persistent UNet tensors (skip/concat buffers) once; returns views through `S`.
struct UNetStatic {
    View concat[4][3];  // up block i, resnet j: [h | skip]
    View mid_in;        // output of the last down block resnet that is not a skip... (it IS a skip; see build)
    int hs[4], ws[4];
};

// Emits one UNet pass: eps(fp32 [NB,h8,w8,4]) = UNet(latents fp32, temb of step `si`).
static void build_unet(Builder& B, const float* latents, float* eps_out, int si, UNetStatic& S,
                       const std::function<void(const std::function<View(int)>&, const View&)>& after_mid = nullptr) {
    Engine* e = B.e;
    const int NB = e->NB;
    const int widths[4] = {320, 640, 1280, 1280};
    auto temb = [&](const std::string& resnet) -> const float* {
        auto it = e->temb.find(resnet);
        if (it == e->temb.end()) { B.bad("time embedding projection missing for " + resnet); return nullptr; }
        return it->second[si];
    };
    // skip k (push order) is consumed by up-resnet (11-k): up block i = (11-k)/3, resnet j = (11-k)%3, and lives in
    // the second half of that resnet's concat buffer. Channel split of concat[i][j]: [h : ch][skip : cs].
    const int up_out[4] = {1280, 1280, 640, 320};
    const int prev_out[4] = {1280, 1280, 1280, 640};  // width of h entering up block i
    auto skip_view = [&](int k) -> View {
        const int r = 11 - k, i = r / 3, j = r % 3;
        const int ch = (j == 0) ? prev_out[i] : up_out[i];
        const View& cb = S.concat[i][j];
        return cb.slice(ch, cb.c - ch);
    };
    auto h_view = [&](int i, int j) -> View {
        const int ch = (j == 0) ? prev_out[i] : up_out[i];
        return S.concat[i][j].slice(0, ch);
    };
    int k = 0;
    // conv_in -> skip 0
    {
        View s0 = skip_view(k++);
        const float* w = B.wf("unet.conv_in.weight");
        const float* b = B.wf("unet.conv_in.bias");
        if (!B.rc) {
            const View o = s0;
            B.out->push_back(mk([=](cudaStream_t st) {
                return launch_conv3x3_small_cin(latents, 0, o.nb, o.h, o.w, 4, w, b, o.p, o.ld, o.c, 0, st);
            }, "conv_in"));
        }
    }
    View x = skip_view(0);
    for (int i = 0; i < 4; ++i) {
        const std::string bp = "unet.down_blocks." + std::to_string(i);
        for (int j = 0; j < 2; ++j) {
            const std::string rp = bp + ".resnets." + std::to_string(j);
            if (i < 3) {
                const size_t m = e->arena.mark();
                View r = B.alloc(NB, x.h, x.w, widths[i]);
                B.resnet(x, rp, temb(rp), r);
                View s = skip_view(k++);
                B.transformer(r, bp + ".attentions." + std::to_string(j), s);
                x = s;
                e->arena.release(m);
            } else {
                View s = skip_view(k++);
                B.resnet(x, rp, temb(rp), s);
                x = s;
            }
        }
        if (i < 3) {
            View s = skip_view(k++);
            B.conv_s2(x, bp + ".downsamplers.0.conv", s, true);
            x = s;
        }
    }
    // mid block -> h of up block 0, resnet 0
    {
        const size_t m = e->arena.mark();
        View a = B.alloc(NB, x.h, x.w, 1280);
        B.resnet(x, "unet.mid_block.resnets.0", temb("unet.mid_block.resnets.0"), a);
        View t = B.alloc(NB, x.h, x.w, 1280);
        B.transformer(a, "unet.mid_block.attentions.0", t);
        B.resnet(t, "unet.mid_block.resnets.1", temb("unet.mid_block.resnets.1"), h_view(0, 0));
        e->arena.release(m);
    }
    if (after_mid) after_mid(skip_view, h_view(0, 0));   // ControlNet residuals are added to the 12 skips and the mid output here
    for (int i = 0; i < 4; ++i) {
        const std::string bp = "unet.up_blocks." + std::to_string(i);
        for (int j = 0; j < 3; ++j) {
            const std::string rp = bp + ".resnets." + std::to_string(j);
            const View& in = S.concat[i][j];
            // where does this layer's result go? next resnet's h slot, or (after the last resnet) the upsampler
            const size_t m = e->arena.mark();
            View dst;
            const bool last = (j == 2);
            if (!last) dst = h_view(i, j + 1);
            else dst = B.alloc(NB, in.h, in.w, up_out[i]);
            if (i > 0) {
                View r = B.alloc(NB, in.h, in.w, up_out[i]);
                B.resnet(in, rp, temb(rp), r);
                B.transformer(r, bp + ".attentions." + std::to_string(j), dst);
            } else {
                B.resnet(in, rp, temb(rp), dst);
            }
            if (last) {
                if (i < 3) {
                    const View& nxt = S.concat[i + 1][0];
                    View up = B.alloc(NB, nxt.h, nxt.w, up_out[i]);
                    B.upsample(dst, up);
                    B.conv(up, bp + ".upsamplers.0.conv", 9, h_view(i + 1, 0), nullptr, nullptr, ACT_NONE);
                } else {
                    View g = B.alloc(NB, in.h, in.w, 320);
                    B.groupnorm(dst, "unet.conv_norm_out", 1e-5f, 1, g);
                    B.gemm(g.act(), 9, B.wb("unet.conv_out.weight"), 4, 9 * 320, eps_out, 4, 1, B.wf("unet.conv_out.bias"),
                           nullptr, nullptr, 0, ACT_NONE);
                }
            }
            e->arena.release(m);
        }
    }
}

// TAESD Block: relu(conv(relu(conv(relu(conv(x))))) + x)
static void taesd_block(Builder& B, const View& x, const std::string& p, const View& o) {
    const size_t m = B.e->arena.mark();
    View a = B.alloc(x.nb, x.h, x.w, 64), b = B.alloc(x.nb, x.h, x.w, 64);
    B.conv(x, p + ".conv.0", 9, a, nullptr, nullptr, ACT_RELU_FLAG);
    B.conv(a, p + ".conv.2", 9, b, nullptr, nullptr, ACT_RELU_FLAG);
    B.conv(b, p + ".conv.4", 9, o, nullptr, &x, ACT_RELU_FLAG);
    B.e->arena.release(m);
}

static void build_taesd_encoder(Builder& B, const uint8_t* rgb, float* latents_out) {
    Engine* e = B.e;
    const int NB = e->NB;
    int h = e->H, w = e->W;
    const std::string p = "vae.encoder.layers.";
    View x = B.alloc(NB, h, w, 64);
    {
        const float* wt = B.wf(p + "0.weight");
        const float* b = B.wf(p + "0.bias");
        if (!B.rc) {
            const View o = x;
            B.out->push_back(mk([=](cudaStream_t st) {
                return launch_conv3x3_small_cin(rgb, 1, o.nb, o.h, o.w, 3, wt, b, o.p, o.ld, 64, 0, st);
            }, "conv_rgb"));
        }
    }
    View y = B.alloc(NB, h, w, 64);
    taesd_block(B, x, p + "1", y);
    x = y;
    int layer = 2;
    for (int s = 0; s < 3; ++s) {
        h = ds(h); w = ds(w);
        View d = B.alloc(NB, h, w, 64);
        B.conv_s2(x, p + std::to_string(layer++), d, false);
        x = d;
        for (int j = 0; j < 3; ++j) {
            View o = B.alloc(NB, h, w, 64);
            taesd_block(B, x, p + std::to_string(layer++), o);
            x = o;
        }
    }
    B.gemm(x.act(), 9, B.wb(p + "14.weight"), 4, 9 * 64, latents_out, 4, 1, B.wf(p + "14.bias"), nullptr, nullptr, 0, ACT_NONE);
}

static void build_taesd_decoder(Builder& B, const float* z, float* image_out /*[NB,H,W,4] fp32, 3 used*/) {
    Engine* e = B.e;
    const int NB = e->NB;
    int h = e->h8, w = e->w8;
    const std::string p = "vae.decoder.layers.";
    View x = B.alloc(NB, h, w, 64);
    {
        const float* wt = B.wf(p + "0.weight");
        const float* b = B.wf(p + "0.bias");
        if (!B.rc) {
            const View o = x;
            B.out->push_back(mk([=](cudaStream_t st) {
                return launch_conv3x3_small_cin(z, 2, o.nb, o.h, o.w, 4, wt, b, o.p, o.ld, 64, 1, st);
            }, "conv_z"));
        }
    }
    int layer = 2;
    for (int s = 0; s < 3; ++s) {
        for (int j = 0; j < 3; ++j) {
            View o = B.alloc(NB, h, w, 64);
            taesd_block(B, x, p + std::to_string(layer++), o);
            x = o;
        }
        layer++;  // nn.Upsample
        h *= 2; w *= 2;
        View up = B.alloc(NB, h, w, 64);
        B.upsample(x, up);
        View c = B.alloc(NB, h, w, 64);
        B.conv(up, p + std::to_string(layer++), 9, c, nullptr, nullptr, ACT_NONE, false);
        x = c;
    }
    View o = B.alloc(NB, h, w, 64);
    taesd_block(B, x, p + std::to_string(layer++), o);
    B.gemm(o.act(), 9, B.wb(p + std::to_string(layer) + ".weight"), 3, 9 * 64, image_out, 4, 1,
           B.wf(p + std::to_string(layer) + ".bias"), nullptr, nullptr, 0, ACT_NONE);
}

// ------------------------------------------------------------------------------------------------ AutoencoderKL
// diffusers AutoencoderKL of SD1.5: widths (128, 256, 512, 512), GroupNorm eps 1e-6, mid-block
// self-attention with ONE head of 512 channels. All 3x3 / 1x1 convolutions and the attention projections run on the tcgen05
// GEMM kernel; the 4096 x 4096 attention is two GEMMs around a row softmax (S = Q K^T in fp32, P = softmax(S / sqrt(512)) in
// bf16, O = P V); the V bias is folded into the out-projection bias (softmax rows sum to one).
static void kl_attention(Builder& B, const View& x, const std::string& p, const View& o) {
    Engine* e = B.e;
    const size_t m = e->arena.mark();
    const int C = x.c, HW = x.h * x.w, NB = x.nb;
    const int HWp = (HW + 63) / 64 * 64;   // key count padded to the GEMM's 64-wide k-blocks: P gets zero columns there
    View t = B.alloc(NB, x.h, x.w, C);
    B.groupnorm(x, p + ".group_norm", 1e-6f, 0, t);
    View qk = B.alloc(NB, x.h, x.w, 2 * C);            // q | k side by side
    B.gemm(t.act_rows(), 1, B.wb(p + ".to_q.weight"), C, C, qk.p, qk.ld, 0, B.wf(p + ".to_q.bias"), nullptr, nullptr, 0, ACT_NONE);
    B.gemm(t.act_rows(), 1, B.wb(p + ".to_k.weight"), C, C, qk.p + C, qk.ld, 0, B.wf(p + ".to_k.bias"), nullptr, nullptr, 0, ACT_NONE);
    // folded out-projection bias (computed once, at plan-build time)
    float* bfold = nullptr;
    {
        auto it = e->derived.find(p + ".folded_out_bias");
        if (it == e->derived.end()) {
            const bf16* wo = B.wb(p + ".to_out.0.weight");
            const float* bv = B.wf(p + ".to_v.bias");
            const float* bo = B.wf(p + ".to_out.0.bias");
            if (B.rc) return;
            if (cudaMalloc(&bfold, (size_t)C * 4) != cudaSuccess || launch_fold_v_bias(wo, bv, bo, bfold, C, e->stream) ||
                cudaStreamSynchronize(e->stream) != cudaSuccess) {
                B.bad("fold_v_bias failed");
                return;
            }
            e->derived.emplace(p + ".folded_out_bias", bfold);
        } else {
            bfold = it->second;
        }
    }
    bf16* vt = reinterpret_cast<bf16*>(e->arena.alloc((size_t)C * HWp * 2));                    // V^T [C][HWp], one image
    float* S = reinterpret_cast<float*>(e->arena.alloc((size_t)HW * HW * 4));
    bf16* P = reinterpret_cast<bf16*>(e->arena.alloc((size_t)HW * HWp * 2));
    View a = B.alloc(NB, x.h, x.w, C);
    if (!vt || !S || !P) { B.bad("activation arena exhausted"); return; }
    const bf16* wv = B.wb(p + ".to_v.weight");
    if (B.rc) return;
    for (int b = 0; b < NB; ++b) {
        const bf16* tb = t.p + (long)b * HW * t.ld;
        const bf16* qb = qk.p + (long)b * HW * qk.ld;
        if (HWp != HW)   // the padded key columns of V^T meet zero probabilities, but must be finite: clear them every frame
            B.out->push_back(mk([=](cudaStream_t st) {
                VSD_CHECK_CUDA(cudaMemset2DAsync(vt + HW, (size_t)HWp * 2, 0, (size_t)(HWp - HW) * 2, (size_t)C, st));
                return 0;
            }, "memset"));
        ActView aw{wv, 1, 1, C, C, C};
        B.gemm(aw, 1, tb, HW, t.ld, vt, HWp, 0, nullptr, nullptr, nullptr, 0, ACT_NONE | ACT_A_STATIC_FLAG);     // V^T = Wv t^T
        ActView aq{qb, 1, 1, HW, C, qk.ld};
        B.gemm(aq, 1, qb + C, HW, qk.ld, S, HW, 1, nullptr, nullptr, nullptr, 0, ACT_NONE | ACT_NO_STATIC_FLAG);    // S = Q K^T
        {
            const float scale = 1.0f / sqrtf((float)C);
            B.out->push_back(mk([=](cudaStream_t st) { return launch_softmax_rows(S, HW, P, HWp, HW, HW, scale, st); }, "softmax"));
        }
        ActView ap{P, 1, 1, HW, HWp, HWp};
        B.gemm(ap, 1, vt, C, HWp, a.p + (long)b * HW * a.ld, a.ld, 0, nullptr, nullptr, nullptr, 0, ACT_NONE | ACT_NO_STATIC_FLAG);  // O = P V
    }
    B.gemm(a.act_rows(), 1, B.wb(p + ".to_out.0.weight"), C, C, o.p, o.ld, 0, bfold, nullptr, x.p, x.ld, ACT_NONE);   // + residual
    e->arena.release(m);
}

static void kl_mid(Builder& B, const View& x, const std::string& p, const View& o) {
    const size_t m = B.e->arena.mark();
    View a = B.alloc(x.nb, x.h, x.w, x.c), b = B.alloc(x.nb, x.h, x.w, x.c);
    B.resnet(x, p + ".resnets.0", nullptr, a, 1e-6f);
    kl_attention(B, a, p + ".attentions.0", b);
    B.resnet(b, p + ".resnets.1", nullptr, o, 1e-6f);
    B.e->arena.release(m);
}

// rgb u8 -> latents (fp32 [NB][h8][w8][4]) = latent_dist.sample(noise) * scaling_factor   (lcm_controlnet.py:298-313)
static void build_kl_encoder(Builder& B, const uint8_t* rgb, float* latents_out) {
    Engine* e = B.e;
    const int NB = e->NB;
    int h = e->H, w = e->W;
    const std::string p = "vae_kl.encoder.";
    const int widths[4] = {128, 256, 512, 512};
    View x = B.alloc(NB, h, w, 128);
    {
        const float* wt = B.wf(p + "conv_in.weight");
        const float* b = B.wf(p + "conv_in.bias");
        if (!B.rc) {
            const View o = x;
            B.out->push_back(mk([=](cudaStream_t st) {
                return launch_conv3x3_small_cin(rgb, 3, o.nb, o.h, o.w, 3, wt, b, o.p, o.ld, 128, 0, st);
            }, "conv_rgb"));
        }
    }
    for (int i = 0; i < 4; ++i) {
        const std::string bp = p + "down_blocks." + std::to_string(i);
        for (int j = 0; j < 2; ++j) {
            View o = B.alloc(NB, h, w, widths[i]);
            B.resnet(x, bp + ".resnets." + std::to_string(j), nullptr, o, 1e-6f);
            x = o;
        }
        if (i < 3) {
            h /= 2; w /= 2;
            View d = B.alloc(NB, h, w, widths[i]);
            B.conv_s2(x, bp + ".downsamplers.0.conv", d, true, 0);   // F.pad(0,1,0,1) + conv stride 2 pad 0
            x = d;
        }
    }
    View mo = B.alloc(NB, h, w, 512);
    kl_mid(B, x, p + "mid_block", mo);
    View n = B.alloc(NB, h, w, 512);
    B.groupnorm(mo, p + "conv_norm_out", 1e-6f, 1, n);
    const long lpx = (long)NB * h * w;
    float* enc8 = B.alloc_f32((size_t)lpx * 8);
    B.gemm(n.act(), 9, B.wb(p + "conv_out.weight"), 8, 9 * 512, enc8, 8, 1, B.wf(p + "conv_out.bias"), nullptr, nullptr, 0, ACT_NONE);
    const float* wq = B.wf("vae_kl.quant_conv.weight");
    const float* bq = B.wf("vae_kl.quant_conv.bias");
    if (B.rc) return;
    const float* noise = e->vae_noise;
    B.out->push_back(mk([=](cudaStream_t st) { return launch_kl_sample(enc8, wq, bq, noise, latents_out, lpx, 0.18215f, st); },
                        "kl_sample"));
}

// latents (as the scheduler leaves them) -> image fp32 [NB][H][W][4] (3 used), already in [-1, 1]   (lcm_controlnet.py:594-596)
static void build_kl_decoder(Builder& B, const float* z, float* image_out) {
    Engine* e = B.e;
    const int NB = e->NB;
    int h = e->h8, w = e->w8;
    const std::string p = "vae_kl.decoder.";
    const long lpx = (long)NB * h * w;
    float* zq = B.alloc_f32((size_t)lpx * 4);
    {
        const float* wp = B.wf("vae_kl.post_quant_conv.weight");
        const float* bp = B.wf("vae_kl.post_quant_conv.bias");
        if (B.rc) return;
        B.out->push_back(mk([=](cudaStream_t st) { return launch_kl_post_quant(z, wp, bp, zq, lpx, 1.0f / 0.18215f, st); },
                            "post_quant"));
    }
    View x = B.alloc(NB, h, w, 512);
    {
        const float* wt = B.wf(p + "conv_in.weight");
        const float* b = B.wf(p + "conv_in.bias");
        if (!B.rc) {
            const View o = x;
            B.out->push_back(mk([=](cudaStream_t st) {
                return launch_conv3x3_small_cin(zq, 0, o.nb, o.h, o.w, 4, wt, b, o.p, o.ld, 512, 0, st);
            }, "conv_z"));
        }
    }
    View mo = B.alloc(NB, h, w, 512);
    kl_mid(B, x, p + "mid_block", mo);
    x = mo;
    const int rev[4] = {512, 512, 256, 128};
    for (int i = 0; i < 4; ++i) {
        const std::string bp = p + "up_blocks." + std::to_string(i);
        for (int j = 0; j < 3; ++j) {
            View o = B.alloc(NB, h, w, rev[i]);
            B.resnet(x, bp + ".resnets." + std::to_string(j), nullptr, o, 1e-6f);
            x = o;
        }
        if (i < 3) {
            h *= 2; w *= 2;
            View up = B.alloc(NB, h, w, rev[i]);
            B.upsample(x, up);
            View c = B.alloc(NB, h, w, rev[i]);
            B.conv(up, bp + ".upsamplers.0.conv", 9, c, nullptr, nullptr, ACT_NONE);
            x = c;
        }
    }
    View n = B.alloc(NB, h, w, 128);
    B.groupnorm(x, p + "conv_norm_out", 1e-6f, 1, n);
    B.gemm(n.act(), 9, B.wb(p + "conv_out.weight"), 3, 9 * 128, image_out, 4, 1, B.wf(p + "conv_out.bias"), nullptr, nullptr, 0,
           ACT_NONE);
}

// ------------------------------------------------------------------------------------------------ ControlNet
// diffusers ControlNetModel, called by the reference before every UNet pass
// (lcm_controlnet.py:558-566) on the Sobel edge map of the input frame (videopipeline.py:109, lcm/canny_gpu.py).
struct CNStatic {
    View feat[12];   // conv_in(+cond) output and the 11 down-block outputs (inputs of the 1x1 "zero" convolutions)
    View mid;        // mid-block output
    View cond;       // conditioning embedding [NB, h8, w8, 320], computed once per frame
    float* control = nullptr;        // [NB, H, W, 3] fp32 control image in [0, 1]
    float* mag = nullptr;            // [NB, H, W] Sobel magnitude
    unsigned int* maxbits = nullptr; // [NB] per-image maximum (float bits)
};

// Once per frame: Sobel edge map of the RGB frame -> ControlNetConditioningEmbedding (3->16->16->32->32->96->96->256->320).
static void build_control_frontend(Builder& B, const uint8_t* rgb, CNStatic& C) {
    Engine* e = B.e;
    const int NB = e->NB, H = e->H, W = e->W;
    {
        const CNStatic c = C;
        B.out->push_back(mk([=](cudaStream_t st) {
            return launch_sobel_control(rgb, c.mag, c.maxbits, c.control, NB, H, W, 0.11f, 0.8f, st);
        }, "sobel"));
    }
    const std::string p = "controlnet.controlnet_cond_embedding.";
    const size_t m = e->arena.mark();
    View x = B.alloc(NB, H, W, 16);
    {
        const float* wt = B.wf(p + "conv_in.weight");
        const float* b = B.wf(p + "conv_in.bias");
        if (!B.rc) {
            const View o = x;
            const float* ctl = C.control;
            B.out->push_back(mk([=](cudaStream_t st) {
                return launch_conv3x3_small_cin(ctl, 0, o.nb, o.h, o.w, 3, wt, b, o.p, o.ld, 16, 2 /*SiLU*/, st);
            }, "cond_conv_in"));
        }
    }
    const int widths[4] = {16, 32, 96, 256};
    int h = H, w = W;
    for (int i = 0; i < 3; ++i) {
        for (int half = 0; half < 2; ++half) {
            const int cin = widths[i], cout = half ? widths[i + 1] : widths[i], stride = half ? 2 : 1;
            const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
            View y = B.alloc(NB, ho, wo, cout);
            const std::string name = p + "blocks." + std::to_string(i * 2 + half);
            const bf16* wt = B.wb(name + ".weight");
            const float* b = B.wf(name + ".bias");
            if (!B.rc) {
                const View xi = x, yo = y;
                B.out->push_back(mk([=](cudaStream_t st) {
                    return launch_conv3x3_direct(xi.p, xi.ld, xi.nb, xi.h, xi.w, xi.c, wt, b, yo.p, yo.ld, yo.c, stride, 1, st);
                }, "cond_conv"));
            }
            x = y; h = ho; w = wo;
        }
    }
    if (!B.rc && (h != e->h8 || w != e->w8)) B.bad("conditioning embedding does not reach the latent resolution");
    B.conv(x, p + "conv_out", 9, C.cond, nullptr, nullptr, ACT_NONE);
    e->arena.release(m);
}

// One ControlNet pass at schedule step `si`: fills C.feat[0..11] and C.mid (pre zero-conv features).
static void build_controlnet(Builder& B, const float* latents, int si, CNStatic& C) {
    Engine* e = B.e;
    const int widths[4] = {320, 640, 1280, 1280};
    auto temb = [&](const std::string& resnet) -> const float* {
        auto it = e->temb.find(resnet);
        if (it == e->temb.end()) { B.bad("time embedding projection missing for " + resnet); return nullptr; }
        return it->second[si];
    };
    int k = 0;
    {   // conv_in(latents) + conditioning embedding
        const float* w = B.wf("controlnet.conv_in.weight");
        const float* b = B.wf("controlnet.conv_in.bias");
        if (!B.rc) {
            const View o = C.feat[0], cd = C.cond;
            B.out->push_back(mk([=](cudaStream_t st) {
                return launch_conv3x3_small_cin(latents, 0, o.nb, o.h, o.w, 4, w, b, o.p, o.ld, o.c, 0, st, cd.p, cd.ld);
            }, "conv_in"));
        }
        k = 1;
    }
    View x = C.feat[0];
    for (int i = 0; i < 4; ++i) {
        const std::string bp = "controlnet.down_blocks." + std::to_string(i);
        for (int j = 0; j < 2; ++j) {
            const std::string rp = bp + ".resnets." + std::to_string(j);
            if (i < 3) {
                const size_t m = e->arena.mark();
                View r = B.alloc(x.nb, x.h, x.w, widths[i]);
                B.resnet(x, rp, temb(rp), r);
                B.transformer(r, bp + ".attentions." + std::to_string(j), C.feat[k]);
                e->arena.release(m);
            } else {
                B.resnet(x, rp, temb(rp), C.feat[k]);
            }
            x = C.feat[k++];
        }
        if (i < 3) {
            B.conv_s2(x, bp + ".downsamplers.0.conv", C.feat[k], true);
            x = C.feat[k++];
        }
    }
    const size_t m = e->arena.mark();
    View a = B.alloc(x.nb, x.h, x.w, 1280);
    B.resnet(x, "controlnet.mid_block.resnets.0", temb("controlnet.mid_block.resnets.0"), a);
    View t = B.alloc(x.nb, x.h, x.w, 1280);
    B.transformer(a, "controlnet.mid_block.attentions.0", t);
    B.resnet(t, "controlnet.mid_block.resnets.1", temb("controlnet.mid_block.resnets.1"), C.mid);
    e->arena.release(m);
}

// skip_k += scale_k * (zero_conv_k(feat_k) + b_k), in place (lcm_controlnet.py:568-577: down_block_additional_residuals)
static void build_controlnet_residuals(Builder& B, CNStatic& C, const std::function<View(int)>& skip_view, const View& mid_out) {
    Engine* e = B.e;
    Scope sc_("cn_residuals");
    for (int k = 0; k < 12; ++k) {
        const std::string n = "controlnet.controlnet_down_blocks." + std::to_string(k);
        const View dst = skip_view(k);
        B.gemm(C.feat[k].act(), 1, B.wb(n + ".weight"), dst.c, C.feat[k].c, dst.p, dst.ld, 0, B.wf(n + ".bias"), nullptr, dst.p, dst.ld,
               ACT_NONE, e->cn_scales + k);
    }
    B.gemm(C.mid.act(), 1, B.wb("controlnet.controlnet_mid_block.weight"), mid_out.c, C.mid.c, mid_out.p, mid_out.ld, 0,
           B.wf("controlnet.controlnet_mid_block.bias"), nullptr, mid_out.p, mid_out.ld, ACT_NONE, e->cn_scales + 12);
}

static int run_plan(Engine* e, const std::vector<Launch>& plan, cudaStream_t st) {
    bool busy[2] = {false, false};
    auto join = [&](int k) -> int {
        VSD_CHECK_CUDA(cudaEventRecord(e->ev_join[k], e->side[k]));
        VSD_CHECK_CUDA(cudaStreamWaitEvent(st, e->ev_join[k], 0));
        busy[k] = false;
        return 0;
    };
    for (const auto& l : plan) {
        if (l.side && e->side[l.side - 1]) {
            const int k = l.side - 1;
            VSD_CHECK_CUDA(cudaEventRecord(e->ev_fork[k], st));          // fork: everything issued so far on the main stream
            VSD_CHECK_CUDA(cudaStreamWaitEvent(e->side[k], e->ev_fork[k], 0));
            int rc = l(e->side[k]);
            if (rc) return rc;
            busy[k] = true;
            continue;
        }
        for (int k = 0; k < 2; ++k)
            if ((l.join >> k & 1) && busy[k]) {
                int rc = join(k);
                if (rc) return rc;
            }
        int rc = l(st);
        if (rc) return rc;
    }
    for (int k = 0; k < 2; ++k)   // never leave a side stream dangling (stream capture requires the join)
        if (busy[k]) {
            int rc = join(k);
            if (rc) return rc;
        }
    return 0;
}

static void free_graphs(Engine* e) {
    if (e->graph_yuv) cudaGraphExecDestroy(e->graph_yuv);
    if (e->graph_rgb) cudaGraphExecDestroy(e->graph_rgb);
    e->graph_yuv = e->graph_rgb = nullptr;
}

static const char* kResnetPrefixes[22] = {
    "unet.down_blocks.0.resnets.0", "unet.down_blocks.0.resnets.1", "unet.down_blocks.1.resnets.0",
    "unet.down_blocks.1.resnets.1", "unet.down_blocks.2.resnets.0", "unet.down_blocks.2.resnets.1",
    "unet.down_blocks.3.resnets.0", "unet.down_blocks.3.resnets.1", "unet.mid_block.resnets.0",
    "unet.mid_block.resnets.1", "unet.up_blocks.0.resnets.0", "unet.up_blocks.0.resnets.1",
    "unet.up_blocks.0.resnets.2", "unet.up_blocks.1.resnets.0", "unet.up_blocks.1.resnets.1",
    "unet.up_blocks.1.resnets.2", "unet.up_blocks.2.resnets.0", "unet.up_blocks.2.resnets.1",
    "unet.up_blocks.2.resnets.2", "unet.up_blocks.3.resnets.0", "unet.up_blocks.3.resnets.1",
    "unet.up_blocks.3.resnets.2"};

static std::vector<std::string> transformer_prefixes(bool with_controlnet) {
    std::vector<std::string> v;
    for (int m = 0; m < (with_controlnet ? 2 : 1); ++m) {
        const std::string root = m == 0 ? "unet." : "controlnet.";
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j)
                v.push_back(root + "down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j) + ".transformer_blocks.0");
        v.push_back(root + "mid_block.attentions.0.transformer_blocks.0");
        if (m == 0)
            for (int i = 1; i < 4; ++i)
                for (int j = 0; j < 3; ++j)
                    v.push_back(root + "up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j) + ".transformer_blocks.0");
    }
    return v;
}
static bool has_controlnet_weights(const Engine* e) { return e->w.find("controlnet.conv_in.weight") != e->w.end(); }

// ------------------------------------------------------------------------------------------------ CLIP text encoder
// transformers CLIPTextModel, SD1.5 text tower (12 layers, 768 wide, 12 heads, quick-GELU MLP 3072, causal mask, final LN);
// the reference runs it in _encode_prompt (lcm_controlnet.py:175-179). Weights: "text_encoder." + transformers keys.
// 77 tokens = one 128-row tile: every linear is the tcgen05 GEMM (bias / quick-GELU / residual fused), attention is a small
// CUDA-core kernel (12 heads x 77 x 77), 86 launches per prompt.
static void free_clip(Engine* e) {
    void* ptrs[7] = {e->clip.ids, e->clip.x, e->clip.xn, e->clip.qkv, e->clip.att, e->clip.h, e->clip.out};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    e->clip = Engine::Clip();
}

static int build_clip(Engine* e) {
    constexpr int T = 77, C = 768, L = 12, HEADS = 12, FF = 3072, VOCAB = 49408;
    Engine::Clip& c = e->clip;
    if (c.built) return 0;
    const std::string root = "text_encoder.text_model.";
    ENG_REQUIRE(e->w.count(root + "embeddings.token_embedding.weight"), "text encoder weights are not loaded");
    VSD_CHECK_CUDA(cudaMalloc(&c.ids, 128 * sizeof(int)));
    auto ab = [&](bf16** p, size_t n) -> int { VSD_CHECK_CUDA(cudaMalloc(p, n * 2)); VSD_CHECK_CUDA(cudaMemset(*p, 0, n * 2)); return 0; };
    VSD_CHECK_CUDA(cudaMalloc(&c.x, (size_t)128 * C * 4));
    VSD_CHECK_CUDA(cudaMemset(c.x, 0, (size_t)128 * C * 4));
    if (ab(&c.xn, (size_t)128 * C) || ab(&c.qkv, (size_t)128 * 3 * C) || ab(&c.att, (size_t)128 * C) ||
        ab(&c.h, (size_t)128 * FF) || ab(&c.out, (size_t)128 * C))
        return -2;
    const int saved_autotune = e->autotune;
    e->autotune = 0;                       // once per prompt change: the shape heuristics are plenty
    Builder B{e, &c.plan};
    Scope sc_("clip");
    {
        const bf16* tok = B.wb(root + "embeddings.token_embedding.weight");
        const bf16* pos = B.wb(root + "embeddings.position_embedding.weight");
        const int* ids = c.ids;
        float* x = c.x;
        if (!B.rc)
            c.plan.push_back(mk([=](cudaStream_t st) { return launch_clip_embed(ids, tok, pos, x, T, C, VOCAB, st); }, "embed"));
    }
    const ActView axn{c.xn, 1, 1, T, C, C}, aatt{c.att, 1, 1, T, C, C}, ah{c.h, 1, 1, T, FF, FF};
    auto ln = [&](const float* in, bf16* out, const std::string& name) {     // the residual stream x stays fp32
        const float* g = B.wf(name + ".weight");
        const float* b = B.wf(name + ".bias");
        if (B.rc) return;
        c.plan.push_back(mk([=](cudaStream_t st) { return launch_layernorm_f32in(in, out, g, b, T, C, 1e-5f, st); }, "ln"));
    };
    for (int l = 0; l < L && !B.rc; ++l) {
        const std::string p = root + "encoder.layers." + std::to_string(l) + ".";
        ln(c.x, c.xn, p + "layer_norm1");
        const char* names[3] = {"q_proj", "k_proj", "v_proj"};
        for (int j = 0; j < 3; ++j)      // q | k | v side by side in one [T][2304] buffer
            B.gemm(axn, 1, B.wb(p + "self_attn." + names[j] + ".weight"), C, C, c.qkv + j * C, 3 * C, 0,
                   B.wf(p + "self_attn." + names[j] + ".bias"), nullptr, nullptr, 0, ACT_NONE);
        {
            const bf16* qkv = c.qkv;
            bf16* att = c.att;
            c.plan.push_back(mk([=](cudaStream_t st) { return launch_clip_attention(qkv, att, T, HEADS, st); }, "attn"));
        }
        B.gemm(aatt, 1, B.wb(p + "self_attn.out_proj.weight"), C, C, c.x, C, 1, B.wf(p + "self_attn.out_proj.bias"), nullptr,
               reinterpret_cast<const bf16*>(c.x), C, ACT_NONE | ACT_RES_F32_FLAG);                     // x += out_proj(att), fp32
        ln(c.x, c.xn, p + "layer_norm2");
        B.gemm(axn, 1, B.wb(p + "mlp.fc1.weight"), FF, C, c.h, FF, 0, B.wf(p + "mlp.fc1.bias"), nullptr, nullptr, 0, ACT_QUICK_GELU);
        B.gemm(ah, 1, B.wb(p + "mlp.fc2.weight"), C, FF, c.x, C, 1, B.wf(p + "mlp.fc2.bias"), nullptr,
               reinterpret_cast<const bf16*>(c.x), C, ACT_NONE | ACT_RES_F32_FLAG);                     // x += mlp(xn), fp32
    }
    ln(c.x, c.out, root + "final_layer_norm");
    e->autotune = saved_autotune;
    if (B.rc) {
        set_error("text encoder plan failed: " + B.fail);
        free_clip(e);
        return B.rc;
    }
    c.built = true;
    return 0;
}

// ids: 77 token ids (host). out: fp32 [77][768] last_hidden_state (host).
static int encode_prompt(Engine* e, const int* ids_host, float* out_host) {
    int rc = build_clip(e);
    if (rc) return rc;
    Engine::Clip& c = e->clip;
    VSD_CHECK_CUDA(cudaMemcpyAsync(c.ids, ids_host, 77 * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    rc = run_plan(e, c.plan, e->stream);
    if (rc) return rc;
    std::vector<uint16_t> t((size_t)77 * 768);
    VSD_CHECK_CUDA(cudaMemcpyAsync(t.data(), c.out, t.size() * 2, cudaMemcpyDeviceToHost, e->stream));
    VSD_CHECK_CUDA(sync_engine(e));
    for (size_t i = 0; i < t.size(); ++i) {
        const uint32_t u = (uint32_t)t[i] << 16;
        memcpy(&out_host[i], &u, 4);
    }
    return vsd_check_pipeline_fault();
}

static void free_resize(Engine* e) {
    void* ptrs[7] = {e->rz.hb, e->rz.hk, e->rz.vb, e->rz.vk, e->rz.src, e->rz.tmp, e->rz.yuv};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    e->rz = Engine::Resize();
}

// (Re)allocates every buffer for a (batch, height, width) configuration. Plans are built by finalize().
static int configure(Engine* e, int nb, int H, int W) {
    ENG_REQUIRE(nb >= 1 && nb <= 64, "batch must be in [1, 64]");
    ENG_REQUIRE(H % 8 == 0 && W % 8 == 0 && H >= 16 && W >= 16, "height and width must be multiples of 8 (>= 16)");
    VSD_CHECK_CUDA(sync_engine(e));
    free_graphs(e);
    free_resize(e);
    e->arena.destroy();
    if (e->splitk_ws) cudaFree(e->splitk_ws);
    for (int k = 0; k < 2; ++k) {
        if (e->splitk_ws_side[k]) cudaFree(e->splitk_ws_side[k]);
        e->splitk_ws_side[k] = nullptr;
    }
    e->splitk_ws = nullptr;
    e->NB = nb; e->H = H; e->W = W; e->h8 = H / 8; e->w8 = W / 8;
    e->configured = false; e->schedule_set = false; e->context_set = false;
    e->xattn.clear(); e->temb.clear(); e->eps.clear(); e->lat.clear(); e->den.clear();
    const double scale = (double)nb * H * W / (512.0 * 512.0);
    const size_t bytes = (size_t)(640.0 * 1048576.0 * scale) + (size_t)384 * 1048576;
    int rc = e->arena.init(bytes);
    if (rc) return rc;
    e->splitk_bytes = (size_t)(16.0 * 1048576.0 * (scale < 1 ? 1 : scale)) + (size_t)48 * 1048576;
    VSD_CHECK_CUDA(cudaMalloc(&e->splitk_ws, e->splitk_bytes));
    for (int k = 0; k < 2; ++k) VSD_CHECK_CUDA(cudaMalloc(&e->splitk_ws_side[k], e->splitk_bytes));
    Arena& A = e->arena;
    const size_t px = (size_t)nb * H * W, lpx = (size_t)nb * e->h8 * e->w8;
    e->d_y = (uint8_t*)A.alloc(px); e->d_u = (uint8_t*)A.alloc(px / 4); e->d_v = (uint8_t*)A.alloc(px / 4);
    e->d_rgb_in = (uint8_t*)A.alloc(px * 3);
    e->d_oy = (uint8_t*)A.alloc(px); e->d_ou = (uint8_t*)A.alloc(px / 4); e->d_ov = (uint8_t*)A.alloc(px / 4);
    e->d_rgb_out = (uint8_t*)A.alloc(px * 3);
    e->init_latents = (float*)A.alloc(lpx * 16); e->noisy = (float*)A.alloc(lpx * 16);
    e->init_noise = (float*)A.alloc(lpx * 16);
    e->vae_noise = (float*)A.alloc(lpx * 16);   // zero-initialised arena: sample() = mean until vsd_set_vae_noise
    e->step_noise = (float*)A.alloc(lpx * 16 * kMaxSteps);  // LCM_ORIGIN_STEPS = 50 bounds the timestep table (lcm_controlnet.py:905-938)
    e->image = (float*)A.alloc(px * 16);
    e->ctx_bf16 = (bf16*)A.alloc((size_t)nb * 128 * 768 * 2);
    e->t_emb_in = (float*)A.alloc(320 * 4); e->t_h = (float*)A.alloc(1280 * 4); e->t_emb = (float*)A.alloc(1280 * 4);
    ENG_REQUIRE(e->t_emb != nullptr, "arena too small for the static buffers");
    // cross-attention caches
    for (const auto& tb : transformer_prefixes(has_controlnet_weights(e))) {
        auto it = e->w.find(tb + ".attn2.to_q.weight");
        ENG_REQUIRE(it != e->w.end(), "load the UNet weights before configure(): " + tb);
        Engine::XAttn xa;
        xa.prefix = tb;
        xa.C = (int)it->second.shape[1];
        xa.d = xa.C / 8;
        xa.dkp = attn_dk_pad(xa.d);
        xa.k2 = (bf16*)A.alloc((size_t)nb * 128 * 8 * xa.dkp * 2);
        xa.v2t = (bf16*)A.alloc((size_t)xa.C * nb * 128 * 2);
        ENG_REQUIRE(xa.v2t != nullptr, "arena too small for the context caches");
        e->xattn.push_back(xa);
    }
    e->arena_static_mark = A.mark();
    e->configured = true;
    return 0;
}

// Time-embedding path (Appendix A.2 steps 1-2 and the 22 per-resnet projections), once per schedule.
static int set_schedule(Engine* e, int steps, const int* timesteps, const float* scalars /*steps x 6*/, float an_a,
                        float an_b, const float* w_emb256, int has_step_noise) {
    ENG_REQUIRE(e->configured, "configure() first");
    ENG_REQUIRE(steps >= 1 && steps <= kMaxSteps, "1..50 steps");
    VSD_CHECK_CUDA(sync_engine(e));
    free_graphs(e);
    e->arena.release(e->arena_static_mark);
    e->temb.clear(); e->eps.clear(); e->lat.clear(); e->den.clear();
    e->steps = steps; e->an_a = an_a; e->an_b = an_b; e->has_step_noise = has_step_noise;
    e->sc.resize(steps);
    for (int i = 0; i < steps; ++i) {
        const float* s = scalars + i * 6;
        e->sc[i] = StepScalars{s[0], s[1], s[2], s[3], s[4], s[5], timesteps[i]};
    }
    e->w_emb.assign(w_emb256, w_emb256 + 256);
    Arena& A = e->arena;
    const size_t lpx = (size_t)e->NB * e->h8 * e->w8;
    for (int i = 0; i < steps; ++i) {
        e->eps.push_back((float*)A.alloc(lpx * 16));
        e->lat.push_back((float*)A.alloc(lpx * 16));
        e->den.push_back((float*)A.alloc(lpx * 16));
    }
    float* d_wemb = (float*)A.alloc(256 * 4);
    float* d_sin = (float*)A.alloc(320 * 4);
    ENG_REQUIRE(d_sin != nullptr, "arena exhausted");
    VSD_CHECK_CUDA(cudaMemcpyAsync(d_wemb, w_emb256, 256 * 4, cudaMemcpyHostToDevice, e->stream));
    auto need_f = [&](const std::string& n) -> const float* {
        auto it = e->w.find(n);
        return (it == e->w.end() || it->second.is_bf16) ? nullptr : (const float*)it->second.p;
    };
    const float* w_cond = need_f("unet.time_embedding.cond_proj.weight");
    const float* w1 = need_f("unet.time_embedding.linear_1.weight");
    const float* b1 = need_f("unet.time_embedding.linear_1.bias");
    const float* w2 = need_f("unet.time_embedding.linear_2.weight");
    const float* b2 = need_f("unet.time_embedding.linear_2.bias");
    ENG_REQUIRE(w_cond && w1 && b1 && w2 && b2, "time_embedding weights missing");
    for (int r = 0; r < 22; ++r) e->temb[kResnetPrefixes[r]] = std::vector<float*>();
    const bool use_cn = e->cn_enabled;
    if (use_cn) ENG_REQUIRE(has_controlnet_weights(e), "ControlNet enabled but no controlnet.* weights are loaded");
    std::vector<std::string> cn_resnets;
    if (use_cn) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 2; ++j) cn_resnets.push_back("controlnet.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j));
        cn_resnets.push_back("controlnet.mid_block.resnets.0");
        cn_resnets.push_back("controlnet.mid_block.resnets.1");
        for (auto& n : cn_resnets) e->temb[n] = std::vector<float*>();
    }
    const float* cw1 = use_cn ? need_f("controlnet.time_embedding.linear_1.weight") : nullptr;
    const float* cb1 = use_cn ? need_f("controlnet.time_embedding.linear_1.bias") : nullptr;
    const float* cw2 = use_cn ? need_f("controlnet.time_embedding.linear_2.weight") : nullptr;
    const float* cb2 = use_cn ? need_f("controlnet.time_embedding.linear_2.bias") : nullptr;
    if (use_cn) ENG_REQUIRE(cw1 && cb1 && cw2 && cb2, "controlnet.time_embedding weights missing");
    std::vector<float> sinus(320);
    for (int i = 0; i < steps; ++i) {
        // Timesteps(320, flip_sin_to_cos=True, freq_shift=0): [cos | sin], fp32
        const float t = (float)timesteps[i];
        for (int k = 0; k < 160; ++k) {
            const float f = expf(-logf(10000.0f) * (float)k / 160.0f);
            const float a = t * f;
            sinus[k] = cosf(a);
            sinus[160 + k] = sinf(a);
        }
        VSD_CHECK_CUDA(cudaMemcpyAsync(d_sin, sinus.data(), 320 * 4, cudaMemcpyHostToDevice, e->stream));
        VSD_CHECK_CUDA(sync_engine(e));
        // t_emb_in = sinusoid + cond_proj(w_emb)   (bias pointer reused as the additive term)
        int rc = launch_gemv_f32(w_cond, d_wemb, d_sin, e->t_emb_in, 320, 256, 0, 0, e->stream);
        if (rc) return rc;
        rc = launch_gemv_f32(w1, e->t_emb_in, b1, e->t_h, 1280, 320, 0, 1, e->stream);
        if (rc) return rc;
        rc = launch_gemv_f32(w2, e->t_h, b2, e->t_emb, 1280, 1280, 0, 0, e->stream);
        if (rc) return rc;
        for (int r = 0; r < 22; ++r) {
            const std::string p = kResnetPrefixes[r];
            auto it = e->w.find(p + ".time_emb_proj.weight");
            const float* pb = need_f(p + ".time_emb_proj.bias");
            ENG_REQUIRE(it != e->w.end() && pb, "time_emb_proj missing: " + p);
            const int cout = (int)it->second.shape[0];
            float* dst = (float*)A.alloc((size_t)e->NB * cout * 4);
            ENG_REQUIRE(dst != nullptr, "arena exhausted");
            rc = launch_gemv_f32((const float*)it->second.p, e->t_emb, pb, dst, cout, 1280, 1, 0, e->stream);
            if (rc) return rc;
            for (int b = 1; b < e->NB; ++b)
                VSD_CHECK_CUDA(cudaMemcpyAsync(dst + (size_t)b * cout, dst, (size_t)cout * 4, cudaMemcpyDeviceToDevice, e->stream));
            e->temb[p].push_back(dst);
        }
        if (use_cn) {   // the ControlNet has its own time MLP (no guidance conditioning): emb = L2(SiLU(L1(sinusoid)))
            rc = launch_gemv_f32(cw1, d_sin, cb1, e->t_h, 1280, 320, 0, 1, e->stream);
            if (rc) return rc;
            rc = launch_gemv_f32(cw2, e->t_h, cb2, e->t_emb, 1280, 1280, 0, 0, e->stream);
            if (rc) return rc;
            for (auto& p : cn_resnets) {
                auto it = e->w.find(p + ".time_emb_proj.weight");
                const float* pb = need_f(p + ".time_emb_proj.bias");
                ENG_REQUIRE(it != e->w.end() && pb, "time_emb_proj missing: " + p);
                const int cout = (int)it->second.shape[0];
                float* dst = (float*)A.alloc((size_t)e->NB * cout * 4);
                ENG_REQUIRE(dst != nullptr, "arena exhausted");
                rc = launch_gemv_f32((const float*)it->second.p, e->t_emb, pb, dst, cout, 1280, 1, 0, e->stream);
                if (rc) return rc;
                for (int b = 1; b < e->NB; ++b)
                    VSD_CHECK_CUDA(cudaMemcpyAsync(dst + (size_t)b * cout, dst, (size_t)cout * 4, cudaMemcpyDeviceToDevice, e->stream));
                e->temb[p].push_back(dst);
            }
        }
    }
    VSD_CHECK_CUDA(sync_engine(e));
    e->schedule_set = true;

    // ---- build the launch plans on top of the static + schedule allocations
    e->plan_pre_yuv.clear(); e->plan_pre_rgb.clear(); e->plan_core.clear(); e->plan_post.clear(); e->plan_unet.clear();
    Builder B{e, &e->plan_pre_yuv};
    {
        Engine* ee = e;
        e->plan_pre_yuv.push_back(mk([ee](cudaStream_t st) {
            return launch_yuv420_to_rgb(ee->d_y, ee->d_u, ee->d_v, ee->d_rgb_in, ee->NB, ee->H, ee->W, st);
        }, "yuv2rgb"));
    }
    B.out = &e->plan_core;
    const size_t enc_mark = A.mark();
    {
        Scope sc_("taesd_enc");
        if (e->vae_kind == 1) build_kl_encoder(B, e->d_rgb_in, e->init_latents);
        else build_taesd_encoder(B, e->d_rgb_in, e->init_latents);
    }
    A.release(enc_mark);
    {
        Engine* ee = e;
        const long n = (long)lpx * 4;
        e->plan_core.push_back(mk([ee, n](cudaStream_t st) {
            return launch_add_noise(ee->init_latents, ee->init_noise, ee->noisy, ee->an_a, ee->an_b, n, st);
        }, "add_noise"));
    }
    // persistent skip / concat buffers (shared by all steps)
    UNetStatic S;
    {
        int hh = e->h8, ww = e->w8;
        for (int i = 0; i < 4; ++i) { S.hs[i] = hh; S.ws[i] = ww; hh = ds(hh); ww = ds(ww); }
        const int cat_c[4][3] = {{2560, 2560, 2560}, {2560, 2560, 1920}, {1920, 1280, 960}, {960, 640, 640}};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 3; ++j) S.concat[i][j] = B.alloc(e->NB, S.hs[3 - i], S.ws[3 - i], cat_c[i][j]);
        // up block i resnet j runs at level (3 - i); the skips consumed there were produced at that level,
        // except the j == 2 skip of blocks 0..2, which is the *previous* level's downsampler output... check:
        // pop order: skip 11,10 (level 3 resnets), 9 (down of level 2 -> lives at level 3) => all level 3. OK.
    }
    CNStatic CN;
    if (use_cn) {
        const int cn_c[12] = {320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280};
        const int cn_l[12] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3};
        for (int k = 0; k < 12; ++k) CN.feat[k] = B.alloc(e->NB, S.hs[cn_l[k]], S.ws[cn_l[k]], cn_c[k]);
        CN.mid = B.alloc(e->NB, S.hs[3], S.ws[3], 1280);
        CN.cond = B.alloc(e->NB, e->h8, e->w8, 320);
        const size_t px = (size_t)e->NB * e->H * e->W;
        CN.control = B.alloc_f32(px * 3);
        CN.mag = B.alloc_f32(px);
        CN.maxbits = reinterpret_cast<unsigned int*>(B.alloc_f32(64));
        B.out = &e->plan_core;
        const size_t fm = A.mark();
        {
            Scope sc_("control");
            build_control_frontend(B, e->d_rgb_in, CN);
        }
        A.release(fm);
    }
    const size_t unet_mark = A.mark();
    for (int i = 0; i < steps; ++i) {
        A.release(unet_mark);
        std::vector<Launch> up;
        B.out = &up;
        const float* lat_in = (i == 0) ? e->noisy : e->lat[i - 1];
        {
            Scope sc_("step" + std::to_string(i));
            if (use_cn) {
                // The ControlNet reads only the step's input latents: it runs as a parallel graph branch (side stream 2)
                // beside the UNet encoder + mid block and joins before its residuals are added. Its temporaries keep their
                // own arena region (the UNet's are allocated above the ControlNet's high-water mark).
                const size_t saved_peak = A.peak;
                A.peak = A.off;
                {
                    Scope sc2_("cn");
                    B.begin_side(2);
                    build_controlnet(B, lat_in, i, CN);
                    B.end_side();
                }
                if (Builder::branches_enabled()) A.off = A.peak;
                if (saved_peak > A.peak) A.peak = saved_peak;
                build_unet(B, lat_in, e->eps[i], i, S, [&](const std::function<View(int)>& skip_view, const View& mid_out) {
                    B.join_next(2);
                    build_controlnet_residuals(B, CN, skip_view, mid_out);
                });
            } else {
                build_unet(B, lat_in, e->eps[i], i, S);
            }
        }
        const StepScalars s = e->sc[i];
        Engine* ee = e;
        const long n = (long)lpx * 4;
        const int hn = has_step_noise;
        float* z = e->step_noise + (size_t)i * lpx * 4;
        float* xp = e->lat[i]; float* dn = e->den[i]; const float* ep = e->eps[i];
        e->plan_unet.push_back(up);
        for (auto& l : up) e->plan_core.push_back(l);
        e->plan_core.push_back(mk([=](cudaStream_t st) {
            (void)ee;
            return launch_lcm_step(ep, lat_in, z, xp, dn, s.sqrt_a, s.sqrt_1ma, s.c_skip, s.c_out, s.sqrt_ap, s.sqrt_1map, hn,
                                   n, st);
        }, "lcm_step"));
    }
    A.release(unet_mark);
    B.out = &e->plan_core;
    {
        Scope sc_("taesd_dec");
        if (e->vae_kind == 1) build_kl_decoder(B, e->den[steps - 1], e->image);
        else build_taesd_decoder(B, e->den[steps - 1], e->image);
    }
    A.release(unet_mark);
    {
        Engine* ee = e;
        e->plan_post.push_back(mk([ee](cudaStream_t st) {
            return launch_pack_rgb_yuv420(ee->image, 4, ee->d_rgb_out, ee->d_oy, ee->d_ou, ee->d_ov, ee->NB, ee->H, ee->W,
                                          ee->vae_kind == 1 ? 0 : 1 /* TAESD's decoder ends with x*2-1 */, st);
        }, "pack"));
    }
    if (B.rc) {
        set_error("plan build failed: " + B.fail);
        e->schedule_set = false;
        return B.rc;
    }
    return 0;
}

// Projects a context (77 x 768 fp32, host) into every transformer layer's cross-attention K / V^T cache slot b.
static int set_context(Engine* e, int b, const float* ctx_host) {
    ENG_REQUIRE(e->configured, "configure() first");
    ENG_REQUIRE(b >= 0 && b < e->NB, "context slot out of range");
    std::vector<uint16_t> t((size_t)128 * 768, 0);
    for (long i = 0; i < 77L * 768; ++i) t[i] = f32_to_bf16_rne(ctx_host[i]);
    bf16* dctx = e->ctx_bf16 + (size_t)b * 128 * 768;
    VSD_CHECK_CUDA(cudaMemcpyAsync(dctx, t.data(), t.size() * 2, cudaMemcpyHostToDevice, e->stream));
    VSD_CHECK_CUDA(sync_engine(e));
    for (auto& xa : e->xattn) {
        auto wk = e->w.find(xa.prefix + ".attn2.to_k.weight");
        auto wv = e->w.find(xa.prefix + ".attn2.to_v.weight");
        ENG_REQUIRE(wk != e->w.end() && wv != e->w.end(), "attn2 weights missing: " + xa.prefix);
        GemmOp op;
        // K2[77][8*dkp] = ctx * Wk_pad^T
        ActView actx{dctx, 1, 1, 77, 768, 768};
        int rc = build_gemm_op(&op, actx, 1, (const bf16*)wk->second.p, 8 * xa.dkp, 768, xa.k2 + (size_t)b * 128 * 8 * xa.dkp,
                               8 * xa.dkp, 0, nullptr, nullptr, nullptr, 0, ACT_NONE, e->splitk_ws, e->splitk_bytes, 0, 1);
        if (rc) return rc;
        rc = launch_gemm_op(op, e->stream);
        if (rc) return rc;
        // V2^T[C][77] = Wv * ctx^T, written at column offset b*128 of the [C][NB*128] cache
        ActView aw{wv->second.p, 1, 1, xa.C, 768, 768};
        rc = build_gemm_op(&op, aw, 1, dctx, 77, 768, xa.v2t + (size_t)b * 128, e->NB * 128, 0, nullptr, nullptr, nullptr, 0,
                           ACT_NONE, e->splitk_ws, e->splitk_bytes, 0, 1);
        if (rc) return rc;
        rc = launch_gemm_op(op, e->stream);
        if (rc) return rc;
    }
    VSD_CHECK_CUDA(sync_engine(e));
    e->context_set = true;
    return 0;
}

static int capture(Engine* e, bool yuv, cudaGraphExec_t* exec) {
    cudaGraph_t g = nullptr;
    VSD_CHECK_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    if (yuv) rc = run_plan(e, e->plan_pre_yuv, e->stream);
    if (!rc) rc = run_plan(e, e->plan_core, e->stream);
    if (!rc) rc = run_plan(e, e->plan_post, e->stream);
    cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    VSD_CHECK_CUDA(ce);
    size_t nodes = 0;
    VSD_CHECK_CUDA(cudaGraphGetNodes(g, nullptr, &nodes));
    if (yuv) e->launches_per_frame_yuv = (long)nodes;
    VSD_CHECK_CUDA(cudaGraphInstantiate(exec, g, 0));
    cudaGraphDestroy(g);
    return 0;
}

// Production entry points: queue the device-side fault words behind the frame (pinned host destination), synchronise,
// and fail loudly when a bounded wait timed out somewhere in the frame (the frame's pixels are garbage then).
static int finish_frame(Engine* e) {
    for (int i = 0; i < 2; ++i)
        VSD_CHECK_CUDA(cudaMemcpyAsync(e->fault_host + i, e->fault_dev[i], sizeof(unsigned int), cudaMemcpyDeviceToHost, e->stream));
    VSD_CHECK_CUDA(sync_engine(e, true));
    if (e->fault_host[0] | e->fault_host[1]) return vsd_check_pipeline_fault();   // reads, reports and clears
    return 0;
}

static int run_frame(Engine* e, bool yuv) {
    ENG_REQUIRE(e->schedule_set, "set_schedule() first");
    ENG_REQUIRE(e->context_set, "set_context() first");
    cudaGraphExec_t* ex = yuv ? &e->graph_yuv : &e->graph_rgb;
    if (!*ex) {
        int rc = capture(e, yuv, ex);
        if (rc) return rc;
    }
    VSD_CHECK_CUDA(cudaGraphLaunch(*ex, e->stream));
    return 0;
}

}  // namespace vsd

using namespace vsd;

extern "C" {

struct vsd_ctx { Engine e; };

vsd_ctx* vsd_create(int device) {
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    if (ensure_init()) return nullptr;
    vsd_ctx* c = new vsd_ctx();
    c->e.device = device;
    bool ok = cudaStreamCreateWithFlags(&c->e.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; k < 2 && ok; ++k)
        ok = cudaStreamCreateWithFlags(&c->e.side[k], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->e.ev_fork[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->e.ev_join[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        set_error("cudaStreamCreate failed");
        delete c;
        return nullptr;
    }
    c->e.fault_dev[0] = trap_code_addr_gemm();
    c->e.fault_dev[1] = trap_code_addr_attn();
    if (!c->e.fault_dev[0] || !c->e.fault_dev[1] ||
        cudaHostAlloc(reinterpret_cast<void**>(&c->e.fault_host), 4 * sizeof(unsigned int), cudaHostAllocDefault) != cudaSuccess) {
        set_error("fault-word setup failed");
        vsd_destroy(c);
        return nullptr;
    }
    memset(c->e.fault_host, 0, 4 * sizeof(unsigned int));
    return c;
}

/* A lane: its own stream, buffers, launch plan and CUDA graph, sharing the parent's weights (read-only). Several
 * lanes on one GPU keep more than one frame in flight (sessions pinned to the same GPU, or one stream pipelined).
 * The parent must outlive its lanes and must have all weights loaded before lanes are created. */
vsd_ctx* vsd_create_lane(vsd_ctx* parent) {
    if (!parent) { set_error("null parent"); return nullptr; }
    vsd_ctx* c = vsd_create(parent->e.device);
    if (!c) return nullptr;
    c->e.w = parent->e.w;
    c->e.owns_weights = false;
    c->e.lane_index = parent->e.lanes_created++;
    c->e.root = &parent->e;
    c->e.tuned = parent->e.tuned;
    c->e.autotune = parent->e.autotune;
    return c;
}

void vsd_destroy(vsd_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->e.device);
    cudaStreamSynchronize(c->e.stream);
    free_graphs(&c->e);
    if (c->e.owns_weights)
        for (auto& kv : c->e.w) {
            ln_unfold_forget(kv.second.p);
            cudaFree(kv.second.p);
        }
    c->e.arena.destroy();
    if (c->e.splitk_ws) cudaFree(c->e.splitk_ws);
    for (int k = 0; k < 2; ++k) {
        if (c->e.splitk_ws_side[k]) cudaFree(c->e.splitk_ws_side[k]);
        if (c->e.ev_fork[k]) cudaEventDestroy(c->e.ev_fork[k]);
        if (c->e.ev_join[k]) cudaEventDestroy(c->e.ev_join[k]);
        if (c->e.side[k]) cudaStreamDestroy(c->e.side[k]);
    }
    if (c->e.done_ev) cudaEventDestroy(c->e.done_ev);
    trace_forget(&c->e);
    if (c->e.flush_buf) cudaFree(c->e.flush_buf);
    if (c->e.tune_w) cudaFree(c->e.tune_w);
    free_resize(&c->e);
    free_clip(&c->e);
    for (auto& kv : c->e.derived) cudaFree(kv.second);
    if (c->e.cn_scales) cudaFree(c->e.cn_scales);
    if (c->e.ev0) cudaEventDestroy(c->e.ev0);
    if (c->e.ev1) cudaEventDestroy(c->e.ev1);
    if (c->e.fault_host) cudaFreeHost(c->e.fault_host);
    cudaStreamDestroy(c->e.stream);
    delete c;
}

#define CTX_GUARD(c)                                             \
    if (!(c)) { set_error("null context"); return -1; }          \
    if (cudaSetDevice((c)->e.device) != cudaSuccess) { set_error("cudaSetDevice failed"); return -2; }

int vsd_load_weight(vsd_ctx* c, const char* name, const float* host_f32, const int64_t* shape, int ndim) {
    CTX_GUARD(c);
    if (!c->e.owns_weights) { set_error("weights of a lane belong to its parent context"); return -1; }
    return load_weight(&c->e, name, host_f32, shape, ndim);
}

int vsd_num_weights(vsd_ctx* c) { return c ? (int)c->e.w.size() : -1; }

int vsd_set_autotune(vsd_ctx* c, int enabled) {
    if (!c) return -1;
    c->e.autotune = enabled < 0 ? 0 : enabled;   // 0 off, n = tune for n frames in flight
    return 0;
}

/* Pre-populates the autotuner cache from vsd_tuning_report text (e.g. saved by an earlier process). */
int vsd_tuning_load(vsd_ctx* c, const char* text) {
    if (!c || !text) return -1;
    int n = 0;
    const char* p = text;
    while (*p) {
        char key[200];
        Engine::Tuned t{0, 1, 0, 1, 0, 0.f};
        int consumed = 0;
        if (sscanf(p, "%199s bn=%d splits=%d occ=%d kbs=%d halo=%d us=%f%n", key, &t.bn, &t.splits, &t.occ, &t.kbs, &t.halo, &t.us,
                   &consumed) >= 6 &&
            t.bn > 0) {
            c->e.tuned[key] = t;
            ++n;
        }
        const char* nl = strchr(p, '\n');
        if (!nl) break;
        p = nl + 1;
    }
    return n;
}

/* Writes "key bn splits occ us" lines of the tuned GEMM shapes into buf; returns the number of entries. */
int vsd_tuning_report(vsd_ctx* c, char* buf, long cap) {
    if (!c) return -1;
    std::string s;
    for (auto& kv : c->e.tuned) {
        char line[256];
        snprintf(line, sizeof(line), "%s bn=%d splits=%d occ=%d kbs=%d halo=%d us=%.2f\n", kv.first.c_str(), kv.second.bn,
                 kv.second.splits, kv.second.occ, kv.second.kbs, kv.second.halo, kv.second.us);
        s += line;
    }
    if (buf && cap > 0) {
        const size_t n = std::min((size_t)cap - 1, s.size());
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int)c->e.tuned.size();
}

int vsd_configure(vsd_ctx* c, int batch, int height, int width) {
    CTX_GUARD(c);
    return configure(&c->e, batch, height, width);
}

int vsd_set_schedule(vsd_ctx* c, int steps, const int* timesteps, const float* scalars, float add_noise_a,
                     float add_noise_b, const float* w_embedding256, int has_step_noise) {
    CTX_GUARD(c);
    return set_schedule(&c->e, steps, timesteps, scalars, add_noise_a, add_noise_b, w_embedding256, has_step_noise);
}

/* ControlNet on/off and its 13 per-residual scales (guess mode: logspace(-1, 0, 13) * controlnet_scale, computed by the
 * host like diffusers does). Toggling `enabled` invalidates the schedule (call vsd_set_schedule again). */
int vsd_set_controlnet(vsd_ctx* c, int enabled, const float* scales13) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    if (enabled) ENG_REQUIRE(has_controlnet_weights(e), "no controlnet.* weights are loaded");
    if (!e->cn_scales) {
        VSD_CHECK_CUDA(cudaMalloc(&e->cn_scales, 16 * sizeof(float)));
        VSD_CHECK_CUDA(cudaMemset(e->cn_scales, 0, 16 * sizeof(float)));
    }
    if (scales13) {
        VSD_CHECK_CUDA(cudaMemcpyAsync(e->cn_scales, scales13, 13 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        VSD_CHECK_CUDA(sync_engine(e));
    }
    if ((enabled != 0) != e->cn_enabled) {
        VSD_CHECK_CUDA(sync_engine(e));
        free_graphs(e);
        e->cn_enabled = enabled != 0;
        e->schedule_set = false;
    }
    return 0;
}

int vsd_set_context(vsd_ctx* c, int slot, const float* context_77x768) {
    CTX_GUARD(c);
    return set_context(&c->e, slot, context_77x768);
}

/* 0 = AutoencoderTiny (default, weights "vae."), 1 = AutoencoderKL (weights "vae_kl."). Invalidates the schedule. */
int vsd_set_vae(vsd_ctx* c, int kind) {
    CTX_GUARD(c);
    ENG_REQUIRE(kind == 0 || kind == 1, "vae kind must be 0 (AutoencoderTiny) or 1 (AutoencoderKL)");
    Engine* e = &c->e;
    if (e->vae_kind != kind) {
        VSD_CHECK_CUDA(sync_engine(e));
        free_graphs(e);
        e->vae_kind = kind;
        e->schedule_set = false;
    }
    return 0;
}

/* Noise of `latent_dist.sample()` (AutoencoderKL only): fp32 [batch][h/8][w/8][4] (NHWC), host. */
int vsd_set_vae_noise(vsd_ctx* c, const float* noise_nhwc) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->configured, "configure() first");
    const size_t lpx = (size_t)e->NB * e->h8 * e->w8;
    VSD_CHECK_CUDA(cudaMemcpyAsync(e->vae_noise, noise_nhwc, lpx * 16, cudaMemcpyHostToDevice, e->stream));
    VSD_CHECK_CUDA(sync_engine(e));
    return 0;
}

int vsd_encode_prompt(vsd_ctx* c, const int* token_ids_77, float* context_77x768) {
    CTX_GUARD(c);
    return encode_prompt(&c->e, token_ids_77, context_77x768);
}

int vsd_set_noise(vsd_ctx* c, const float* init_noise_nhwc, const float* step_noise_nhwc) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->schedule_set, "set_schedule() first");
    const size_t lpx = (size_t)e->NB * e->h8 * e->w8;
    VSD_CHECK_CUDA(cudaMemcpyAsync(e->init_noise, init_noise_nhwc, lpx * 16, cudaMemcpyHostToDevice, e->stream));
    if (step_noise_nhwc && e->has_step_noise)
        VSD_CHECK_CUDA(cudaMemcpyAsync(e->step_noise, step_noise_nhwc, lpx * 16 * e->steps, cudaMemcpyHostToDevice, e->stream));
    VSD_CHECK_CUDA(sync_engine(e));
    return 0;
}

int vsd_upload_yuv420(vsd_ctx* c, const uint8_t* y, const uint8_t* u, const uint8_t* v) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->configured, "configure() first");
    const size_t px = (size_t)e->NB * e->H * e->W;
    VSD_CHECK_CUDA(cudaMemcpyAsync(e->d_y, y, px, cudaMemcpyHostToDevice, e->stream));
    VSD_CHECK_CUDA(cudaMemcpyAsync(e->d_u, u, px / 4, cudaMemcpyHostToDevice, e->stream));
    VSD_CHECK_CUDA(cudaMemcpyAsync(e->d_v, v, px / 4, cudaMemcpyHostToDevice, e->stream));
    return 0;
}

int vsd_run_yuv420(vsd_ctx* c) {
    CTX_GUARD(c);
    return run_frame(&c->e, true);
}

int vsd_download_yuv420(vsd_ctx* c, uint8_t* y, uint8_t* u, uint8_t* v) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    const size_t px = (size_t)e->NB * e->H * e->W;
    VSD_CHECK_CUDA(cudaMemcpyAsync(y, e->d_oy, px, cudaMemcpyDeviceToHost, e->stream));
    VSD_CHECK_CUDA(cudaMemcpyAsync(u, e->d_ou, px / 4, cudaMemcpyDeviceToHost, e->stream));
    VSD_CHECK_CUDA(cudaMemcpyAsync(v, e->d_ov, px / 4, cudaMemcpyDeviceToHost, e->stream));
    return 0;
}

int vsd_sync(vsd_ctx* c) {
    CTX_GUARD(c);
    VSD_CHECK_CUDA(sync_engine(&c->e));
    return vsd_check_pipeline_fault();
}

int vsd_infer_yuv420(vsd_ctx* c, const uint8_t* y, const uint8_t* u, const uint8_t* v, uint8_t* out_y, uint8_t* out_u,
                     uint8_t* out_v) {
    int rc = vsd_upload_yuv420(c, y, u, v);
    if (rc) return rc;
    rc = vsd_run_yuv420(c);
    if (rc) return rc;
    rc = vsd_download_yuv420(c, out_y, out_u, out_v);
    if (rc) return rc;
    return finish_frame(&c->e);
}

int vsd_infer_rgb(vsd_ctx* c, const uint8_t* rgb_in, uint8_t* rgb_out) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->configured, "configure() first");
    const size_t px = (size_t)e->NB * e->H * e->W;
    VSD_CHECK_CUDA(cudaMemcpyAsync(e->d_rgb_in, rgb_in, px * 3, cudaMemcpyHostToDevice, e->stream));
    int rc = run_frame(e, false);
    if (rc) return rc;
    VSD_CHECK_CUDA(cudaMemcpyAsync(rgb_out, e->d_rgb_out, px * 3, cudaMemcpyDeviceToHost, e->stream));
    return finish_frame(e);
}

int vsd_set_resize(vsd_ctx* c, int in_w, int in_h, int x0, int y0, int cw, int ch, const int* h_bounds, const int* h_coeffs,
                   int h_ksize, const int* v_bounds, const int* v_coeffs, int v_ksize) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->configured, "configure() first");
    ENG_REQUIRE(in_w > 0 && in_h > 0 && cw > 0 && ch > 0 && x0 >= 0 && y0 >= 0 && x0 + cw <= in_w && y0 + ch <= in_h,
                "crop rectangle outside the input frame");
    VSD_CHECK_CUDA(sync_engine(e));
    free_resize(e);
    Engine::Resize& r = e->rz;
    r.in_w = in_w; r.in_h = in_h; r.x0 = x0; r.y0 = y0; r.cw = cw; r.ch = ch; r.hks = h_ksize; r.vks = v_ksize;
    const size_t hb_b = (size_t)e->W * 2 * 4, hk_b = (size_t)e->W * h_ksize * 4, vb_b = (size_t)e->H * 2 * 4, vk_b = (size_t)e->H * v_ksize * 4;
    VSD_CHECK_CUDA(cudaMalloc(&r.hb, hb_b)); VSD_CHECK_CUDA(cudaMalloc(&r.hk, hk_b));
    VSD_CHECK_CUDA(cudaMalloc(&r.vb, vb_b)); VSD_CHECK_CUDA(cudaMalloc(&r.vk, vk_b));
    VSD_CHECK_CUDA(cudaMalloc(&r.src, (size_t)e->NB * in_w * in_h * 3));
    VSD_CHECK_CUDA(cudaMalloc(&r.tmp, (size_t)e->NB * ch * e->W * 3));
    if (in_w % 2 == 0 && in_h % 2 == 0) VSD_CHECK_CUDA(cudaMalloc(&r.yuv, (size_t)e->NB * in_w * in_h * 3 / 2));
    VSD_CHECK_CUDA(cudaMemcpy(r.hb, h_bounds, hb_b, cudaMemcpyHostToDevice));
    VSD_CHECK_CUDA(cudaMemcpy(r.hk, h_coeffs, hk_b, cudaMemcpyHostToDevice));
    VSD_CHECK_CUDA(cudaMemcpy(r.vb, v_bounds, vb_b, cudaMemcpyHostToDevice));
    VSD_CHECK_CUDA(cudaMemcpy(r.vk, v_coeffs, vk_b, cudaMemcpyHostToDevice));
    return 0;
}

/* Arbitrary-size packed RGB24 frames in ([batch][in_h][in_w][3], the geometry given to vsd_set_resize), working-size frames
 * out: center crop + Lanczos resize on the GPU (videopipeline.py:92-107), then the same path as vsd_infer_rgb. */
int vsd_infer_rgb_resized(vsd_ctx* c, const uint8_t* rgb_src, uint8_t* rgb_out) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->rz.src != nullptr, "vsd_set_resize() first");
    const Engine::Resize& r = e->rz;
    VSD_CHECK_CUDA(cudaMemcpyAsync(r.src, rgb_src, (size_t)e->NB * r.in_w * r.in_h * 3, cudaMemcpyHostToDevice, e->stream));
    int rc = launch_crop_resize(r.src, r.in_w, r.in_h, r.x0, r.y0, r.cw, r.ch, r.tmp, e->d_rgb_in, e->W, e->H, r.hb, r.hk, r.hks,
                                r.vb, r.vk, r.vks, e->NB, e->stream);
    if (rc) return rc;
    rc = run_frame(e, false);
    if (rc) return rc;
    const size_t px = (size_t)e->NB * e->H * e->W;
    VSD_CHECK_CUDA(cudaMemcpyAsync(rgb_out, e->d_rgb_out, px * 3, cudaMemcpyDeviceToHost, e->stream));
    return finish_frame(e);
}

/* Same with YUV420P planes at the source geometry ([batch][in_h][in_w], [batch][in_h/2][in_w/2] x2): colour conversion at the
 * source size (what frame.to_image() does, server.py:104-108), crop + Lanczos, frame, working-size YUV420P planes out. */
int vsd_infer_yuv420_resized(vsd_ctx* c, const uint8_t* y, const uint8_t* u, const uint8_t* v, uint8_t* out_y, uint8_t* out_u,
                             uint8_t* out_v) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->rz.src != nullptr, "vsd_set_resize() first");
    ENG_REQUIRE(e->rz.yuv != nullptr, "YUV420 input needs even source width and height");
    const Engine::Resize& r = e->rz;
    const size_t spx = (size_t)e->NB * r.in_w * r.in_h;
    uint8_t *dy = r.yuv, *du = r.yuv + spx, *dv = r.yuv + spx + spx / 4;
    VSD_CHECK_CUDA(cudaMemcpyAsync(dy, y, spx, cudaMemcpyHostToDevice, e->stream));
    VSD_CHECK_CUDA(cudaMemcpyAsync(du, u, spx / 4, cudaMemcpyHostToDevice, e->stream));
    VSD_CHECK_CUDA(cudaMemcpyAsync(dv, v, spx / 4, cudaMemcpyHostToDevice, e->stream));
    int rc = launch_yuv420_to_rgb(dy, du, dv, r.src, e->NB, r.in_h, r.in_w, e->stream);
    if (!rc) rc = launch_crop_resize(r.src, r.in_w, r.in_h, r.x0, r.y0, r.cw, r.ch, r.tmp, e->d_rgb_in, e->W, e->H, r.hb, r.hk, r.hks,
                                     r.vb, r.vk, r.vks, e->NB, e->stream);
    if (!rc) rc = run_frame(e, false);
    if (rc) return rc;
    rc = vsd_download_yuv420(c, out_y, out_u, out_v);
    if (rc) return rc;
    return finish_frame(e);
}

/* bring-up / tests: the resized working-size input frame(s) of the last vsd_infer_rgb_resized call */
int vsd_debug_read_rgb_in(vsd_ctx* c, uint8_t* host) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    VSD_CHECK_CUDA(sync_engine(e));
    VSD_CHECK_CUDA(cudaMemcpy(host, e->d_rgb_in, (size_t)e->NB * e->H * e->W * 3, cudaMemcpyDeviceToHost));
    return 0;
}

void* vsd_stream(vsd_ctx* c) { return c ? (void*)c->e.stream : nullptr; }

long vsd_launches_per_frame(vsd_ctx* c, int yuv) {
    if (!c) return -1;
    const Engine& e = c->e;
    if (yuv && e.launches_per_frame_yuv > 0) return e.launches_per_frame_yuv;  // kernel nodes of the captured graph
    // before capture: plan entries (a lower bound: GroupNorm and split-K entries launch two kernels each)
    return (long)e.plan_core.size() + (long)e.plan_post.size() + (yuv ? (long)e.plan_pre_yuv.size() : 0);
}

/* Number of GEMM shapes the autotuner had to time on the device since vsd_create (0 when a loaded table covered the plan). */
long vsd_tuning_misses(vsd_ctx* c) { return c ? c->e.tune_misses : -1; }

long vsd_arena_peak_bytes(vsd_ctx* c) { return c ? (long)c->e.arena.peak : -1; }

/* ---- debug taps used by the parity tests ---- */
int vsd_debug_read(vsd_ctx* c, const char* what, int index, float* host, long nfloats) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    const std::string w = what;
    const float* src = nullptr;
    if (w == "init_latents") src = e->init_latents;
    else if (w == "noisy") src = e->noisy;
    else if (w == "image") src = e->image;
    else if (w == "eps" && index >= 0 && index < (int)e->eps.size()) src = e->eps[index];
    else if (w == "latents" && index >= 0 && index < (int)e->lat.size()) src = e->lat[index];
    else if (w == "denoised" && index >= 0 && index < (int)e->den.size()) src = e->den[index];
    ENG_REQUIRE(src != nullptr, "unknown debug tap " + w);
    VSD_CHECK_CUDA(sync_engine(e));
    VSD_CHECK_CUDA(cudaMemcpy(host, src, (size_t)nfloats * 4, cudaMemcpyDeviceToHost));
    return 0;
}

/* Runs one UNet pass eagerly: latents (host fp32 NHWC [NB,h8,w8,4]) at schedule step `step` -> eps (host). */
int vsd_debug_unet(vsd_ctx* c, const float* latents_nhwc, int step, float* eps_nhwc) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->schedule_set && e->context_set, "schedule and context must be set");
    ENG_REQUIRE(step >= 0 && step < e->steps, "step out of range");
    const size_t lpx = (size_t)e->NB * e->h8 * e->w8;
    float* dst = (step == 0) ? e->noisy : e->lat[step - 1];
    VSD_CHECK_CUDA(cudaMemcpyAsync(dst, latents_nhwc, lpx * 16, cudaMemcpyHostToDevice, e->stream));
    int rc = run_plan(e, e->plan_unet[step], e->stream);
    if (rc) return rc;
    VSD_CHECK_CUDA(cudaMemcpyAsync(eps_nhwc, e->eps[step], lpx * 16, cudaMemcpyDeviceToHost, e->stream));
    VSD_CHECK_CUDA(sync_engine(e));
    return vsd_check_pipeline_fault();
}

/* Section profiler: groups consecutive plan entries by the first `depth` components of their tag, captures each
 * group as its own CUDA graph and times `reps` replays with events. Writes "tag launches us" lines into buf. */
int vsd_debug_profile_sections(vsd_ctx* c, int depth, int reps, char* buf, long cap) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->schedule_set && e->context_set, "schedule and context must be set");
    std::vector<const Launch*> all;
    for (auto& l : e->plan_pre_yuv) all.push_back(&l);
    for (auto& l : e->plan_core) all.push_back(&l);
    for (auto& l : e->plan_post) all.push_back(&l);
    auto key = [&](const std::string& t) {
        size_t pos = 0;
        for (int d = 0; d < depth; ++d) {
            size_t nx = t.find('.', pos);
            if (nx == std::string::npos) return t;
            pos = nx + 1;
        }
        return t.substr(0, pos ? pos - 1 : 0);
    };
    cudaEvent_t e0, e1;
    VSD_CHECK_CUDA(cudaEventCreate(&e0));
    VSD_CHECK_CUDA(cudaEventCreate(&e1));
    std::string out;
    size_t i = 0;
    while (i < all.size()) {
        const std::string k = key(all[i]->tag);
        size_t j = i;
        while (j < all.size() && key(all[j]->tag) == k) ++j;
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ex = nullptr;
        VSD_CHECK_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (size_t q = i; q < j && !rc; ++q) rc = (*all[q])(e->stream);
        cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        if (rc) return rc;
        VSD_CHECK_CUDA(ce);
        size_t nodes = 0;
        cudaGraphGetNodes(g, nullptr, &nodes);
        VSD_CHECK_CUDA(cudaGraphInstantiate(&ex, g, 0));
        VSD_CHECK_CUDA(cudaGraphLaunch(ex, e->stream));   // warm
        VSD_CHECK_CUDA(cudaEventRecord(e0, e->stream));
        for (int r = 0; r < reps; ++r) VSD_CHECK_CUDA(cudaGraphLaunch(ex, e->stream));
        VSD_CHECK_CUDA(cudaEventRecord(e1, e->stream));
        VSD_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        char line[256];
        snprintf(line, sizeof(line), "%s %zu %.2f\n", k.c_str(), nodes, ms * 1000.f / reps);
        out += line;
        cudaGraphExecDestroy(ex);
        cudaGraphDestroy(g);
        i = j;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (buf && cap > 0) {
        const size_t n = std::min((size_t)cap - 1, out.size());
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return 0;
}

/* Runs the whole frame eagerly (no CUDA graph), for debugging and for ncu launch lists. */
int vsd_debug_run_eager(vsd_ctx* c, int yuv) {
    CTX_GUARD(c);
    Engine* e = &c->e;
    ENG_REQUIRE(e->schedule_set && e->context_set, "schedule and context must be set");
    int rc = 0;
    if (yuv) rc = run_plan(e, e->plan_pre_yuv, e->stream);
    if (!rc) rc = run_plan(e, e->plan_core, e->stream);
    if (!rc) rc = run_plan(e, e->plan_post, e->stream);
    if (rc) return rc;
    VSD_CHECK_CUDA(sync_engine(e));
    return vsd_check_pipeline_fault();
}

}  // extern "C"
