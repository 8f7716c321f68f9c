// C-ABI entry points for single operators (declared in include/videosd.h, section "operator entry points").
// These are what the -m gpu parity tests call, one kernel family at a time, with device pointers owned by
// the caller (torch tensors' data_ptr()). The frame-level API lives in vsd_engine.cu.
#include "vsd_internal.h"
#include "../../include/videosd.h"
#include <mutex>
#include <vector>

namespace vsd {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }

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

}  // namespace vsd

using namespace vsd;

extern "C" {

const char* vsd_last_error(void) { return get_error(); }

int vsd_abi_version(void) { return VSD_ABI_VERSION; }

int vsd_check_pipeline_fault(void) {
    unsigned int a = read_trap_code_gemm();
    unsigned int b = read_trap_code_attn();
    if (a || b) {
        char buf[160];
        snprintf(buf, sizeof(buf), "tensor-core pipeline wait timed out: gemm=0x%08x attn=0x%08x", a, b);
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
    rc = build_gemm_op(&op, a, taps, reinterpret_cast<const bf16*>(wt), n, taps * c, out, ldo, out_f32, bias, rowvec,
                       reinterpret_cast<const bf16*>(residual), ldr, act, g_ws, g_ws_bytes, block_n, splits);
    if (rc) return rc;
    return launch_gemm_op(op, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
