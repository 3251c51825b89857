// lpx_api.cu — library lifecycle, error reporting and small utilities of the C ABI (include/lpx.h).
#include <atomic>
#include <cstring>

#include "lpx_common.cuh"
#include "lpx_runtime.hpp"

namespace lpx {

static thread_local std::string t_error;
static std::atomic<long long> g_launches{0};
static thread_local int t_bnb_instance = 0;

void set_error(const std::string& msg) { t_error = msg; }
void count_launch(int n) { g_launches += n; }
void set_bnb_instance(int k) { t_bnb_instance = k; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    set_error(buf);
    return LPX_E_CUDA;
}

Runtime& rt() {
    static Runtime r;
    return r;
}

static int init_on(int device) {
    Runtime& r = rt();
    std::lock_guard<std::recursive_mutex> lk(r.mu);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error(std::string("no CUDA device available (liblpx has no CPU fallback): ") +
                  (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
        return LPX_E_CUDA;
    }
    if (device < 0) LPX_CUDA(cudaGetDevice(&device));
    if (r.ready && r.device == device) return LPX_OK;
    if (r.ready) lpx_shutdown();
    LPX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LPX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char buf[256];
        std::snprintf(buf, sizeof buf, "device %d (%s) is sm_%d%d; liblpx is built for sm_100a only", device, prop.name,
                      prop.major, prop.minor);
        set_error(buf);
        return LPX_E_CUDA;
    }
    r.device = device;
    r.sms = prop.multiProcessorCount;
    r.smem_optin = (int)prop.sharedMemPerBlockOptin;
    LPX_CUDA(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
    LPX_CUDA(cudaStreamCreateWithFlags(&r.h2d, cudaStreamNonBlocking));
    LPX_CUDA(cudaStreamCreateWithFlags(&r.d2h, cudaStreamNonBlocking));
    r.ready = true;
    return LPX_OK;
}

int ensure_device() {
    Runtime& r = rt();
    if (r.ready) {
        // another library (torch) may have switched the current device behind our back
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != r.device) cudaSetDevice(r.device);
        return LPX_OK;
    }
    return init_on(-1);
}

int sm_count() { return rt().sms; }
int max_smem_optin() { return rt().smem_optin; }

void* ws_dev(Slot s, size_t bytes) {
    Runtime& r = rt();
    if (bytes == 0) bytes = 16;
    if (r.dev_bytes[s] >= bytes) return r.dev[s];
    if (r.dev[s]) {
        cudaDeviceSynchronize();
        cudaFree(r.dev[s]);
        r.dev[s] = nullptr;
        r.dev_bytes[s] = 0;
    }
    size_t want = bytes + bytes / 8;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc(workspace)", __FILE__, __LINE__);
        return nullptr;
    }
    r.dev[s] = p;
    r.dev_bytes[s] = want;
    return p;
}

void* ws_pin(Slot s, size_t bytes) {
    Runtime& r = rt();
    if (bytes == 0) bytes = 16;
    if (r.pin_bytes[s] >= bytes) return r.pin[s];
    if (r.pin[s]) {
        cudaDeviceSynchronize();
        cudaFreeHost(r.pin[s]);
        r.pin[s] = nullptr;
        r.pin_bytes[s] = 0;
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaHostAlloc(workspace)", __FILE__, __LINE__);
        return nullptr;
    }
    r.pin[s] = p;
    r.pin_bytes[s] = bytes;
    return p;
}

}  // namespace lpx

using namespace lpx;

extern "C" {

void lpx_default_options(lpx_options* opt) {
    if (!opt) return;
    std::memset(opt, 0, sizeof *opt);
    opt->max_iterations = 10000;
    opt->kernel = LPX_KERNEL_AUTO;
}

const char* lpx_version(void) { return "lpx 0.1 (sm_100a, FP64 tableau simplex / B&B)"; }
const char* lpx_last_error(void) { return t_error.c_str(); }

const char* lpx_status_message(int status) {
    switch (status) {
        case LPX_OPTIMAL: return "OPTIMAL";
        case LPX_UNBOUNDED: return "UNBOUNDED";
        case LPX_INFEASIBLE: return "INFEASIBLE";
        case LPX_RUNNING: return "RUNNING";
        case LPX_S_GE_ROW:
            return "Constraint contains '>=' sign. The Primal Simplex method cannot handle this. Please try the Dual "
                   "Simplex algorithm instead.";
        case LPX_S_NEG_RHS:
            return "Constraint has a negative RHS value. The Primal Simplex method cannot handle this. Please try the "
                   "Dual Simplex algorithm instead.";
        case LPX_S_ITER_LIMIT: return "Iteration limit exceeded.";
        case LPX_S_REV_UNSUPPORTED:
            return "Revised Primal Simplex currently supports only <= constraints with RHS >= 0. Use Dual Simplex for "
                   "models with >= or =.";
        case LPX_S_SINGULAR: return "Singular basis encountered.";
        case LPX_E_BAD_ARGS: return "bad arguments";
        case LPX_E_CUDA: return "CUDA error";
        case LPX_E_CAPACITY: return "problem exceeds kernel capacity";
        case LPX_E_NCCL: return "NCCL error";
    }
    return "unknown status";
}

int lpx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int lpx_init(int device) { return init_on(device); }

void lpx_shutdown(void) {
    Runtime& r = rt();
    std::lock_guard<std::recursive_mutex> lk(r.mu);
    if (!r.ready) return;
    cudaDeviceSynchronize();
    knapsack_release_cache();
    pooled_release_cache();
    for (int s = 0; s < WS_COUNT; s++) {
        if (r.dev[s]) cudaFree(r.dev[s]);
        if (r.pin[s]) cudaFreeHost(r.pin[s]);
        r.dev[s] = r.pin[s] = nullptr;
        r.dev_bytes[s] = r.pin_bytes[s] = 0;
    }
    if (r.stream) cudaStreamDestroy(r.stream);
    if (r.h2d) cudaStreamDestroy(r.h2d);
    if (r.d2h) cudaStreamDestroy(r.d2h);
    r.stream = r.h2d = r.d2h = nullptr;
    r.ready = false;
}

void* lpx_host_alloc(size_t bytes) {
    if (ensure_device() != LPX_OK) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__);
        return nullptr;
    }
    return p;
}

void lpx_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int lpx_tableau_dims(int m, int n, const int* rel, int* rows, int* cols) {
    if (m < 1 || n < 1) {
        set_error("lpx_tableau_dims: m and n must be >= 1");
        return LPX_E_BAD_ARGS;
    }
    int mm = 0;
    for (int i = 0; i < m; i++) mm += (rel && rel[i] == 2) ? 2 : 1;
    if (rows) *rows = mm + 1;
    if (cols) *cols = n + mm + 1;
    return LPX_OK;
}

long long lpx_kernel_launches(void) { return g_launches.load(); }
void lpx_reset_counters(void) { g_launches = 0; }
int lpx_bnb_instance(void) { return t_bnb_instance; }

}  // extern "C"
