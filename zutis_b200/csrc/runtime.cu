// Error reporting and device facts for libzutis_b200 (no global mutable state beyond caches).
#include "common.cuh"

#include <mutex>

namespace zutis {

static thread_local char g_err[512] = "ok";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return ZUTIS_OK;
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    (void)cudaGetLastError();
    return ZUTIS_ERR_CUDA;
}

int check_launch(const char* what) { return check_cuda(cudaGetLastError(), what); }

struct DeviceFacts {
    int known = 0, sms = 0, major = 0, minor = 0;
};
static DeviceFacts g_facts[64];
static std::mutex g_facts_mu;

static int facts_for_current(DeviceFacts* out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= 64) {
        (void)cudaGetLastError();
        return fail(ZUTIS_ERR_NO_DEVICE, "no CUDA device available (%s); libzutis_b200 has no CPU fallback",
                    e == cudaSuccess ? "bad ordinal" : cudaGetErrorString(e));
    }
    std::lock_guard<std::mutex> lk(g_facts_mu);
    DeviceFacts& f = g_facts[dev];
    if (!f.known) {
        cudaDeviceProp p;
        int st = check_cuda(cudaGetDeviceProperties(&p, dev), "cudaGetDeviceProperties");
        if (st != ZUTIS_OK) return st;
        f.sms = p.multiProcessorCount; f.major = p.major; f.minor = p.minor; f.known = 1;
    }
    *out = f;
    return ZUTIS_OK;
}

int sm_count() {
    DeviceFacts f;
    if (facts_for_current(&f) != ZUTIS_OK) return 148;
    return f.sms;
}

int current_device_ok() {
    DeviceFacts f;
    int st = facts_for_current(&f);
    if (st != ZUTIS_OK) return st;
    if (f.major != 10)
        return fail(ZUTIS_ERR_NO_DEVICE, "device is sm_%d%d; libzutis_b200 is built for sm_100a only", f.major, f.minor);
    return ZUTIS_OK;
}

}  // namespace zutis

extern "C" {

const char* zutis_last_error_string(void) { return zutis::g_err; }

int zutis_abi_version(void) { return 1; }

int zutis_device_check(int device) {
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) {
        (void)cudaGetLastError();
        return zutis::fail(ZUTIS_ERR_NO_DEVICE, "no CUDA device available; libzutis_b200 has no CPU fallback");
    }
    if (zutis::check_cuda(cudaSetDevice(device), "cudaSetDevice") != ZUTIS_OK) return ZUTIS_ERR_NO_DEVICE;
    int st = zutis::current_device_ok();
    cudaSetDevice(prev);
    return st;
}

}  // extern "C"
