// Error reporting + device info for the tt_b200 C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace tt {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

}  // namespace tt

extern "C" int tt_abi_version(void) { return TT_ABI_VERSION; }

extern "C" const char *tt_last_error(void) { return tt::g_err; }

extern "C" int tt_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return tt::cuda_status(e, "cudaGetDevice");
    int sms = 0, major = 0, minor = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (sm_count_host) *sm_count_host = sms;
    if (cc_major_host) *cc_major_host = major;
    if (cc_minor_host) *cc_minor_host = minor;
    if (major != 10) {
        tt::set_error("tt_b200 is built for sm_100a only; device reports cc %d.%d", major, minor);
        return TT_E_DEVICE;
    }
    return 0;
}
