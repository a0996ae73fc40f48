// C-ABI plumbing: error text, version, launch checking.
#include "common.cuh"
#include "../../include/mfnerf_b200.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace mfn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int debug_sync() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MFN_DEBUG_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}

int check_launch(const char* what, cudaStream_t stream) {
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess && debug_sync()) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return MFN_ERR_CUDA;
    }
    return MFN_OK;
}

}  // namespace mfn

extern "C" const char* mfn_last_error(void) { return mfn::g_err; }
extern "C" int mfn_version(void) { return MFN_VERSION; }
extern "C" int mfn_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}
