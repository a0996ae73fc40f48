// C-ABI plumbing: error text, version, launch checking.
#include "common.cuh"
#include "../../include/mfnerf_b200.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

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

static std::atomic<long long> g_launches{0};
void note_launch(int k) { g_launches.fetch_add(k, std::memory_order_relaxed); }

// profiling table: mfn_profile_set(name, start, stop) registers one entry; an empty name clears the table
constexpr int kMaxProf = 32;
struct ProfEntry { char name[48]; cudaEvent_t start, stop; };
static ProfEntry g_prof[kMaxProf];
static int g_n_prof = 0;

static void prof_record(cudaEvent_t ev, cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    cudaEventRecordWithFlags(ev, st, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}
ProfScope::ProfScope(const char* name, cudaStream_t stream) : st(stream), slot(-1) {
    for (int i = 0; i < g_n_prof; ++i)
        if (strcmp(name, g_prof[i].name) == 0) { slot = i; prof_record(g_prof[i].start, st); break; }
}
ProfScope::~ProfScope() { if (slot >= 0 && slot < g_n_prof) prof_record(g_prof[slot].stop, st); }

int check_launch(const char* what, cudaStream_t stream) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
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

extern "C" int64_t mfn_launch_count(void) { return (int64_t)mfn::g_launches.load(); }

extern "C" void* mfn_event_create(void) {
    cudaEvent_t ev = nullptr;
    if (cudaEventCreate(&ev) != cudaSuccess) { (void)cudaGetLastError(); mfn::set_error("mfn_event_create: cudaEventCreate failed"); return nullptr; }
    return (void*)ev;
}
extern "C" int mfn_event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); return MFN_OK; }
extern "C" int mfn_event_elapsed_ms(void* start, void* stop, float* ms_host) {
    if (!start || !stop || !ms_host) { mfn::set_error("mfn_event_elapsed_ms: null pointer"); return MFN_ERR_ARG; }
    cudaError_t e = cudaEventSynchronize((cudaEvent_t)stop);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms_host, (cudaEvent_t)start, (cudaEvent_t)stop);
    if (e != cudaSuccess) { (void)cudaGetLastError(); mfn::set_error("mfn_event_elapsed_ms: %s", cudaGetErrorString(e)); return MFN_ERR_CUDA; }
    return MFN_OK;
}
extern "C" int mfn_profile_set(const char* kernel_name, void* ev_start, void* ev_stop) {
    if (!kernel_name || !kernel_name[0]) { mfn::g_n_prof = 0; return MFN_OK; }
    if (!ev_start || !ev_stop) { mfn::set_error("mfn_profile_set: null event"); return MFN_ERR_ARG; }
    int slot = mfn::g_n_prof;
    for (int i = 0; i < mfn::g_n_prof; ++i) if (strcmp(kernel_name, mfn::g_prof[i].name) == 0) slot = i;
    if (slot >= mfn::kMaxProf) { mfn::set_error("mfn_profile_set: more than %d profiled kernels", mfn::kMaxProf); return MFN_ERR_ARG; }
    strncpy(mfn::g_prof[slot].name, kernel_name, sizeof(mfn::g_prof[slot].name) - 1);
    mfn::g_prof[slot].name[sizeof(mfn::g_prof[slot].name) - 1] = 0;
    mfn::g_prof[slot].start = (cudaEvent_t)ev_start; mfn::g_prof[slot].stop = (cudaEvent_t)ev_stop;
    if (slot == mfn::g_n_prof) ++mfn::g_n_prof;
    return MFN_OK;
}
