// Error plumbing and device queries for libddm_b200.so.
#include "ddm_common.cuh"

#include <stdarg.h>

namespace ddm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return DDM_ERR_CUDA;
}

}  // namespace ddm

DDM_API int ddm_abi_version(void) { return DDM_ABI_VERSION; }

DDM_API const char *ddm_last_error(void) { return ddm::g_err; }

DDM_API int ddm_device_info(int device, int *sm_count, int *sm_clock_khz, int *cc_major, int *cc_minor)
{
    int v = 0;
    if (sm_count) {
        DDM_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
        *sm_count = v;
    }
    if (sm_clock_khz) {
        DDM_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device));
        *sm_clock_khz = v;
    }
    if (cc_major) {
        DDM_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
        *cc_major = v;
    }
    if (cc_minor) {
        DDM_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
        *cc_minor = v;
    }
    return DDM_OK;
}
