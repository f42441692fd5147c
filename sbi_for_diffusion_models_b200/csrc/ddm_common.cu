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

// One thread waits (bounded) for a word in mapped host memory that the host raises right after the launch
// call has returned.  If the kernel times out instead, the launch call did not return while the kernel was
// running: launches are serialised (CUDA_LAUNCH_BLOCKING, Nsight Compute, compute-sanitizer, a debugger).
__global__ void probe_wait_kernel(volatile const unsigned int *flag, unsigned long long timeout_ns,
                                  volatile unsigned int *timed_out)
{
    unsigned long long t0 = 0ull, now = 0ull;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*flag == 0u) {
        __nanosleep(2000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) {
            *timed_out = 1u;
            return;
        }
    }
}

}  // namespace ddm

DDM_API int ddm_probe_launch_blocking(int64_t timeout_us, int *blocking)
{
    DDM_REQUIRE(blocking != nullptr && timeout_us > 0 && timeout_us <= 5000000, "ddm_probe_launch_blocking: bad arguments");
    unsigned int *words = nullptr;
    DDM_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&words), 2 * sizeof(unsigned int), cudaHostAllocMapped));
    words[0] = 0u;
    words[1] = 0u;
    unsigned int *dev_words = nullptr;
    cudaStream_t st = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(reinterpret_cast<void **>(&dev_words), words, 0);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        ddm::probe_wait_kernel<<<1, 1, 0, st>>>(dev_words, (unsigned long long)timeout_us * 1000ull, dev_words + 1);
        e = cudaGetLastError();
        __atomic_store_n(&words[0], 1u, __ATOMIC_RELEASE);  // the launch call is back: release the kernel
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    const int timed_out = (int)__atomic_load_n(&words[1], __ATOMIC_ACQUIRE);
    if (st) cudaStreamDestroy(st);
    cudaFreeHost(words);
    if (e != cudaSuccess) return ddm::cuda_fail(e, "ddm_probe_launch_blocking");
    *blocking = timed_out ? 1 : 0;
    return DDM_OK;
}

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
