// Shared device/host helpers for libddm_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ddm_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libddm_b200 targets sm_100a (B200) only"
#endif

#define DDM_API extern "C" __attribute__((visibility("default")))

namespace ddm {

// ---- error plumbing ---------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define DDM_CUDA_TRY(expr)                                        \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return ::ddm::cuda_fail(_e, #expr); \
    } while (0)

#define DDM_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            ::ddm::set_error(__VA_ARGS__);  \
            return DDM_ERR_INVALID;         \
        }                                   \
    } while (0)

// ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------------------
// counter = (lo32(trial), hi32(trial), step / 4, 0), key = (lo32(seed), hi32(seed)).
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

struct PhiloxKey {
    uint32_t k0, k1;
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              PhiloxKey key, uint32_t (&out)[4])
{
    uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;  // IMAD.WIDE.U32
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;  // one LOP3
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0;
        c1 = (uint32_t)p1;
        c2 = n2;
        c3 = (uint32_t)p0;
        k0 += kPhiloxW0;  // key schedule depends on the seed only: hoisted to uniform regs
        k1 += kPhiloxW1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// ---- uniform bits -> standard normals (Box-Muller on the MUFU unit) -------------------
// Every arithmetic step is an explicit *_rn intrinsic or a single approx instruction so
// that the simulator kernel and the normal-dump kernel produce the same bits no matter
// how the surrounding code is scheduled (no context-dependent FMA contraction).
__device__ __forceinline__ float mufu_lg2(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sqrt(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sin(float x)
{
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_cos(float x)
{
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Two 32-bit words -> two independent N(0,1) draws.
//   radius word: low 23 bits -> f in [1,2) -> u = 2 - f in (0,1] -> r = sqrt(-2 ln u)
//   angle  word: low 23 bits -> g in [1,2) -> phi = 2 pi (g - 1.5) in [-pi, pi)
__device__ __forceinline__ void box_muller(uint32_t wr, uint32_t wa, float &z0, float &z1)
{
    const float f = __uint_as_float((wr & 0x007FFFFFu) | 0x3F800000u);
    const float u = __fsub_rn(2.0f, f);
    const float r = mufu_sqrt(__fmul_rn(mufu_lg2(u), -1.3862943611198906f));  // -2 ln 2 * lg2 u
    const float g = __uint_as_float((wa & 0x007FFFFFu) | 0x3F800000u);
    const float phi = __fmaf_rn(g, 6.283185307179586f, -9.42477796076938f);
    z0 = __fmul_rn(r, mufu_cos(phi));
    z1 = __fmul_rn(r, mufu_sin(phi));
}

// The four normals of steps 4*blk .. 4*blk+3 of global trial `trial`.
__device__ __forceinline__ void philox_normals4(uint32_t trial_lo, uint32_t trial_hi, uint32_t blk,
                                                PhiloxKey key, float (&z)[4])
{
    uint32_t w[4];
    philox4x32_10(trial_lo, trial_hi, blk, 0u, key, w);
    box_muller(w[0], w[1], z[0], z[1]);
    box_muller(w[2], w[3], z[2], z[3]);
}

}  // namespace ddm
