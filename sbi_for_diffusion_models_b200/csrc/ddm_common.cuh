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
// counter = (lo32(trial), hi32(trial), block, 0), key = (lo32(seed), hi32(seed)); the simulator draws
// the normals of steps 6 * block .. 6 * block + 5 from one block (see normals6 below).
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// The key schedule (k + r * W) depends on the seed only: the host expands it once and the
// kernels read the ten round keys straight from the constant bank (one LOP3 operand each).
struct PhiloxKey {
    uint32_t k0[10], k1[10];
};

inline PhiloxKey make_philox_key(uint64_t seed)
{
    PhiloxKey k;
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = (uint32_t)seed + (uint32_t)r * kPhiloxW0;
        k.k1[r] = (uint32_t)(seed >> 32) + (uint32_t)r * kPhiloxW1;
    }
    return k;
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKey &key, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;  // IMAD.WIDE.U32
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[r];  // one LOP3
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
        c0 = n0;
        c1 = (uint32_t)p1;
        c2 = n2;
        c3 = (uint32_t)p0;
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

// One Philox block (128 bits) feeds SIX normals: six 21-bit fields at bit offsets 0, 21, ..., 105,
// i.e. three Box-Muller pairs (radius field, angle field).  A field is dropped into the low mantissa
// bits of 1.0f with ONE LOP3: (w & 0x001FFFFF) | one gives f in [1, 1.25) on a 2^-23 grid.  `one`
// (= 0x3F800000) must sit in a register, because a LOP3 can carry only one immediate; callers pass
// it from a kernel parameter so that the compiler cannot fold it back into a second immediate.
constexpr int kNormalsPerBlock = 6;

__device__ __forceinline__ float field_to_1_125(uint32_t w, uint32_t one)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, 0x001FFFFF, %2, 0xEA;" : "=r"(r) : "r"(w), "r"(one));
    return __uint_as_float(r);
}

// Two 21-bit fields -> two independent N(0,1) draws.
//   radius field: f in [1,1.25) -> u = 5 - 4f in [2^-21, 1] (exact) -> r = sqrt(-2 ln u) <= 5.4
//   angle  field: g in [1,1.25) -> phi = 2 pi (4 (g - 1) - 0.5) = 8 pi g - 9 pi in [-pi, pi)
__device__ __forceinline__ void box_muller(uint32_t wr, uint32_t wa, uint32_t one, float &z0, float &z1)
{
    const float f = field_to_1_125(wr, one);
    const float u = __fmaf_rn(f, -4.0f, 5.0f);
    const float r = mufu_sqrt(__fmul_rn(mufu_lg2(u), -1.3862943611198906f));  // -2 ln 2 * lg2 u
    const float g = field_to_1_125(wa, one);
    const float phi = __fmaf_rn(g, 25.132741228718345f, -28.274333882308138f);
    z0 = __fmul_rn(r, mufu_cos(phi));
    z1 = __fmul_rn(r, mufu_sin(phi));
}

// 128 bits -> six normals (fields that straddle two words come out of one funnel shift)
__device__ __forceinline__ void normals6(const uint32_t (&w)[4], uint32_t one, float (&z)[6])
{
    box_muller(w[0], __funnelshift_r(w[0], w[1], 21), one, z[0], z[1]);
    box_muller(__funnelshift_r(w[1], w[2], 10), __funnelshift_r(w[1], w[2], 31), one, z[2], z[3]);
    box_muller(__funnelshift_r(w[2], w[3], 20), w[3] >> 9, one, z[4], z[5]);
}

// The six normals of steps 6*blk .. 6*blk+5 of global trial `trial`.
__device__ __forceinline__ void philox_normals6(uint32_t trial_lo, uint32_t trial_hi, uint32_t blk,
                                                const PhiloxKey &key, uint32_t one, float (&z)[6])
{
    uint32_t w[4];
    philox4x32_10(trial_lo, trial_hi, blk, 0u, key, w);
    normals6(w, one, z);
}

// ---- Philox with the trial-constant part of rounds 1-2 hoisted -------------------------
// With counter (g_lo, g_hi, blk, 0) only `blk` changes along a trial.  Round 1 multiplies
// M0 * g_lo (constant per trial) and round 2 multiplies M1 * (hi(M0 g_lo) ^ k1) (also constant),
// so four words per trial replace two IMAD.WIDE and one LOP3 in every block.  Same output bits
// as philox4x32_10 (exact integer algebra).
struct PhiloxTrial {
    uint32_t a;  // g_hi ^ k0[0]
    uint32_t d;  // hi(M1 * c2') ^ k0[1]          with c2' = hi(M0 g_lo) ^ k1[0]
    uint32_t e;  // lo(M1 * c2')
    uint32_t f;  // lo(M0 g_lo) ^ k1[1]
};

__device__ __forceinline__ PhiloxTrial philox_trial_setup(uint32_t g_lo, uint32_t g_hi, const PhiloxKey &key)
{
    const uint64_t p0 = (uint64_t)kPhiloxM0 * g_lo;
    const uint32_t c2p = (uint32_t)(p0 >> 32) ^ key.k1[0];  // c3 = 0
    const uint64_t p1 = (uint64_t)kPhiloxM1 * c2p;
    PhiloxTrial t;
    t.a = g_hi ^ key.k0[0];
    t.d = (uint32_t)(p1 >> 32) ^ key.k0[1];
    t.e = (uint32_t)p1;
    t.f = (uint32_t)p0 ^ key.k1[1];
    return t;
}

__device__ __forceinline__ void philox4x32_10_trial(const PhiloxTrial &t, uint32_t blk, const PhiloxKey &key,
                                                    uint32_t (&out)[4])
{
    // round 1 (only M1 * blk varies)
    const uint64_t q1 = (uint64_t)kPhiloxM1 * blk;
    const uint32_t r1c0 = (uint32_t)(q1 >> 32) ^ t.a;
    const uint32_t r1c1 = (uint32_t)q1;
    // round 2 (only M0 * c0' varies)
    const uint64_t q0 = (uint64_t)kPhiloxM0 * r1c0;
    uint32_t c0 = t.d ^ r1c1;
    uint32_t c1 = t.e;
    uint32_t c2 = (uint32_t)(q0 >> 32) ^ t.f;
    uint32_t c3 = (uint32_t)q0;
#pragma unroll
    for (int r = 2; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
        c0 = n0;
        c1 = (uint32_t)p1;
        c2 = n2;
        c3 = (uint32_t)p0;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

__device__ __forceinline__ void philox_normals6_trial(const PhiloxTrial &t, uint32_t blk, const PhiloxKey &key,
                                                      uint32_t one, float (&z)[6])
{
    uint32_t w[4];
    philox4x32_10_trial(t, blk, key, w);
    normals6(w, one, z);
}

}  // namespace ddm
