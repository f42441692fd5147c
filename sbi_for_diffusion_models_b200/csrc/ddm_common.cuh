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
// counter = (lo32(trial), block, hi32(trial), 0), key = (lo32(seed), hi32(seed)); the simulator draws
// the normals of steps 6 * block .. 6 * block + 5 from one block (see normals6 below).  The block index sits
// in counter word 1 -- a word the first round only XORs -- so that most of rounds 1-3 is constant along a
// trial (PhiloxTrial below).
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// The key schedule (k + r * W) depends on the seed only: the host expands it once and the
// kernels read the ten round keys straight from the constant bank (one LOP3 operand each).
struct PhiloxKey {
    uint32_t k0[10], k1[10];
    // rounds 3..9 once more, interleaved (k0[3], k1[3], k0[4], ...) and 16-byte aligned: the simulator's hot loop
    // fetches them with four 128-bit uniform loads per chunk
    alignas(16) uint32_t hot[16];
};

inline PhiloxKey make_philox_key(uint64_t seed)
{
    PhiloxKey k;
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = (uint32_t)seed + (uint32_t)r * kPhiloxW0;
        k.k1[r] = (uint32_t)(seed >> 32) + (uint32_t)r * kPhiloxW1;
    }
    for (int r = 3; r < 10; ++r) {
        k.hot[2 * (r - 3)] = k.k0[r];
        k.hot[2 * (r - 3) + 1] = k.k1[r];
    }
    k.hot[14] = k.hot[15] = 0u;
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
    philox4x32_10(trial_lo, blk, trial_hi, 0u, key, w);
    normals6(w, one, z);
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2) -----------------------------------
// One instruction, two IEEE round-to-nearest results (no flush to zero): per lane the same bits as the
// scalar __fmaf_rn / __fmul_rn / __fadd_rn, at half the issue slots.  The simulator kernel is
// issue-bound, so the Box-Muller affine steps and products of TWO pairs of fields go through these.
struct f32x2 {
    unsigned long long v;
};
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 a, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}

// Two Box-Muller transforms side by side: the same arithmetic as box_muller() on (wr0, wa0) and on
// (wr1, wa1), with the five affine / product steps of both issued as packed pairs, and the result already
// multiplied by the simulator's noise scale (one more packed product: nz = z * scale, rt_choice_model.py:186).
//   out: (z(wr0,wa0).cos, z(wr1,wa1).cos) in zc, (..sin, ..sin) in zs -- scaled.
struct BmConsts {  // constant pairs, built once per kernel (registers or the constant bank)
    f32x2 m4, p5, lg, ang_a, ang_b, scale;
};
__device__ __forceinline__ BmConsts make_bm_consts(float noise_scale)
{
    BmConsts c;
    c.m4 = pack2(-4.0f, -4.0f);
    c.p5 = pack2(5.0f, 5.0f);
    c.lg = pack2(-1.3862943611198906f, -1.3862943611198906f);
    c.ang_a = pack2(25.132741228718345f, 25.132741228718345f);
    c.ang_b = pack2(-28.274333882308138f, -28.274333882308138f);
    c.scale = pack2(noise_scale, noise_scale);
    return c;
}
__device__ __forceinline__ void box_muller_x2_scaled(uint32_t wr0, uint32_t wa0, uint32_t wr1, uint32_t wa1, uint32_t one,
                                                     const BmConsts &k, float &c0, float &s0, float &c1, float &s1)
{
    const f32x2 u = fma2(pack2(field_to_1_125(wr0, one), field_to_1_125(wr1, one)), k.m4, k.p5);
    float u0, u1;
    unpack2(u, u0, u1);
    const f32x2 t = mul2(pack2(mufu_lg2(u0), mufu_lg2(u1)), k.lg);
    float t0, t1;
    unpack2(t, t0, t1);
    const f32x2 r = pack2(mufu_sqrt(t0), mufu_sqrt(t1));
    const f32x2 phi = fma2(pack2(field_to_1_125(wa0, one), field_to_1_125(wa1, one)), k.ang_a, k.ang_b);
    float p0, p1;
    unpack2(phi, p0, p1);
    const f32x2 zc = mul2(mul2(r, pack2(mufu_cos(p0), mufu_cos(p1))), k.scale);
    const f32x2 zs = mul2(mul2(r, pack2(mufu_sin(p0), mufu_sin(p1))), k.scale);
    unpack2(zc, c0, c1);
    unpack2(zs, s0, s1);
}

// ---- Philox with the trial-constant part of rounds 1-3 hoisted -------------------------
// With counter (g_lo, blk, g_hi, 0) only `blk` changes along a trial, and round 1 merely XORs it into word 0:
//   round 1: both products (M0 g_lo, M1 g_hi) are constant;            n0 = A ^ blk
//   round 2: M1 * n2 is constant, M0 * n0 varies;                      m2 = hi(M0 n0) ^ F, m3 = lo(M0 n0)
//   round 3: M0 * m0 is constant, M1 * m2 varies;                      q0 = hi(M1 m2) ^ D, q1 = lo(M1 m2), q2 = H ^ m3, q3 = E
// so five words per trial replace four IMAD.WIDE and three LOP3 in every block (16 + 18 instructions per block
// instead of 20 + 20).  Same output bits as philox4x32_10 on that counter (exact integer algebra).
struct PhiloxTrial {
    uint32_t a;  // hi(M1 g_hi) ^ k0[0]
    uint32_t f;  // lo(M0 g_lo) ^ k1[1]
    uint32_t d;  // lo(M1 n2) ^ k0[2]                 with n2 = hi(M0 g_lo) ^ k1[0]
    uint32_t h;  // hi(M0 m0) ^ k1[2]                 with m0 = hi(M1 n2) ^ lo(M1 g_hi) ^ k0[1]
    uint32_t e;  // lo(M0 m0)
};

__device__ __forceinline__ PhiloxTrial philox_trial_setup(uint32_t g_lo, uint32_t g_hi, const PhiloxKey &key)
{
    const uint64_t p0 = (uint64_t)kPhiloxM0 * g_lo;
    const uint64_t p1 = (uint64_t)kPhiloxM1 * g_hi;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ key.k1[0];  // c3 = 0
    const uint64_t r1 = (uint64_t)kPhiloxM1 * n2;
    const uint32_t m0 = (uint32_t)(r1 >> 32) ^ n1 ^ key.k0[1];
    const uint64_t r0 = (uint64_t)kPhiloxM0 * m0;
    PhiloxTrial t;
    t.a = (uint32_t)(p1 >> 32) ^ key.k0[0];
    t.f = (uint32_t)p0 ^ key.k1[1];
    t.d = (uint32_t)r1 ^ key.k0[2];
    t.h = (uint32_t)(r0 >> 32) ^ key.k1[2];
    t.e = (uint32_t)r0;
    return t;
}

// The block index is blk ^ sub: callers whose `blk` has its low bits clear pass the small constant part as `sub`
// (blk + sub == blk ^ sub then), which folds into the one three-input LOP3 of round 1.
__device__ __forceinline__ void philox4x32_10_trial(const PhiloxTrial &t, uint32_t blk, const PhiloxKey &key,
                                                    uint32_t (&out)[4], uint32_t sub = 0u)
{
    // round 1 (blk enters by XOR only), round 2 (only M0 * n0 varies)
    const uint64_t q0 = (uint64_t)kPhiloxM0 * (t.a ^ blk ^ sub);
    const uint32_t m2 = (uint32_t)(q0 >> 32) ^ t.f;
    const uint32_t m3 = (uint32_t)q0;
    // round 3 (only M1 * m2 varies)
    const uint64_t q1 = (uint64_t)kPhiloxM1 * m2;
    uint32_t c0 = (uint32_t)(q1 >> 32) ^ t.d;
    uint32_t c1 = (uint32_t)q1;
    uint32_t c2 = t.h ^ m3;
    uint32_t c3 = t.e;
#pragma unroll
    for (int r = 3; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.hot[2 * (r - 3)];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.hot[2 * (r - 3) + 1];
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

// Blocks blk + sub and blk + sub + 1 -> the twelve SCALED normals (z * noise_scale) of the twelve steps from
// 6 (blk + sub) on, same bits as philox_normals6_trial + __fmul_rn per normal: Box-Muller pair j of the first
// block runs side by side with pair j of the second one.  Requires blk % (sub + 2) == 0 with sub + 2 a power of
// two (the simulator: blk a multiple of its blocks per chunk, sub even).
__device__ __forceinline__ void philox_scaled_normals12_trial(const PhiloxTrial &t, uint32_t blk, uint32_t sub,
                                                              const PhiloxKey &key, uint32_t one, const BmConsts &k,
                                                              float (&nz)[12])
{
    uint32_t a[4], b[4];
    philox4x32_10_trial(t, blk, key, a, sub);
    philox4x32_10_trial(t, blk, key, b, sub + 1u);
    box_muller_x2_scaled(a[0], __funnelshift_r(a[0], a[1], 21), b[0], __funnelshift_r(b[0], b[1], 21), one, k, nz[0], nz[1],
                         nz[6], nz[7]);
    box_muller_x2_scaled(__funnelshift_r(a[1], a[2], 10), __funnelshift_r(a[1], a[2], 31), __funnelshift_r(b[1], b[2], 10),
                         __funnelshift_r(b[1], b[2], 31), one, k, nz[2], nz[3], nz[8], nz[9]);
    box_muller_x2_scaled(__funnelshift_r(a[2], a[3], 20), a[3] >> 9, __funnelshift_r(b[2], b[3], 20), b[3] >> 9, one, k, nz[4],
                         nz[5], nz[10], nz[11]);
}

}  // namespace ddm
