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
// counter = (lo32(trial), block >> 1, hi32(trial), block & 1), key = (lo32(seed), hi32(seed)); the simulator
// draws the normals of steps 6 * block .. 6 * block + 5 from one block (see normals6 below).  The index of a
// PAIR of blocks sits in counter word 1 and the block's place in its pair in word 3 -- the two words the first
// round only XORs -- so that most of rounds 1-3 is constant along a trial or shared by the two blocks of a pair
// (PhiloxTrial below).
#ifndef DDM_PHILOX_ROUNDS
#define DDM_PHILOX_ROUNDS 10   // the library ships Philox4x32-10; other values only for the timing experiment of DESIGN 3.1
                               // (the dump kernels and the oracle stay at ten rounds, so every replay test fails)
#endif
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

// One Philox block (128 bits) feeds SIX normals: six 21-bit fields = three Box-Muller pairs (radius field, angle
// field).  A field becomes the low mantissa bits of 1.0f, i.e. f in [1, 1.25) on a 2^-23 grid.  The fields are cut so
// that four of the six cost ONE instruction each (per 64-bit half {lo, hi} of the block):
//   low  field = lo[0:21]              one LOP3:   (lo & 0x001FFFFF) | one
//   top  field = hi[11:32]             one LEA.HI: one + (hi >> 11)
//   mid  field = lo[21:32] ++ hi[0:10] funnel shift by 21, then the LOP3 (bit 10 of hi is not used)
// `one` (= 0x3F800000) must sit in a register, because a LOP3 can carry only one immediate; callers pass
// it from a kernel parameter so that the compiler cannot fold it back into a second immediate.
constexpr int kNormalsPerBlock = 6;

__device__ __forceinline__ float field_to_1_125(uint32_t w, uint32_t one)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, 0x001FFFFF, %2, 0xEA;" : "=r"(r) : "r"(w), "r"(one));
    return __uint_as_float(r);
}
__device__ __forceinline__ float top_field_to_1_125(uint32_t w, uint32_t one) { return __uint_as_float(one + (w >> 11)); }
__device__ __forceinline__ float mid_field_to_1_125(uint32_t lo, uint32_t hi, uint32_t one)
{
    return field_to_1_125(__funnelshift_r(lo, hi, 21), one);
}

// Two fields (already floats in [1, 1.25)) -> two independent N(0,1) draws.
//   radius field: f in [1,1.25) -> u = 5 - 4f in [2^-21, 1] (exact) -> r = sqrt(-2 ln u) <= 5.4
//   angle  field: g in [1,1.25) -> phi = 2 pi (4 (g - 1) - 0.5) = 8 pi g - 9 pi in [-pi, pi)
__device__ __forceinline__ void box_muller(float f, float g, float &z0, float &z1)
{
    const float u = __fmaf_rn(f, -4.0f, 5.0f);
    const float r = mufu_sqrt(__fmul_rn(mufu_lg2(u), -1.3862943611198906f));  // -2 ln 2 * lg2 u
    const float phi = __fmaf_rn(g, 25.132741228718345f, -28.274333882308138f);
    z0 = __fmul_rn(r, mufu_cos(phi));
    z1 = __fmul_rn(r, mufu_sin(phi));
}

// 128 bits -> six normals: (low, top) of words 0-1, (mid of words 0-1, mid of words 2-3), (low, top) of words 2-3
__device__ __forceinline__ void normals6(const uint32_t (&w)[4], uint32_t one, float (&z)[6])
{
    box_muller(field_to_1_125(w[0], one), top_field_to_1_125(w[1], one), z[0], z[1]);
    box_muller(mid_field_to_1_125(w[0], w[1], one), mid_field_to_1_125(w[2], w[3], one), z[2], z[3]);
    box_muller(field_to_1_125(w[2], one), top_field_to_1_125(w[3], one), z[4], z[5]);
}

// The six normals of steps 6*blk .. 6*blk+5 of global trial `trial`.
__device__ __forceinline__ void philox_normals6(uint32_t trial_lo, uint32_t trial_hi, uint32_t blk,
                                                const PhiloxKey &key, uint32_t one, float (&z)[6])
{
    uint32_t w[4];
    philox4x32_10(trial_lo, blk >> 1, trial_hi, blk & 1u, key, w);
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

// Two Box-Muller transforms side by side: the same arithmetic as box_muller() on (fr0, fa0) and on
// (fr1, fa1) -- fields already in [1, 1.25) --, with the five affine / product steps of both issued as packed pairs, and the result already
// multiplied by the simulator's noise scale (one more packed product: nz = z * scale, rt_choice_model.py:186).
//   out: (z(fr0,fa0).cos, z(fr1,fa1).cos) in zc, (..sin, ..sin) in zs -- scaled.
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
__device__ __forceinline__ void box_muller_x2_scaled(float fr0, float fa0, float fr1, float fa1,
                                                     const BmConsts &k, float &c0, float &s0, float &c1, float &s1)
{
    const f32x2 u = fma2(pack2(fr0, fr1), k.m4, k.p5);
    float u0, u1;
    unpack2(u, u0, u1);
    const f32x2 t = mul2(pack2(mufu_lg2(u0), mufu_lg2(u1)), k.lg);
    float t0, t1;
    unpack2(t, t0, t1);
    const f32x2 r = pack2(mufu_sqrt(t0), mufu_sqrt(t1));
    const f32x2 phi = fma2(pack2(fa0, fa1), k.ang_a, k.ang_b);
    float p0, p1;
    unpack2(phi, p0, p1);
    const f32x2 zc = mul2(mul2(r, pack2(mufu_cos(p0), mufu_cos(p1))), k.scale);
    const f32x2 zs = mul2(mul2(r, pack2(mufu_sin(p0), mufu_sin(p1))), k.scale);
    unpack2(zc, c0, c1);
    unpack2(zs, s0, s1);
}

// ---- Philox with the trial-constant part of rounds 1-3 hoisted -------------------------
// With counter (g_lo, pair, g_hi, s) -- pair = blk >> 1, s = blk & 1 -- only `pair` and `s` change along a trial,
// and round 1 merely XORs them into words 0 and 2:
//   round 1: both products (M0 g_lo, M1 g_hi) are constant;        n0 = A ^ pair,  n2 = N2 ^ s (two constants)
//   round 2: M1 * n2 is constant per s, M0 * n0 varies with pair;  m2 = hi(M0 n0) ^ F, m3 = lo(M0 n0)   (both blocks)
//   round 3: M0 * m0 is constant per s, M1 * m2 varies with pair;  q0 = hi(M1 m2) ^ D[s], q1 = lo(M1 m2),
//                                                                  q2 = H[s] ^ m3,        q3 = E[s]
// so a PAIR of blocks costs two wide multiplies and two LOP3 for rounds 1-3 together, then two LOP3 per block, then
// seven full rounds per block: 30 + 34 instructions per pair instead of 40 + 40.  Same output bits as philox4x32_10
// on that counter (exact integer algebra); eight words per trial.
struct PhiloxTrial {
    uint32_t a;     // hi(M1 g_hi) ^ k0[0]
    uint32_t f;     // lo(M0 g_lo) ^ k1[1]
    uint32_t d[2];  // lo(M1 n2[s]) ^ k0[2]                 with n2[s] = hi(M0 g_lo) ^ s ^ k1[0]
    uint32_t h[2];  // hi(M0 m0[s]) ^ k1[2]                 with m0[s] = hi(M1 n2[s]) ^ lo(M1 g_hi) ^ k0[1]
    uint32_t e[2];  // lo(M0 m0[s])
};

__device__ __forceinline__ PhiloxTrial philox_trial_setup(uint32_t g_lo, uint32_t g_hi, const PhiloxKey &key)
{
    const uint64_t p0 = (uint64_t)kPhiloxM0 * g_lo;
    const uint64_t p1 = (uint64_t)kPhiloxM1 * g_hi;
    const uint32_t n1 = (uint32_t)p1;
    PhiloxTrial t;
    t.a = (uint32_t)(p1 >> 32) ^ key.k0[0];
    t.f = (uint32_t)p0 ^ key.k1[1];
#pragma unroll
    for (uint32_t s = 0; s < 2; ++s) {
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ s ^ key.k1[0];
        const uint64_t r1 = (uint64_t)kPhiloxM1 * n2;
        const uint32_t m0 = (uint32_t)(r1 >> 32) ^ n1 ^ key.k0[1];
        const uint64_t r0 = (uint64_t)kPhiloxM0 * m0;
        t.d[s] = (uint32_t)r1 ^ key.k0[2];
        t.h[s] = (uint32_t)(r0 >> 32) ^ key.k1[2];
        t.e[s] = (uint32_t)r0;
    }
    return t;
}

// rounds 4..10 from the state after round 3
__device__ __forceinline__ void philox_rounds_4_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey &key,
                                                   uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 3; r < DDM_PHILOX_ROUNDS; ++r) {
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

// Blocks 2 (pair ^ sub) and 2 (pair ^ sub) + 1.  Callers whose `pair` has its low bits clear pass the small constant
// part as `sub` (pair + sub == pair ^ sub then), which folds into the one three-input LOP3 of round 1.
__device__ __forceinline__ void philox4x32_10_pair(const PhiloxTrial &t, uint32_t pair, uint32_t sub, const PhiloxKey &key,
                                                   uint32_t (&out0)[4], uint32_t (&out1)[4])
{
    const uint64_t q0 = (uint64_t)kPhiloxM0 * (t.a ^ pair ^ sub);   // rounds 1-2 (only M0 * n0 varies)
    const uint32_t m2 = (uint32_t)(q0 >> 32) ^ t.f;
    const uint32_t m3 = (uint32_t)q0;
    const uint64_t q1 = (uint64_t)kPhiloxM1 * m2;                    // round 3 (only M1 * m2 varies)
    philox_rounds_4_10((uint32_t)(q1 >> 32) ^ t.d[0], (uint32_t)q1, t.h[0] ^ m3, t.e[0], key, out0);
    philox_rounds_4_10((uint32_t)(q1 >> 32) ^ t.d[1], (uint32_t)q1, t.h[1] ^ m3, t.e[1], key, out1);
}

// one block (the small-batch kernel: its warps take single blocks)
__device__ __forceinline__ void philox4x32_10_trial(const PhiloxTrial &t, uint32_t blk, const PhiloxKey &key,
                                                    uint32_t (&out)[4])
{
    const bool s = (blk & 1u) != 0u;
    const uint64_t q0 = (uint64_t)kPhiloxM0 * (t.a ^ (blk >> 1));
    const uint32_t m2 = (uint32_t)(q0 >> 32) ^ t.f;
    const uint32_t m3 = (uint32_t)q0;
    const uint64_t q1 = (uint64_t)kPhiloxM1 * m2;
    philox_rounds_4_10((uint32_t)(q1 >> 32) ^ (s ? t.d[1] : t.d[0]), (uint32_t)q1, (s ? t.h[1] : t.h[0]) ^ m3,
                       s ? t.e[1] : t.e[0], key, out);
}

__device__ __forceinline__ void philox_normals6_trial(const PhiloxTrial &t, uint32_t blk, const PhiloxKey &key,
                                                      uint32_t one, float (&z)[6])
{
    uint32_t w[4];
    philox4x32_10_trial(t, blk, key, w);
    normals6(w, one, z);
}

// Blocks 2 (pair ^ sub) and 2 (pair ^ sub) + 1 -> the twelve SCALED normals (z * noise_scale) of the twelve steps
// from 12 (pair ^ sub) on, same bits as philox_normals6_trial + __fmul_rn per normal: Box-Muller pair j of the
// first block runs side by side with pair j of the second one.
__device__ __forceinline__ void philox_scaled_normals12_trial(const PhiloxTrial &t, uint32_t pair, uint32_t sub,
                                                              const PhiloxKey &key, uint32_t one, const BmConsts &k,
                                                              float (&nz)[12])
{
    uint32_t a[4], b[4];
    philox4x32_10_pair(t, pair, sub, key, a, b);
    box_muller_x2_scaled(field_to_1_125(a[0], one), top_field_to_1_125(a[1], one), field_to_1_125(b[0], one),
                         top_field_to_1_125(b[1], one), k, nz[0], nz[1], nz[6], nz[7]);
    box_muller_x2_scaled(mid_field_to_1_125(a[0], a[1], one), mid_field_to_1_125(a[2], a[3], one),
                         mid_field_to_1_125(b[0], b[1], one), mid_field_to_1_125(b[2], b[3], one), k, nz[2], nz[3], nz[8], nz[9]);
    box_muller_x2_scaled(field_to_1_125(a[2], one), top_field_to_1_125(a[3], one), field_to_1_125(b[2], one),
                         top_field_to_1_125(b[3], one), k, nz[4], nz[5], nz[10], nz[11]);
}

}  // namespace ddm
