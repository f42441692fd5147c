// NumPy-compatible pulse-side generator on the device.
//
// Replaces the per-trial Python loop of the reference
// (/root/reference/src/sbi_for_diffusion_models/models/rt_choice_model.py:62-91 calling
// models/choice_model.py:43-60; reached from proposals.py:30-40) which draws, per trial,
// one double for the correct side and P doubles for the per-pulse successes from a NumPy
// PCG64 generator.  Trial i owns draws [i (P+1), (i+1)(P+1)) of that stream, so every trial
// can start from the stream state jumped ahead by i (P+1) draws: the output is the SAME
// matrix NumPy produces, bit for bit (integer arithmetic only; the double compare
// u < p is done as k < ceil(p 2^53) on the 53-bit integer k).
//
// PCG64 (O'Neill 2014, as in numpy/random/src/pcg64): 128-bit LCG
//   state' = state * 0x2360ED051FC65DA44385DF649FCCF645 + inc,
// output XSL-RR: rotr64(hi ^ lo, state' >> 122).
//
// Layout: one lane per trial generates its 1+P draws into sign bits; the warp then writes
// the 32 rows cooperatively so global stores are full 128-byte lines.
#include "ddm_common.cuh"

namespace ddm {

typedef unsigned __int128 u128;

constexpr int kJumpBits = 56;  // first_trial * (P + 1) must stay below 2^56 draws

struct PcgJumpTable {
    // LCG composed 2^k times: state -> mult[k] * state + plus[k]
    uint64_t mult_hi[kJumpBits], mult_lo[kJumpBits], plus_hi[kJumpBits], plus_lo[kJumpBits];
};

__host__ __device__ __forceinline__ u128 make128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }

__host__ __device__ __forceinline__ u128 pcg_mult()
{
    return make128(0x2360ED051FC65DA4ull, 0x4385DF649FCCF645ull);
}

__device__ __forceinline__ uint64_t pcg_next53(u128 &state, u128 inc)
{
    state = state * pcg_mult() + inc;
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    const uint64_t out = (x >> rot) | (x << ((64u - rot) & 63u));
    return out >> 11;  // the integer k of NumPy's double k * 2^-53
}

constexpr int kMaxMaskWords = 8;  // P <= 256 handled with sign bits in registers

template <int MW>
__global__ void __launch_bounds__(256) pulses_kernel(PcgJumpTable tab, uint64_t st_hi, uint64_t st_lo,
                                                     uint64_t inc_hi, uint64_t inc_lo,
                                                     unsigned long long first_trial, long long n, int P,
                                                     unsigned long long threshold, float *out, long long ld)
{
    const unsigned lane = threadIdx.x & 31u;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const u128 inc = make128(inc_hi, inc_lo);

    for (long long row0 = warp_global * 32; row0 < n; row0 += n_warps * 32) {
        const long long i = row0 + lane;
        uint32_t bits[MW];
#pragma unroll
        for (int w = 0; w < MW; ++w) bits[w] = 0u;
        if (i < n) {
            // jump to this trial's first draw
            u128 st = make128(st_hi, st_lo);
            unsigned long long delta = (first_trial + (unsigned long long)i) * (unsigned long long)(P + 1);
            for (int k = 0; delta != 0ull; ++k, delta >>= 1)
                if (delta & 1ull)
                    st = st * make128(tab.mult_hi[k], tab.mult_lo[k]) + make128(tab.plus_hi[k], tab.plus_lo[k]);
            const bool side_pos = pcg_next53(st, inc) < (1ull << 52);  // rng.random() < 0.5
#pragma unroll
            for (int w = 0; w < MW; ++w) {
                uint32_t acc = 0u;
                for (int b = 0; b < 32; ++b) {
                    const int j = w * 32 + b;
                    if (j < P) {
                        const bool ok = pcg_next53(st, inc) < threshold;  // rng.random() < p_success
                        acc |= (uint32_t)(ok == side_pos) << b;           // where(ok, side, -side) > 0
                    }
                }
                bits[w] = acc;
            }
        }
        // cooperative, coalesced row writes
        const int rows = (int)((n - row0) < 32 ? (n - row0) : 32);
        for (int r = 0; r < rows; ++r) {
            float *dst = out + (row0 + r) * ld;
#pragma unroll
            for (int w = 0; w < MW; ++w) {
                const uint32_t m = __shfl_sync(0xFFFFFFFFu, bits[w], r);
                const int j = w * 32 + (int)lane;
                if (j < P) dst[j] = ((m >> lane) & 1u) ? 1.0f : -1.0f;
            }
        }
    }
}

static void build_jump_table(u128 inc, PcgJumpTable &tab)
{
    u128 m = pcg_mult(), c = inc;
    for (int k = 0; k < kJumpBits; ++k) {
        tab.mult_hi[k] = (uint64_t)(m >> 64);
        tab.mult_lo[k] = (uint64_t)m;
        tab.plus_hi[k] = (uint64_t)(c >> 64);
        tab.plus_lo[k] = (uint64_t)c;
        c = (m + 1) * c;
        m = m * m;
    }
}

}  // namespace ddm

using namespace ddm;

DDM_API int ddm_pulses_pcg64(uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                             uint64_t first_trial, int64_t n, int64_t P, uint64_t threshold,
                             float *out_dev, int64_t ld, void *stream)
{
    DDM_REQUIRE(n >= 0, "ddm_pulses_pcg64: n_trials must be >= 0");
    DDM_REQUIRE(P >= 0, "ddm_pulses_pcg64: n_pulses must be >= 0");
    DDM_REQUIRE(P <= 32 * kMaxMaskWords, "ddm_pulses_pcg64: P=%lld > %d unsupported", (long long)P,
                32 * kMaxMaskWords);
    if (n == 0 || P == 0) return DDM_OK;  // generate_pulse_sides returns before drawing when P <= 0
    DDM_REQUIRE(out_dev != nullptr && ld >= P, "ddm_pulses_pcg64: bad output / ld");
    const unsigned __int128 last = ((unsigned __int128)first_trial + (unsigned __int128)n) * (unsigned __int128)(P + 1);
    DDM_REQUIRE((last >> kJumpBits) == 0, "ddm_pulses_pcg64: stream offset beyond 2^%d draws", kJumpBits);

    PcgJumpTable tab;
    build_jump_table(make128(inc_hi, inc_lo), tab);
    const long long warps = (n + 31) / 32;
    long long grid = (warps + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mw = (int)((P + 31) / 32);
#define DDM_PULSES(MW)                                                                              \
    pulses_kernel<MW><<<(unsigned)grid, 256, 0, st>>>(tab, state_hi, state_lo, inc_hi, inc_lo,      \
                                                       first_trial, n, (int)P, threshold, out_dev, ld)
    switch (mw) {
        case 1: DDM_PULSES(1); break;
        case 2: DDM_PULSES(2); break;
        case 3: DDM_PULSES(3); break;
        case 4: DDM_PULSES(4); break;
        default: DDM_PULSES(kMaxMaskWords); break;
    }
#undef DDM_PULSES
    DDM_CUDA_TRY(cudaGetLastError());
    return DDM_OK;
}

DDM_API int ddm_pcg64_advance(uint64_t *state_hi, uint64_t *state_lo, uint64_t inc_hi, uint64_t inc_lo,
                              uint64_t draws)
{
    DDM_REQUIRE(state_hi && state_lo, "ddm_pcg64_advance: null state");
    u128 st = make128(*state_hi, *state_lo);
    u128 m = pcg_mult(), c = make128(inc_hi, inc_lo);
    for (uint64_t d = draws; d != 0; d >>= 1) {
        if (d & 1) st = st * m + c;
        c = (m + 1) * c;
        m = m * m;
    }
    *state_hi = (uint64_t)(st >> 64);
    *state_lo = (uint64_t)st;
    return DDM_OK;
}
