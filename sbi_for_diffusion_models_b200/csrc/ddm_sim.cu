// Pulse-driven drift-diffusion trial simulator for B200 (sm_100a).
//
// Replaces the lock-step torch loop of the reference
// (/root/reference/src/sbi_for_diffusion_models/models/rt_choice_model.py:112-221) with one
// persistent kernel:
//
//   * one LANE per trial, each lane on its own clock.  A lane whose trial has crossed a
//     bound (or run out of decision window) is refilled with the next trial from a global
//     queue at the next chunk boundary, so warps stay full although trial lengths vary
//     from 1 to n_max steps (mean ~5e3 of 16000 under the pipeline prior);
//   * noise is counter-based: Philox4x32-10 keyed by the seed with counter
//     (global trial index, step / 6) -> six 21-bit fields -> Box-Muller on the MUFU unit.  Which lane, warp,
//     launch or GPU runs a trial does not change its result;
//   * the fp32 arithmetic of a step is the reference's, operation for operation, with FMA
//     contraction forbidden (explicit *_rn intrinsics), so that feeding the same normals
//     to the reference reproduces the output bit for bit;
//   * pulse rows are loaded once per trial by the whole warp (coalesced 128-byte requests)
//     and packed to sign bits with warp ballots (80 pulses -> three registers);  a row that
//     holds anything other than +-1 is flagged and read back from global memory at kick
//     time instead, still bit-exact;
//   * a chunk is 24 Euler steps (four Philox blocks of six normals each).  The fast path only tracks the running
//     max / min of the accumulator; the per-step first-passage search runs once per trial,
//     in the chunk where max >= B, min <= 0 or the window ends.
#include "ddm_common.cuh"

#include <math_constants.h>

namespace ddm {

#ifndef DDM_SIM_THREADS
#define DDM_SIM_THREADS 256
#endif
#ifndef DDM_SIM_MIN_BLOCKS
#define DDM_SIM_MIN_BLOCKS 3  // measured on B200 (24-step chunks): 3 blocks (66 regs) 1.156e12 steps/s, 4: 1.143e12, 5 (spills): 1.105e12
#endif
constexpr int kThreads = DDM_SIM_THREADS;  // small CTAs retire sooner in the drain phase of a launch
constexpr unsigned kFull = 0xFFFFFFFFu;

struct SimParams {
    const float *theta;
    long long ld_theta;
    const float *pulses;
    long long ld_pulses;
    const float *noise;
    long long ld_noise;
    float *x_out;
    int *steps_out;
    unsigned long long *ws;
    unsigned int n_trials;
    int n_pulses;  // columns of the pulse matrix the schedule can reach
    int n_max;
    float t_max, t_nd_hi;
    // the four scalars of the hot loop side by side: one 128-bit uniform load per chunk
    alignas(16) float dt;
    float noise_scale;
    uint32_t one_bits;  // 0x3F800000, kept in a register for the one-LOP3 mantissa insert
    int spp;
    PhiloxKey key;
    unsigned long long trial_offset;
    int log_rt;
    const unsigned long long *ready;  // streaming mode: trials [0, *ready) have arrived in HBM (else null)
    unsigned long long wait_timeout_ns;  // streaming mode: give up on rows that have not arrived after this long
    float *x_peers[DDM_MAX_PEERS];  // fused all-gather: the same (rt, choice) also goes to these (peer-mapped) blocks
    int n_peers;
};

#ifndef DDM_SIM_NB
#define DDM_SIM_NB 4  // Philox blocks (x6 Euler steps) per chunk: 24 steps
#endif

// torch.clamp: NaN propagates (fminf/fmaxf would drop it)
__device__ __forceinline__ float clamp_keep_nan(float x, float lo, float hi)
{
    return x < lo ? lo : (x > hi ? hi : x);
}

// Streaming mode: sleep until the copy engine has delivered `need` rows (bounded: by default 20 s without
// progress sets the error word; ddm_sim_set_stream_timeout_us).  Only the STREAM instantiation of the kernel contains this spin
// loop: its presence makes ptxas give up uniform-register round keys in the hot loop (+3
// instructions per Euler step), which the resident-input kernel must not pay.
__device__ __forceinline__ int wait_for_rows(const unsigned long long *ready, unsigned long long need,
                                          unsigned long long *error_word, unsigned long long timeout_ns)
{
    const volatile unsigned long long *rdy = ready;
    unsigned long long t0 = 0ull, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*rdy < need) {
        __nanosleep(500);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) {
            atomicExch(error_word, 1ull);
            return 1;
        }
    }
    __threadfence();
    return 0;
}

// PACKED: p.theta points at 32-byte records [theta bits x 5, pulse sign masks x 3] made on the host by
// ddm_pack_z_host (the PCIe ingest path): no pulse matrix, every lane loads its own record.
template <int MASKW, bool INJECT, bool ALIGNED, int NB, bool STREAM, bool PACKED = false>
__global__ void __launch_bounds__(kThreads, STREAM ? 4 : DDM_SIM_MIN_BLOCKS) sim_kernel(const SimParams p)
{
    static_assert(!PACKED || (MASKW == 3 && !INJECT), "packed records carry three sign masks and native noise");
    constexpr int NPB = kNormalsPerBlock;
    constexpr int STEPS = NPB * NB;
    static_assert(STEPS % 8 == 0, "chunks must keep pulse kicks on multiples of 8 steps");
    static_assert(NB % 2 == 0 && (NB & (NB - 1)) == 0, "Philox blocks are consumed in pairs; pair + b / 2 == pair ^ (b / 2) needs a power of two");
    constexpr int MW = MASKW > 0 ? MASKW : 1;
    const unsigned lane = threadIdx.x & 31u;
    const BmConsts bm = make_bm_consts(p.noise_scale);
    float nz12[2 * NPB];

    // ---- per-lane trial state ---------------------------------------------------------
    // An idle lane carries a = NaN and rem = -1: none of the end-of-trial tests below can fire for it (NaN compares
    // false, fmaxf / fminf skip NaN, (unsigned)rem is huge), so the hot loop never has to ask whether a lane is busy.
    float a = CUDART_NAN_F, nlam = 0.f, B = 1.f, v = 0.f, tnd = 0.f;
    int rem = -1;    // steps of the decision window still ahead at the chunk's first step (nsteps - t); < 0: idle
    int nsteps = 0;  // decision window in steps
    int dk = 0;      // steps from the chunk's first step to the next pulse kick (tk - t)
    int pidx = 0;    // column of the next pulse (maintained only where the pulse VALUE is read from memory)
    uint32_t pair = 0u;  // Philox block pair of the chunk's first step (= t / 12)
    float kv = 0.f;     // signed value v * s[pidx] of the next kick
    PhiloxTrial pt{0u, 0u, {0u, 0u}, {0u, 0u}, {0u, 0u}};
    // sign masks, REVERSED and INVERTED: bit 31 of sgn[0] is set iff the NEXT pulse is -1, then bit 30, ..., then
    // sgn[1], sgn[2].  A kick shifts the chain left by one, so the next kick value is ONE LOP3: v ^ (sgn[0] & 2^31).
    uint32_t sgn[MW];
#pragma unroll
    for (int w = 0; w < MW; ++w) sgn[w] = 0u;
    uint32_t trial = 0;  // index into this launch's arrays
    bool generic = (MASKW == 0);  // kick reads the pulse value from global memory
    bool exhausted = false;       // warp-uniform: the queue has run dry
    unsigned long long useful = 0ull;
    unsigned int chunks = 0u;

    auto next_pulse = [&]() {   // a kick has happened: advance to the next pulse (rt_choice_model.py:190-192)
        if (MASKW > 0) {
#pragma unroll
            for (int w = 0; w + 1 < MW; ++w) sgn[w] = __funnelshift_l(sgn[w + 1], sgn[w], 1);
            sgn[MW - 1] <<= 1;
        }
    };
    auto signed_kick = [&]() -> float { return __uint_as_float(__float_as_uint(v) ^ (sgn[0] & 0x80000000u)); };  // v * (+-1) exactly
    auto loaded_kick = [&]() -> float {   // generic rows: the pulse value comes from memory
        float s = 0.0f;
        if (rem >= 0 && pidx < p.n_pulses) s = __ldcg(p.pulses + (long long)trial * p.ld_pulses + pidx);
        return __fmul_rn(v, s);
    };

    // ---- refill idle lanes from the global queue; returns true when the whole warp is (still) idle ----------
    // Called once before the loop and then only where lanes become idle: in the end-of-trial branch of a chunk.
    auto refill = [&]() -> bool {
        unsigned idle = __ballot_sync(kFull, rem < 0);
        if (idle != 0u && !exhausted) {
            const int want = __popc(idle);
            unsigned long long base = 0ull;
            if (lane == 0) base = atomicAdd(&p.ws[DDM_WS_QUEUE], (unsigned long long)want);
            base = __shfl_sync(kFull, base, 0);
            exhausted = (base + (unsigned long long)want >= (unsigned long long)p.n_trials);
            if (STREAM && base < (unsigned long long)p.n_trials) {
                // streaming mode: the copy engine is still delivering z; wait (bounded) until every
                // trial this warp just claimed has landed.  Copies never wait on this kernel.
                unsigned long long need = base + (unsigned long long)want;
                if (need > (unsigned long long)p.n_trials) need = (unsigned long long)p.n_trials;
                int failed = 0;
                if (lane == 0) failed = wait_for_rows(p.ready, need, &p.ws[DDM_WS_ERROR], p.wait_timeout_ns);
                failed = __shfl_sync(kFull, failed, 0);
                if (failed) {
                    exhausted = true;
                    base = (unsigned long long)p.n_trials;  // nobody gets a trial
                }
            }
            const unsigned long long mine = base + (unsigned long long)__popc(idle & ((1u << lane) - 1u));
            const bool got = rem < 0 && mine < (unsigned long long)p.n_trials;
            if (got) trial = (uint32_t)mine;

            if (MASKW > 0 && !PACKED) {
                // whole warp loads each new trial's pulse row: coalesced, then ballot -> sign bits
                unsigned todo = __ballot_sync(kFull, got);
                while (todo != 0u) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const uint32_t tj = __shfl_sync(kFull, trial, j);
                    const float *row = p.pulses + (long long)tj * p.ld_pulses;
                    bool odd = false;
#pragma unroll
                    for (int w = 0; w < MASKW; ++w) {
                        const int c = w * 32 + (int)lane;
                        const float s = (c < p.n_pulses) ? __ldcg(row + c) : 1.0f;
                        const unsigned neg = __ballot_sync(kFull, !(s > 0.0f));   // bit c: pulse c is not +1
                        odd = odd || (fabsf(s) != 1.0f);
                        if ((int)lane == j) sgn[w] = __brev(neg);
                    }
                    const unsigned any_odd = __ballot_sync(kFull, odd);
                    if ((int)lane == j) generic = (any_odd != 0u);
                }
            }
            if (got) {
                float th0, th1, th2, th3, th4;
                if (PACKED) {
                    const uint4 *rec = reinterpret_cast<const uint4 *>(p.theta) + 2ll * trial;
                    const uint4 lo = __ldcg(rec), hi = __ldcg(rec + 1);
                    th0 = __uint_as_float(lo.x), th1 = __uint_as_float(lo.y), th2 = __uint_as_float(lo.z);
                    th3 = __uint_as_float(lo.w), th4 = __uint_as_float(hi.x);
                    sgn[0] = __brev(~hi.y), sgn[1 % MW] = __brev(~hi.z), sgn[2 % MW] = __brev(~hi.w);   // records: bit c set = +1
                    generic = false;
                } else {
                    const float *th = p.theta + (long long)trial * p.ld_theta;
                    th0 = __ldcg(th + 0), th1 = __ldcg(th + 1), th2 = __ldcg(th + 2);
                    th3 = __ldcg(th + 3), th4 = __ldcg(th + 4);
                }
                // rt_choice_model.py:131-135
                const float a0 = clamp_keep_nan(th0, 0.0f, 1.0f);
                nlam = -th1;
                v = fabsf(th2);
                B = fabsf(th3);
                B = (B < 1e-6f) ? 1e-6f : B;
                tnd = clamp_keep_nan(th4, 0.0f, p.t_nd_hi);
                // :141  floor((T_MAX - t_nd) / dt) with a true IEEE division (CPU semantics)
                const float win = floorf(__fdiv_rn(__fsub_rn(p.t_max, tnd), p.dt));
                nsteps = !(win > 0.0f) ? 0 : (win >= (float)p.n_max ? p.n_max : (int)win);
                a = __fmul_rn(a0, B);  // :144
                rem = nsteps;
                pair = 0u;
                dk = 0;
                pidx = 0;
                kv = (MASKW == 0 || generic) ? loaded_kick() : signed_kick();
                if (!INJECT) {
                    const unsigned long long g = p.trial_offset + (unsigned long long)trial;
                    pt = philox_trial_setup((uint32_t)g, (uint32_t)(g >> 32), p.key);
                }
                if (MASKW > 0 && generic) atomicAdd(&p.ws[DDM_WS_GENERIC_ROWS], 1ull);
            }
            idle = __ballot_sync(kFull, rem < 0);
        }
        return idle == kFull;
    };

    if (refill()) goto done;
    for (;;) {
        // ---- one chunk of STEPS Euler steps -------------------------------------------
        auto kick = [&](float acc) -> float {
            // rt_choice_model.py:192  a += v * s[:, p_idx] * active   (then advance to the next pulse)
            acc = __fadd_rn(acc, kv);
            next_pulse();
            pidx += 1;
            dk += p.spp;
            kv = (MASKW == 0 || generic) ? loaded_kick() : signed_kick();
            return acc;
        };

        // ALIGNED (steps_per_pulse % 8 == 0 and >= STEPS): at most one kick per chunk, at i = dk, a
        // multiple of 8.  The kick sites are predicated adds; the pulse bookkeeping runs once per
        // chunk instead of once per site.
        const int dk0 = dk;
        float av[STEPS];
        float acc = a;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float z[NPB];
            if (INJECT) {
#pragma unroll
                for (int j = 0; j < NPB; ++j) {
                    const int step = nsteps - rem + NPB * b + j;
                    z[j] = (rem >= 0 && step < p.n_max)
                               ? __ldg(p.noise + (long long)step * p.ld_noise + trial)
                               : 0.0f;
                }
            } else if ((b & 1) == 0) {
                // two Philox blocks at a time: their Box-Muller pairs run as packed fp32 pairs, noise scale included
                philox_scaled_normals12_trial(pt, pair, (uint32_t)(b >> 1), p.key, p.one_bits, bm, nz12);
            }
#pragma unroll
            for (int j = 0; j < NPB; ++j) {
                const int i = NPB * b + j;
                const float nz = INJECT ? __fmul_rn(z[j], p.noise_scale) : nz12[NPB * (b & 1) + j];  // :186
                const float leak = __fmul_rn(__fmul_rn(nlam, acc), p.dt);      // (-lam*a)*dt
                acc = __fadd_rn(__fadd_rn(acc, leak), nz);                     // :187
                // :190-192; with steps_per_pulse % 8 == 0 a kick can only fall on i % 8 == 0
                if (ALIGNED) {
                    if ((i & 7) == 0) {
                        if (dk0 == i) acc = __fadd_rn(acc, kv);  // a += v * s[:, p_idx] * active (one predicated FADD)
                    }
                } else {
                    if (dk == i) acc = kick(acc);
                }
                av[i] = acc;
            }
        }
        a = acc;
        chunks += 1u;
        if (ALIGNED) {  // advance to the next pulse if this chunk held a kick
            const bool kicked = dk0 < STEPS;
            if (kicked) next_pulse();            // predicated shifts, no branch
            dk += kicked ? p.spp : 0;
            if (MASKW == 0 || generic) {         // rare (rows with a pulse value other than +-1) or MASKW == 0
                if (kicked) {
                    pidx += 1;
                    kv = loaded_kick();
                }
            } else {
                kv = signed_kick();
            }
        }
        dk -= STEPS;

        // fmaxf / fminf ignore NaN operands, like the reference's comparisons (always false).  Running max / min
        // per group of 7 steps (three 3-input FMNMX each): the exact search below only visits the group(s) that can
        // hold the crossing.
        constexpr int GS = 7, G = (STEPS + GS - 1) / GS;
        float ghi[G], glo[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int s0 = GS * g, s1 = (GS * g + GS < STEPS) ? GS * g + GS : STEPS;
            ghi[g] = glo[g] = av[s0];
#pragma unroll
            for (int i = s0 + 1; i < s1; i += 2) {
                if (i + 1 < s1) {
                    ghi[g] = fmaxf(fmaxf(ghi[g], av[i]), av[i + 1]);
                    glo[g] = fminf(fminf(glo[g], av[i]), av[i + 1]);
                } else {
                    ghi[g] = fmaxf(ghi[g], av[i]);
                    glo[g] = fminf(glo[g], av[i]);
                }
            }
        }
        float hi = ghi[0], lo = glo[0];
#pragma unroll
        for (int g = 1; g < G; ++g) {
            hi = fmaxf(hi, ghi[g]);
            lo = fminf(lo, glo[g]);
        }

        // idle lanes: hi = lo = NaN and (unsigned)rem >= 2^31, so all three tests are false without asking
        const bool ending = (hi >= B) | (lo <= 0.0f) | ((unsigned)rem <= (unsigned)STEPS);   // no short-circuit: no branch
        rem -= STEPS;
        pair += (uint32_t)(NB / 2);
        if (__any_sync(kFull, ending)) {
            const int rem0 = rem + STEPS;   // window left at the chunk's first step
            // ---- some trial of the warp ends inside the chunk: exact first-passage search -------
            // (warp-uniform control flow: typically ONE lane ends, in ONE group of 7 steps)
            int first = STEPS;       // first step of the chunk at which the accumulator is outside (0, B)
            float afirst = 1.0f;     // its value there
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const bool look = ending && first == STEPS && (ghi[g] >= B || glo[g] <= 0.0f);
                if (__any_sync(kFull, look)) {
                    const int s0 = GS * g, s1 = (GS * g + GS < STEPS) ? GS * g + GS : STEPS;
#pragma unroll
                    for (int i = s1 - 1; i >= s0; --i) {
                        if (look && (av[i] >= B || av[i] <= 0.0f)) {   // :195-196
                            first = i;
                            afirst = av[i];
                        }
                    }
                }
            }
            if (ending) {
                int hit_step, choice;
                if (first < rem0) {                  // inside the decision window (t + i < n_steps)
                    hit_step = nsteps - rem0 + first + 1;  // :201
                    choice = (afirst <= 0.0f) ? 0 : 1;     // lower bound wins ties, :202-203
                } else {                             // window over without a crossing, :206-215
                    hit_step = nsteps;
                    choice = 2;
                }
                // :218, then pack_x_rt_choice :338-342
                float rt = __fadd_rn(tnd, __fmul_rn((float)hit_step, p.dt));
                rt = clamp_keep_nan(rt, 1e-6f, p.t_max);
                rt = (rt < 1e-6f) ? 1e-6f : rt;
                if (p.log_rt) rt = logf(rt);
                const float2 res = make_float2(rt, (float)choice);
                reinterpret_cast<float2 *>(p.x_out)[trial] = res;
                // fused all-gather (ddm_sim_gather_f32): 8 bytes per trial and peer over NVLink, posted stores
                for (int d = 0; d < p.n_peers; ++d) reinterpret_cast<float2 *>(p.x_peers[d])[trial] = res;
                if (p.steps_out) p.steps_out[trial] = hit_step;
                useful += (unsigned long long)hit_step;
                a = CUDART_NAN_F;   // idle (see the state comment above)
                rem = -1;
            }
            if (refill()) break;
        }
    }
done:

    // ---- counters --------------------------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) useful += __shfl_xor_sync(kFull, useful, o);
    if (lane == 0) {
        atomicAdd(&p.ws[DDM_WS_USEFUL_STEPS], useful);
        atomicAdd(&p.ws[DDM_WS_LANE_STEPS], (unsigned long long)chunks * (unsigned long long)(STEPS * 32));
    }
}

// ---- small batches: producer / consumer CTAs ---------------------------------------------
// With few trials the persistent kernel above cannot fill the machine: every warp is alone on its scheduler and a
// trial is a serial chain (Philox + Box-Muller + update: measured 54 cycles per Euler step), so a launch drains at
// the pace of its longest trial (16 000 steps: 0.44 ms) however few trials there are -- the regime of
// simulate_observed_session and of single SBC sessions (T = 50).
// Here a CTA takes 32 trials that advance in lock-step: ONE consumer warp (lane = trial) runs nothing but the
// recurrence -- four dependent fp32 ops per step, noise read from shared memory -- while SEVEN producer warps
// generate the noise of the next 42 steps (lane = trial, warp = Philox block of the chunk) into the other half
// of a double buffer.  Same Philox indexing, same operations in the same order: bit-identical to sim_kernel
// (tested).  Measured: 0.38 ms against 0.44 ms for 50 .. 1000 trials (46 cycles per step: the lone consumer warp
// issues ~14 instructions per step), break-even near 1e4 trials; used for launches of up to 8192 trials.
constexpr int kSmallTrials = 32;
constexpr int kSmallProducers = 7;                                   // warps = Philox blocks per chunk
constexpr int kSmallSteps = kSmallProducers * kNormalsPerBlock;      // 42 steps per chunk
constexpr int kSmallThreads = 32 * (1 + kSmallProducers);

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__global__ void __launch_bounds__(kSmallThreads) sim_small_kernel(const SimParams p)
{
    __shared__ float nz_s[2][kSmallSteps][kSmallTrials];
    __shared__ unsigned int busy_s;   // bit l: trial l of this CTA is still running (consumer -> producers)
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned int trial = blockIdx.x * kSmallTrials + lane;
    const bool have = trial < p.n_trials;
    constexpr int kFullBar = 1, kEmptyBar = 3;   // + buffer index; barrier 0 is __syncthreads
    if (threadIdx.x == 0) busy_s = 0xFFFFFFFFu;
    __syncthreads();

    if (warp > 0) {
        // ---------------- producers: block (7 k + warp - 1) of trial `lane`, chunk k -----------------
        const unsigned long long g = p.trial_offset + (unsigned long long)trial;
        const PhiloxTrial pt = philox_trial_setup((uint32_t)g, (uint32_t)(g >> 32), p.key);
        const int w = (int)warp - 1;
        for (int k = 0;; ++k) {
            const int b = k & 1;
            if (k >= 2) named_bar_sync(kEmptyBar + b, kSmallThreads);   // the consumer has read chunk k - 2
            const unsigned int busy = *reinterpret_cast<volatile unsigned int *>(&busy_s);
            if (busy == 0u) break;
            if ((busy >> lane) & 1u) {
                float z[kNormalsPerBlock];
                philox_normals6_trial(pt, (uint32_t)(kSmallProducers * k + w), p.key, p.one_bits, z);
#pragma unroll
                for (int j = 0; j < kNormalsPerBlock; ++j)
                    nz_s[b][kNormalsPerBlock * w + j][lane] = __fmul_rn(z[j], p.noise_scale);   // :186
            }
            named_bar_arrive(kFullBar + b, kSmallThreads);
        }
        return;
    }

    // -------------------- consumer: the recurrence of 32 trials, lane = trial --------------------
    float a = 0.f, nlam = 0.f, B = 1.f, v = 0.f, tnd = 0.f, kv = 0.f;
    int nsteps = 0;
    const float *prow = p.pulses + (long long)(have ? trial : 0) * p.ld_pulses;
    if (have) {
        const float *th = p.theta + (long long)trial * p.ld_theta;
        const float th0 = __ldcg(th + 0), th1 = __ldcg(th + 1), th2 = __ldcg(th + 2), th3 = __ldcg(th + 3), th4 = __ldcg(th + 4);
        const float a0 = clamp_keep_nan(th0, 0.0f, 1.0f);   // rt_choice_model.py:131-135
        nlam = -th1;
        v = fabsf(th2);
        B = fabsf(th3);
        B = (B < 1e-6f) ? 1e-6f : B;
        tnd = clamp_keep_nan(th4, 0.0f, p.t_nd_hi);
        const float win = floorf(__fdiv_rn(__fsub_rn(p.t_max, tnd), p.dt));   // :141
        nsteps = !(win > 0.0f) ? 0 : (win >= (float)p.n_max ? p.n_max : (int)win);
        a = __fmul_rn(a0, B);                                                   // :144
        kv = __fmul_rn(v, p.n_pulses > 0 ? __ldcg(prow) : 0.0f);
    }
    bool done = !have || nsteps == 0;
    int hit_step = done ? 0 : -1, choice = 2;
    int t = 0, tk = 0, pidx = 0;   // all trials of the warp share the clock, so kicks are warp-uniform
    int chunks = 0;
    for (int k = 0;; ++k) {
        const int b = k & 1;
        const unsigned int busy = __ballot_sync(kFull, !done);
        if (lane == 0) *reinterpret_cast<volatile unsigned int *>(&busy_s) = busy;
        if (k >= 1) named_bar_arrive(kEmptyBar + ((k - 1) & 1), kSmallThreads);   // chunk k - 1 has been read
        if (busy == 0u) {
            // let the producers (one chunk ahead, waiting for the OTHER buffer's release next) see busy == 0
            named_bar_arrive(kEmptyBar + b, kSmallThreads);
            break;
        }
        named_bar_sync(kFullBar + b, kSmallThreads);
        chunks += 1;
        // six steps at a time: their noise is fetched up front (off the dependent chain) and the pulse-kick test is
        // one warp-uniform branch per group instead of one per step
#pragma unroll 1
        for (int i0 = 0; i0 < kSmallSteps; i0 += kNormalsPerBlock) {
            float nzv[kNormalsPerBlock];
#pragma unroll
            for (int j = 0; j < kNormalsPerBlock; ++j) nzv[j] = nz_s[b][i0 + j][lane];
            auto step = [&](float nz, bool may_kick) {
                const float leak = __fmul_rn(__fmul_rn(nlam, a), p.dt);   // (-lam*a)*dt
                a = __fadd_rn(__fadd_rn(a, leak), nz);                    // :187
                if (may_kick && t == tk) {                                // :190-192 (warp-uniform)
                    a = __fadd_rn(a, kv);
                    tk += p.spp;
                    pidx += 1;
                    kv = __fmul_rn(v, (have && pidx < p.n_pulses) ? __ldcg(prow + pidx) : 0.0f);
                }
                const bool up = a >= B, dn = a <= 0.0f;                   // :195-196
                if (!done && t < nsteps && (up || dn)) {
                    hit_step = t + 1;                                     // :201
                    choice = dn ? 0 : 1;                                  // lower bound wins ties
                    done = true;
                }
                t += 1;
            };
            if (tk - t >= kNormalsPerBlock) {
#pragma unroll
                for (int j = 0; j < kNormalsPerBlock; ++j) step(nzv[j], false);
            } else {
#pragma unroll
                for (int j = 0; j < kNormalsPerBlock; ++j) step(nzv[j], true);
            }
        }
        if (!done && t >= nsteps) {   // window over without a crossing, :206-215
            hit_step = nsteps;
            choice = 2;
            done = true;
        }
    }
    if (have) {
        if (hit_step < 0) hit_step = nsteps;
        float rt = __fadd_rn(tnd, __fmul_rn((float)hit_step, p.dt));   // :218, then pack_x_rt_choice :338-342
        rt = clamp_keep_nan(rt, 1e-6f, p.t_max);
        rt = (rt < 1e-6f) ? 1e-6f : rt;
        if (p.log_rt) rt = logf(rt);
        reinterpret_cast<float2 *>(p.x_out)[trial] = make_float2(rt, (float)choice);
        if (p.steps_out) p.steps_out[trial] = hit_step;
    }
    unsigned long long useful = have ? (unsigned long long)hit_step : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) useful += __shfl_xor_sync(kFull, useful, o);
    if (lane == 0) {
        atomicAdd(&p.ws[DDM_WS_USEFUL_STEPS], useful);
        atomicAdd(&p.ws[DDM_WS_LANE_STEPS], (unsigned long long)chunks * (unsigned long long)(kSmallSteps * 32));
    }
}

// ---- z rows -> 32-byte records on the device (the device->host mirror of the packed ingest) -------------
// One warp per row: coalesced loads of the row, ballots -> three sign masks, lane 0 writes the record
// [theta bits x 5, masks x 3] (same format as ddm_pack_z_host); rows holding a value other than +-1 are counted.
__global__ void __launch_bounds__(256) pack_z_kernel(const float *__restrict__ z, long long ld, long long n_rows, int n_pulses,
                                                     uint32_t *__restrict__ packed, unsigned long long *__restrict__ generic_rows)
{
    const unsigned lane = threadIdx.x & 31u;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        const float *row = z + r * ld;
        uint32_t m[3];
        bool odd = false;
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const int c = 32 * w + (int)lane;
            const float s = c < n_pulses ? __ldcs(row + 5 + c) : 1.0f;
            m[w] = __ballot_sync(kFull, s > 0.0f);
            odd = odd || (fabsf(s) != 1.0f);
        }
        const bool any_odd = __any_sync(kFull, odd);
        const float th = lane < 5 ? __ldcs(row + lane) : 0.0f;
        uint32_t word = __float_as_uint(th);
        if (lane >= 5 && lane < 8) word = m[lane - 5];
        if (lane < 8) packed[r * 8 + lane] = word;
        if (lane == 0 && any_odd) atomicAdd(generic_rows, 1ull);
    }
}

// ---- dump kernels: the noise stream as a tensor ---------------------------------------
template <bool WORDS>
__global__ void __launch_bounds__(256) philox_dump_kernel(PhiloxKey key, uint32_t one, unsigned long long trial_offset,
                                                          long long n_trials, long long n_steps,
                                                          void *out, long long ld)
{
    constexpr int PER = WORDS ? 4 : kNormalsPerBlock;
    const long long n_blk = (n_steps + PER - 1) / PER;
    const long long total = n_blk * n_trials;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long blk = idx / n_trials;
        const long long i = idx - blk * n_trials;
        const unsigned long long g = trial_offset + (unsigned long long)i;
        if (WORDS) {
            uint32_t w[4];
            philox4x32_10((uint32_t)g, (uint32_t)(blk >> 1), (uint32_t)(g >> 32), (uint32_t)(blk & 1), key, w);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (blk * 4 + j < n_steps) static_cast<uint32_t *>(out)[(blk * 4 + j) * ld + i] = w[j];
        } else {
            float z[kNormalsPerBlock];
            philox_normals6((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)blk, key, one, z);
#pragma unroll
            for (int j = 0; j < kNormalsPerBlock; ++j)
                if (blk * kNormalsPerBlock + j < n_steps)
                    static_cast<float *>(out)[(blk * kNormalsPerBlock + j) * ld + i] = z[j];
        }
    }
}

// ---- launch ----------------------------------------------------------------------------
template <int MASKW, bool INJECT, bool ALIGNED, bool STREAM, bool PACKED = false>
static int launch_sim(const SimParams &p, int sm_count, cudaStream_t stream)
{
    auto kern = sim_kernel<MASKW, INJECT, ALIGNED, DDM_SIM_NB, STREAM, PACKED>;
    int per_sm = 0;
    DDM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const long long warps_needed = ((long long)p.n_trials + 31) / 32;
    const long long resident_warps = (long long)sm_count * per_sm * (kThreads / 32);
    int block = kThreads;
    long long grid;
    if (warps_needed >= resident_warps) {
        grid = (long long)sm_count * per_sm;  // persistent: every resident slot, lanes refill
    } else {
        // few trials: spread the warps over as many SMs as possible
        long long wpb = (warps_needed + sm_count - 1) / sm_count;
        if (wpb < 1) wpb = 1;
        if (wpb > kThreads / 32) wpb = kThreads / 32;
        block = (int)wpb * 32;
        grid = (warps_needed + wpb - 1) / wpb;
    }
    kern<<<(unsigned)grid, block, 0, stream>>>(p);
    DDM_CUDA_TRY(cudaGetLastError());
    return DDM_OK;
}

}  // namespace ddm

using namespace ddm;

static unsigned long long g_stream_timeout_ns = 20000000000ull;
static long long g_small_max_trials = 8192;    // launches of at most this many trials take sim_small_kernel

DDM_API int ddm_sim_set_small_batch_max(int64_t max_trials)
{
    DDM_REQUIRE(max_trials >= 0 && max_trials <= 0x7FFFFFFFll, "ddm_sim_set_small_batch_max: %lld out of range", (long long)max_trials);
    g_small_max_trials = (long long)max_trials;
    return DDM_OK;
}

DDM_API int ddm_sim_set_stream_timeout_us(int64_t timeout_us)
{
    DDM_REQUIRE(timeout_us >= 1000 && timeout_us <= 600000000ll, "ddm_sim_set_stream_timeout_us: %lld outside [1 ms, 600 s]",
                (long long)timeout_us);
    g_stream_timeout_ns = (unsigned long long)timeout_us * 1000ull;
    return DDM_OK;
}

DDM_API size_t ddm_sim_workspace_bytes(void) { return DDM_WS_WORDS * sizeof(unsigned long long); }

static int sim_impl(const float *theta_dev, int64_t ld_theta, const float *pulses_dev,
                    int64_t ld_pulses, int64_t N, int64_t P, int64_t n_max,
                    int64_t steps_per_pulse, float dt, float t_max, float t_nd_hi,
                    float noise_scale, uint64_t seed, uint64_t trial_offset,
                    const float *noise_dev, int64_t ld_noise, int log_rt, float *x_out_dev,
                    int32_t *steps_out_dev, void *workspace_dev, const uint64_t *ready_dev, void *stream,
                    float *const *x_peers = nullptr, int n_peers = 0)
{
    DDM_REQUIRE(n_peers >= 0 && n_peers <= DDM_MAX_PEERS && (n_peers == 0 || x_peers != nullptr),
                "ddm_sim_gather_f32: n_peers=%d outside [0, %d]", n_peers, DDM_MAX_PEERS);
    for (int d = 0; d < n_peers; ++d)
        DDM_REQUIRE(x_peers[d] != nullptr && (reinterpret_cast<uintptr_t>(x_peers[d]) & 7u) == 0,
                    "ddm_sim_gather_f32: peer block %d must be a non-null, 8-byte aligned device pointer", d);
    DDM_REQUIRE(N >= 0 && N <= 0x7FFFFFFFll, "ddm_sim_f32: N=%lld outside [0, 2^31)", (long long)N);
    // (an idle lane counts its window register down from -1 by 24 per chunk for as long as the longest trial of its
    // warp runs: n_max <= 2^30 keeps that above INT_MIN)
    DDM_REQUIRE(n_max >= 0 && n_max <= 0x40000000ll, "ddm_sim_f32: n_max=%lld outside [0, 2^30]", (long long)n_max);
    DDM_REQUIRE(steps_per_pulse >= 1 && steps_per_pulse <= 0x7FFFFFFFll,
                "ddm_sim_f32: steps_per_pulse=%lld must be >= 1", (long long)steps_per_pulse);
    const int64_t need = (n_max + steps_per_pulse - 1) / steps_per_pulse;
    DDM_REQUIRE(P >= need, "ddm_sim_f32: pulse matrix has P=%lld columns but the schedule needs %lld",
                (long long)P, (long long)need);
    DDM_REQUIRE(dt > 0.0f && t_max > 0.0f, "ddm_sim_f32: dt and t_max must be positive");
    DDM_REQUIRE(workspace_dev != nullptr && (reinterpret_cast<uintptr_t>(workspace_dev) & 7u) == 0,
                "ddm_sim_f32: workspace must be a non-null, 8-byte aligned device pointer");
    if (N > 0) {
        DDM_REQUIRE(theta_dev && pulses_dev && x_out_dev, "ddm_sim_f32: null theta / pulses / x_out");
        DDM_REQUIRE(ld_theta == 0 || ld_theta >= 5, "ddm_sim_f32: ld_theta=%lld (0 = broadcast, else >= 5)", (long long)ld_theta);
        DDM_REQUIRE(ld_pulses == 0 || ld_pulses >= need, "ddm_sim_f32: ld_pulses=%lld < %lld",
                    (long long)ld_pulses, (long long)need);
        DDM_REQUIRE((reinterpret_cast<uintptr_t>(x_out_dev) & 7u) == 0, "ddm_sim_f32: x_out must be 8-byte aligned");
        DDM_REQUIRE(noise_dev == nullptr || ld_noise >= N, "ddm_sim_f32: ld_noise=%lld < N", (long long)ld_noise);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DDM_CUDA_TRY(cudaMemsetAsync(workspace_dev, 0, ddm_sim_workspace_bytes(), st));
    if (N == 0) return DDM_OK;

    int dev = 0, sms = 0;
    DDM_CUDA_TRY(cudaGetDevice(&dev));
    DDM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    SimParams p;
    p.theta = theta_dev;
    p.ld_theta = ld_theta;
    p.pulses = pulses_dev;
    p.ld_pulses = ld_pulses;
    p.noise = noise_dev;
    p.ld_noise = ld_noise;
    p.x_out = x_out_dev;
    p.steps_out = steps_out_dev;
    p.ws = static_cast<unsigned long long *>(workspace_dev);
    p.n_trials = (unsigned int)N;
    p.n_pulses = (int)need;
    p.n_max = (int)n_max;
    p.spp = (int)steps_per_pulse;
    p.dt = dt;
    p.t_max = t_max;
    p.t_nd_hi = t_nd_hi;
    p.noise_scale = noise_scale;
    p.key = make_philox_key(seed);
    p.trial_offset = trial_offset;
    p.log_rt = log_rt ? 1 : 0;
    p.one_bits = 0x3F800000u;
    p.ready = reinterpret_cast<const unsigned long long *>(ready_dev);
    p.wait_timeout_ns = g_stream_timeout_ns;
    p.n_peers = n_peers;
    for (int d = 0; d < DDM_MAX_PEERS; ++d) p.x_peers[d] = d < n_peers ? x_peers[d] : nullptr;

    const bool inject = noise_dev != nullptr;
    if (!inject && ready_dev == nullptr && n_peers == 0 && N <= g_small_max_trials) {
        // small batch: producer / consumer CTAs (bit-identical results, ~10x lower latency)
        sim_small_kernel<<<(unsigned)((N + kSmallTrials - 1) / kSmallTrials), kSmallThreads, 0, st>>>(p);
        DDM_CUDA_TRY(cudaGetLastError());
        return DDM_OK;
    }
    const bool aligned = (steps_per_pulse % 8) == 0 && steps_per_pulse >= kNormalsPerBlock * DDM_SIM_NB;
    const bool packed = need <= 96;
#define DDM_PICK(MW, INJ, AL, ST) return launch_sim<MW, INJ, AL, ST>(p, sms, st)
    if (ready_dev != nullptr) {  // streaming ingest: native noise only
        if (packed) { if (aligned) DDM_PICK(3, false, true, true); else DDM_PICK(3, false, false, true); }
        else        { if (aligned) DDM_PICK(0, false, true, true); else DDM_PICK(0, false, false, true); }
    }
    if (packed) {
        if (inject) { if (aligned) DDM_PICK(3, true, true, false); else DDM_PICK(3, true, false, false); }
        else        { if (aligned) DDM_PICK(3, false, true, false); else DDM_PICK(3, false, false, false); }
    } else {
        if (inject) { if (aligned) DDM_PICK(0, true, true, false); else DDM_PICK(0, true, false, false); }
        else        { if (aligned) DDM_PICK(0, false, true, false); else DDM_PICK(0, false, false, false); }
    }
#undef DDM_PICK
}

DDM_API int ddm_sim_f32(const float *theta_dev, int64_t ld_theta, const float *pulses_dev,
                        int64_t ld_pulses, int64_t N, int64_t P, int64_t n_max,
                        int64_t steps_per_pulse, float dt, float t_max, float t_nd_hi,
                        float noise_scale, uint64_t seed, uint64_t trial_offset,
                        const float *noise_dev, int64_t ld_noise, int log_rt, float *x_out_dev,
                        int32_t *steps_out_dev, void *workspace_dev, void *stream)
{
    return sim_impl(theta_dev, ld_theta, pulses_dev, ld_pulses, N, P, n_max, steps_per_pulse, dt, t_max, t_nd_hi,
                    noise_scale, seed, trial_offset, noise_dev, ld_noise, log_rt, x_out_dev, steps_out_dev,
                    workspace_dev, nullptr, stream);
}

DDM_API int ddm_sim_gather_f32(const float *theta_dev, int64_t ld_theta, const float *pulses_dev,
                               int64_t ld_pulses, int64_t N, int64_t P, int64_t n_max,
                               int64_t steps_per_pulse, float dt, float t_max, float t_nd_hi,
                               float noise_scale, uint64_t seed, uint64_t trial_offset, int log_rt,
                               float *x_out_dev, float *const *x_peer_blocks, int n_peers, void *workspace_dev,
                               void *stream)
{
    return sim_impl(theta_dev, ld_theta, pulses_dev, ld_pulses, N, P, n_max, steps_per_pulse, dt, t_max, t_nd_hi,
                    noise_scale, seed, trial_offset, nullptr, 0, log_rt, x_out_dev, nullptr, workspace_dev, nullptr, stream,
                    x_peer_blocks, n_peers);
}

DDM_API int ddm_sim_stream_f32(const float *theta_dev, int64_t ld_theta, const float *pulses_dev,
                               int64_t ld_pulses, int64_t N, int64_t P, int64_t n_max,
                               int64_t steps_per_pulse, float dt, float t_max, float t_nd_hi,
                               float noise_scale, uint64_t seed, uint64_t trial_offset, int log_rt,
                               float *x_out_dev, void *workspace_dev, const uint64_t *ready_dev, void *stream)
{
    DDM_REQUIRE(ready_dev != nullptr && (reinterpret_cast<uintptr_t>(ready_dev) & 7u) == 0,
                "ddm_sim_stream_f32: ready_dev must be a non-null, 8-byte aligned device pointer");
    return sim_impl(theta_dev, ld_theta, pulses_dev, ld_pulses, N, P, n_max, steps_per_pulse, dt, t_max, t_nd_hi,
                    noise_scale, seed, trial_offset, nullptr, 0, log_rt, x_out_dev, nullptr, workspace_dev,
                    ready_dev, stream);
}

DDM_API int ddm_sim_packed_f32(const uint32_t *packed_dev, int64_t N, int64_t n_max, int64_t steps_per_pulse, float dt,
                               float t_max, float t_nd_hi, float noise_scale, uint64_t seed, uint64_t trial_offset,
                               int log_rt, float *x_out_dev, int32_t *steps_out_dev, void *workspace_dev,
                               const uint64_t *ready_dev, void *stream)
{
    DDM_REQUIRE(N >= 0 && N <= 0x7FFFFFFFll, "ddm_sim_packed_f32: N=%lld outside [0, 2^31)", (long long)N);
    DDM_REQUIRE(n_max >= 0 && n_max <= 0x40000000ll, "ddm_sim_packed_f32: n_max=%lld outside [0, 2^30]", (long long)n_max);
    DDM_REQUIRE(steps_per_pulse >= 1 && steps_per_pulse <= 0x7FFFFFFFll,
                "ddm_sim_packed_f32: steps_per_pulse=%lld must be >= 1", (long long)steps_per_pulse);
    const int64_t need = (n_max + steps_per_pulse - 1) / steps_per_pulse;
    DDM_REQUIRE(need <= 96, "ddm_sim_packed_f32: the schedule needs %lld pulses, a record holds 96 sign bits",
                (long long)need);
    DDM_REQUIRE(dt > 0.0f && t_max > 0.0f, "ddm_sim_packed_f32: dt and t_max must be positive");
    DDM_REQUIRE(workspace_dev != nullptr && (reinterpret_cast<uintptr_t>(workspace_dev) & 7u) == 0,
                "ddm_sim_packed_f32: workspace must be a non-null, 8-byte aligned device pointer");
    DDM_REQUIRE(ready_dev == nullptr || (reinterpret_cast<uintptr_t>(ready_dev) & 7u) == 0,
                "ddm_sim_packed_f32: ready_dev must be 8-byte aligned");
    if (N > 0) {
        DDM_REQUIRE(packed_dev && x_out_dev, "ddm_sim_packed_f32: null records / x_out");
        DDM_REQUIRE((reinterpret_cast<uintptr_t>(packed_dev) & 15u) == 0, "ddm_sim_packed_f32: records must be 16-byte aligned");
        DDM_REQUIRE((reinterpret_cast<uintptr_t>(x_out_dev) & 7u) == 0, "ddm_sim_packed_f32: x_out must be 8-byte aligned");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DDM_CUDA_TRY(cudaMemsetAsync(workspace_dev, 0, ddm_sim_workspace_bytes(), st));
    if (N == 0) return DDM_OK;
    int dev = 0, sms = 0;
    DDM_CUDA_TRY(cudaGetDevice(&dev));
    DDM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SimParams p;
    p.theta = reinterpret_cast<const float *>(packed_dev);
    p.ld_theta = 8;
    p.pulses = nullptr;
    p.ld_pulses = 0;
    p.noise = nullptr;
    p.ld_noise = 0;
    p.x_out = x_out_dev;
    p.steps_out = steps_out_dev;
    p.ws = static_cast<unsigned long long *>(workspace_dev);
    p.n_trials = (unsigned int)N;
    p.n_pulses = (int)need;
    p.n_max = (int)n_max;
    p.spp = (int)steps_per_pulse;
    p.dt = dt;
    p.t_max = t_max;
    p.t_nd_hi = t_nd_hi;
    p.noise_scale = noise_scale;
    p.key = make_philox_key(seed);
    p.trial_offset = trial_offset;
    p.log_rt = log_rt ? 1 : 0;
    p.one_bits = 0x3F800000u;
    p.ready = reinterpret_cast<const unsigned long long *>(ready_dev);
    p.wait_timeout_ns = g_stream_timeout_ns;
    p.n_peers = 0;
    for (int d = 0; d < DDM_MAX_PEERS; ++d) p.x_peers[d] = nullptr;
    const bool aligned = (steps_per_pulse % 8) == 0 && steps_per_pulse >= kNormalsPerBlock * DDM_SIM_NB;
    if (ready_dev != nullptr) {
        if (aligned) return launch_sim<3, false, true, true, true>(p, sms, st);
        return launch_sim<3, false, false, true, true>(p, sms, st);
    }
    if (aligned) return launch_sim<3, false, true, false, true>(p, sms, st);
    return launch_sim<3, false, false, false, true>(p, sms, st);
}

static int dump_common(bool words, uint64_t seed, uint64_t trial_offset, int64_t N, int64_t n_steps,
                       void *out_dev, int64_t ld_out, void *stream)
{
    DDM_REQUIRE(N >= 0 && n_steps >= 0, "philox dump: negative size");
    DDM_REQUIRE(ld_out >= N, "philox dump: ld_out=%lld < N=%lld", (long long)ld_out, (long long)N);
    if (N == 0 || n_steps == 0) return DDM_OK;
    DDM_REQUIRE(out_dev != nullptr, "philox dump: null output");
    const long long per = words ? 4 : kNormalsPerBlock;
    const long long total = ((n_steps + per - 1) / per) * N;
    long long grid = (total + 255) / 256;
    if (grid > 148 * 64) grid = 148 * 64;
    const PhiloxKey key = make_philox_key(seed);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (words)
        philox_dump_kernel<true><<<(unsigned)grid, 256, 0, st>>>(key, 0x3F800000u, trial_offset, N, n_steps, out_dev, ld_out);
    else
        philox_dump_kernel<false><<<(unsigned)grid, 256, 0, st>>>(key, 0x3F800000u, trial_offset, N, n_steps, out_dev, ld_out);
    DDM_CUDA_TRY(cudaGetLastError());
    return DDM_OK;
}

DDM_API int ddm_philox_normals_f32(uint64_t seed, uint64_t trial_offset, int64_t N, int64_t n_steps,
                                   float *out_dev, int64_t ld_out, void *stream)
{
    return dump_common(false, seed, trial_offset, N, n_steps, out_dev, ld_out, stream);
}

DDM_API int ddm_philox_words_u32(uint64_t seed, uint64_t trial_offset, int64_t N, int64_t n_steps,
                                 uint32_t *out_dev, int64_t ld_out, void *stream)
{
    return dump_common(true, seed, trial_offset, N, n_steps, out_dev, ld_out, stream);
}

DDM_API int ddm_pack_z_dev(const float *z_dev, int64_t ld, int64_t N, int64_t n_pulses, uint32_t *packed_dev,
                           uint64_t *generic_rows_dev, void *stream)
{
    DDM_REQUIRE(N >= 0 && n_pulses >= 0 && n_pulses <= 96 && ld >= 5 + n_pulses,
                "ddm_pack_z_dev: bad arguments (N=%lld, n_pulses=%lld in [0,96], ld=%lld >= 5 + n_pulses)", (long long)N,
                (long long)n_pulses, (long long)ld);
    DDM_REQUIRE(generic_rows_dev != nullptr && (reinterpret_cast<uintptr_t>(generic_rows_dev) & 7u) == 0,
                "ddm_pack_z_dev: generic_rows_dev must be a non-null, 8-byte aligned device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DDM_CUDA_TRY(cudaMemsetAsync(generic_rows_dev, 0, sizeof(uint64_t), st));
    if (N == 0) return DDM_OK;
    DDM_REQUIRE(z_dev && packed_dev, "ddm_pack_z_dev: null pointer");
    long long blocks = (N + 7) / 8;   // 8 warps per block, one row per warp and pass
    if (blocks > 148 * 64) blocks = 148 * 64;
    pack_z_kernel<<<(unsigned)blocks, 256, 0, st>>>(z_dev, ld, N, (int)n_pulses, packed_dev,
                                                   reinterpret_cast<unsigned long long *>(generic_rows_dev));
    DDM_CUDA_TRY(cudaGetLastError());
    return DDM_OK;
}
