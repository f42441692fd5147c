// Bookkeeping of the many-chain slice sampler (samplers.py::VectorizedSliceSampler) as two kernels around the
// potential call of an iteration -- the device side of what the reference hands to sbi's MCMCPosterior
// (/root/reference/src/sbi_for_diffusion_models/mnle.py:77-93; its notebook uses sbi's slice_np_vectorized).
//
// Every chain runs its own state machine (Neal 2003: stepping out with a limit, then shrinkage):
//   slice_propose : the point chain c needs evaluated next -> row c of the query matrix
//   [ potential(query) -> f ]                                 (mnle_loglik_sum_batched_tc_f32 + prior, in torch)
//   slice_update  : chain c's move given f[c]; finished coordinate updates tune the width, advance the coordinate,
//                   count sweeps, record a draw and start the next update
// The arithmetic is written with explicit round-to-nearest intrinsics in the order of the torch reference
// implementation in samplers.py, so both paths give the same bits (tested).
#include "ddm_common.cuh"

namespace ddm {

struct SliceState {
    float *x;          // (N, D) current points
    float *lp;         // (N) log-probability at x
    float *width;      // (N, D) slice widths
    float *tuned;      // (N, D) updates averaged into width so far
    float *lo, *hi;    // (N) bracket
    float *x0;         // (N) coordinate value the bracket was built around
    float *log_y;      // (N) slice level
    float *J, *K;      // (N) expansions left on the left / right
    long long *d;      // (N) coordinate being updated
    long long *sweeps; // (N) finished sweeps
    long long *taken;  // (N) draws recorded
    long long *phase;  // (N) 1 stepping out left, 2 right, 3 shrinking
    long long *nshr;   // (N) shrinkage proposals of this update
    float *out;        // (S, N, D) recorded draws
    long long N;
    int D, S, thin, warmup, total, max_step_out, max_shrink;
};

__device__ __forceinline__ float slice_query(const SliceState &s, long long c, float u0)
{
    const long long ph = s.phase[c];
    const float lo = s.lo[c], hi = s.hi[c];
    if (ph == 1) return lo;
    if (ph == 2) return hi;
    return __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u0));
}

__global__ void __launch_bounds__(256) slice_propose_kernel(const SliceState s, const float *__restrict__ u, float *__restrict__ q)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= s.N) return;
    const int d = (int)s.d[c];
    const float v = slice_query(s, c, u[c]);
    for (int j = 0; j < s.D; ++j) q[c * s.D + j] = (j == d) ? v : s.x[c * s.D + j];
}

__global__ void __launch_bounds__(256) slice_update_kernel(const SliceState s, const float *__restrict__ u, const float *__restrict__ f_in)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= s.N) return;
    const long long N = s.N;
    const float u0 = u[c], u1 = u[N + c], u2 = u[2 * N + c], u3 = u[3 * N + c];
    float f = f_in[c];
    if (f != f) f = -INFINITY;   // nan_to_num(nan = -inf)
    const int D = s.D;
    int d = (int)s.d[c];
    long long ph = s.phase[c];
    float lo = s.lo[c], hi = s.hi[c];
    const float w = s.width[c * D + d];
    const float q_val = slice_query(s, c, u0);
    const bool inside = f > s.log_y[c];
    long long sweeps = s.sweeps[c];
    const bool live = sweeps < s.total;
    // stepping out
    const bool g1 = ph == 1 && inside, g2 = ph == 2 && inside;
    float J = s.J[c] - (g1 ? 1.0f : 0.0f), K = s.K[c] - (g2 ? 1.0f : 0.0f);
    const bool leave1 = ph == 1 && !(g1 && J > 0.0f), leave2 = ph == 2 && !(g2 && K > 0.0f);
    if (g1) lo = __fsub_rn(lo, w);
    if (g2) hi = __fadd_rn(hi, w);
    // shrinkage
    const long long nshr = s.nshr[c] + (ph == 3 ? 1 : 0);
    const bool give_up = ph == 3 && !inside && nshr >= s.max_shrink;
    const bool acc = ph == 3 && inside && live;
    const bool rej = ph == 3 && !inside && !give_up;
    const bool left = q_val < s.x0[c];
    if (rej && left) lo = q_val;
    if (rej && !left) hi = q_val;
    float lp = s.lp[c];
    if (acc) {
        s.x[c * D + d] = q_val;
        lp = f;
    }
    if (leave1) ph = K > 0.0f ? 2 : 3;
    if (leave2) ph = 3;
    // a finished update
    const bool fin = acc || (give_up && live);
    if (fin && sweeps < s.warmup) {
        const float cnt = s.tuned[c * D + d];
        float nw = __fadd_rn(w, __fdiv_rn(__fsub_rn(__fsub_rn(hi, lo), w), __fadd_rn(cnt, 1.0f)));
        s.width[c * D + d] = nw < 1e-6f ? 1e-6f : nw;
        s.tuned[c * D + d] = cnt + 1.0f;
    }
    const bool sweep_end = fin && d == D - 1;
    if (fin) d = (d + 1) % D;
    if (sweep_end) sweeps += 1;
    const long long past = sweeps - s.warmup;
    long long taken = s.taken[c];
    if (sweep_end && past > 0 && past % s.thin == 0 && taken < s.S) {
        for (int j = 0; j < D; ++j) s.out[(taken * N + c) * D + j] = s.x[c * D + j];
        taken += 1;
    }
    float log_y = s.log_y[c], x0 = s.x0[c];
    long long ns = nshr;
    if (fin && sweeps < s.total) {   // start the update of the next coordinate
        const float wn = s.width[c * D + d];
        x0 = s.x[c * D + d];
        lo = __fsub_rn(x0, __fmul_rn(wn, u2));
        hi = __fadd_rn(lo, wn);
        const float m = (float)s.max_step_out;
        float Jn = floorf(__fmul_rn(m, u3));
        Jn = Jn < 0.0f ? 0.0f : (Jn > m - 1.0f ? m - 1.0f : Jn);
        J = Jn;
        K = (m - 1.0f) - Jn;
        ph = J > 0.0f ? 1 : (K > 0.0f ? 2 : 3);
        log_y = __fadd_rn(lp, logf(u1));
        ns = 0;
    }
    s.lp[c] = lp;
    s.lo[c] = lo;
    s.hi[c] = hi;
    s.x0[c] = x0;
    s.log_y[c] = log_y;
    s.J[c] = J;
    s.K[c] = K;
    s.d[c] = d;
    s.sweeps[c] = sweeps;
    s.taken[c] = taken;
    s.phase[c] = ph;
    s.nshr[c] = ns;
}

}  // namespace ddm

using namespace ddm;

static int load_state(const void *const *ptrs, const int64_t *ints, SliceState *s)
{
    DDM_REQUIRE(ptrs != nullptr && ints != nullptr, "slice sampler: null argument table");
    for (int i = 0; i < 16; ++i) DDM_REQUIRE(ptrs[i] != nullptr, "slice sampler: null state pointer %d", i);
    s->x = (float *)ptrs[0];
    s->lp = (float *)ptrs[1];
    s->width = (float *)ptrs[2];
    s->tuned = (float *)ptrs[3];
    s->lo = (float *)ptrs[4];
    s->hi = (float *)ptrs[5];
    s->x0 = (float *)ptrs[6];
    s->log_y = (float *)ptrs[7];
    s->J = (float *)ptrs[8];
    s->K = (float *)ptrs[9];
    s->d = (long long *)ptrs[10];
    s->sweeps = (long long *)ptrs[11];
    s->taken = (long long *)ptrs[12];
    s->phase = (long long *)ptrs[13];
    s->nshr = (long long *)ptrs[14];
    s->out = (float *)ptrs[15];
    s->N = ints[0];
    s->D = (int)ints[1];
    s->S = (int)ints[2];
    s->thin = (int)ints[3];
    s->warmup = (int)ints[4];
    s->total = (int)ints[5];
    s->max_step_out = (int)ints[6];
    s->max_shrink = (int)ints[7];
    DDM_REQUIRE(s->N >= 0 && s->D >= 1 && s->D <= 64 && s->S >= 1 && s->thin >= 1 && s->max_step_out >= 1,
                "slice sampler: bad sizes (N=%lld, D=%d, S=%d, thin=%d)", (long long)s->N, s->D, s->S, s->thin);
    return DDM_OK;
}

DDM_API int ddm_slice_propose_f32(const void *const *state_ptrs, const int64_t *state_ints, const float *u_dev, float *query_dev,
                                  void *stream)
{
    SliceState s;
    const int rc = load_state(state_ptrs, state_ints, &s);
    if (rc != DDM_OK) return rc;
    if (s.N == 0) return DDM_OK;
    DDM_REQUIRE(u_dev && query_dev, "ddm_slice_propose_f32: null pointer");
    slice_propose_kernel<<<(unsigned)((s.N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(s, u_dev, query_dev);
    DDM_CUDA_TRY(cudaGetLastError());
    return DDM_OK;
}

DDM_API int ddm_slice_update_f32(const void *const *state_ptrs, const int64_t *state_ints, const float *u_dev, const float *f_dev,
                                 void *stream)
{
    SliceState s;
    const int rc = load_state(state_ptrs, state_ints, &s);
    if (rc != DDM_OK) return rc;
    if (s.N == 0) return DDM_OK;
    DDM_REQUIRE(u_dev && f_dev, "ddm_slice_update_f32: null pointer");
    slice_update_kernel<<<(unsigned)((s.N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(s, u_dev, f_dev);
    DDM_CUDA_TRY(cudaGetLastError());
    return DDM_OK;
}
