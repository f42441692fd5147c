// fp32 CUDA-core dense layer over a 64-row tile held in shared memory: shared by the forward
// kernel (mnle_simt.cu) and the forward-mode gradient kernel (mnle_grad.cu).
#pragma once

#include "mnle_common.cuh"

namespace mnle {

constexpr int kTM = 64;         // rows per CTA tile
constexpr int kThreads = 256;
constexpr int kLdIn = 92;       // padded leading dims, multiples of 4 floats: rows are read with 128-bit loads
constexpr int kLdH = 132;       // (context tiles must be zero in columns [kCtx, kLdIn))
constexpr int kKC = 16;         // weight k-chunk staged in shared memory (two buffers: the next chunk is in flight)
constexpr int kLdW = 20;        // w_s[n][kk]: 20 n mod 32 walks all eight 4-bank groups -> conflict-free 128-bit loads

struct SimtSmem {
    float in[kTM * kLdIn];
    float ha[kTM * kLdH];
    float hb[kTM * kLdH];
    float w[2 * kHidden * kLdW];
    float red[8];
};

// kMaskRelu / kMaskSigmoid (backward passes): the product is multiplied by the derivative of the
// activation whose OUTPUT currently sits in out_s, and replaces it there.
// 4-byte asynchronous global -> shared copy (LDGSTS) and its completion
__device__ __forceinline__ void cp_async_f32(float *dst_smem, const float *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src_gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

enum Act { kNone = 0, kRelu = 1, kSigmoid = 2, kMaskRelu = 3, kMaskSigmoid = 4 };

// out[r][n] = act(sum_k in[r][k] * W[n][k] + b[n]) for r < 64, n < n_valid (<= 16 * NJ).
// in_s rows must be 16-byte aligned (ld_in a multiple of 4) and finite up to the next multiple of 4 past K.
// DUAL: rows come in groups of kDualRows = 6 (one primal row followed by its five tangent rows
// d/d theta_i); the layer is linear in the tangents, so they get no bias (and ACT must be kNone:
// the caller applies the activation and its derivative afterwards).
constexpr int kDualRows = 6;
// TRANS: W is read transposed, out[r][n] = sum_k in[r][k] * W[k][n] with row stride ldw (the
// backward-data product of a layer whose forward weights are W[k][n]); bias may then be null.
// out_s must not alias in_s.
template <int NJ, int ACT, bool DUAL = false, bool TRANS = false>
__device__ __forceinline__ void dense(const float *__restrict__ W, const float *__restrict__ bias, int K,
                                      int n_valid, const float *in_s, int ld_in, float *out_s, int ld_out,
                                      float *w_s, int ldw = 0)
{
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    // Weights go global -> shared with cp.async, one 16-k chunk ahead of the FMAs: the L2 latency of a
    // chunk (the top stall of the first version, long_scoreboard 2.5 per issue) is hidden behind the
    // previous chunk's arithmetic.  Chunk layout w_s[buf][n][kk] (zero beyond n_valid / K): thread
    // (tx, ty) reads four consecutive kk of column n = tx + 16 j with one 128-bit load.
    auto stage = [&](int buf, int k0) {
        float *dst = w_s + buf * (kHidden * kLdW);
        if (!TRANS) {
            const int kk = tid & 15;
            for (int n = tid >> 4; n < 16 * NJ; n += kThreads / 16) {
                if (n < n_valid && k0 + kk < K) cp_async_f32(dst + n * kLdW + kk, W + (size_t)n * K + k0 + kk);
                else dst[n * kLdW + kk] = 0.f;
            }
        } else {
            const int n = tid & 127;
            if (n < 16 * NJ) {
                for (int kk = tid >> 7; kk < kKC; kk += kThreads / 128) {
                    if (n < n_valid && k0 + kk < K) cp_async_f32(dst + n * kLdW + kk, W + (size_t)(k0 + kk) * ldw + n);
                    else dst[n * kLdW + kk] = 0.f;
                }
            }
        }
        cp_async_commit();
    };
    __syncthreads();  // hazards of the caller on in_s / out_s, and the previous layer's chunks are consumed
    stage(0, 0);
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += kKC, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();  // this chunk has landed for everyone; everyone is done with the other buffer
        if (k0 + kKC < K) stage(buf ^ 1, k0 + kKC);
        const float *w_c = w_s + buf * (kHidden * kLdW);
        const int kend = min(kKC, (K - k0 + 3) & ~3);  // in_s is finite (zero) up to the next multiple of 4
#pragma unroll 2
        for (int kk = 0; kk < kend; kk += 4) {
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(in_s + (ty * 4 + i) * ld_in + k0 + kk);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 wv = *reinterpret_cast<const float4 *>(w_c + (tx + 16 * j) * kLdW + kk);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][j] = fmaf(a[i].x, wv.x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, wv.y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, wv.z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, wv.w, acc[i][j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int n = tx + 16 * j;
        if (n < n_valid) {
            const float bv = (TRANS && bias == nullptr) ? 0.f : __ldg(bias + n);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v = acc[i][j] + ((!DUAL || (ty * 4 + i) % kDualRows == 0) ? bv : 0.f);
                if (ACT == kRelu) v = fmaxf(v, 0.f);
                if (ACT == kSigmoid) v = 1.0f / (1.0f + expf(-v));
                if (ACT == kMaskRelu) v = out_s[(ty * 4 + i) * ld_out + n] > 0.f ? v : 0.f;
                if (ACT == kMaskSigmoid) {
                    const float h = out_s[(ty * 4 + i) * ld_out + n];
                    v *= h * (1.0f - h);
                }
                out_s[(ty * 4 + i) * ld_out + n] = v;
            }
        }
    }
    __syncthreads();
}

}  // namespace mnle
