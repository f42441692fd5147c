// MNLE log-likelihood on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators).
//
// The 128x128 and 128x71 layers of the categorical net and of the ten spline conditioners are
// dense GEMMs over (trial, chain) rows -- the only tensor-core-shaped work on this path.  fp32
// accuracy (sums within 1e-4 relative, north_star) is kept by splitting every operand into
// bf16 hi + lo and issuing three MMAs per product (hi*hi + hi*lo + lo*hi, fp32 accumulate in
// TMEM): 16 mantissa bits per operand at 3/2 of the cost of a tf32 product.
//
// This file: the operand layout / descriptor self-test (mnle_tc_selftest) and the fused kernel.
#include "mnle_common.cuh"
#include "tc_ptx.cuh"

namespace mnle {

using namespace tc;

// Shared-memory image of a K-major bf16 operand tile [R rows][K]: for every group of 8 k's
// (16 bytes) all R rows are contiguous.  offset(r, k) = (k/8) * R*16 + r*16 + (k%8)*2 bytes.
//   => SBO (next 8-row group) = 128 B, LBO (next 8-k group) = R*16 B.
__host__ __device__ constexpr uint32_t tile_offset(int R, int r, int k) { return (uint32_t)((k >> 3) * R * 16 + r * 16 + (k & 7) * 2); }

// ------------------------------------------------------------------------------ self-test ---
// D (128 x N) = A (128 x 128) * B (N x 128)^T through exactly the building blocks of the fused
// kernel.  passes = 1: bf16(A) * bf16(B) only; passes = 3: hi/lo split.  lbo / sbo can be
// overridden (0 = the layout's own values) to probe descriptor conventions on hardware.
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B, int N,
                                                          int passes, uint32_t lbo_a, uint32_t lbo_b, uint32_t sbo,
                                                          float *__restrict__ D)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *a_hi = smem;                 // 128 x 128 bf16 = 32 KB
    unsigned char *a_lo = smem + 32768;
    unsigned char *b_hi = smem + 65536;         // up to 128 x 128 bf16
    unsigned char *b_lo = smem + 98304;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 131072);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 131072 + 16);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    // operands -> smem images (thread r owns row r of A and, if r < N, row r of B)
    for (int kg = 0; kg < 16; ++kg) {
        uint16_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_bf16(A[tid * 128 + kg * 8 + j], hi[j], lo[j]);
        uint4 vh = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), hi[4] | (hi[5] << 16), hi[6] | (hi[7] << 16));
        uint4 vl = make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), lo[4] | (lo[5] << 16), lo[6] | (lo[7] << 16));
        *reinterpret_cast<uint4 *>(a_hi + tile_offset(128, tid, kg * 8)) = vh;
        *reinterpret_cast<uint4 *>(a_lo + tile_offset(128, tid, kg * 8)) = vl;
        if (tid < N) {
#pragma unroll
            for (int j = 0; j < 8; ++j) split_bf16(B[tid * 128 + kg * 8 + j], hi[j], lo[j]);
            vh = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), hi[4] | (hi[5] << 16), hi[6] | (hi[7] << 16));
            vl = make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), lo[4] | (lo[5] << 16), lo[6] | (lo[7] << 16));
            *reinterpret_cast<uint4 *>(b_hi + tile_offset(N, tid, kg * 8)) = vh;
            *reinterpret_cast<uint4 *>(b_lo + tile_offset(N, tid, kg * 8)) = vl;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16_f32(128, N);
        const uint32_t la = lbo_a ? lbo_a : 128u * 16u, lb = lbo_b ? lbo_b : (uint32_t)N * 16u, sb = sbo ? sbo : 128u;
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; ++pass) {
            const unsigned char *pa = (pass == 2) ? a_lo : a_hi;
            const unsigned char *pb = (pass == 1) ? b_lo : b_hi;
            for (int ks = 0; ks < 8; ++ks) {
                const uint64_t da = umma_desc_kmajor(smem_u32(pa) + ks * 2 * 128 * 16, la, sb);
                const uint64_t db = umma_desc_kmajor(smem_u32(pb) + ks * 2 * N * 16, lb, sb);
                umma_bf16(tmem, da, db, idesc, acc);
                acc = 1;
            }
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    // thread (warp w, lane l) reads TMEM lane 32 w + l = output row
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace mnle

using namespace mnle;

DDM_API int mnle_tc_selftest(const float *a_dev, const float *b_dev, int N, int passes, uint32_t lbo_a, uint32_t lbo_b,
                             uint32_t sbo, float *d_dev, void *stream)
{
    DDM_REQUIRE(a_dev && b_dev && d_dev, "mnle_tc_selftest: null pointer");
    DDM_REQUIRE(N >= 16 && N <= 128 && N % 16 == 0, "mnle_tc_selftest: N=%d must be a multiple of 16 in [16,128]", N);
    DDM_REQUIRE(passes == 1 || passes == 3, "mnle_tc_selftest: passes must be 1 or 3");
    const int smem = 131072 + 64;
    DDM_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a_dev, b_dev, N, passes, lbo_a, lbo_b, sbo, d_dev);
    DDM_CUDA_TRY(cudaGetLastError());
    return DDM_OK;
}
