// Host-side packing of z rows for the PCIe-bound ingest path.
//
// The reference hands the simulator z = [theta (5), pulse sides (P)] as fp32 rows (340 bytes per
// trial at P = 80; data_simulator.py:22-29).  Pulse sides are +-1, so when z lives in HOST memory
// the link carries 10x more bytes than information: 1e8 trials are 34 GB over a ~55 GB/s PCIe 5 x16
// link, longer than the kernel needs to simulate them.  ddm_pack_z_host turns each row into one
// 32-byte record [theta bits x 5, pulse sign masks x 3] on the host cores (multi-threaded, AVX-512 mask
// compares, else AVX2 / SSE2 compares + movemask) so that the copy engine moves 32 bytes per trial;
// ddm_sim_packed_f32 (ddm_sim.cu) consumes the records directly.  Rows holding anything other than
// +-1 are counted: the caller sends such batches through the fp32 path instead.
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime.h>

#include "../../include/ddm_b200.h"

#define DDM_API extern "C" __attribute__((visibility("default")))
namespace ddm {
void set_error(const char *fmt, ...);  // ddm_common.cu
}

#include <unistd.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace ddm {

// bit j of the result = (s[j] > 0) for j < n (n <= 32), bits n..31 = 1 (the kernel's default for
// columns past the schedule); *odd |= any |s[j]| != 1
static inline uint32_t sign_mask_scalar(const float *s, int n, bool *odd)
{
    uint32_t m = 0xFFFFFFFFu;
    for (int j = 0; j < n; ++j) {
        const float v = s[j];
        if (!(v > 0.0f)) m &= ~(1u << j);
        if (!(v == 1.0f || v == -1.0f)) *odd = true;
    }
    return m;
}

static void pack_rows_sse2(const float *z, int64_t ld, int64_t r0, int64_t r1, int n_pulses, uint32_t *out,
                           int64_t *generic_rows)
{
    const __m128 zero = _mm_setzero_ps(), one = _mm_set1_ps(1.0f);
    const __m128 absmask = _mm_castsi128_ps(_mm_set1_epi32(0x7FFFFFFF));
    int64_t gen = 0;
    for (int64_t r = r0; r < r1; ++r) {
        const float *row = z + r * ld;
        uint32_t *rec = out + r * 8;
        for (int i = 0; i < 5; ++i) memcpy(&rec[i], &row[i], 4);
        const float *s = row + 5;
        bool odd = false;
        for (int w = 0; w < 3; ++w) {
            const int base = 32 * w;
            const int n = n_pulses - base < 0 ? 0 : (n_pulses - base > 32 ? 32 : n_pulses - base);
            uint32_t m = 0xFFFFFFFFu;
            int j = 0;
            if (n > 0) m = (n == 32) ? 0u : (0xFFFFFFFFu << n);
            int oddbits = 0;
            for (; j + 4 <= n; j += 4) {
                const __m128 v = _mm_loadu_ps(s + base + j);
                m |= (uint32_t)_mm_movemask_ps(_mm_cmpgt_ps(v, zero)) << j;
                oddbits |= _mm_movemask_ps(_mm_cmpneq_ps(_mm_and_ps(v, absmask), one));
            }
            if (j < n) {
                bool o = false;
                const uint32_t tail = sign_mask_scalar(s + base + j, n - j, &o);
                m |= (tail & ((1u << (n - j)) - 1u)) << j;
                odd = odd || o;
            }
            odd = odd || oddbits != 0;
            rec[5 + w] = m;
        }
        gen += odd ? 1 : 0;
    }
    *generic_rows = gen;
}

__attribute__((target("avx2"))) static void pack_rows_avx2(const float *z, int64_t ld, int64_t r0, int64_t r1,
                                                           int n_pulses, uint32_t *out, int64_t *generic_rows)
{
    const __m256 zero = _mm256_setzero_ps(), one = _mm256_set1_ps(1.0f);
    const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7FFFFFFF));
    int64_t gen = 0;
    for (int64_t r = r0; r < r1; ++r) {
        const float *row = z + r * ld;
        uint32_t *rec = out + r * 8;
        for (int i = 0; i < 5; ++i) memcpy(&rec[i], &row[i], 4);
        const float *s = row + 5;
        bool odd = false;
        for (int w = 0; w < 3; ++w) {
            const int base = 32 * w;
            const int n = n_pulses - base < 0 ? 0 : (n_pulses - base > 32 ? 32 : n_pulses - base);
            uint32_t m = 0xFFFFFFFFu;
            if (n > 0) m = (n == 32) ? 0u : (0xFFFFFFFFu << n);
            int j = 0, oddbits = 0;
            for (; j + 8 <= n; j += 8) {
                const __m256 v = _mm256_loadu_ps(s + base + j);
                m |= (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(v, zero, _CMP_GT_OQ)) << j;
                oddbits |= _mm256_movemask_ps(_mm256_cmp_ps(_mm256_and_ps(v, absmask), one, _CMP_NEQ_UQ));
            }
            if (j < n) {
                bool o = false;
                const uint32_t tail = sign_mask_scalar(s + base + j, n - j, &o);
                m |= (tail & ((1u << (n - j)) - 1u)) << j;
                odd = odd || o;
            }
            odd = odd || oddbits != 0;
            rec[5 + w] = m;
        }
        gen += odd ? 1 : 0;
    }
    *generic_rows = gen;
}

// AVX-512 (F + DQ + VL): a compare writes its 16 result bits straight into a mask register, so a row of
// 80 pulses is five 64-byte loads and ten compares; the record leaves as one 32-byte non-temporal store
// (the staging block is write-only for the host: no read-for-ownership traffic).  Rows further along are
// prefetched explicitly: the 340-byte row stride defeats the adjacent-line prefetcher on some parts.
__attribute__((target("avx512f,avx512dq,avx512vl,avx2"))) static void pack_rows_avx512(
    const float *z, int64_t ld, int64_t r0, int64_t r1, int n_pulses, uint32_t *out, int64_t *generic_rows)
{
    const __m512i absmask = _mm512_set1_epi32(0x7FFFFFFF), one = _mm512_set1_epi32(0x3F800000);
    const __m512 zero = _mm512_setzero_ps();
    __mmask16 lm[6];  // lanes of 16-column group g that the schedule reaches
    for (int g = 0; g < 6; ++g) {
        const int n = n_pulses - 16 * g;
        lm[g] = (__mmask16)(n >= 16 ? 0xFFFF : (n <= 0 ? 0 : ((1u << n) - 1u)));
    }
    const int groups = (n_pulses + 15) / 16;
    const bool aligned_out = (reinterpret_cast<uintptr_t>(out) & 31u) == 0;
    constexpr int64_t kAhead = 12;  // rows (~4 KB)
    int64_t gen = 0;
    for (int64_t r = r0; r < r1; ++r) {
        const float *row = z + r * ld;
        if (r + kAhead < r1) {
            const char *pf = reinterpret_cast<const char *>(row + kAhead * ld);
            _mm_prefetch(pf, _MM_HINT_T0);
            _mm_prefetch(pf + 64, _MM_HINT_T0);
            _mm_prefetch(pf + 128, _MM_HINT_T0);
            _mm_prefetch(pf + 192, _MM_HINT_T0);
            _mm_prefetch(pf + 256, _MM_HINT_T0);
            _mm_prefetch(pf + 320, _MM_HINT_T0);
        }
        const float *s = row + 5;
        uint32_t gt[6];
        __mmask16 odd = 0;
#pragma GCC unroll 6
        for (int g = 0; g < 6; ++g) {
            if (g < groups) {
                const __m512 v = _mm512_maskz_loadu_ps(lm[g], s + 16 * g);  // masked-out lanes are not touched
                gt[g] = (uint32_t)(_mm512_mask_cmp_ps_mask(lm[g], v, zero, _CMP_GT_OQ) | (__mmask16)~lm[g]);
                odd |= _mm512_mask_cmpneq_epi32_mask(lm[g], _mm512_and_si512(_mm512_castps_si512(v), absmask), one);
            } else {
                gt[g] = 0xFFFFu;
            }
        }
        // record: theta bits x 5 (the first 8 floats of the row, lanes 5..7 replaced), masks x 3
        __m256i rec = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(row));
        const __m256i masks = _mm256_setr_epi32(0, 0, 0, 0, 0, (int)(gt[0] | (gt[1] << 16)), (int)(gt[2] | (gt[3] << 16)),
                                                (int)(gt[4] | (gt[5] << 16)));
        rec = _mm256_mask_blend_epi32((__mmask8)0xE0, rec, masks);
        if (aligned_out) _mm256_stream_si256(reinterpret_cast<__m256i *>(out + r * 8), rec);
        else _mm256_storeu_si256(reinterpret_cast<__m256i *>(out + r * 8), rec);
        gen += odd ? 1 : 0;
    }
    _mm_sfence();
    *generic_rows = gen;
}

// Persistent worker pool: a chunk of 2^18 rows packs in about a millisecond, so spawning threads per
// call (~30-50 us each) would cost as much as the work.  Workers sleep on a condition variable between
// calls; the calling thread takes jobs too.  A forked child starts a fresh pool.
class Pool {
public:
    void run(int n_jobs, int n_threads, const std::function<void(int)> &f)
    {
        std::lock_guard<std::mutex> call(call_mutex_);
        std::unique_lock<std::mutex> lk(m_);
        if (pid_ != getpid()) {  // after fork(): the parent's workers do not exist here
            for (auto &t : threads_) t.detach();
            threads_.clear();
            pid_ = getpid();
        }
        while ((int)threads_.size() < n_threads - 1) threads_.emplace_back([this] { worker(); });
        job_ = &f;
        n_jobs_ = n_jobs;
        next_ = 0;
        pending_ = n_jobs;
        cv_work_.notify_all();
        while (next_ < n_jobs_) {
            const int j = next_++;
            lk.unlock();
            f(j);
            lk.lock();
            --pending_;
        }
        cv_done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void worker()
    {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_work_.wait(lk, [this] { return job_ != nullptr && next_ < n_jobs_; });
            const int j = next_++;
            const std::function<void(int)> *f = job_;
            lk.unlock();
            (*f)(j);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
    std::mutex call_mutex_, m_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> threads_;
    const std::function<void(int)> *job_ = nullptr;
    int n_jobs_ = 0, next_ = 0, pending_ = 0;
    pid_t pid_ = getpid();
};

static Pool &pool()
{
    static Pool *p = new Pool();  // never destroyed: workers may be asleep in it at exit
    return *p;
}

}  // namespace ddm

static bool have_avx512()
{
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
           getenv("DDM_PACK_NO_AVX512") == nullptr;
}

DDM_API int ddm_pack_simd_bits(void) { return have_avx512() ? 512 : (__builtin_cpu_supports("avx2") ? 256 : 128); }

DDM_API int64_t ddm_pack_z_host(const float *z_host, int64_t ld, int64_t N, int64_t n_pulses, uint32_t *packed_host,
                                int n_threads)
{
    if (N < 0 || n_pulses < 0 || n_pulses > 96 || ld < 5 + n_pulses || (N > 0 && (!z_host || !packed_host))) {
        ddm::set_error("ddm_pack_z_host: bad arguments (N=%lld, n_pulses=%lld in [0,96], ld=%lld >= 5 + n_pulses)",
                       (long long)N, (long long)n_pulses, (long long)ld);
        return DDM_ERR_INVALID;
    }
    if (N == 0) return 0;
    const bool avx2 = __builtin_cpu_supports("avx2");
    // a row must hold 8 floats for the record's 32-byte load (5 + n_pulses >= 8)
    const bool avx512 = have_avx512() && n_pulses >= 3;
    int nt = n_threads < 1 ? 1 : n_threads;
    const int64_t min_rows = 1 << 14;  // below that a thread costs more than it packs
    if ((int64_t)nt > (N + min_rows - 1) / min_rows) nt = (int)((N + min_rows - 1) / min_rows);
    std::vector<int64_t> gen((size_t)nt, 0);
    auto work = [&](int t) {
        const int64_t r0 = N * t / nt, r1 = N * (t + 1) / nt;
        if (avx512) ddm::pack_rows_avx512(z_host, ld, r0, r1, (int)n_pulses, packed_host, &gen[(size_t)t]);
        else if (avx2) ddm::pack_rows_avx2(z_host, ld, r0, r1, (int)n_pulses, packed_host, &gen[(size_t)t]);
        else ddm::pack_rows_sse2(z_host, ld, r0, r1, (int)n_pulses, packed_host, &gen[(size_t)t]);
    };
    if (nt == 1) work(0);
    else ddm::pool().run(nt, nt, work);
    int64_t total = 0;
    for (int64_t g : gen) total += g;
    return total;
}

// ---- records -> fp32 rows on the host (device->host direction) -----------------------------------------
// The mirror image of the packer: the GPU packs the rows it produced (ddm_pack_z_dev), 32 bytes per trial cross
// the link, and the host cores rebuild z = [theta (5), pulses (+-1)] -- for a caller that wants the training set
// in host memory (reference data_simulator.py:53-60) the link carries 40 bytes per trial instead of 348.
namespace ddm {

static void unpack_rows_scalar(const uint32_t *packed, int64_t r0, int64_t r1, int n_pulses, float *z, int64_t ld)
{
    for (int64_t r = r0; r < r1; ++r) {
        const uint32_t *rec = packed + r * 8;
        float *row = z + r * ld;
        memcpy(row, rec, 20);
        for (int j = 0; j < n_pulses; ++j) row[5 + j] = ((rec[5 + (j >> 5)] >> (j & 31)) & 1u) ? 1.0f : -1.0f;
    }
}

// Contiguous rows of 5 + n_pulses floats: sixteen rows are a whole number of 64-byte lines whatever the row
// length, so they are rebuilt in a line-aligned stack block and leave with non-temporal stores (no
// read-for-ownership of the destination: 340 instead of 680 bytes of DRAM traffic per trial).
__attribute__((target("avx512f,avx512dq,avx512vl,avx2"))) static void unpack_rows_avx512(const uint32_t *packed, int64_t r0,
                                                                                        int64_t r1, int n_pulses, float *z)
{
    const int W = 5 + n_pulses;
    const __m512 pos = _mm512_set1_ps(1.0f), neg = _mm512_set1_ps(-1.0f);
    alignas(64) float block[16 * (5 + 96) + 16];
    int64_t r = r0;
    for (; r + 16 <= r1; r += 16) {
        for (int i = 0; i < 16; ++i) {
            const uint32_t *rec = packed + (r + i) * 8;
            float *row = block + i * W;
            memcpy(row, rec, 20);
            for (int g = 0; 16 * g < n_pulses; ++g) {
                const __mmask16 k = (__mmask16)(rec[5 + (g >> 1)] >> (16 * (g & 1)));
                const int left = n_pulses - 16 * g;
                const __m512 v = _mm512_mask_blend_ps(k, neg, pos);
                if (left >= 16) _mm512_storeu_ps(row + 5 + 16 * g, v);
                else _mm512_mask_storeu_ps(row + 5 + 16 * g, (__mmask16)((1u << left) - 1u), v);
            }
        }
        float *dst = z + r * W;   // 64-byte aligned: r is a multiple of 16 rows from an aligned base
        for (int i = 0; i < W; ++i) _mm512_stream_ps(dst + 16 * i, _mm512_load_ps(block + 16 * i));
    }
    _mm_sfence();
    if (r < r1) unpack_rows_scalar(packed, r, r1, n_pulses, z, W);
}

}  // namespace ddm

DDM_API int ddm_unpack_z_host(const uint32_t *packed_host, int64_t N, int64_t n_pulses, float *z_host, int64_t ld, int n_threads)
{
    if (N < 0 || n_pulses < 0 || n_pulses > 96 || ld < 5 + n_pulses || (N > 0 && (!packed_host || !z_host))) {
        ddm::set_error("ddm_unpack_z_host: bad arguments (N=%lld, n_pulses=%lld in [0,96], ld=%lld >= 5 + n_pulses)", (long long)N,
                       (long long)n_pulses, (long long)ld);
        return DDM_ERR_INVALID;
    }
    if (N == 0) return DDM_OK;
    const bool fast = have_avx512() && ld == 5 + n_pulses && (reinterpret_cast<uintptr_t>(z_host) & 63u) == 0;
    int nt = n_threads < 1 ? 1 : n_threads;
    const int64_t min_rows = 1 << 14;
    if ((int64_t)nt > (N + min_rows - 1) / min_rows) nt = (int)((N + min_rows - 1) / min_rows);
    auto work = [&](int t) {
        int64_t r0 = N * t / nt, r1 = N * (t + 1) / nt;
        r0 &= ~int64_t(15);                       // thread ranges start on 16-row (= whole-line) boundaries
        if (t + 1 < nt) r1 &= ~int64_t(15);
        if (fast) ddm::unpack_rows_avx512(packed_host, r0, r1, (int)n_pulses, z_host);
        else ddm::unpack_rows_scalar(packed_host, r0, r1, (int)n_pulses, z_host, ld);
    };
    if (nt == 1) work(0);
    else ddm::pool().run(nt, nt, work);
    return DDM_OK;
}

// The whole ingest of one batch without returning to the caller between chunks: pack chunk k on the host
// cores, enqueue its copy and the 8-byte copy that raises *ready_dev, go on with chunk k + 1 while the
// copy engine works.  (Driving this loop from Python cost ~150 us per chunk, a fifth of the packing time.)
DDM_API int ddm_ingest_packed(const float *z_host, int64_t ld, int64_t N, int64_t n_pulses, int64_t chunk_rows,
                              uint32_t *staging_host, uint32_t *packed_dev, uint64_t *ready_dev,
                              const uint64_t *marks_host, int n_threads, void *copy_stream, int64_t *generic_rows)
{
    if (N < 0 || chunk_rows < 1 || (N > 0 && (!z_host || !staging_host || !packed_dev || !ready_dev || !marks_host))) {
        ddm::set_error("ddm_ingest_packed: bad arguments (N=%lld, chunk_rows=%lld)", (long long)N, (long long)chunk_rows);
        return DDM_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(copy_stream);
    int64_t generic = 0;
    int64_t k = 0;
    for (int64_t a = 0; a < N; a += chunk_rows, ++k) {
        const int64_t rows = (N - a < chunk_rows) ? N - a : chunk_rows;
        const int64_t got = ddm_pack_z_host(z_host + a * ld, ld, rows, n_pulses, staging_host + a * 8, n_threads);
        if (got < 0) return (int)got;
        generic += got;
        cudaError_t e = cudaMemcpyAsync(packed_dev + a * 8, staging_host + a * 8, (size_t)rows * 32, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ready_dev, marks_host + k, sizeof(uint64_t), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) {
            ddm::set_error("ddm_ingest_packed: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
            return DDM_ERR_CUDA;
        }
    }
    if (generic_rows) *generic_rows = generic;
    return DDM_OK;
}
