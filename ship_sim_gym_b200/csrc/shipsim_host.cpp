// shipsim_host.cpp -- host half of shipsim_step_host: observation rows [previous frame | frame] (ship_env.py:112-113)
// are put together from the frames that crossed PCIe.  Plain C++ (no CUDA): streaming stores, widest vector unit the
// CPU has (runtime dispatch), because the rows are written once and never read here -- no write-allocate traffic.
// Measured on the B200 box's host (16 threads, 4,096 x 1,000 rows): memcpy 9.3 ms, SSE2 5.5 ms, AVX-512 4.8 ms
// (profiles/host_assembly_bench.cpp).
#include "shipsim_host.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace shipsim {

static constexpr int kF = 16;       // floats per frame

static void rows_plain(float *obs, const float *fr, const uint8_t *cut, size_t a, size_t b, size_t N)
{
    for (size_t row = a; row < b; ++row) {
        float *dst = obs + row * (2 * kF);
        if (cut && cut[row]) for (int i = 0; i < kF; ++i) dst[i] = -1.f;
        else std::memcpy(dst, fr + row * kF, kF * sizeof(float));
        std::memcpy(dst + kF, fr + (row + N) * kF, kF * sizeof(float));
    }
}

#if defined(__x86_64__)
static void rows_sse2(float *obs, const float *fr, const uint8_t *cut, size_t a, size_t b, size_t N)
{
    const __m128 neg = _mm_set1_ps(-1.f);
    for (size_t row = a; row < b; ++row) {
        float *dst = obs + row * (2 * kF);
        const float *pf = fr + row * kF, *cf = fr + (row + N) * kF;
        const bool reset_row = cut && cut[row];
        for (int i = 0; i < kF; i += 4) _mm_stream_ps(dst + i, reset_row ? neg : _mm_load_ps(pf + i));
        for (int i = 0; i < kF; i += 4) _mm_stream_ps(dst + kF + i, _mm_load_ps(cf + i));
    }
    _mm_sfence();
}

__attribute__((target("avx512f"))) static void rows_avx512(float *obs, const float *fr, const uint8_t *cut, size_t a, size_t b, size_t N)
{
    const __m512 neg = _mm512_set1_ps(-1.f);
    for (size_t row = a; row < b; ++row) {
        float *dst = obs + row * (2 * kF);
        const bool reset_row = cut && cut[row];
        _mm512_stream_ps(dst, reset_row ? neg : _mm512_load_ps(fr + row * kF));
        _mm512_stream_ps(dst + kF, _mm512_load_ps(fr + (row + N) * kF));
    }
    _mm_sfence();
}
#endif

void assemble_history_rows(float *obs, const float *frames, const uint8_t *cut, size_t row_begin, size_t row_end, size_t N)
{
#if defined(__x86_64__)
    static const int level = __builtin_cpu_supports("avx512f") ? 2 : 1;
    if ((((uintptr_t)obs | (uintptr_t)frames) & 63) == 0 && level == 2) return rows_avx512(obs, frames, cut, row_begin, row_end, N);
    if ((((uintptr_t)obs | (uintptr_t)frames) & 15) == 0) return rows_sse2(obs, frames, cut, row_begin, row_end, N);
#endif
    rows_plain(obs, frames, cut, row_begin, row_end, N);
}

// ---------------------------------------------------------------------------------------------------------------------
// Compacted wire format -> rows.  Scalar bookkeeping per row (a handful of integer operations and, on 20 % of the rows,
// a few 4-byte patches of the env's current frame, which stays in the core's cache) around two 64-byte streaming stores.
// ---------------------------------------------------------------------------------------------------------------------
static inline float bits_f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

template <int LEVEL>        // 0 plain, 1 SSE2 (callers guarantee the alignment the level needs)
static void expand_rows_impl(float *obs, float *rew, uint8_t *done, const uint32_t *rec, const uint32_t *off, const float *var, float *cur,
                             int kc, size_t N, size_t blk_begin, size_t blk_end, float step_penalty, bool cut_on_done, int history)
{
    const size_t nblk = (N + 31) / 32;
    const float rtab[4] = {step_penalty, 1.f, -1.f, 0.f};
    const int row_f = kF * history;
    for (int k = 0; k < kc; ++k) {
        for (size_t blk = blk_begin; blk < blk_end; ++blk) {
            const float *v = var + off[(size_t)k * nblk + blk];
            const size_t e1 = std::min(N, blk * 32 + 32);
            for (size_t e = blk * 32; e < e1; ++e) {
                const size_t row = (size_t)k * N + e;
                const uint32_t *r = rec + row * 4;
                const uint32_t word = r[3];
                const bool dn = (word >> 5) & 1u;
                float *c = cur + e * kF;
                float *dst = obs ? obs + row * row_f : nullptr;
                if (dst && history == 2) {                      // the previous frame goes out before it is overwritten
                    const bool reset_row = cut_on_done && dn;
#if defined(__x86_64__)
                    if (LEVEL == 1) for (int i = 0; i < kF; i += 4) _mm_stream_ps(dst + i, reset_row ? _mm_set1_ps(-1.f) : _mm_load_ps(c + i));
                    else
#endif
                    { if (reset_row) for (int i = 0; i < kF; ++i) dst[i] = -1.f; else std::memcpy(dst, c, kF * sizeof(float)); }
                    dst += kF;
                }
                c[0] = bits_f(r[0]); c[1] = bits_f(r[1]); c[3] = bits_f(r[2]);
                c[2] = (float)(((int)(word & 7u) - 2) * 5);
                for (uint32_t m = (word >> 8) & 0xfffu; m; m &= m - 1u) c[4 + __builtin_ctz(m)] = *v++;
                if (dst) {
#if defined(__x86_64__)
                    if (LEVEL == 1) for (int i = 0; i < kF; i += 4) _mm_stream_ps(dst + i, _mm_load_ps(c + i));
                    else
#endif
                    std::memcpy(dst, c, kF * sizeof(float));
                }
                if (rew) rew[row] = rtab[(word >> 3) & 3u];
                if (done) done[row] = dn ? 1 : 0;
            }
        }
    }
#if defined(__x86_64__)
    if (LEVEL >= 1) _mm_sfence();
#endif
}

#if defined(__x86_64__)
// AVX-512: the frame is one register.  Slots 4..15 take the changed values straight from the stream with an expanding
// load under the change mask (vexpandps), slots 0..3 come from the record; two streaming stores per row.
__attribute__((target("avx512f"))) static void expand_rows_avx512(float *obs, float *rew, uint8_t *done, const uint32_t *rec,
                                                                  const uint32_t *off, const float *var, float *cur, int kc, size_t N,
                                                                  size_t blk_begin, size_t blk_end, float step_penalty, bool cut_on_done,
                                                                  int history)
{
    const size_t nblk = (N + 31) / 32;
    const float rtab[4] = {step_penalty, 1.f, -1.f, 0.f};
    alignas(32) static const float rudtab[8] = {-10.f, -5.f, 0.f, 5.f, 10.f, 15.f, 20.f, 25.f};
    const __m512 neg = _mm512_set1_ps(-1.f);
    const int row_f = kF * history;
    for (int k = 0; k < kc; ++k) {
        for (size_t blk = blk_begin; blk < blk_end; ++blk) {
            const float *v = var + off[(size_t)k * nblk + blk];
            const size_t e1 = std::min(N, blk * 32 + 32);
            for (size_t e = blk * 32; e < e1; ++e) {
                const size_t row = (size_t)k * N + e;
                const __m128i q = _mm_loadu_si128((const __m128i *)(rec + row * 4));
                const uint32_t word = (uint32_t)_mm_extract_epi32(q, 3);
                const bool dn = (word >> 5) & 1u;
                float *c = cur + e * kF;
                __m512 f = _mm512_load_ps(c);
                float *dst = obs ? obs + row * row_f : nullptr;
                if (dst && history == 2) {
                    _mm512_stream_ps(dst, cut_on_done && dn ? neg : f);
                    dst += kF;
                }
                __m128 head = _mm_castsi128_ps(_mm_shuffle_epi32(q, _MM_SHUFFLE(2, 2, 1, 0)));      // x, y, angle, angle
                head = _mm_insert_ps(head, _mm_load_ss(rudtab + (word & 7u)), 0x20);                // x, y, rudder, angle
                const __mmask16 m = (__mmask16)(((word >> 8) & 0xfffu) << 4);
                f = _mm512_mask_expandloadu_ps(f, m, v);
                f = _mm512_insertf32x4(f, head, 0);
                v += __builtin_popcount(m);
                _mm512_store_ps(c, f);
                if (dst) _mm512_stream_ps(dst, f);
                if (rew) rew[row] = rtab[(word >> 3) & 3u];
                if (done) done[row] = dn ? 1 : 0;
            }
        }
    }
    _mm_sfence();
}
#endif

void expand_delta_rows(float *obs, float *rew, uint8_t *done, const uint32_t *rec, const uint32_t *off, const float *var, float *cur,
                       int kc, size_t N, size_t blk_begin, size_t blk_end, float step_penalty, bool cut_on_done, int history)
{
#if defined(__x86_64__)
    static const int level = __builtin_cpu_supports("avx512f") ? 2 : 1;
    const uintptr_t al = (uintptr_t)obs | (uintptr_t)cur;
    if ((al & 63) == 0 && level == 2)
        return expand_rows_avx512(obs, rew, done, rec, off, var, cur, kc, N, blk_begin, blk_end, step_penalty, cut_on_done, history);
    if ((al & 15) == 0)
        return expand_rows_impl<1>(obs, rew, done, rec, off, var, cur, kc, N, blk_begin, blk_end, step_penalty, cut_on_done, history);
#endif
    expand_rows_impl<0>(obs, rew, done, rec, off, var, cur, kc, N, blk_begin, blk_end, step_penalty, cut_on_done, history);
}

}  // namespace shipsim
