// shipsim_host.cpp -- host half of shipsim_step_host: observation rows [previous frame | frame] (ship_env.py:112-113)
// are put together from the frames that crossed PCIe.  Plain C++ (no CUDA): streaming stores, widest vector unit the
// CPU has (runtime dispatch), because the rows are written once and never read here -- no write-allocate traffic.
// Measured on the B200 box's host (16 threads, 4,096 x 1,000 rows): memcpy 9.3 ms, SSE2 5.5 ms, AVX-512 4.8 ms
// (profiles/host_assembly_bench.cpp).
#include "shipsim_host.h"

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

}  // namespace shipsim
