#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
__attribute__((target("avx512f"))) static void rows_avx512(float *dst, const float *fr, const unsigned char *done, size_t a, size_t b, size_t N)
{
    const __m512 neg = _mm512_set1_ps(-1.f);
    for (size_t r = a; r < b; ++r) {
        float *d = dst + r * 32;
        _mm512_stream_ps(d, done[r] ? neg : _mm512_load_ps(fr + r * 16));
        _mm512_stream_ps(d + 16, _mm512_load_ps(fr + (r + N) * 16));
    }
}
__attribute__((target("avx2"))) static void rows_avx2(float *dst, const float *fr, const unsigned char *done, size_t a, size_t b, size_t N)
{
    const __m256 neg = _mm256_set1_ps(-1.f);
    for (size_t r = a; r < b; ++r) {
        float *d = dst + r * 32;
        const float *p = fr + r * 16, *q = fr + (r + N) * 16;
        _mm256_stream_ps(d, done[r] ? neg : _mm256_load_ps(p));
        _mm256_stream_ps(d + 8, done[r] ? neg : _mm256_load_ps(p + 8));
        _mm256_stream_ps(d + 16, _mm256_load_ps(q));
        _mm256_stream_ps(d + 24, _mm256_load_ps(q + 8));
    }
}
int main(int argc, char **argv)
{
    const size_t N = 4096, K = 1000, rows = N * K;
    const int nt = argc > 1 ? atoi(argv[1]) : 16;
    float *fr = (float *)aligned_alloc(4096, (rows + N) * 64);
    float *dst = (float *)aligned_alloc(4096, rows * 128);
    unsigned char *done = (unsigned char *)calloc(rows, 1);
    memset(fr, 1, (rows + N) * 64);
    memset(dst, 0, rows * 128);
    printf("avx512f %d avx2 %d\n", __builtin_cpu_supports("avx512f"), __builtin_cpu_supports("avx2"));
    for (int mode = 0; mode < 5; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t)
                th.emplace_back([&, t] {
                    const size_t a = rows * t / nt, b = rows * (t + 1) / nt;
                    if (mode == 0) {
                        for (size_t r = a; r < b; ++r) {
                            float *d = dst + r * 32;
                            if (done[r]) for (int i = 0; i < 16; ++i) d[i] = -1.f;
                            else memcpy(d, fr + r * 16, 64);
                            memcpy(d + 16, fr + (r + N) * 16, 64);
                        }
                    } else if (mode == 1) {
                        for (size_t r = a; r < b; ++r) {
                            float *d = dst + r * 32;
                            const float *p = fr + r * 16, *q = fr + (r + N) * 16;
                            for (int i = 0; i < 4; ++i) _mm_stream_ps(d + 4 * i, done[r] ? _mm_set1_ps(-1.f) : _mm_load_ps(p + 4 * i));
                            for (int i = 0; i < 4; ++i) _mm_stream_ps(d + 16 + 4 * i, _mm_load_ps(q + 4 * i));
                        }
                    } else if (mode == 3) {
                        if (__builtin_cpu_supports("avx512f")) rows_avx512(dst, fr, done, a, b, N);
                    } else if (mode == 4) {
                        if (__builtin_cpu_supports("avx2")) rows_avx2(dst, fr, done, a, b, N);
                    } else {   // second halves only (what a strided DMA would leave to the host): copy 64 B per row
                        for (size_t r = a; r < b; ++r) {
                            float *d = dst + r * 32;
                            const float *p = r >= N ? dst + (r - N) * 32 + 16 : fr + r * 16;
                            for (int i = 0; i < 4; ++i) _mm_stream_ps(d + 4 * i, done[r] ? _mm_set1_ps(-1.f) : _mm_load_ps(p + 4 * i));
                        }
                    }
                });
            for (auto &x : th) x.join();
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            printf("threads %d mode %d: %.2f ms\n", nt, mode, ms);
        }
    return 0;
}
