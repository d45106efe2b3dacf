// Host-only ceiling of the expansion half of shipsim_step_host (no GPU involved): synthetic records for 4,096 envs x
// 1,000 steps at the measured density of changed values, expanded by T threads under different job shapes and page
// sizes.  Build + run:  g++ -O2 -pthread -o /tmp/heb profiles/host_expand_bench.cpp ship_sim_gym_b200/csrc/shipsim_host.cpp && /tmp/heb 16
#include "../ship_sim_gym_b200/csrc/shipsim_host.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <sys/mman.h>
#include <thread>
#include <vector>

int main(int argc, char **argv)
{
    const size_t N = 4096, K = 1000, nblk = N / 32;
    const int nt = argc > 1 ? atoi(argv[1]) : 16;
    std::vector<uint32_t> rec(N * K * 4), off(K * nblk);
    std::vector<float> var(N * K * 2);
    std::mt19937 g(1);
    size_t pos = 0;
    for (size_t k = 0; k < K; ++k)
        for (size_t b = 0; b < nblk; ++b) {
            off[k * nblk + b] = (uint32_t)pos;
            for (size_t e = b * 32; e < b * 32 + 32; ++e) {
                uint32_t mask = 0;
                for (int j = 0; j < 12; ++j) if (g() % 100 < 8) mask |= 1u << j;
                uint32_t *r = &rec[(k * N + e) * 4];
                r[0] = g(); r[1] = g(); r[2] = g(); r[3] = 2u | ((g() % 50 == 0) << 5) | (mask << 8);
                pos += __builtin_popcount(mask);
            }
        }
    printf("values per env-step %.2f\n", (double)pos / (N * K));
    for (int huge = 0; huge < 2; ++huge) {
        const size_t bytes = N * K * 128;
        float *obs = (float *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        madvise(obs, bytes, huge ? MADV_HUGEPAGE : MADV_NOHUGEPAGE);
        memset(obs, 0, bytes);
        float *cur = (float *)aligned_alloc(64, N * 64);
        std::vector<float> rew(N * K);
        std::vector<uint8_t> done(N * K);
        for (int chunks : {1, 16, 32})
            for (int jpw : {1, 2, 4}) {
                double best = 1e9;
                for (int rep = 0; rep < 3; ++rep) {
                    memset(cur, 0, N * 64);
                    auto t0 = std::chrono::steady_clock::now();
                    for (int c = 0; c < chunks; ++c) {
                        const size_t k0 = K * c / chunks, k1 = K * (c + 1) / chunks;
                        const int jobs = nt * jpw;
                        std::vector<std::thread> th;
                        for (int t = 0; t < nt; ++t)
                            th.emplace_back([&, t] {
                                for (int j = t; j < jobs; j += nt)
                                    shipsim::expand_delta_rows(obs + k0 * N * 32, rew.data() + k0 * N, done.data() + k0 * N, rec.data() + k0 * N * 4,
                                                               off.data() + k0 * nblk, var.data(), cur, (int)(k1 - k0), N, nblk * j / jobs,
                                                               nblk * (j + 1) / jobs, -0.01f, true, 2);
                            });
                        for (auto &x : th) x.join();
                    }
                    best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
                }
                printf("hugepages %d threads %d chunks %2d jobs/thread %d: %.2f ms\n", huge, nt, chunks, jpw, best);
            }
        munmap(obs, bytes);
    }
    return 0;
}
