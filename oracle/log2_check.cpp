// Checks the product's glibc-log2 restatement (diverseseq_b200/csrc/log2_glibc.h) against
// the platform libm `log2` — the function Rust's f64::log2 resolves to on Linux
// (/root/reference/src/record.rs:96).  TEST INFRASTRUCTURE ONLY (part of the oracle library).
#include "../diverseseq_b200/csrc/log2_glibc.h"

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {
inline uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline double sample(uint64_t i, uint64_t r) {
    double x;
    switch (i & 7) {
        case 0: memcpy(&x, &r, 8); return x;                                        // any bit pattern
        case 1: return (double)(r >> 11) * 0x1p-53;                                  // uniform [0,1)
        case 2: return 1.0 + ((double)(int64_t)(r >> 11) - 0x1p52) * 0x1p-57;        // near 1
        case 3: return (double)((r >> 40) + 1) / (double)(((r >> 8) & 0xffffffff) + 1);  // count ratios
        case 4: return std::ldexp((double)(r >> 11) * 0x1p-53, -(int)(r & 63));      // small
        case 5: { uint64_t u = r & 0x000fffffffffffffULL; memcpy(&x, &u, 8); return x; }  // subnormal
        case 6: return (double)((r & 0xfffff) + 1) / 4194304.0;                      // count/total
        default: return (double)(r >> 11) * 0x1p-53 * 16.0;
    }
}
}  // namespace

extern "C" {

double dvso_log2_port(double x) { return dvs_log2(x); }
double dvso_log2_libm(double x) { return std::log2(x); }

// number of inputs (out of n per thread) where port and libm differ bitwise (NaN == NaN)
uint64_t dvso_log2_port_mismatches(uint64_t seed, uint64_t n, int threads, double* first_bad) {
    std::atomic<uint64_t> bad{0};
    std::atomic<bool> have{false};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            uint64_t s = seed + 7919ULL * (uint64_t)t, b = 0;
            for (uint64_t i = 0; i < n; ++i) {
                double x = sample(i, splitmix(s));
                double a = std::log2(x), c = dvs_log2(x);
                uint64_t ua, uc;
                memcpy(&ua, &a, 8);
                memcpy(&uc, &c, 8);
                if (ua != uc && !(a != a && c != c)) {
                    ++b;
                    if (first_bad && !have.exchange(true)) *first_bad = x;
                }
            }
            bad += b;
        });
    for (auto& th : pool) th.join();
    return bad.load();
}

// the same sample stream, for device-vs-host comparisons
void dvso_log2_samples(uint64_t seed, uint64_t n, double* xs, double* ys) {
    uint64_t s = seed;
    for (uint64_t i = 0; i < n; ++i) {
        xs[i] = sample(i, splitmix(s));
        ys[i] = std::log2(xs[i]);
    }
}
}
