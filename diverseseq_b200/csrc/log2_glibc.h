// dvs_log2(): operation-for-operation restatement of glibc 2.39 `log2` (x86-64,
// FMA/AVX2 ifunc variant — the one Rust's `f64::log2` reaches on any FMA-capable
// Linux host).  The reference calls it once per non-zero k-mer frequency in
// `entropy` (/root/reference/src/record.rs:86-106, line 96).
//
// glibc's log2 is <1 ULP but NOT correctly rounded, so bit-identical entropies
// (and hence bit-identical nmost/max selection decisions) require the identical
// sequence of IEEE-754 double operations, including which multiply-adds the
// compiler fused.  The sequence below was transcribed from the instruction
// stream of that variant (objdump of libm.so.6) and the
// constants come from libm's __log2_data via tools/extract_log2_table.py.
// Algorithm: ARM optimized-routines log2.c (table of 64 {1/c, log2 c}, order-6
// polynomial, separate order-10 polynomial near 1).
//
// Every operation is an explicit fma / mul / add so neither nvcc (-fmad) nor gcc
// (-ffp-contract) can change the rounding.  Usable from host C++ and CUDA.
#pragma once
#include <stdint.h>
#include <string.h>
#include "log2_glibc_table.h"

#if defined(__CUDACC__)
#define DVS_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define DVS_HD static inline
#endif

struct DvsLog2Data {
    double a[6];
    double b[10];
    double tab[128];
};

#if defined(__CUDACC__)
// __constant__ copy for device code, plain copy for host code compiled by nvcc.
__device__ __constant__ DvsLog2Data dvs_log2_data_dev = {
    {DVS_LOG2_POLY_A}, {DVS_LOG2_POLY_B}, {DVS_LOG2_TAB}};
#endif
static const DvsLog2Data dvs_log2_data_host = {
    {DVS_LOG2_POLY_A}, {DVS_LOG2_POLY_B}, {DVS_LOG2_TAB}};

namespace dvs_log2_detail {
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ uint64_t bits_(double x) { return (uint64_t)__double_as_longlong(x); }
__device__ __forceinline__ double dbl_(uint64_t u) { return __longlong_as_double((long long)u); }
#else
// host: compiled with -ffp-contract=off; __builtin_fma maps to a hardware fma
// (build uses -mfma) or to libm's correctly rounded fma().
static inline double fma_(double a, double b, double c) { return __builtin_fma(a, b, c); }
static inline double mul_(double a, double b) { volatile double r = a * b; return r; }
static inline double add_(double a, double b) { volatile double r = a + b; return r; }
static inline double sub_(double a, double b) { volatile double r = a - b; return r; }
static inline uint64_t bits_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double dbl_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#endif
}  // namespace dvs_log2_detail

DVS_HD double dvs_log2(double x) {
    using namespace dvs_log2_detail;
#if defined(__CUDA_ARCH__)
    const DvsLog2Data& D = dvs_log2_data_dev;
#else
    const DvsLog2Data& D = dvs_log2_data_host;
#endif
    const double InvLn2hi = DVS_LOG2_INVLN2HI;
    const double InvLn2lo = DVS_LOG2_INVLN2LO;
    uint64_t ix = bits_(x);
    uint32_t top = (uint32_t)(ix >> 48);

    // |x - 1| small: LO = bits(1 - 0x1.5b51p-5), HI - LO = 0x210aa00000000
    if (ix - 0x3feea4af00000000ULL < 0x000210aa00000000ULL) {
        if (ix == 0x3ff0000000000000ULL) return 0.0;
        const double* B = D.b;
        double r = sub_(x, 1.0);
        double hi = mul_(InvLn2hi, r);
        double r2 = mul_(r, r);
        double u = fma_(InvLn2hi, r, -hi);
        double r4 = mul_(r2, r2);
        double b01 = fma_(r, B[1], B[0]);
        double lo = fma_(r, InvLn2lo, u);
        double y = fma_(b01, r2, hi);
        double d = sub_(hi, y);
        double w = fma_(b01, r2, d);
        double b23 = fma_(r, B[3], B[2]);
        lo = add_(w, lo);
        double b45 = fma_(r, B[5], B[4]);
        double c1 = fma_(b45, r2, b23);
        double b67 = fma_(r, B[7], B[6]);
        double b89 = fma_(r, B[9], B[8]);
        double c2 = fma_(b89, r2, b67);
        double c = fma_(c2, r4, c1);
        lo = fma_(c, r4, lo);
        return add_(y, lo);
    }
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        // zero, negative, inf, nan, subnormal
        if ((ix << 1) == 0) return dbl_(0xfff0000000000000ULL);  // log2(+-0) = -inf
        if (ix == 0x7ff0000000000000ULL) return x;                 // log2(inf) = inf
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u)
            return dbl_(0xfff8000000000000ULL);                    // x<0 or nan -> nan
        ix = bits_(mul_(x, 0x1p52));                               // subnormal: scale up
        ix -= 52ULL << 52;
    }
    const double* A = D.a;
    uint64_t tmp = ix - 0x3fe6000000000000ULL;
    int i = (int)((tmp >> 46) & 63);
    int k = (int)((int64_t)tmp >> 52);
    uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
    double invc = D.tab[2 * i];
    double logc = D.tab[2 * i + 1];
    double z = dbl_(iz);
    double kd = (double)k;

    double t3 = add_(kd, logc);
    double r = fma_(z, invc, -1.0);
    double q01 = fma_(r, A[1], A[0]);
    double t1 = mul_(InvLn2hi, r);
    double u = fma_(InvLn2hi, r, -t1);
    double hi = add_(t1, t3);
    double lo = sub_(t3, hi);
    double t2 = fma_(r, InvLn2lo, u);
    double r2 = mul_(r, r);
    lo = add_(lo, t1);
    lo = add_(lo, t2);
    double q23 = fma_(r, A[3], A[2]);
    double r4 = mul_(r2, r2);
    double q45 = fma_(r, A[5], A[4]);
    double s = fma_(q23, r2, q01);
    double p = fma_(q45, r4, s);
    double y = fma_(r2, p, lo);
    return add_(y, hi);
}

#if defined(__CUDACC__)
// The table-driven path of dvs_log2 on its own, for the throughput kernels: branch-free so that several
// evaluations interleave in one thread (FP64 latency on sm_100 is ~14 cycles), table read from a
// shared-memory copy (`tab[i] = {1/c_i, log2 c_i}`; the __constant__ copy would serialise a warp's 32
// different indices).  Same operations in the same order as above, hence the same bits, for every input
// that dvs_log2 sends down this path; `special` is raised for the inputs it handles elsewhere
// (|x - 1| small, zero, negative, subnormal, inf, nan), for which the value returned here is meaningless.
__device__ __forceinline__ double dvs_log2_main(double x, const double2* __restrict__ tab, int& special) {
    using namespace dvs_log2_detail;
    constexpr double A0[6] = {DVS_LOG2_POLY_A};
    const double InvLn2hi = DVS_LOG2_INVLN2HI;
    const double InvLn2lo = DVS_LOG2_INVLN2LO;
    // all constants of the reference code have zero low words, so the integer part only needs the high
    // word of x (the exponent and the top 20 mantissa bits)
    const uint32_t hi32 = (uint32_t)__double2hiint(x);
    const uint32_t top = hi32 >> 16;
    special |= (hi32 - 0x3feea4afu < 0x000210aau) ? 1 : 0;
    special |= (top - 0x0010u >= 0x7ff0u - 0x0010u) ? 1 : 0;
    const uint32_t tmp = hi32 - 0x3fe60000u;
    const int i = (int)((tmp >> 14) & 63u);
    const int k = (int)tmp >> 20;
    const double z = __hiloint2double((int)(hi32 - (tmp & 0xfff00000u)), __double2loint(x));
    const double2 ic = tab[i];
    const double invc = ic.x, logc = ic.y;
    const double kd = (double)k;

    const double t3 = add_(kd, logc);
    const double r = fma_(z, invc, -1.0);
    const double q01 = fma_(r, A0[1], A0[0]);
    const double t1 = mul_(InvLn2hi, r);
    const double u = fma_(InvLn2hi, r, -t1);
    const double hi = add_(t1, t3);
    double lo = sub_(t3, hi);
    const double t2 = fma_(r, InvLn2lo, u);
    const double r2 = mul_(r, r);
    lo = add_(lo, t1);
    lo = add_(lo, t2);
    const double q23 = fma_(r, A0[3], A0[2]);
    const double r4 = mul_(r2, r2);
    const double q45 = fma_(r, A0[5], A0[4]);
    const double s = fma_(q23, r2, q01);
    const double p = fma_(q45, r4, s);
    const double y = fma_(r2, p, lo);
    return add_(y, hi);
}
// 64 x {1/c, log2 c} into shared memory (call from >= 64 threads, then __syncthreads())
__device__ __forceinline__ void dvs_log2_stage_table(double2* tab) {
    if (threadIdx.x < 64)
        tab[threadIdx.x] = make_double2(dvs_log2_data_dev.tab[2 * threadIdx.x], dvs_log2_data_dev.tab[2 * threadIdx.x + 1]);
}
#endif
