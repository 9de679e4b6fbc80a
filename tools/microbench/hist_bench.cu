// Micro-benchmarks that size the k-mer counting kernel's bottleneck on B200.
// Not part of the product; run under gpurun:  nvcc ... -o hist_bench hist_bench.cu && ./hist_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint64_t sm64(uint64_t z) { z += 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
__global__ void k_gen(uint8_t* p, size_t n, int skew) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    uint64_t h = sm64(i);
    uint64_t out = 0;
    for (int j = 0; j < 8; ++j) {
        uint32_t r = (h >> (8 * j)) & 0xFF;
        uint32_t b = skew ? (r < 110 ? 2 : r < 200 ? 0 : r < 230 ? 1 : 3) : (r & 3);  // skew: A 43% T 35% C 12% G 10%
        out |= (uint64_t)b << (8 * j);
    }
    *(uint64_t*)(p + i) = out;
}
__device__ __forceinline__ uint32_t pack4(uint32_t w) { return (w * 0x40100401u) >> 24; }
__device__ __forceinline__ uint32_t pack16(uint4 v) { return (pack4(v.x) << 24) | (pack4(v.y) << 16) | (pack4(v.z) << 8) | pack4(v.w); }

// MODE 0: smem atomics, every k-mer (bins = 4^K)          MODE 1: no atomics (xor sink)  -> load/ALU ceiling
// MODE 2: smem atomics on (K+1)-mers at stride 2           MODE 3: global RED into table of 4^K bins
// MODE 4: smem atomics, COPIES sub-histograms by warp      MODE 5: (K+2)-mers at stride 3 with u16-packed counters
template <int MODE, int K, int COPIES>
__global__ void k_hist(const uint8_t* __restrict__ seq, size_t n, uint32_t* __restrict__ out, size_t chunk) {
    extern __shared__ uint32_t hist[];
    constexpr uint32_t KK = (MODE == 2 || MODE == 6) ? K + 1 : (MODE == 5 ? K + 2 : K);
    constexpr uint32_t BINS = 1u << (2 * KK);
    constexpr uint32_t WORDS = (MODE == 5) ? BINS / 2 : BINS;
    constexpr uint32_t mask = BINS - 1u;
    uint32_t sink = 0;
    for (size_t c0 = (size_t)blockIdx.x * chunk; c0 < n; c0 += (size_t)gridDim.x * chunk) {
        if (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6) {
            for (uint32_t i = threadIdx.x; i < WORDS * COPIES; i += blockDim.x) hist[i] = 0;
            __syncthreads();
        }
        uint32_t* myhist = hist + ((MODE == 4) ? ((threadIdx.x >> 5) % COPIES) * BINS : 0);
        size_t cend = c0 + chunk < n ? c0 + chunk : n;
        for (size_t a = c0 + (size_t)threadIdx.x * 16; a < cend; a += (size_t)blockDim.x * 16) {
            uint4 cur = __ldg((const uint4*)(seq + a));
            uint4 prev = __ldg((const uint4*)(seq + (a >= 16 ? a - 16 : 0)));
            uint32_t pc = pack16(cur), pp = pack16(prev);
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) sink ^= __funnelshift_r(pc, pp, 2 * (15 - j)) & mask;
            } else if (MODE == 0 || MODE == 4) {
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(&myhist[__funnelshift_r(pc, pp, 2 * (15 - j)) & mask], 1u);
            } else if (MODE == 2) {
#pragma unroll
                for (int j = 1; j < 16; j += 2) atomicAdd(&hist[__funnelshift_r(pc, pp, 2 * (15 - j)) & mask], 1u);
            } else if (MODE == 5) {
                // 16 positions per thread is not a multiple of 3: take j = (phase .. 15 step 3), phase by block of 16
                int ph = (int)((a / 16) % 3);
#pragma unroll
                for (int j0 = 0; j0 < 16; j0 += 3) {
                    int j = j0 + ((3 - ph) % 3 == 0 ? 0 : (3 - ph) % 3 == 1 ? 1 : 2);
                    if (j < 16) {
                        uint32_t idx = __funnelshift_r(pc, pp, 2 * (15 - j)) & mask;
                        atomicAdd(&hist[idx >> 1], 1u << (16 * (idx & 1)));
                    }
                }
            } else if (MODE == 6) {
                // lane-group replication: lanes [g*32/COPIES, (g+1)*32/COPIES) use replica g, which owns the
                // banks [g*32/COPIES, ...): address = (idx / BPG) * 32 + idx % BPG + g * BPG, BPG = 32/COPIES
                constexpr uint32_t BPG = 32 / COPIES;
                const uint32_t goff = ((threadIdx.x & 31) / BPG) * BPG;
#pragma unroll
                for (int j = 1; j < 16; j += 2) {
                    uint32_t idx = __funnelshift_r(pc, pp, 2 * (15 - j)) & mask;
                    atomicAdd(&hist[(idx / BPG) * 32 + (idx % BPG) + goff], 1u);
                }
            } else if (MODE == 3) {
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(&out[__funnelshift_r(pc, pp, 2 * (15 - j)) & mask], 1u);
            }
        }
        if (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6) {
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < WORDS * COPIES; i += blockDim.x) { uint32_t c = hist[i]; if (c) atomicAdd(&out[i % WORDS], c); }
            __syncthreads();
        }
    }
    if (MODE == 1 && sink == 0x12345) out[0] = sink;
}

// raw ATOMS throughput: PATTERN 0 conflict-free (lane-distinct banks, distinct addresses), 1 random bins, 2 same address
template <int PATTERN>
__global__ void k_atoms(uint32_t* out, int iters, uint32_t bins) {
    extern __shared__ uint32_t hist[];
    for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t idx;
            if (PATTERN == 0) idx = ((it * 8 + u) * 32 + (threadIdx.x & 31) + (threadIdx.x >> 5) * 32 * 7) & (bins - 1);
            else if (PATTERN == 1) { x = x * 1664525u + 1013904223u; idx = (x >> 8) & (bins - 1); }
            else idx = 5;
            atomicAdd(&hist[idx], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = hist[5];
}

// dependent FP64 add chain latency (what bounds the exact sequential entropy sum)
__global__ void k_dadd_chain(const double* in, double* out, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc = __dadd_rn(acc, in[i & 1023]);
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}

template <class F> float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

template <int MODE, int K, int COPIES>
void run_hist(const char* name, const uint8_t* seq, size_t n, uint32_t* out, int threads, int ctas_per_sm, int sms) {
    constexpr uint32_t KK = (MODE == 2 || MODE == 6) ? K + 1 : (MODE == 5 ? K + 2 : K);
    size_t smem = (MODE == 1 || MODE == 3) ? 0 : (size_t)(MODE == 5 ? (1u << (2 * KK)) / 2 : (1u << (2 * KK))) * 4 * COPIES;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_hist<MODE, K, COPIES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = sms * ctas_per_sm;
    CK(cudaMemset(out, 0, (size_t)4 << (2 * 12)));
    float ms = time_ms([&] { k_hist<MODE, K, COPIES><<<grid, threads, smem>>>(seq, n, out, 1 << 20); });
    CK(cudaGetLastError());
    printf("%-44s K=%2d thr=%4d cta/sm=%d smem=%6zuB : %8.3f ms  %8.1f Gbp/s\n", name, K, threads, ctas_per_sm, smem, ms, n / ms / 1e6);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d MHz\n", p.name, sms, p.clockRate / 1000);
    size_t n = (size_t)1 << 30;  // 1 Gbase, > L2
    uint8_t* seq; uint32_t* out;
    CK(cudaMalloc(&seq, n + 64)); CK(cudaMalloc(&out, (size_t)4 << (2 * 12)));
    for (int skew = 0; skew < 2; ++skew) {
        k_gen<<<(unsigned)((n / 8 + 255) / 256), 256>>>(seq, n, skew); CK(cudaDeviceSynchronize());
        printf("---- data: %s ----\n", skew ? "skewed (A43 T35 C12 G10)" : "uniform");
        run_hist<1, 6, 1>("no atomics (load+pack ceiling)", seq, n, out, 512, 3, sms);
        run_hist<1, 6, 1>("no atomics (load+pack ceiling)", seq, n, out, 1024, 2, sms);
        run_hist<0, 6, 1>("smem atomics / k-mer", seq, n, out, 256, 8, sms);
        run_hist<0, 6, 1>("smem atomics / k-mer", seq, n, out, 512, 3, sms);
        run_hist<0, 6, 1>("smem atomics / k-mer", seq, n, out, 512, 4, sms);
        run_hist<0, 6, 1>("smem atomics / k-mer", seq, n, out, 1024, 2, sms);
        run_hist<0, 4, 1>("smem atomics / k-mer", seq, n, out, 512, 4, sms);
        run_hist<0, 7, 1>("smem atomics / k-mer", seq, n, out, 512, 3, sms);
        run_hist<0, 7, 1>("smem atomics / k-mer", seq, n, out, 1024, 2, sms);
        run_hist<4, 6, 2>("smem atomics, 2 copies by warp", seq, n, out, 512, 4, sms);
        run_hist<4, 6, 4>("smem atomics, 4 copies by warp", seq, n, out, 512, 3, sms);
        run_hist<4, 6, 8>("smem atomics, 8 copies by warp", seq, n, out, 1024, 1, sms);
        run_hist<2, 6, 1>("(k+1)-mers stride 2 (half the atomics)", seq, n, out, 512, 3, sms);
        run_hist<2, 6, 1>("(k+1)-mers stride 2 (half the atomics)", seq, n, out, 1024, 2, sms);
        run_hist<5, 6, 1>("(k+2)-mers stride 3, u16-packed", seq, n, out, 1024, 1, sms);
        run_hist<6, 6, 1>("stride 2, lane-group replicas x1 (=MODE 2 layout)", seq, n, out, 1024, 1, sms);
        run_hist<6, 6, 2>("stride 2, lane-group replicas x2 (16 banks each)", seq, n, out, 1024, 1, sms);
        run_hist<6, 5, 1>("stride 2, replicas x1, k=5", seq, n, out, 1024, 2, sms);
        run_hist<6, 5, 2>("stride 2, replicas x2, k=5", seq, n, out, 1024, 2, sms);
        run_hist<6, 5, 4>("stride 2, replicas x4, k=5 (8 banks each)", seq, n, out, 1024, 2, sms);
        run_hist<6, 5, 8>("stride 2, replicas x8, k=5 (4 banks each)", seq, n, out, 1024, 1, sms);
        run_hist<3, 8, 1>("global RED, 256 KB table", seq, n, out, 512, 4, sms);
        run_hist<3, 10, 1>("global RED, 4 MB table", seq, n, out, 512, 4, sms);
        run_hist<3, 12, 1>("global RED, 64 MB table", seq, n, out, 512, 4, sms);
    }
    // raw ATOMS
    for (int thr : {256, 512, 1024}) {
        int iters = 2000, grid = sms * (2048 / thr);
        double ops = (double)grid * thr * iters * 8;
        CK(cudaFuncSetAttribute(k_atoms<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
        float m0 = time_ms([&] { k_atoms<0><<<grid, thr, 16384>>>(out, iters, 4096); });
        float m1 = time_ms([&] { k_atoms<1><<<grid, thr, 16384>>>(out, iters, 4096); });
        float m2 = time_ms([&] { k_atoms<2><<<grid, thr, 16384>>>(out, iters, 4096); });
        printf("ATOMS thr=%4d: conflict-free %.1f G/s  random-4096 %.1f G/s  same-addr %.1f G/s  (per SM per clk @%d MHz: %.2f / %.2f / %.2f)\n", thr,
               ops / m0 / 1e6, ops / m1 / 1e6, ops / m2 / 1e6, p.clockRate / 1000, ops / m0 / 1e6 / sms / (p.clockRate / 1e6), ops / m1 / 1e6 / sms / (p.clockRate / 1e6), ops / m2 / 1e6 / sms / (p.clockRate / 1e6));
    }
    {
        double* din; double* dout; CK(cudaMalloc(&din, 8192)); CK(cudaMalloc(&dout, 8 * 1024)); CK(cudaMemset(din, 0, 8192));
        int nadd = 1 << 20;
        float ms = time_ms([&] { k_dadd_chain<<<1, 32>>>(din, dout, nadd); });
        printf("dependent DADD chain: %.2f ns per add (%.1f cycles at %d MHz nominal)\n", ms * 1e6 / nadd, ms * 1e6 / nadd * (p.clockRate / 1e6), p.clockRate / 1000);
    }
    return 0;
}
