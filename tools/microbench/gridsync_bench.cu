// Cost of one grid-wide barrier on a persistent one-CTA-per-SM kernel: cooperative-groups grid.sync()
// against a hand-written sense-reversing barrier (one atomic per CTA + acquire spin).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gridsync_bench gridsync_bench.cu && ./gridsync_bench
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512) k_cg(int iters, unsigned* sink) {
    cg::grid_group g = cg::this_grid();
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += i;
        g.sync();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

__device__ __forceinline__ void my_barrier(unsigned* count, unsigned* gen, unsigned nblocks, unsigned& local_gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = local_gen + 1;
        __threadfence();
        if (atomicAdd(count, 1u) == nblocks - 1) {
            *count = 0;
            __threadfence();
            asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(gen), "r"(target) : "memory");
        } else {
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(gen) : "memory");
            } while (v != target);
        }
    }
    local_gen += 1;
    __syncthreads();
}

__global__ void __launch_bounds__(512) k_my(int iters, unsigned* count, unsigned* gen, unsigned* sink) {
    unsigned lg = 0, acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += i;
        my_barrier(count, gen, gridDim.x, lg);
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

static void flag_bench();
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    unsigned *sink, *bar;
    cudaMalloc(&sink, 4);
    cudaMalloc(&bar, 8);
    cudaMemset(bar, 0, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int per = 1; per <= 2; ++per) {
        int grid = p.multiProcessorCount * per, iters = 2000;
        void* args[] = {&iters, &sink};
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(512), args, 0, 0);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
        }
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("cg grid.sync   grid=%d: %.3f us per barrier\n", grid, ms * 1000 / iters);
        unsigned* cnt = bar;
        unsigned* gen = bar + 1;
        void* args2[] = {&iters, &cnt, &gen, &sink};
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(bar, 0, 8);
            cudaEventRecord(a);
            cudaLaunchCooperativeKernel((void*)k_my, dim3(grid), dim3(512), args2, 0, 0);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
        }
        cudaEventElapsedTime(&ms, a, b);
        printf("custom barrier grid=%d: %.3f us per barrier (%s)\n", grid, ms * 1000 / iters, cudaGetErrorString(cudaGetLastError()));
    }
    flag_bench();
    return 0;
}

// ---- flag latency: one-way visibility of a st.relaxed.gpu seen by a polling ld.relaxed.gpu in another CTA ----
// ping-pong between CTA 0 and CTA `peer` (one thread each): 2 one-way latencies per iteration; and the
// all-to-all form used by the SM-replicated selection rounds: every CTA publishes a tagged 16-byte slot,
// then polls all the others' (threads = slots) — one "exchange" per iteration.
__device__ __forceinline__ void st128(void* p, unsigned long long lo, unsigned long long hi) {
    asm volatile("{ .reg .b128 v; mov.b128 v, {%1, %2}; st.relaxed.gpu.global.b128 [%0], v; }" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void ld128(const void* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{ .reg .b128 v; ld.relaxed.gpu.global.b128 v, [%2]; mov.b128 {%0, %1}, v; }" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
__global__ void k_pingpong(unsigned long long* slots, int iters, int peer) {
    if (threadIdx.x != 0) return;
    unsigned long long lo, hi;
    if (blockIdx.x == 0) {
        for (int i = 1; i <= iters; ++i) {
            st128(slots, 1, (unsigned long long)i);
            do ld128(slots + 4, lo, hi); while (hi != (unsigned long long)i);
        }
    } else if ((int)blockIdx.x == peer) {
        for (int i = 1; i <= iters; ++i) {
            do ld128(slots, lo, hi); while (hi != (unsigned long long)i);
            st128(slots + 4, 1, (unsigned long long)i);
        }
    }
}
__global__ void __launch_bounds__(512) k_exchange(unsigned long long* slots, int iters, int work) {
    __shared__ double sink[512];
    const unsigned b = blockIdx.x, G = gridDim.x, t = threadIdx.x;
    double acc = t;
    for (int i = 1; i <= iters; ++i) {
        for (int w = 0; w < work; ++w) acc = acc * 1.0000001 + 1e-9;  // dependent FP64 chain = "compute"
        sink[t] = acc;
        __syncthreads();
        if (t == 0) st128(slots + ((i & 1) * 256 + b) * 4, (unsigned long long)__double_as_longlong(sink[1]), (unsigned long long)i);
        if (t < G) {
            unsigned long long lo, hi;
            do ld128(slots + ((i & 1) * 256 + t) * 4, lo, hi); while (hi != (unsigned long long)i);
            acc += (double)(lo & 1);
        }
        __syncthreads();
    }
    if (acc == 123.456) slots[0] = 1;
}

// leader form: every CTA publishes its slot, CTA 0 gathers them all and publishes one result slot that
// the other CTAs poll (2 one-way latencies, but only G + G pollers instead of G * G)
__global__ void __launch_bounds__(512) k_exchange_leader(unsigned long long* slots, int iters, int work) {
    __shared__ double sink[512];
    const unsigned b = blockIdx.x, G = gridDim.x, t = threadIdx.x;
    unsigned long long* result = slots + 2 * 256 * 4;
    double acc = t;
    for (int i = 1; i <= iters; ++i) {
        for (int w = 0; w < work; ++w) acc = acc * 1.0000001 + 1e-9;
        sink[t] = acc;
        __syncthreads();
        if (t == 0) st128(slots + ((i & 1) * 256 + b) * 4, (unsigned long long)__double_as_longlong(sink[1]), (unsigned long long)i);
        if (b == 0) {
            if (t < G) {
                unsigned long long lo, hi;
                do ld128(slots + ((i & 1) * 256 + t) * 4, lo, hi); while (hi != (unsigned long long)i);
                acc += (double)(lo & 1);
            }
            __syncthreads();
            if (t == 0) st128(result + (i & 1) * 4, 7, (unsigned long long)i);
        } else if (t == 0) {
            unsigned long long lo, hi;
            do ld128(result + (i & 1) * 4, lo, hi); while (hi != (unsigned long long)i);
            acc += (double)(lo & 1);
        }
        __syncthreads();
    }
    if (acc == 123.456) slots[0] = 1;
}
static void flag_bench() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    unsigned long long* slots;
    cudaMalloc(&slots, 2 * 256 * 32 + 64);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms;
    for (int peer : {1, 2, 37, 74, 147}) {
        int iters = 20000;
        cudaMemset(slots, 0, 2 * 256 * 32);
        void* args[] = {&slots, &iters, &peer};
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void*)k_pingpong, dim3(p.multiProcessorCount), dim3(32), args, 0, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("ping-pong CTA0 <-> CTA%d: %.0f ns one-way (%s)\n", peer, ms * 1e6 / iters / 2, cudaGetErrorString(cudaGetLastError()));
    }
    for (int work : {0, 100}) {
        int iters = 5000;
        cudaMemset(slots, 0, 2 * 256 * 32);
        void* args[] = {&slots, &iters, &work};
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void*)k_exchange, dim3(p.multiProcessorCount), dim3(512), args, 0, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("all-to-all tagged exchange, %d CTAs, %d dependent DFMA of work: %.0f ns per exchange\n", p.multiProcessorCount, work, ms * 1e6 / iters);
        cudaMemset(slots, 0, 2 * 256 * 32 + 64);
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void*)k_exchange_leader, dim3(p.multiProcessorCount), dim3(512), args, 0, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("leader gather + broadcast,   %d CTAs, %d dependent DFMA of work: %.0f ns per exchange (%s)\n", p.multiProcessorCount, work, ms * 1e6 / iters, cudaGetErrorString(cudaGetLastError()));
    }
}
