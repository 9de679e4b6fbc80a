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
    return 0;
}
