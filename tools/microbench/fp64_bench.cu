// FP64 throughput on B200: DFMA (SIMT) vs DMMA (mma.sync f64 shapes) - decides the Euclidean Gram tile kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_bench fp64_bench.cu && ./fp64_bench
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_dfma(double* out, int iters, double x, double y) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 independent m8n8k4 accumulators per warp
__global__ void k_dmma884(double* out, int iters, double x, double y) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    double a = x + threadIdx.x, b = y + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef WIDE_SHAPES
// m16n8k8: A 4 regs, B 2 regs, C 4 regs per lane
__global__ void k_dmma1688(double* out, int iters, double x, double y) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
    double a0 = x + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = y + threadIdx.x, b1 = b0 + 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int thr : {256, 512, 1024}) {
        const int grid = sms * (2048 / thr);
        float ms;
        k_dfma<<<grid, thr>>>(out, 100, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_dfma<<<grid, thr>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA      thr=%4d: %7.2f TFLOP/s\n", thr, 2.0 * grid * thr * 16.0 * iters / ms / 1e9);
        k_dmma884<<<grid, thr>>>(out, 100, 1.0, 1e-9); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_dmma884<<<grid, thr>>>(out, iters, 1.0, 1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA 8x8x4 thr=%4d: %7.2f TFLOP/s\n", thr, 2.0 * (grid * (double)thr / 32) * 8.0 * 256.0 * iters / ms / 1e9);
#ifdef WIDE_SHAPES
        k_dmma1688<<<grid, thr>>>(out, 100, 1.0, 1e-9); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_dmma1688<<<grid, thr>>>(out, iters, 1.0, 1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA 16x8x8 thr=%4d: %7.2f TFLOP/s\n", thr, 2.0 * (grid * (double)thr / 32) * 4.0 * 1024.0 * iters / ms / 1e9);
#endif
    }
    return 0;
}
