// coresident.cu - does a second kernel's CTA fit on an SM beside one CTA of the counting kernel's shape (1,024 threads,
// 56 registers, 144 KB of shared memory), and what decides it?  Kernel A ("holder") occupies every SM and waits;
// kernel B ("probe", 128 threads) is launched on another stream, records its %smid and stays until all its CTAs have
// arrived (or 20 ms have passed).  Reported: how many probe CTAs arrived while the holders were still there and on
// how many distinct SMs they sat.  Variables: registers of either kernel (forced with live accumulators), the
// probe's shared memory, and the holder's shared-memory carve-out preference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coresident coresident.cu && ./coresident
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// NACC live accumulators force the register allocation up to the __maxnreg__ cap
template <int NACC>
__device__ __forceinline__ float burn(const float* src, int rounds) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = src[i];
    for (int r = 0; r < rounds; ++r)
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = acc[i] * 1.0001f + acc[(i + 1) % NACC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    return s;
}

template <int REGS, int NACC>
__global__ void __maxnreg__(REGS) k_hold(volatile unsigned* release, unsigned* arrived, unsigned long long max_ns,
                                         const float* src, float* sink, int rounds) {
    extern __shared__ unsigned char raw[];
    const float v = burn<NACC>(src, rounds);
    if (v == 12345.f) sink[threadIdx.x] = v;
    if (threadIdx.x == 0) {
        raw[0] = 1;
        atomicAdd(arrived, 1u);
        const unsigned long long t0 = now();
        while (!*release && now() - t0 < max_ns) __nanosleep(1000);
    }
    __syncthreads();
}

template <int REGS, int NACC>
__global__ void __maxnreg__(REGS) k_probe(unsigned* smid_out, volatile unsigned* holders_alive, unsigned* saw_alive,
                                          unsigned* arrived, unsigned want, int smem, const float* src, float* sink,
                                          int rounds) {
    extern __shared__ unsigned char raw[];
    const float v = burn<NACC>(src, rounds);
    if (v == 12345.f) sink[threadIdx.x] = v;
    if (threadIdx.x == 0) {
        if (smem) raw[0] = 1;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        smid_out[blockIdx.x] = smid;
        saw_alive[blockIdx.x] = *holders_alive;
        atomicAdd(arrived, 1u);
        const unsigned long long t0 = now();
        while (*(volatile unsigned*)arrived < want && now() - t0 < 20ull * 1000 * 1000) __nanosleep(1000);
    }
    __syncthreads();
}

__global__ void k_mark(unsigned* w, unsigned v) { *w = v; }

template <int RA, int NA, int RB, int NB>
void run(int smemB, int carveA, int carveB) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned *d, *h;
    float* fsrc;
    CK(cudaMalloc(&d, 4096 * 4));
    CK(cudaMemset(d, 0, 4096 * 4));
    CK(cudaMalloc(&fsrc, 4096 * 4));
    CK(cudaMemset(fsrc, 0, 4096 * 4));
    CK(cudaMallocHost(&h, 4096 * 4));
    cudaStream_t sa, sb;
    CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
    const int smemA = 147456;
    auto ka = k_hold<RA, NA>;
    auto kb = k_probe<RB, NB>;
    CK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, smemA));
    if (smemB > 48 * 1024) CK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smemB));
    CK(cudaFuncSetAttribute(ka, cudaFuncAttributePreferredSharedMemoryCarveout, carveA));
    CK(cudaFuncSetAttribute(kb, cudaFuncAttributePreferredSharedMemoryCarveout, carveB));
    cudaFuncAttributes fa, fb;
    CK(cudaFuncGetAttributes(&fa, ka));
    CK(cudaFuncGetAttributes(&fb, kb));
    // d[0] = release, d[1] = holders arrived, d[2] = holders alive, d[3] = probes arrived, d[16..] smid, d[1024..] saw_alive
    k_mark<<<1, 1, 0, sa>>>(d + 2, 1u);
    ka<<<sms, 1024, smemA, sa>>>(d, d + 1, 100ull * 1000 * 1000, fsrc, fsrc + 2048, 1);
    k_mark<<<1, 1, 0, sa>>>(d + 2, 0u);  // runs when every holder has gone
    for (;;) {  // wait until all holders are resident
        CK(cudaMemcpyAsync(h, d + 1, 4, cudaMemcpyDeviceToHost, sb));
        CK(cudaStreamSynchronize(sb));
        if (h[0] >= (unsigned)sms) break;
    }
    kb<<<sms, 128, smemB, sb>>>(d + 16, d + 2, d + 1024, d + 3, (unsigned)sms, smemB, fsrc, fsrc + 2048, 1);
    k_mark<<<1, 1, 0, sb>>>(d, 1u);  // release the holders once the probe has finished
    CK(cudaStreamSynchronize(sb));
    CK(cudaStreamSynchronize(sa));
    CK(cudaMemcpy(h, d, 4096 * 4, cudaMemcpyDeviceToHost));
    std::vector<unsigned> smid(h + 16, h + 16 + sms);
    std::sort(smid.begin(), smid.end());
    const int distinct = (int)(std::unique(smid.begin(), smid.end()) - smid.begin());
    int alive = 0;
    for (int i = 0; i < sms; ++i) alive += h[1024 + i] ? 1 : 0;
    printf("holder 1024 thr x %2d regs, 144 KB, carve-out %3d | probe 128 thr x %2d regs, %5d B, carve-out %3d -> %3d of %d "
           "probe CTAs arrived beside the holders, on %3d distinct SMs\n", fa.numRegs, carveA, fb.numRegs, smemB, carveB,
           alive, sms, distinct);
    cudaFree(d); cudaFree(fsrc); cudaFreeHost(h); cudaStreamDestroy(sa); cudaStreamDestroy(sb);
}

int main() {
    const int dflt = cudaSharedmemCarveoutDefault, mx = cudaSharedmemCarveoutMaxShared;
    run<56, 48, 64, 58>(0, dflt, dflt);        // registers only: 1024 x 56 + 128 x 64 = 65,536
    run<56, 48, 64, 58>(8000, dflt, dflt);     // fits the holder's default carve-out
    run<56, 48, 64, 58>(30000, dflt, dflt);    // needs a larger carve-out than the holder chose
    run<56, 48, 64, 58>(68000, dflt, dflt);
    run<56, 48, 64, 58>(30000, mx, mx);        // ... which the holder asks for up front
    run<56, 48, 64, 58>(68000, mx, mx);
    run<56, 48, 64, 58>(68000, mx, dflt);
    run<56, 48, 40, 28>(68000, mx, mx);
    run<48, 40, 64, 58>(68000, mx, mx);
    run<64, 58, 64, 58>(8000, dflt, dflt);     // no registers left: must not fit
    return 0;
}
