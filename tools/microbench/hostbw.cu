// hostbw.cu - the host-side bounds of the end-to-end path (host buffers of 1 byte per base -> GPUs):
//   (1) host DRAM read bandwidth with T threads streaming over private buffers (what the 2-bit packers can at best
//       consume), (2) pinned host -> device copy bandwidth with 1..N GPUs copying at once, (3) both at once.
//   nvcc -O3 -o hostbw hostbw.cu -lpthread && ./hostbw [threads] [GB per buffer]
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <vector>

static double now_s() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

struct ReadJob {
    const uint64_t* p;
    size_t words;
    int reps;
    uint64_t sum;
};
static void* read_worker(void* a) {
    ReadJob* j = (ReadJob*)a;
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int r = 0; r < j->reps; ++r)
        for (size_t i = 0; i + 4 <= j->words; i += 4) {
            s0 += j->p[i]; s1 += j->p[i + 1]; s2 += j->p[i + 2]; s3 += j->p[i + 3];
        }
    j->sum = s0 + s1 + s2 + s3;
    return nullptr;
}

static double host_read(int threads, std::vector<uint64_t*>& bufs, size_t words, int reps) {
    std::vector<pthread_t> th(threads);
    std::vector<ReadJob> jobs(threads);
    const double t0 = now_s();
    for (int t = 0; t < threads; ++t) {
        jobs[t] = {bufs[t], words, reps, 0};
        pthread_create(&th[t], nullptr, read_worker, &jobs[t]);
    }
    uint64_t s = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], nullptr);
        s += jobs[t].sum;
    }
    const double dt = now_s() - t0;
    if (s == 42) printf(" ");
    return (double)threads * words * 8 * reps / dt / 1e9;
}

int main(int argc, char** argv) {
    const int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN);
    const int threads = argc > 1 ? atoi(argv[1]) : ncpu;
    const double gb = argc > 2 ? atof(argv[2]) : 1.0;
    const size_t words = (size_t)(gb * 1e9 / 8);
    int ngpu = 0;
    cudaGetDeviceCount(&ngpu);
    printf("host: %d CPUs online, %d reader threads, %.1f GB per buffer; %d GPUs\n", ncpu, threads, gb, ngpu);

    std::vector<uint64_t*> bufs(threads);
    for (int t = 0; t < threads; ++t) {
        bufs[t] = (uint64_t*)malloc(words * 8);
        memset(bufs[t], t + 1, words * 8);
    }
    for (int t : {1, 2, 4, 8, threads})
        if (t <= threads) printf("host DRAM read, %2d threads: %7.1f GB/s\n", t, host_read(t, bufs, words, 3));

    // pinned H2D, n GPUs at once
    std::vector<void*> hp(ngpu), dp(ngpu);
    std::vector<cudaStream_t> st(ngpu);
    const size_t bytes = (size_t)2e9;
    for (int g = 0; g < ngpu; ++g) {
        cudaSetDevice(g);
        cudaMallocHost(&hp[g], bytes);
        memset(hp[g], g + 1, bytes);
        cudaMalloc(&dp[g], bytes);
        cudaStreamCreate(&st[g]);
    }
    auto h2d = [&](int n, int reps) {
        for (int g = 0; g < n; ++g) {  // warm
            cudaSetDevice(g);
            cudaMemcpyAsync(dp[g], hp[g], bytes, cudaMemcpyHostToDevice, st[g]);
        }
        for (int g = 0; g < n; ++g) { cudaSetDevice(g); cudaStreamSynchronize(st[g]); }
        const double t0 = now_s();
        for (int r = 0; r < reps; ++r)
            for (int g = 0; g < n; ++g) {
                cudaSetDevice(g);
                cudaMemcpyAsync(dp[g], hp[g], bytes, cudaMemcpyHostToDevice, st[g]);
            }
        for (int g = 0; g < n; ++g) { cudaSetDevice(g); cudaStreamSynchronize(st[g]); }
        return (double)n * bytes * reps / (now_s() - t0) / 1e9;
    };
    for (int n = 1; n <= ngpu; n *= 2) {
        const double bw = h2d(n, 4);
        printf("pinned H2D, %d GPU%s at once: %7.1f GB/s total, %6.1f GB/s per GPU\n", n, n > 1 ? "s" : " ", bw, bw / n);
    }
    if (ngpu) {  // the packers read while all GPUs copy
        struct Arg { std::vector<uint64_t*>* b; size_t w; int t; double bw; } arg{&bufs, words, threads, 0};
        pthread_t bg;
        pthread_create(&bg, nullptr, [](void* a) -> void* {
            Arg* x = (Arg*)a;
            x->bw = host_read(x->t, *x->b, x->w, 6);
            return nullptr;
        }, &arg);
        const double bw = h2d(ngpu, 8);
        pthread_join(bg, nullptr);
        printf("both at once: H2D %d GPUs %7.1f GB/s total (%5.1f per GPU) while %d threads read %7.1f GB/s\n", ngpu, bw,
               bw / ngpu, threads, arg.bw);
    }
    return 0;
}
