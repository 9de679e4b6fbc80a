// Packed host->device upload of sequence bytes.
//
// The reference hands the hot path 1 byte per base (src/zarr_io.rs:309-313).  Over PCIe Gen5
// (~55 GB/s measured) the 42 GB benchmark set costs 0.76 s, 30x the GPU compute.  DNA bytes are
// 0..3 almost everywhere, so the host threads pack 4 bases per byte (hostpack.cpp, AVX2) while
// earlier blocks are already in flight, the GPU unpacks them back to the reference's byte layout in
// HBM (k_unpack, HBM-bound) and patches the rare bytes >= 4 from an exception list.  The device
// ends up with exactly the caller's bytes; nothing downstream changes.
//
// Pipeline: the stream is cut into 256 MB blocks of 4 MB sub-chunks.  Worker threads take
// sub-chunks in order from an atomic counter and pack into a ring of pinned staging blocks; the
// calling thread ships every completed block (cudaMemcpyAsync of packed bytes + exceptions,
// k_unpack, k_patch) and releases its ring slot with an event.  Blocks whose exception list
// overflows (>= 1/64 invalid) are sent raw instead.
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"

namespace dvs {

struct PackExc {
    uint32_t* pos;
    uint8_t* val;
    std::atomic<uint32_t>* count;
    uint32_t cap;
};
void pack_bytes(const uint8_t* src, size_t n, uint8_t* dst_block, uint32_t rel0, PackExc& ex);

constexpr size_t kUpBlock = 256ull << 20;   // input bytes per block
constexpr size_t kUpSub = 4ull << 20;       // input bytes per worker sub-chunk
constexpr int kUpRing = 3;                  // staging blocks in flight
constexpr uint32_t kUpExcCap = (uint32_t)(kUpBlock / 64);

struct UploadStage {
    uint8_t* h_packed = nullptr;   // pinned, kUpRing * kUpBlock/4
    uint32_t* h_pos = nullptr;     // pinned, kUpRing * kUpExcCap
    uint8_t* h_val = nullptr;      // pinned, kUpRing * kUpExcCap
    uint8_t* d_packed = nullptr;   // device mirrors
    uint32_t* d_pos = nullptr;
    uint8_t* d_val = nullptr;
    cudaEvent_t done[kUpRing] = {};
    bool ready = false;
};

// 16 output bytes per thread from 4 packed bytes
__global__ void k_unpack(const uint8_t* __restrict__ packed, uint8_t* __restrict__ out, size_t n) {
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (g >= n) return;
    const uint32_t p = *reinterpret_cast<const uint32_t*>(packed + (g >> 2));
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t b = (p >> (8 * j)) & 0xFFu;
        w[j] = (b & 3u) | (((b >> 2) & 3u) << 8) | (((b >> 4) & 3u) << 16) | (((b >> 6) & 3u) << 24);
    }
    if (g + 16 <= n) {
        *reinterpret_cast<uint4*>(out + g) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (size_t i = g; i < n; ++i) out[i] = (uint8_t)((w[(i - g) >> 2] >> (8 * ((i - g) & 3))) & 0xFFu);
    }
}

__global__ void k_patch(uint8_t* __restrict__ out, const uint32_t* __restrict__ pos, const uint8_t* __restrict__ val,
                        uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[pos[i]] = val[i];
}

static int stage_init(dvs_ctx* ctx, UploadStage& st) {
    if (st.ready) return DVS_OK;
    DVS_CUDA_TRY(cudaHostAlloc((void**)&st.h_packed, kUpRing * (kUpBlock / 4), cudaHostAllocDefault));
    DVS_CUDA_TRY(cudaHostAlloc((void**)&st.h_pos, (size_t)kUpRing * kUpExcCap * sizeof(uint32_t), cudaHostAllocDefault));
    DVS_CUDA_TRY(cudaHostAlloc((void**)&st.h_val, (size_t)kUpRing * kUpExcCap, cudaHostAllocDefault));
    DVS_CUDA_TRY(cudaMalloc((void**)&st.d_packed, kUpRing * (kUpBlock / 4)));
    DVS_CUDA_TRY(cudaMalloc((void**)&st.d_pos, (size_t)kUpRing * kUpExcCap * sizeof(uint32_t)));
    DVS_CUDA_TRY(cudaMalloc((void**)&st.d_val, (size_t)kUpRing * kUpExcCap));
    for (int i = 0; i < kUpRing; ++i) DVS_CUDA_TRY(cudaEventCreateWithFlags(&st.done[i], cudaEventDisableTiming | cudaEventBlockingSync));
    st.ready = true;
    (void)ctx;
    return DVS_OK;
}

void upload_stage_free(void* p) {
    auto* st = (UploadStage*)p;
    if (!st) return;
    if (st->ready) {
        cudaFreeHost(st->h_packed);
        cudaFreeHost(st->h_pos);
        cudaFreeHost(st->h_val);
        cudaFree(st->d_packed);
        cudaFree(st->d_pos);
        cudaFree(st->d_val);
        for (int i = 0; i < kUpRing; ++i) cudaEventDestroy(st->done[i]);
    }
    delete st;
}

// Copies src[0..total) to d_dst[0..total) through the packed pipeline on ctx->stream.
//
// Work is claimed from both ends of the block list: worker threads pack sub-chunks from the front;
// whenever the stream has drained (PCIe idle because packing is the slower side) the calling thread
// steals the LAST unclaimed block and ships it raw, so the link and the host cores finish together.
int upload_packed(dvs_ctx* ctx, const uint8_t* src, uint8_t* d_dst, size_t total) {
    if (!ctx->upload_stage) ctx->upload_stage = new UploadStage();
    UploadStage& st = *(UploadStage*)ctx->upload_stage;
    DVS_TRY(stage_init(ctx, st));
    cudaStream_t stream = ctx->stream;
    const size_t nblocks = (total + kUpBlock - 1) / kUpBlock;
    const size_t subs_per_block = kUpBlock / kUpSub;

    std::vector<std::atomic<uint32_t>> sub_done(nblocks);
    std::vector<std::atomic<uint32_t>> exc_count(nblocks);
    for (size_t b = 0; b < nblocks; ++b) {
        sub_done[b].store(0);
        exc_count[b].store(0);
    }
    // claim state (under claim_lock): sub-chunks [0, front_sub) belong to the packers, blocks
    // [tail_block, nblocks) were stolen for raw transfer
    std::atomic_flag claim_lock = ATOMIC_FLAG_INIT;
    size_t front_sub = 0, tail_block = nblocks;
    auto lock = [&] { while (claim_lock.test_and_set(std::memory_order_acquire)) std::this_thread::yield(); };
    auto unlock = [&] { claim_lock.clear(std::memory_order_release); };
    std::atomic<size_t> released{(size_t)kUpRing};  // packed blocks [0, released) may be written by workers
    std::atomic<bool> abort{false};

    auto worker = [&] {
        for (;;) {
            lock();
            const size_t sidx = front_sub;
            const bool have = sidx / subs_per_block < tail_block && sidx * kUpSub < total;
            if (have) ++front_sub;
            unlock();
            if (!have || abort.load()) return;
            const size_t b = sidx / subs_per_block;
            while (b >= released.load(std::memory_order_acquire)) {  // ring slot still in flight
                if (abort.load()) return;
                std::this_thread::yield();
            }
            const size_t off = sidx * kUpSub;
            const size_t n = std::min(kUpSub, total - off);
            const int slot = (int)(b % kUpRing);
            PackExc ex{st.h_pos + (size_t)slot * kUpExcCap, st.h_val + (size_t)slot * kUpExcCap, &exc_count[b], kUpExcCap};
            pack_bytes(src + off, n, st.h_packed + (size_t)slot * (kUpBlock / 4), (uint32_t)(off - b * kUpBlock), ex);
            sub_done[b].fetch_add(1, std::memory_order_release);
        }
    };
    // one process per GPU: share the host cores between the local ranks (torchrun exports
    // LOCAL_WORLD_SIZE); DVS_HOST_THREADS overrides
    unsigned nthreads = std::max(1u, std::thread::hardware_concurrency());
    if (const char* lws = getenv("LOCAL_WORLD_SIZE")) {
        // leave one core per rank for the thread that drives the GPU (an 8-GPU box with 32 cores was
        // measured: fully subscribing the cores starved the other ranks' kernel-launching threads)
        const unsigned share = nthreads / (unsigned)std::max(1, atoi(lws));
        nthreads = atoi(lws) > 1 ? std::max(1u, share > 1 ? share - 1 : share) : std::max(1u, share);
    }
    if (const char* ht = getenv("DVS_HOST_THREADS")) nthreads = (unsigned)std::max(1, atoi(ht));
    nthreads = (unsigned)std::min<size_t>(nthreads, std::max<size_t>(1, (total + kUpSub - 1) / kUpSub));
    // Stealing balances the two resources whatever the host looks like: on the 16-core 1-GPU bench host
    // packing (host DRAM bandwidth, ~108 GB/s) and the link finish together either way (0.40 s); with
    // two ranks sharing 24 cores the packers alone fell to 0.71 s per rank while each GPU's own link sat
    // idle.  DVS_UPLOAD_STEAL=0 disables it.
    const char* steal_env = getenv("DVS_UPLOAD_STEAL");
    const bool steal = !(steal_env && steal_env[0] == '0');
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(worker);

    int rc = DVS_OK;
    auto fail = [&](cudaError_t e) {
        set_error("packed upload failed: %s", cudaGetErrorString(e));
        rc = DVS_ERR_CUDA;
    };
    auto send_raw = [&](size_t b) {
        const size_t off = b * kUpBlock, n = std::min(kUpBlock, total - off);
        cudaError_t e = cudaMemcpyAsync(d_dst + off, src + off, n, cudaMemcpyHostToDevice, stream);
        ctx->last_upload_wire_bytes += n;
        if (e != cudaSuccess) fail(e);
    };
    size_t b = 0;  // next packed block to ship
    while (rc == DVS_OK) {
        lock();
        const size_t tail = tail_block;
        unlock();
        if (b >= tail) break;
        const size_t off = b * kUpBlock;
        const size_t n = std::min(kUpBlock, total - off);
        const uint32_t want = (uint32_t)((n + kUpSub - 1) / kUpSub);
        if (sub_done[b].load(std::memory_order_acquire) < want) {
            // block b is still being packed.  If the link is idle, take a raw block from the far end.
            bool stole = false;
            if (steal && cudaStreamQuery(stream) == cudaSuccess) {
                lock();
                // only whole blocks no packer has touched: strictly beyond the block of front_sub
                if (tail_block > b + 1 && tail_block - 1 > (front_sub == 0 ? 0 : (front_sub - 1) / subs_per_block)) {
                    --tail_block;
                    const size_t sb = tail_block;
                    unlock();
                    send_raw(sb);
                    stole = true;
                } else {
                    unlock();
                }
            }
            if (!stole) std::this_thread::sleep_for(std::chrono::microseconds(50));  // the workers own the cores
            continue;
        }
        const int slot = (int)(b % kUpRing);
        const uint32_t nexc = exc_count[b].load();
        cudaError_t e = cudaSuccess;
        if (nexc > kUpExcCap) {
            send_raw(b);  // too many bytes >= 4 in this block: ship it as it is
        } else {
            ctx->last_upload_wire_bytes += (n + 3) / 4 + (uint64_t)nexc * 5;
            uint8_t* dp = st.d_packed + (size_t)slot * (kUpBlock / 4);
            e = cudaMemcpyAsync(dp, st.h_packed + (size_t)slot * (kUpBlock / 4), (n + 3) / 4, cudaMemcpyHostToDevice, stream);
            if (e == cudaSuccess) {
                const size_t groups = (n + 15) / 16;
                k_unpack<<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(dp, d_dst + off, n);
                ctx->launches++;
                e = cudaGetLastError();
            }
            if (e == cudaSuccess && nexc) {
                uint32_t* dpos = st.d_pos + (size_t)slot * kUpExcCap;
                uint8_t* dval = st.d_val + (size_t)slot * kUpExcCap;
                e = cudaMemcpyAsync(dpos, st.h_pos + (size_t)slot * kUpExcCap, nexc * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(dval, st.h_val + (size_t)slot * kUpExcCap, nexc, cudaMemcpyHostToDevice, stream);
                if (e == cudaSuccess) {
                    k_patch<<<(nexc + 255) / 256, 256, 0, stream>>>(d_dst + off, dpos, dval, nexc);
                    ctx->launches++;
                    e = cudaGetLastError();
                }
            }
        }
        if (rc == DVS_OK && e == cudaSuccess) e = cudaEventRecord(st.done[slot], stream);
        if (e != cudaSuccess) {
            fail(e);
            break;
        }
        // the slot of block b is reused by block b + kUpRing: release it once this block has landed.
        // Releasing lags one block behind so the copy engine always has a block queued.
        if (b + 1 >= (size_t)kUpRing - 1) {
            const size_t rel_block = b + 1 - ((size_t)kUpRing - 1);  // oldest block still owning a slot
            e = cudaEventSynchronize(st.done[rel_block % kUpRing]);
            if (e != cudaSuccess) {
                fail(e);
                break;
            }
            released.store(rel_block + 1 + kUpRing, std::memory_order_release);
        }
        ++b;
    }
    if (rc != DVS_OK) abort.store(true);
    released.store(nblocks + kUpRing + 1, std::memory_order_release);
    for (auto& th : pool) th.join();
    return rc;
}

}  // namespace dvs

extern "C" int dvs_debug_pack_host(const uint8_t* src, uint64_t n, uint8_t* packed, uint32_t* exc_pos,
                                   uint8_t* exc_val, uint32_t cap, uint32_t* nexc) {
    std::atomic<uint32_t> count{0};
    dvs::PackExc ex{exc_pos, exc_val, &count, cap};
    dvs::pack_bytes(src, (size_t)n, packed, 0u, ex);
    *nexc = count.load();
    return DVS_OK;
}
