// comm.cu — peer windows over NVLink (see comm.cuh): creation / mapping, device-side barrier, push flags,
// small all-reduce(min), generic all-gather.  Replaces the torch.distributed / NCCL plumbing of round 1
// ("tiny all-reduce over NVLink", "all-gather of the row blocks").
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>

#include "comm.cuh"

namespace dvs {

struct CommHandle {  // what dvs_comm_create hands out for the host to exchange (DVS_COMM_HANDLE_BYTES)
    uint32_t magic;
    int32_t pid;
    int32_t device;
    int32_t rank;
    uint64_t ptr;
    uint64_t bytes;
    int32_t ipc_valid;
    int32_t pad;
    cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(CommHandle) <= DVS_COMM_HANDLE_BYTES, "handle blob too small");
constexpr uint32_t kCommMagic = 0x44565343u;  // "DVSC"

struct CommPeers {
    uint8_t* base[kCommMaxWorld];
    int rank, world;
};

static CommPeers peers_of(const dvs_comm* c) {
    CommPeers p;
    for (int r = 0; r < kCommMaxWorld; ++r) p.base[r] = r < c->world ? c->peer[r] : nullptr;
    p.rank = c->rank;
    p.world = c->world;
    return p;
}

constexpr unsigned long long kWatchdogNs = 20ull * 1000 * 1000 * 1000;  // a wait longer than this is a hang
// DVS_WATCHDOG_MS overrides it (tests)
static unsigned long long watchdog_ns() {
    const char* e = getenv("DVS_WATCHDOG_MS");
    return e && atoll(e) > 0 ? (unsigned long long)atoll(e) * 1000000ull : kWatchdogNs;
}

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_sys_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_sys_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// every rank writes `epoch` into its word of every peer's flag array, then waits for all of its own words
__global__ void k_comm_barrier(CommPeers P, uint64_t epoch, unsigned long long wd) {
    const int t = threadIdx.x;
    if (t >= P.world) return;
    __threadfence_system();
    st_sys_u64(reinterpret_cast<uint64_t*>(P.base[t] + kCommBarrierOff) + P.rank, epoch);
    const uint64_t* mine = reinterpret_cast<const uint64_t*>(P.base[P.rank] + kCommBarrierOff) + t;
    const unsigned long long t0 = gtime();
    while (ld_sys_u64(mine) < epoch) {
        if (gtime() - t0 > wd) {
            *reinterpret_cast<volatile uint32_t*>(P.base[P.rank] + kCommErrOff) = 1u;
            break;
        }
    }
    __threadfence_system();
}

// wait until the pushes of `epoch` from every peer have landed (their flag copies travel behind the data)
__global__ void k_comm_wait_push(CommPeers P, uint64_t epoch, unsigned long long wd) {
    const int t = threadIdx.x;
    if (t >= P.world || t == P.rank) return;
    const uint64_t* mine = reinterpret_cast<const uint64_t*>(P.base[P.rank] + kCommPushOff) + t;
    const unsigned long long t0 = gtime();
    while (ld_sys_u64(mine) < epoch) {
        if (gtime() - t0 > wd) {
            *reinterpret_cast<volatile uint32_t*>(P.base[P.rank] + kCommErrOff) = 2u;
            break;
        }
    }
    __threadfence_system();
}

__global__ void k_comm_set_u64(uint64_t* p, uint64_t v) { *p = v; }

// all-reduce(min) of one u32: value and the exchange's tag travel in one 8-byte word
__global__ void k_comm_min_u32(CommPeers P, uint32_t tag, const uint32_t* in, uint32_t* out, unsigned long long wd) {
    __shared__ uint32_t s_v[kCommMaxWorld];
    const int t = threadIdx.x;
    if (t < P.world) {
        const uint64_t word = ((uint64_t)(*in) << 32) | tag;
        st_sys_u64(reinterpret_cast<uint64_t*>(P.base[t] + kCommMinOff) + (tag & 1u) * kCommMaxWorld + P.rank, word);
        const uint64_t* mine =
            reinterpret_cast<const uint64_t*>(P.base[P.rank] + kCommMinOff) + (tag & 1u) * kCommMaxWorld + t;
        const unsigned long long t0 = gtime();
        uint64_t w;
        while ((uint32_t)(w = ld_sys_u64(mine)) != tag) {
            if (gtime() - t0 > wd) {
                *reinterpret_cast<volatile uint32_t*>(P.base[P.rank] + kCommErrOff) = 3u;
                w = ~0ull;
                break;
            }
        }
        s_v[t] = (uint32_t)(w >> 32);
    }
    __syncthreads();
    if (t == 0) {
        uint32_t m = s_v[0];
        for (int r = 1; r < P.world; ++r) m = min(m, s_v[r]);
        *out = m;
    }
}

int comm_heap_alloc(dvs_comm* c, uint64_t bytes, uint64_t* off) {
    bytes = (bytes + 255) & ~255ull;
    if (bytes == 0) bytes = 256;
    for (size_t i = 0; i < c->blocks.size(); ++i) {
        CommBlock& b = c->blocks[i];
        if (b.used || b.bytes < bytes) continue;
        if (b.bytes > bytes) {
            CommBlock rest{b.off + bytes, b.bytes - bytes, false};
            b.bytes = bytes;
            c->blocks.insert(c->blocks.begin() + i + 1, rest);
        }
        c->blocks[i].used = true;
        *off = c->blocks[i].off;
        return DVS_OK;
    }
    set_error("the peer window is too small: %llu bytes requested from a window of %llu (pass a larger window_bytes "
              "to dvs_comm_create)", (unsigned long long)bytes, (unsigned long long)c->window_bytes);
    return DVS_ERR_ARG;
}

void comm_heap_free(dvs_comm* c, uint64_t off) {
    for (size_t i = 0; i < c->blocks.size(); ++i)
        if (c->blocks[i].off == off && c->blocks[i].used) {
            c->blocks[i].used = false;
            if (i + 1 < c->blocks.size() && !c->blocks[i + 1].used) {
                c->blocks[i].bytes += c->blocks[i + 1].bytes;
                c->blocks.erase(c->blocks.begin() + i + 1);
            }
            if (i > 0 && !c->blocks[i - 1].used) {
                c->blocks[i - 1].bytes += c->blocks[i].bytes;
                c->blocks.erase(c->blocks.begin() + i);
            }
            return;
        }
}

int comm_check_error(dvs_ctx* ctx, dvs_comm* c, const char* what) {
    uint32_t flag = 0;
    DVS_CUDA_TRY(cudaMemcpyAsync(&flag, c->window + kCommErrOff, sizeof flag, cudaMemcpyDeviceToHost, ctx->stream));
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        set_error("%s: a device-side wait on a peer gave up after %.1f s (code %u): a rank is missing or died", what,
                  watchdog_ns() / 1e9, flag);
        return DVS_ERR_CUDA;
    }
    return DVS_OK;
}

int comm_check_error_async(dvs_ctx* ctx, dvs_comm* c, const char* what) {
    const char* env = getenv("DVS_COMM_CHECK");
    return (env && env[0] == '1') ? comm_check_error(ctx, c, what) : DVS_OK;
}

int comm_host_rendezvous(dvs_ctx* ctx, dvs_comm* c) {
    if (!c->host_barrier || c->world == 1) return DVS_OK;
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    DVS_CUDA_TRY(cudaStreamSynchronize(c->side));
    c->host_barrier(c->host_barrier_arg);
    return DVS_OK;
}

// host-synchronised all-reduce(min): values travel through the same slots, the waiting happens on the host
__global__ void k_comm_min_local(CommPeers P, uint32_t tag, uint32_t* out) {
    uint32_t m = 0xFFFFFFFFu;
    for (int r = 0; r < P.world; ++r) {
        const uint64_t w = *(reinterpret_cast<const uint64_t*>(P.base[P.rank] + kCommMinOff) + (tag & 1u) * kCommMaxWorld + r);
        m = min(m, (uint32_t)(w >> 32));
    }
    *out = m;
}
__global__ void k_comm_min_post(CommPeers P, uint32_t tag, const uint32_t* in) {
    const int t = threadIdx.x;
    if (t < P.world)
        st_sys_u64(reinterpret_cast<uint64_t*>(P.base[t] + kCommMinOff) + (tag & 1u) * kCommMaxWorld + P.rank,
                   ((uint64_t)(*in) << 32) | tag);
}

int comm_barrier(dvs_ctx* ctx, dvs_comm* c) {
    if (c->world == 1) return DVS_OK;
    if (c->host_barrier) return comm_host_rendezvous(ctx, c);
    ++c->epoch;
    k_comm_barrier<<<1, 32, 0, ctx->stream>>>(peers_of(c), c->epoch, watchdog_ns());
    DVS_LAUNCHED(ctx);
    return DVS_OK;
}

int comm_push_begin(dvs_ctx* ctx, dvs_comm* c) {
    ++c->push_epoch;
    k_comm_set_u64<<<1, 1, 0, ctx->stream>>>(c->d_epoch_src, c->push_epoch);
    DVS_LAUNCHED(ctx);
    return DVS_OK;
}

int comm_push_commit(dvs_ctx* ctx, dvs_comm* c) {
    for (int d = 1; d < c->world; ++d) {
        const int r = (c->rank + d) % c->world;
        DVS_CUDA_TRY(cudaMemcpyAsync(c->peer[r] + kCommPushOff + (size_t)c->rank * 8, c->d_epoch_src, 8,
                                     cudaMemcpyDeviceToDevice, c->side));
    }
    DVS_CUDA_TRY(cudaEventRecord(c->ev_pushed, c->side));
    return DVS_OK;
}

int comm_push_wait(dvs_ctx* ctx, dvs_comm* c) {
    DVS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, c->ev_pushed, 0));
    if (c->host_barrier) return comm_host_rendezvous(ctx, c);  // every rank's pushes have completed
    if (c->world > 1) {
        k_comm_wait_push<<<1, 32, 0, ctx->stream>>>(peers_of(c), c->push_epoch, watchdog_ns());
        DVS_LAUNCHED(ctx);
    }
    return DVS_OK;
}

int comm_min_u32(dvs_ctx* ctx, dvs_comm* c, const uint32_t* d_in, uint32_t* d_out) {
    ++c->min_tag;
    if (c->host_barrier) {
        k_comm_min_post<<<1, 32, 0, ctx->stream>>>(peers_of(c), (uint32_t)c->min_tag, d_in);
        DVS_LAUNCHED(ctx);
        DVS_TRY(comm_host_rendezvous(ctx, c));
        k_comm_min_local<<<1, 1, 0, ctx->stream>>>(peers_of(c), (uint32_t)c->min_tag, d_out);
        DVS_LAUNCHED(ctx);
        return DVS_OK;
    }
    k_comm_min_u32<<<1, 32, 0, ctx->stream>>>(peers_of(c), (uint32_t)c->min_tag, d_in, d_out, watchdog_ns());
    DVS_LAUNCHED(ctx);
    return DVS_OK;
}

}  // namespace dvs

using namespace dvs;

extern "C" {

int dvs_comm_create(dvs_ctx* ctx, int rank, int world, uint64_t window_bytes, dvs_comm** out, void* handle_out) {
    if (!ctx || !out || !handle_out || world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world) {
        set_error("dvs_comm_create: bad argument (world must be 1..%d)", kCommMaxWorld);
        return DVS_ERR_ARG;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    auto* c = new dvs_comm();
    c->device = ctx->device;
    c->rank = rank;
    c->world = world;
    c->window_bytes = std::max<uint64_t>((window_bytes + 255) & ~255ull, kCommCtrlBytes + (1u << 20));
    cudaError_t e = cudaMalloc((void**)&c->window, c->window_bytes);  // plain cudaMalloc: pool memory cannot be IPC-exported
    if (e == cudaSuccess) e = cudaMemset(c->window, 0, kCommCtrlBytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_epoch_src, 256);
    if (e == cudaSuccess) e = cudaMemset(c->d_epoch_src, 0, 256);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pushed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fan, cudaEventDisableTiming);
    for (int d = 1; d < world && e == cudaSuccess; ++d) {
        e = cudaStreamCreateWithFlags(&c->peer_stream[d], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->peer_ev[d], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_pushed, c->side);
    if (e != cudaSuccess) {
        set_error("dvs_comm_create: %s (window of %llu bytes)", cudaGetErrorString(e), (unsigned long long)c->window_bytes);
        dvs_comm_destroy(c);
        return DVS_ERR_CUDA;
    }
    c->blocks.push_back({kCommCtrlBytes, c->window_bytes - kCommCtrlBytes, false});
    c->peer[rank] = c->window;
    CommHandle h;
    memset(&h, 0, sizeof h);
    h.magic = kCommMagic;
    h.pid = (int32_t)getpid();
    h.device = ctx->device;
    h.rank = rank;
    h.ptr = (uint64_t)(uintptr_t)c->window;
    h.bytes = c->window_bytes;
    h.ipc_valid = cudaIpcGetMemHandle(&h.ipc, c->window) == cudaSuccess ? 1 : 0;
    (void)cudaGetLastError();
    memset(handle_out, 0, DVS_COMM_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof h);
    *out = c;
    return DVS_OK;
}

int dvs_comm_connect(dvs_ctx* ctx, dvs_comm* c, const void* handles) {
    if (!ctx || !c || !handles) {
        set_error("dvs_comm_connect: NULL argument");
        return DVS_ERR_ARG;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    const uint8_t* hb = static_cast<const uint8_t*>(handles);
    for (int r = 0; r < c->world; ++r) {
        CommHandle h;
        memcpy(&h, hb + (size_t)r * DVS_COMM_HANDLE_BYTES, sizeof h);
        if (h.magic != kCommMagic || h.rank != r || h.bytes != c->window_bytes) {
            set_error("dvs_comm_connect: handle %d is not rank %d's window of %llu bytes (every rank must pass the same "
                      "world and window_bytes)", r, r, (unsigned long long)c->window_bytes);
            return DVS_ERR_ARG;
        }
        if (r == c->rank) continue;
        if (h.pid == (int32_t)getpid()) {  // another context of this process: the pointer itself
            if (h.device != c->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("dvs_comm_connect: no peer access from device %d to %d: %s", c->device, h.device,
                              cudaGetErrorString(e));
                    return DVS_ERR_CUDA;
                }
                (void)cudaGetLastError();
            }
            c->peer[r] = reinterpret_cast<uint8_t*>((uintptr_t)h.ptr);
        } else {
            if (!h.ipc_valid) {
                set_error("dvs_comm_connect: rank %d could not export its window (cudaIpcGetMemHandle failed there)", r);
                return DVS_ERR_CUDA;
            }
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                set_error("dvs_comm_connect: cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
                return DVS_ERR_CUDA;
            }
            c->peer[r] = static_cast<uint8_t*>(p);
            c->ipc_opened[r] = true;
        }
    }
    c->connected = true;
    return DVS_OK;
}

int dvs_comm_set_host_barrier(dvs_comm* c, void (*barrier)(void*), void* arg) {
    if (!c) {
        set_error("dvs_comm_set_host_barrier: NULL argument");
        return DVS_ERR_ARG;
    }
    c->host_barrier = barrier;
    c->host_barrier_arg = arg;
    return DVS_OK;
}

int dvs_comm_rank(const dvs_comm* c) { return c->rank; }
int dvs_comm_world(const dvs_comm* c) { return c->world; }

int dvs_comm_barrier(dvs_ctx* ctx, dvs_comm* c) {
    if (!ctx || !c || !c->connected) {
        set_error("dvs_comm_barrier: the communicator is not connected");
        return DVS_ERR_ARG;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    DVS_TRY(comm_barrier(ctx, c));
    return comm_check_error(ctx, c, "dvs_comm_barrier");
}

int dvs_comm_allgatherv(dvs_ctx* ctx, dvs_comm* c, const void* d_src, const uint64_t* bytes_per_rank, void* d_dst) {
    if (!ctx || !c || !c->connected || !bytes_per_rank || !d_dst) {
        set_error("dvs_comm_allgatherv: bad argument / communicator not connected");
        return DVS_ERR_ARG;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    uint64_t mx = 0, off_me = 0;
    std::vector<uint64_t> off(c->world + 1, 0);
    for (int r = 0; r < c->world; ++r) {
        mx = std::max(mx, bytes_per_rank[r]);
        off[r + 1] = off[r] + bytes_per_rank[r];
    }
    off_me = off[c->rank];
    const uint64_t mine = bytes_per_rank[c->rank];
    if (c->world == 1) {
        if (mine) DVS_CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, mine, cudaMemcpyDeviceToDevice, st));
        return DVS_OK;
    }
    // stage in the symmetric heap (the source may be pool memory, which peers cannot map), pull from every peer
    uint64_t hoff = 0;
    DVS_TRY(comm_heap_alloc(c, mx, &hoff));
    int rc = DVS_OK;
    cudaError_t e = cudaSuccess;
    if (mine) e = cudaMemcpyAsync(c->window + hoff, d_src, mine, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) rc = comm_barrier(ctx, c);  // every rank's staging is complete
    for (int d = 0; d < c->world && e == cudaSuccess && rc == DVS_OK; ++d) {
        const int r = (c->rank + d) % c->world;
        if (bytes_per_rank[r])
            e = cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + off[r], c->peer[r] + hoff, bytes_per_rank[r],
                                cudaMemcpyDeviceToDevice, st);
    }
    (void)off_me;
    if (e == cudaSuccess && rc == DVS_OK) rc = comm_barrier(ctx, c);  // nobody reuses its staging before all have read it
    comm_heap_free(c, hoff);
    if (e != cudaSuccess) {
        set_error("dvs_comm_allgatherv: %s", cudaGetErrorString(e));
        return DVS_ERR_CUDA;
    }
    if (rc != DVS_OK) return rc;
    return comm_check_error(ctx, c, "dvs_comm_allgatherv");
}

void dvs_comm_destroy(dvs_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->side) {
        cudaStreamSynchronize(c->side);
        cudaStreamDestroy(c->side);
    }
    for (int r = 0; r < c->world; ++r)
        if (c->ipc_opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
    for (int d = 0; d < kCommMaxWorld; ++d) {
        if (c->peer_stream[d]) {
            cudaStreamSynchronize(c->peer_stream[d]);
            cudaStreamDestroy(c->peer_stream[d]);
        }
        if (c->peer_ev[d]) cudaEventDestroy(c->peer_ev[d]);
    }
    if (c->ev_fan) cudaEventDestroy(c->ev_fan);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_pushed) cudaEventDestroy(c->ev_pushed);
    if (c->d_epoch_src) cudaFree(c->d_epoch_src);
    if (c->window) cudaFree(c->window);
    delete c;
}

}  // extern "C"
