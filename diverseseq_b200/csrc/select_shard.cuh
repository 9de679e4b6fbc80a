// select_shard.cuh — the cross-GPU exchange of the candidate-sharded selection rounds
// Part of select.cu (included inside namespace dvs before the kernels that use it).
#pragma once

// Candidate-sharded rounds over several GPUs (dvs_select_sharded): every GPU holds the
// rows of all records and replays the identical state updates, but scores only the window positions it
// owns (position % world == rank).  The per-round collective is the leaders' all-reduce(min) of
// {first_true, first_unsure}: each leader stores its pair, tagged, into its slot of every peer's window
// (st.relaxed.sys over NVLink, two self-validating 8-byte words) and polls its own slots - 16 bytes per
// peer and round, no NCCL, no host.  world == 1 leaves the single-GPU protocol unchanged.
struct ShardArgs {
    int rank, world;
    unsigned tag_base;                  // (launch number << 20): exchange tags are unique across launches
    unsigned char* xbase[kCommMaxWorld];  // window base of every rank (own window included)
};
constexpr unsigned long long kSelWatchdogNs = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ unsigned long long sel_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// all-reduce(min) of the pair over the ranks; threads 0..world-1 of the calling CTA take part, the result is
// returned through shared memory (s_pair) to the whole CTA after the caller's next barrier.  Returns false
// when a peer did not answer in time.
__device__ __forceinline__ void shard_exchange_min(const ShardArgs& sh, unsigned xs, unsigned ft, unsigned fu,
                                                   unsigned* s_ft, unsigned* s_fu, unsigned* s_dead) {
    const unsigned t = threadIdx.x;
    if (t < (unsigned)sh.world) {
        const unsigned tag = sh.tag_base + xs;
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(sh.xbase[t] + kCommSelOff) +
                                  ((xs & 1u) * kCommMaxWorld + (unsigned)sh.rank) * 2u;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(((unsigned long long)tag << 32) | ft) : "memory");
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(((unsigned long long)tag << 32) | fu) : "memory");
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(sh.xbase[sh.rank] + kCommSelOff) +
                                        ((xs & 1u) * kCommMaxWorld + t) * 2u;
        unsigned long long w0, w1;
        const unsigned long long t0 = sel_gtime();
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w0) : "l"(src) : "memory");
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w1) : "l"(src + 1) : "memory");
            if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) break;
            if (sel_gtime() - t0 > kSelWatchdogNs) {
                *s_dead = 1u;
                w0 = w1 = 0xFFFFFFFFull;
                break;
            }
        }
        atomicMin(s_ft, (unsigned)w0);
        atomicMin(s_fu, (unsigned)w1);
    }
}

