// MinHash bottom-s sketches and all-pairs mash distances.
//
// Replaces mash_sketch / get_kmer_hashes / hash_kmer / murmurhash3_32 of the reference
// (/root/reference/src/distance.rs:17-182) and the pure-Python pair loop mash_distance /
// mash_distances (/root/reference/diverse_seq/distance.py:119-175, 230-291).
//
// Sketch = the `sketch_size` smallest DISTINCT hash values over all fully-valid k-mer windows,
// ascending (fewer if there are not enough).  The reference builds a HashSet of ~L hashes and a
// max-heap; here:
//   1. k_mash_filter streams the sequence (16-byte loads, 2-bit packing as in count.cu), hashes
//      every valid window with the reference's hash (k rounds per window, not rolling) and keeps
//      only hashes <= T_r, a per-record threshold sized so that ~2s candidates survive
//      (hashes are ~uniform on u32);
//   2. the surviving (record<<32 | hash) keys of all records are sorted by one bitonic network
//      (shared-memory tiles + global steps);
//   3. k_mash_compact drops duplicates and writes the first s per record.
// If a record ends up with fewer than s distinct candidates while T_r < 2^32-1 (repetitive
// sequence), or its candidate buffer overflowed, only that record is redone with a larger
// threshold / buffer, so the result is exact for every input.
#include <algorithm>

#include <string.h>

#include "comm.cuh"
#include "common.cuh"

namespace dvs {

constexpr int kMashThreads = 256;
constexpr unsigned kSortTile = 2048;  // keys per shared-memory tile (1024 threads)

struct MashWork {
    uint64_t begin, end;  // 16-byte aligned absolute byte range
    uint32_t rec;         // record index in the seqset
    uint32_t slot;        // index into the active-record arrays
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// distance.rs:28-39, one "block" per byte
__device__ __forceinline__ uint32_t mm_round(uint32_t h, uint32_t b) {
    uint32_t k = b * 0xCC9E2D51u;
    k = rotl32(k, 15);
    k *= 0x1B873593u;
    h ^= k;
    h = rotl32(h, 13);
    return h * 5u + 0xE6546B64u;
}
// distance.rs:41-46
__device__ __forceinline__ uint32_t mm_fmix(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

// reverse complement of a packed k-mer (2 bits/base, first base most significant, k <= 16):
// complement is (b+2)%4 == b^2 (distance.rs:17-19), then reverse the base order
__device__ __forceinline__ uint32_t revcomp_packed(uint32_t v, int k) {
    uint32_t c = v ^ 0xAAAAAAAAu;
    uint32_t r = __brev(c);                                   // reverses bits, also within pairs
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);  // restore bit order inside each base
    return r >> (32 - 2 * k);
}

// hash_kmer (distance.rs:65-87) on a packed k-mer
__device__ __forceinline__ uint32_t hash_packed(uint32_t v, int k, bool canonical) {
    if (canonical) {
        uint32_t rc = revcomp_packed(v, k);
        if (rc < v) v = rc;  // lexicographic order == integer order of the packed value; ties keep kmer
    }
    uint32_t h = 0x9747B28Cu ^ (uint32_t)k;
    for (int i = k - 1; i >= 0; --i) h = mm_round(h, (v >> (2 * i)) & 3u);
    return mm_fmix(h);
}

__device__ __forceinline__ uint32_t mpack4(uint32_t w) { return (w * 0x40100401u) >> 24; }
__device__ __forceinline__ uint32_t mpack16(uint4 v) {
    return (mpack4(v.x) << 24) | (mpack4(v.y) << 16) | (mpack4(v.z) << 8) | mpack4(v.w);
}

struct MashActive {
    uint32_t thresh;   // keep hash <= thresh
    uint32_t cap;      // capacity of this record's candidate region
    uint64_t base;     // start of the region in the key buffer
};

__device__ __forceinline__ void mash_emit(uint32_t h, const MashActive& a, uint32_t rec_slot, uint32_t* cnt,
                                          unsigned long long* keys) {
    if (h <= a.thresh) {
        uint32_t pos = atomicAdd(&cnt[rec_slot], 1u);
        if (pos < a.cap) keys[a.base + pos] = ((unsigned long long)rec_slot << 32) | h;
    }
}

// fast path: num_states == 4, k <= 16
__global__ void __launch_bounds__(kMashThreads)
k_mash_filter(const uint8_t* __restrict__ seqs, const uint64_t* __restrict__ offsets, const MashWork* __restrict__ work,
              uint32_t nwork, uint32_t* __restrict__ next_item, int k, int canonical,
              const MashActive* __restrict__ active, uint32_t* __restrict__ cnt, unsigned long long* __restrict__ keys) {
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x;
    const uint32_t mask = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    for (;;) {
        if (tid == 0) s_item = atomicAdd(next_item, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        __syncthreads();
        if (item >= nwork) break;
        const MashWork w = work[item];
        const MashActive act = active[w.slot];
        const uint64_t start = offsets[w.rec], end = offsets[w.rec + 1];
        for (uint64_t a = w.begin + (uint64_t)tid * 16; a < w.end; a += (uint64_t)kMashThreads * 16) {
            const uint4 cur = __ldg(reinterpret_cast<const uint4*>(seqs + a));
            const uint4 prev = __ldg(reinterpret_cast<const uint4*>(seqs + a - 16));
            uint32_t any = cur.x | cur.y | cur.z | cur.w | prev.x | prev.y | prev.z | prev.w;
            const bool fast = ((any & 0xFCFCFCFCu) == 0) && (a >= start + 16) && (a + 16 <= end);
            if (fast) {
                const uint32_t pc = mpack16(cur), pp = mpack16(prev);
#pragma unroll 4
                for (int j = 0; j < 16; ++j) {
                    uint32_t v = __funnelshift_r(pc, pp, 2 * (15 - j)) & mask;
                    mash_emit(hash_packed(v, k, canonical != 0), act, w.slot, cnt, keys);
                }
            } else {
                const uint32_t wv[8] = {prev.x, prev.y, prev.z, prev.w, cur.x, cur.y, cur.z, cur.w};
                uint32_t run = 0, v = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const uint64_t p = a - 16 + i;
                    uint32_t b = (wv[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                    if (p < start || p >= end) b = 0xFFu;
                    if (b >= 4u) {
                        run = 0;
                        v = 0;
                    } else {
                        v = ((v << 2) | b) & mask;
                        ++run;
                        if (i >= 16 && run >= (uint32_t)k) mash_emit(hash_packed(v, k, canonical != 0), act, w.slot, cnt, keys);
                    }
                }
            }
        }
    }
}

// ---- k = 16, num_states = 4: the ctree default (diverse_seq/cluster.py), rebuilt around the pipe balance ----
// k_mash_filter spends 5 ALU-pipe + 3 FMA-pipe instructions per hash round and is ALU bound (ncu: ALU pipe
// 91 %, FMA pipe 39 %).  Here
//   * the per-base term rotl(b*C1,15)*C2 depends on the base alone: it is looked up ONCE per sequence position
//     (31 positions serve the 16 windows of a lane's block) from a 4-entry shared-memory table, for the base and
//     for its complement, instead of being recomputed in every round of every window;
//   * a round is then  h ^= t;  h = rotl(h,13)*5 + c.  The rotate is taken off the ALU pipe in two rounds out of
//     three: h * 2^13 as a 64-bit product gives (h << 13, h >> 19) in one IMAD.WIDE, and
//     rotl*5 + c = lo*5 + (hi*5 + c) is two more IMADs (the multiplier 2^13 comes from a kernel argument so
//     that the compiler cannot turn the product back into shifts);
//   * canonical k-mers: the reverse complement of the whole 31-base string is formed once, every window's
//     reverse complement is one funnel shift of it, and the rounds pick the forward or the complemented term
//     with one SEL.
// Same hash values bit for bit (distance.rs:21-87); blocks with an invalid byte or a record edge go through
// hash_packed like before.
template <bool CANON>
__global__ void __launch_bounds__(kMashThreads)
k_mash_filter16(const uint8_t* __restrict__ seqs, const uint64_t* __restrict__ offsets, const MashWork* __restrict__ work,
                uint32_t nwork, uint32_t* __restrict__ next_item, const MashActive* __restrict__ active,
                uint32_t* __restrict__ cnt, unsigned long long* __restrict__ keys, uint32_t two13) {
    __shared__ uint32_t s_item;
    __shared__ uint2 s_T[4];  // {term of base b, term of its complement b ^ 2}
    const int tid = threadIdx.x;
    if (tid < 4) {
        auto term = [](uint32_t b) { return rotl32(b * 0xCC9E2D51u, 15) * 0x1B873593u; };
        s_T[tid] = make_uint2(term((uint32_t)tid), term((uint32_t)tid ^ 2u));
    }
    __syncthreads();
    constexpr int k = 16;
    auto revpairs = [](uint32_t x) {  // complement + reverse the base order of 16 packed bases
        uint32_t r = __brev(x ^ 0xAAAAAAAAu);
        return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    };
    for (;;) {
        if (tid == 0) s_item = atomicAdd(next_item, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        __syncthreads();
        if (item >= nwork) break;
        const MashWork w = work[item];
        const MashActive act = active[w.slot];
        const uint64_t start = offsets[w.rec], end = offsets[w.rec + 1];
        for (uint64_t a = w.begin + (uint64_t)tid * 16; a < w.end; a += (uint64_t)kMashThreads * 16) {
            const uint4 cur = __ldg(reinterpret_cast<const uint4*>(seqs + a));
            const uint4 prev = __ldg(reinterpret_cast<const uint4*>(seqs + a - 16));
            uint32_t any = cur.x | cur.y | cur.z | cur.w | prev.x | prev.y | prev.z | prev.w;
            const bool fast = ((any & 0xFCFCFCFCu) == 0) && (a >= start + 16) && (a + 16 <= end);
            if (fast) {
                // terms of the 31 positions q = 0..30 (q = 15 + byte index of the block; the halo is q < 15)
                const uint32_t wv[8] = {prev.x, prev.y, prev.z, prev.w, cur.x, cur.y, cur.z, cur.w};
                uint32_t tf[31], tr[31];
#pragma unroll
                for (int q = 0; q < 31; ++q) {
                    const int bi = q + 1;  // byte index inside the 32 loaded bytes
                    const uint32_t b = (wv[bi >> 2] >> (8 * (bi & 3))) & 0xFFu;
                    const uint2 t = s_T[b];
                    tf[q] = t.x;
                    tr[q] = t.y;
                }
                const uint32_t pc = mpack16(cur), pp = mpack16(prev);
                uint32_t ylo = 0, yhi = 0;
                if (CANON) {  // reverse complement of the 32-base string pp:pc
                    ylo = revpairs(pp);
                    yhi = revpairs(pc);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    bool use_rc = false;
                    if (CANON) {
                        const uint32_t v = j == 15 ? pc : __funnelshift_r(pc, pp, 2 * (15 - j));
                        const uint32_t rc = j == 15 ? yhi : __funnelshift_r(ylo, yhi, 2 * (j + 1));
                        use_rc = rc < v;  // ties keep the k-mer itself
                    }
                    uint32_t h = 0x9747B28Cu ^ (uint32_t)k;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t t = CANON ? (use_rc ? tr[j + 15 - i] : tf[j + i]) : tf[j + i];
                        h ^= t;
                        if (i % 3 != 2) {  // rotl(h,13)*5 + c on the FMA pipe
                            const unsigned long long wide = (unsigned long long)h * two13;
                            const uint32_t lo = (uint32_t)wide, hi = (uint32_t)(wide >> 32);
                            h = lo * 5u + (hi * 5u + 0xE6546B64u);
                        } else {
                            h = rotl32(h, 13) * 5u + 0xE6546B64u;
                        }
                    }
                    mash_emit(mm_fmix(h), act, w.slot, cnt, keys);
                }
            } else {
                const uint32_t wv[8] = {prev.x, prev.y, prev.z, prev.w, cur.x, cur.y, cur.z, cur.w};
                uint32_t run = 0, v = 0;
#pragma unroll 1
                for (int i = 0; i < 32; ++i) {
                    const uint64_t p = a - 16 + i;
                    uint32_t b = (wv[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                    if (p < start || p >= end) b = 0xFFu;
                    if (b >= 4u) {
                        run = 0;
                        v = 0;
                    } else {
                        v = (v << 2) | b;
                        ++run;
                        if (i >= 16 && run >= (uint32_t)k) mash_emit(hash_packed(v, k, CANON), act, w.slot, cnt, keys);
                    }
                }
            }
        }
    }
}

// ---- per-record sort + unique + bottom-s in shared memory ----------------------------------------------------
// The threshold leaves ~2s + 512 candidates per record; one CTA sorts them in shared memory (bitonic, u32),
// drops duplicates and writes the first s - instead of one global bitonic network over the padded key buffer of
// all records (~30 launches over 16 M keys for 1,000 genomes).  A record with more than kSortRecMax candidates
// (a widened threshold, sketch_size = "all") is left to the global path.
constexpr unsigned kSortRecMax = 32768;
constexpr unsigned kSortRecThreads = 1024;
__global__ void __launch_bounds__(kSortRecThreads)
k_mash_sort_record(const unsigned long long* __restrict__ keys, const MashActive* __restrict__ active,
                   const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ slot_rec, uint64_t sketch_size,
                   uint32_t stride, uint32_t* __restrict__ sketches, uint32_t* __restrict__ lens,
                   uint32_t* __restrict__ ndistinct) {
    extern __shared__ uint32_t s_h[];
    __shared__ unsigned s_warp[kSortRecThreads / 32];
    __shared__ unsigned s_base;
    const unsigned slot = blockIdx.x, t = threadIdx.x;
    const MashActive act = active[slot];
    const unsigned n = min(cnt[slot], act.cap);
    if (n > kSortRecMax) return;  // (the host has seen the same count and sends this launch's records elsewhere)
    unsigned P = 2;
    while (P < n) P <<= 1;
    for (unsigned i = t; i < P; i += kSortRecThreads) s_h[i] = i < n ? (uint32_t)keys[act.base + i] : 0xFFFFFFFFu;
    __syncthreads();
    for (unsigned kk = 2; kk <= P; kk <<= 1)
        for (unsigned j = kk >> 1; j > 0; j >>= 1) {
            for (unsigned q = t; q < P / 2; q += kSortRecThreads) {
                const unsigned i = 2 * q - (q & (j - 1));
                const uint32_t x = s_h[i], y = s_h[i + j];
                if ((x > y) == ((i & kk) == 0)) {
                    s_h[i] = y;
                    s_h[i + j] = x;
                }
            }
            __syncthreads();
        }
    // the first n elements are the record's candidates in ascending order (padding sorts last)
    const uint32_t rec = slot_rec[slot];
    uint32_t* out = sketches + (size_t)rec * stride;
    const uint64_t want = min(sketch_size, (uint64_t)stride);
    if (t == 0) s_base = 0;
    __syncthreads();
    for (unsigned c = 0; c < n; c += kSortRecThreads) {
        const unsigned i = c + t;
        const unsigned flag = (i < n) && (i == 0 || s_h[i - 1] != s_h[i]);
        const unsigned lane = t & 31, wid = t >> 5;
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        unsigned woff = 0, tot = 0;
        for (unsigned wdx = 0; wdx < kSortRecThreads / 32; ++wdx) {
            if (wdx < wid) woff += s_warp[wdx];
            tot += s_warp[wdx];
        }
        const unsigned pos = s_base + woff + __popc(bal & ((1u << lane) - 1u));
        if (flag && pos < want) out[pos] = s_h[i];
        __syncthreads();
        if (t == 0) s_base += tot;
        __syncthreads();
    }
    if (t == 0) {
        ndistinct[slot] = s_base;
        lens[rec] = (uint32_t)min((uint64_t)s_base, want);
    }
}

// generic path (any num_states, any k): one thread per window start, bytes re-read from L1/L2
__global__ void __launch_bounds__(kMashThreads)
k_mash_filter_generic(const uint8_t* __restrict__ seqs, const uint64_t* __restrict__ offsets,
                      const MashWork* __restrict__ work, uint32_t nwork, uint32_t* __restrict__ next_item, int k,
                      uint32_t num_states, int canonical, const MashActive* __restrict__ active,
                      uint32_t* __restrict__ cnt, unsigned long long* __restrict__ keys) {
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(next_item, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        __syncthreads();
        if (item >= nwork) break;
        const MashWork w = work[item];
        const MashActive act = active[w.slot];
        const uint64_t start = offsets[w.rec], end = offsets[w.rec + 1];
        // windows whose LAST byte lies in [max(begin,start), min(end_item,end))
        const uint64_t lo = max(w.begin, start), hi = min(w.end, end);
        for (uint64_t last = lo + tid; last < hi; last += kMashThreads) {
            if (last + 1 < start + (uint64_t)k) continue;
            const uint8_t* km = seqs + (last + 1 - k);
            bool ok = true;
            for (int i = 0; i < k; ++i) ok = ok && (km[i] < num_states);
            if (!ok) continue;
            bool use_rc = false;
            if (canonical) {  // distance.rs:66-78
                for (int i = 0; i < k; ++i) {
                    uint8_t f = km[i], r = (uint8_t)((km[k - 1 - i] + 2) % 4);
                    if (f < r) break;
                    if (f > r) {
                        use_rc = true;
                        break;
                    }
                }
            }
            uint32_t h = 0x9747B28Cu ^ (uint32_t)k;
            for (int i = 0; i < k; ++i) {
                uint32_t b = use_rc ? (uint32_t)((km[k - 1 - i] + 2) % 4) : (uint32_t)km[i];
                h = mm_round(h, b);
            }
            mash_emit(mm_fmix(h), act, w.slot, cnt, keys);
        }
    }
}

// ---- bitonic sort of u64 keys (n a power of two, multiple of kSortTile) ----------------------

__device__ __forceinline__ void cmpswap(unsigned long long& a, unsigned long long& b, bool up) {
    if ((a > b) == up) {
        unsigned long long t = a;
        a = b;
        b = t;
    }
}

// sorts each tile completely (stages 2..kSortTile); tile direction follows the global network
__global__ void __launch_bounds__(kSortTile / 2) k_bitonic_tiles(unsigned long long* keys) {
    __shared__ unsigned long long s[kSortTile];
    const unsigned t = threadIdx.x;
    const size_t g0 = (size_t)blockIdx.x * kSortTile;
    s[t] = keys[g0 + t];
    s[t + kSortTile / 2] = keys[g0 + t + kSortTile / 2];
    __syncthreads();
    for (unsigned k = 2; k <= kSortTile; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            unsigned i = 2 * t - (t & (j - 1));  // index with bit j clear
            bool up = (((g0 + i) & k) == 0);
            cmpswap(s[i], s[i + j], up);
            __syncthreads();
        }
    }
    keys[g0 + t] = s[t];
    keys[g0 + t + kSortTile / 2] = s[t + kSortTile / 2];
}

// one global compare-exchange step (j >= kSortTile)
__global__ void k_bitonic_global(unsigned long long* keys, size_t n, size_t k, size_t j) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    size_t i = 2 * t - (t & (j - 1));
    bool up = ((i & k) == 0);
    unsigned long long a = keys[i], b = keys[i + j];
    if ((a > b) == up) {
        keys[i] = b;
        keys[i + j] = a;
    }
}

// steps j = kSortTile/2 .. 1 of stage k, fused in shared memory
__global__ void __launch_bounds__(kSortTile / 2) k_bitonic_merge_tiles(unsigned long long* keys, size_t k) {
    __shared__ unsigned long long s[kSortTile];
    const unsigned t = threadIdx.x;
    const size_t g0 = (size_t)blockIdx.x * kSortTile;
    s[t] = keys[g0 + t];
    s[t + kSortTile / 2] = keys[g0 + t + kSortTile / 2];
    __syncthreads();
    const bool up = ((g0 & k) == 0);
    for (unsigned j = kSortTile >> 1; j > 0; j >>= 1) {
        unsigned i = 2 * t - (t & (j - 1));
        cmpswap(s[i], s[i + j], up);
        __syncthreads();
    }
    keys[g0 + t] = s[t];
    keys[g0 + t + kSortTile / 2] = s[t + kSortTile / 2];
}

// one block per active record: unique + first s of its sorted segment
__global__ void __launch_bounds__(256)
k_mash_compact(const unsigned long long* __restrict__ keys, size_t nkeys, const uint32_t* __restrict__ slot_rec,
               uint64_t sketch_size, uint32_t stride, uint32_t* __restrict__ sketches, uint32_t* __restrict__ lens,
               uint32_t* __restrict__ ndistinct) {
    __shared__ unsigned s_warp[8];
    __shared__ unsigned s_base;
    const unsigned slot = blockIdx.x;
    const uint32_t rec = slot_rec[slot];
    // segment [lo, hi) of keys whose high word == slot
    auto lower = [&](unsigned long long key) {
        size_t a = 0, b = nkeys;
        while (a < b) {
            size_t m = (a + b) >> 1;
            if (keys[m] < key) a = m + 1; else b = m;
        }
        return a;
    };
    const size_t lo = lower((unsigned long long)slot << 32);
    const size_t hi = lower(((unsigned long long)slot + 1) << 32);
    uint32_t* out = sketches + (size_t)rec * stride;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const uint64_t want = min(sketch_size, (uint64_t)stride);
    for (size_t c = lo; c < hi; c += blockDim.x) {
        const size_t i = c + threadIdx.x;
        unsigned flag = 0;
        unsigned long long key = 0;
        if (i < hi) {
            key = keys[i];
            flag = (i == lo) || (keys[i - 1] != key);
        }
        // block exclusive scan of flags
        unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        unsigned bal = __ballot_sync(0xffffffffu, flag);
        unsigned pre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        unsigned woff = 0, tot = 0;
        for (unsigned w = 0; w < 8; ++w) {
            if (w < wid) woff += s_warp[w];
            tot += s_warp[w];
        }
        const unsigned base = s_base;
        const unsigned pos = base + woff + pre;
        if (flag && pos < want) out[pos] = (uint32_t)key;
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + tot;
        __syncthreads();
        if (s_base >= want && want > 0) break;  // enough; ndistinct is a lower bound >= want
    }
    if (threadIdx.x == 0) {
        unsigned nd = s_base;
        ndistinct[slot] = nd;
        lens[rec] = (uint32_t)min((uint64_t)nd, want);
    }
}

// mash_distance of one pair of ascending sketches (distance.py:230-291): merge until `union == s` or one
// side is exhausted, add the tails, cap at s; integer counts are exact, one f64 log at the end
__device__ __forceinline__ double mash_pair(const uint32_t* __restrict__ A, const uint32_t* __restrict__ B, uint32_t la,
                                            uint32_t lb, int k, uint64_t sketch_size, uint64_t& inter_o, uint64_t& uni_o,
                                            int* __restrict__ err) {
    uint64_t inter = 0, uni = 0;
    uint32_t x = 0, y = 0;
    if (la && lb) {
        uint32_t l = A[0], r = B[0];
        while (uni < sketch_size && x < la && y < lb) {
            if (l < r) {
                ++x;
                if (x < la) l = A[x];
            } else if (r < l) {
                ++y;
                if (y < lb) r = B[y];
            } else {
                ++x;
                ++y;
                ++inter;
                if (x < la) l = A[x];
                if (y < lb) r = B[y];
            }
            ++uni;
        }
    }
    if (uni < sketch_size) {
        if (x < la) uni += la - x;
        if (y < lb) uni += lb - y;
        uni = min(uni, sketch_size);
    }
    double d;
    if (uni == 0) {
        *err = 1;  // ZeroDivisionError in the reference
        d = 0.0;
    } else if (inter == uni) {
        d = 0.0;
    } else if (inter == 0) {
        d = 1.0;
    } else {
        double jac = __ddiv_rn((double)inter, (double)uni);
        d = __ddiv_rn(-log(__ddiv_rn(__dmul_rn(2.0, jac), __dadd_rn(1.0, jac))), (double)k);
        if (d > 1.0) d = 1.0;
    }
    inter_o = inter;
    uni_o = uni;
    return d;
}

// all pairs touching rows [row_begin,row_end)
__global__ void k_mash_pairs(const uint32_t* __restrict__ sk, const uint32_t* __restrict__ lens, uint32_t stride,
                             uint32_t n, int k, uint64_t sketch_size, uint32_t row_begin, uint32_t row_end,
                             double* __restrict__ dist, uint32_t* __restrict__ inter_out, uint32_t* __restrict__ uni_out,
                             int* __restrict__ err) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nrows = row_end - row_begin;
    if (idx >= nrows * n) return;
    const uint32_t i = row_begin + (uint32_t)(idx / n), j = (uint32_t)(idx % n);
    const size_t o_ij = (size_t)(i - row_begin) * n + j;
    if (i == j) {
        dist[o_ij] = 0.0;
        if (inter_out) inter_out[o_ij] = 0;
        if (uni_out) uni_out[o_ij] = 0;
        return;
    }
    const bool j_in = (j >= row_begin && j < row_end);
    if (j_in && j > i) return;  // written by the (j,i) thread as the mirror
    uint64_t inter, uni;
    const double d = mash_pair(sk + (size_t)i * stride, sk + (size_t)j * stride, lens[i], lens[j], k, sketch_size, inter,
                               uni, err);
    dist[o_ij] = d;
    if (inter_out) inter_out[o_ij] = (uint32_t)inter;
    if (uni_out) uni_out[o_ij] = (uint32_t)uni;
    if (j_in) {
        const size_t o_ji = (size_t)(j - row_begin) * n + i;
        dist[o_ji] = d;
        if (inter_out) inter_out[o_ji] = (uint32_t)inter;
        if (uni_out) uni_out[o_ji] = (uint32_t)uni;
    }
}

// The pairs of the lower triangle dealt over the GPUs: this GPU takes pairs p = first + step * idx of the
// linear enumeration p = i (i - 1) / 2 + j (j < i) and stores each distance (and its mirror) into the matrix
// of EVERY GPU through the peer windows - the "all-gather of the matrix" happens inside the kernel.
struct MashOuts {
    double* p[kCommMaxWorld];
    int* err[kCommMaxWorld];  // "a pair of empty sketches was seen" flag of every GPU (all ranks raise alike)
    int count;
};
__global__ void k_mash_pairs_tri(const uint32_t* __restrict__ sk, const uint32_t* __restrict__ lens, uint32_t stride,
                                 uint32_t n, int k, uint64_t sketch_size, uint64_t first, uint64_t step, uint64_t npairs,
                                 const MashOuts outs) {
    const uint64_t p = first + step * ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (p >= npairs) return;
    uint32_t i = (uint32_t)((sqrt(8.0 * (double)p + 1.0) + 1.0) * 0.5);
    while ((uint64_t)i * (i - 1) / 2 > p) --i;
    while ((uint64_t)(i + 1) * i / 2 <= p) ++i;
    const uint32_t j = (uint32_t)(p - (uint64_t)i * (i - 1) / 2);
    uint64_t inter, uni;
    int bad = 0;
    const double d = mash_pair(sk + (size_t)i * stride, sk + (size_t)j * stride, lens[i], lens[j], k, sketch_size, inter,
                               uni, &bad);
    for (int r = 0; r < outs.count; ++r) {
        outs.p[r][(size_t)i * n + j] = d;
        outs.p[r][(size_t)j * n + i] = d;
        if (bad) *outs.err[r] = 1;
    }
}

// the global sort path needs sentinels (all ones, they sort last) in every key slot that holds no candidate:
// the tail of each record's region and the padding behind the last region
__global__ void k_mash_pad_keys(unsigned long long* __restrict__ keys, size_t npad, const MashActive* __restrict__ active,
                                const uint32_t* __restrict__ cnt, uint32_t na) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += (size_t)gridDim.x * blockDim.x) {
        // region containing i: the last record whose base <= i
        uint32_t lo = 0, hi = na;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (active[mid].base <= i) lo = mid; else hi = mid;
        }
        const size_t off = i - active[lo].base;
        if (off >= min(cnt[lo], active[lo].cap)) keys[i] = ~0ull;
    }
}

static int bitonic_sort(dvs_ctx* ctx, unsigned long long* keys, size_t n) {
    cudaStream_t st = ctx->stream;
    k_bitonic_tiles<<<(unsigned)(n / kSortTile), kSortTile / 2, 0, st>>>(keys);
    DVS_LAUNCHED(ctx);
    for (size_t k = (size_t)kSortTile * 2; k <= n; k <<= 1) {
        for (size_t j = k >> 1; j >= kSortTile; j >>= 1) {
            k_bitonic_global<<<(unsigned)((n / 2 + 255) / 256), 256, 0, st>>>(keys, n, k, j);
            DVS_LAUNCHED(ctx);
        }
        k_bitonic_merge_tiles<<<(unsigned)(n / kSortTile), kSortTile / 2, 0, st>>>(keys, k);
        DVS_LAUNCHED(ctx);
    }
    return DVS_OK;
}

}  // namespace dvs

using namespace dvs;

extern "C" {

int dvs_mash_sketch(dvs_ctx* ctx, const dvs_seqset* s, int k, uint64_t sketch_size, int num_states, int canonical,
                    dvs_sketches** out) {
    if (!ctx || !s || !out) {
        set_error("dvs_mash_sketch: NULL argument");
        return DVS_ERR_ARG;
    }
    if (k < 1 || num_states < 1 || num_states > 255) {
        set_error("dvs_mash_sketch: unsupported k=%d / num_states=%d", k, num_states);
        return DVS_ERR_ARG;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    const uint32_t nrec = s->nrec;
    std::vector<uint64_t> nk(nrec);
    uint64_t max_nk = 0;
    for (uint32_t r = 0; r < nrec; ++r) {
        uint64_t len = s->h_offsets[r + 1] - s->h_offsets[r];
        nk[r] = len >= (uint64_t)k ? len - k + 1 : 0;
        max_nk = std::max(max_nk, nk[r]);
    }
    auto* sk = new dvs_sketches();
    sk->device = ctx->device;
    sk->nrec = nrec;
    sk->stride = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(sketch_size, max_nk));
    auto fail = [&](int rc) {
        dvs_sketches_free(sk);
        return rc;
    };
    if (sk->data.alloc((size_t)nrec * sk->stride) != DVS_OK || sk->lens.alloc(nrec) != DVS_OK) return fail(DVS_ERR_CUDA);
#define TRY_S(expr)                                                    \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) {                                       \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); \
            return fail(DVS_ERR_CUDA);                                 \
        }                                                              \
    } while (0)
    TRY_S(cudaMemsetAsync(sk->lens.p, 0, nrec * sizeof(uint32_t), st));
    PhaseTimer pt(ctx, DVS_PHASE_SKETCH);

    const bool fast = (num_states == 4 && k <= 16);
    // per-record threshold / capacity; only records with k-mers take part
    std::vector<uint32_t> thresh(nrec), cap(nrec);
    std::vector<uint32_t> act;
    for (uint32_t r = 0; r < nrec; ++r) {
        if (nk[r] == 0 || sketch_size == 0) continue;
        const uint64_t target = 2 * std::min<uint64_t>(sketch_size, nk[r]) + 512;
        if (nk[r] <= target) {
            thresh[r] = 0xFFFFFFFFu;
            cap[r] = (uint32_t)std::min<uint64_t>(nk[r], 0xFFFFFFFFu);
        } else {
            // expected survivors = nk * (T+1) / 2^32 ~= target
            unsigned __int128 t = ((unsigned __int128)target << 32) / nk[r];
            thresh[r] = (uint32_t)std::min<unsigned __int128>(t, 0xFFFFFFFFu);
            cap[r] = (uint32_t)std::min<uint64_t>(2 * target + 1024, nk[r]);
        }
        act.push_back(r);
    }

    const uint64_t chunk = 256 << 10;
    while (!act.empty()) {
        const uint32_t na = (uint32_t)act.size();
        std::vector<MashActive> h_act(na);
        std::vector<MashWork> work;
        uint64_t total_cap = 0;
        for (uint32_t a = 0; a < na; ++a) {
            const uint32_t r = act[a];
            h_act[a] = {thresh[r], cap[r], total_cap};
            total_cap += cap[r];
            uint64_t b = s->h_offsets[r], e = s->h_offsets[r + 1];
            uint64_t a0 = b & ~15ULL, a1 = (e + 15) & ~15ULL;
            for (uint64_t p = a0; p < a1; p += chunk) work.push_back({p, std::min(p + chunk, a1), r, a});
        }
        size_t npad = kSortTile;
        while (npad < total_cap) npad <<= 1;
        DevBuf<unsigned long long> d_keys;
        DevBuf<MashActive> d_act;
        DevBuf<MashWork> d_work;
        DevBuf<uint32_t> d_cnt, d_next, d_slot_rec, d_nd;
        if (d_keys.alloc(npad) != DVS_OK || d_act.alloc(na) != DVS_OK || d_work.alloc(work.size()) != DVS_OK ||
            d_cnt.alloc(na) != DVS_OK || d_next.alloc(1) != DVS_OK || d_slot_rec.alloc(na) != DVS_OK ||
            d_nd.alloc(na) != DVS_OK)
            return fail(DVS_ERR_CUDA);
        TRY_S(cudaMemsetAsync(d_cnt.p, 0, na * sizeof(uint32_t), st));
        TRY_S(cudaMemsetAsync(d_next.p, 0, sizeof(uint32_t), st));
        TRY_S(cudaMemcpyAsync(d_act.p, h_act.data(), na * sizeof(MashActive), cudaMemcpyHostToDevice, st));
        TRY_S(cudaMemcpyAsync(d_work.p, work.data(), work.size() * sizeof(MashWork), cudaMemcpyHostToDevice, st));
        TRY_S(cudaMemcpyAsync(d_slot_rec.p, act.data(), na * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        const unsigned grid = (unsigned)std::min<size_t>(work.size(), (size_t)ctx->sm_count * 8);
        // DVS_MASH_FILTER16=0: the general k <= 16 kernel also for k = 16 (A/B measurements)
        const char* f16_env = getenv("DVS_MASH_FILTER16");
        if (fast && k == 16 && !(f16_env && f16_env[0] == '0')) {
            if (canonical)
                k_mash_filter16<true><<<grid, kMashThreads, 0, st>>>(s->data(), s->offsets.p, d_work.p,
                                                                     (uint32_t)work.size(), d_next.p, d_act.p, d_cnt.p,
                                                                     d_keys.p, 8192u);
            else
                k_mash_filter16<false><<<grid, kMashThreads, 0, st>>>(s->data(), s->offsets.p, d_work.p,
                                                                      (uint32_t)work.size(), d_next.p, d_act.p, d_cnt.p,
                                                                      d_keys.p, 8192u);
        } else if (fast)
            k_mash_filter<<<grid, kMashThreads, 0, st>>>(s->data(), s->offsets.p, d_work.p, (uint32_t)work.size(),
                                                         d_next.p, k, canonical, d_act.p, d_cnt.p, d_keys.p);
        else
            k_mash_filter_generic<<<grid, kMashThreads, 0, st>>>(s->data(), s->offsets.p, d_work.p,
                                                                 (uint32_t)work.size(), d_next.p, k,
                                                                 (uint32_t)num_states, canonical, d_act.p, d_cnt.p,
                                                                 d_keys.p);
        ctx->launches++;
        TRY_S(cudaGetLastError());
        std::vector<uint32_t> h_cnt(na), h_nd(na);
        TRY_S(cudaMemcpyAsync(h_cnt.data(), d_cnt.p, na * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        TRY_S(cudaStreamSynchronize(st));
        bool overflow = false;
        for (uint32_t a = 0; a < na; ++a)
            if (h_cnt[a] > cap[act[a]]) {
                overflow = true;
                cap[act[a]] = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(2ull * h_cnt[a], 2ull * cap[act[a]]), nk[act[a]]);
            }
        if (overflow) continue;  // rare: redo the active set with bigger buffers
        uint32_t max_cnt = 0;
        for (uint32_t a = 0; a < na; ++a) max_cnt = std::max(max_cnt, h_cnt[a]);
        const char* gs_env = getenv("DVS_MASH_GLOBAL_SORT");
        if (max_cnt <= kSortRecMax && !(gs_env && gs_env[0] == '1')) {
            // every record's candidates fit shared memory: one CTA sorts, dedups and truncates a record
            unsigned P = 2;
            while (P < max_cnt) P <<= 1;
            const size_t smem = (size_t)P * sizeof(uint32_t);
            TRY_S(cudaFuncSetAttribute(k_mash_sort_record, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_mash_sort_record<<<na, kSortRecThreads, smem, st>>>(d_keys.p, d_act.p, d_cnt.p, d_slot_rec.p, sketch_size,
                                                                  sk->stride, sk->data.p, sk->lens.p, d_nd.p);
            ctx->launches++;
            TRY_S(cudaGetLastError());
        } else {
            // global bitonic network over the whole key buffer: unused slots must hold sentinels that sort last
            k_mash_pad_keys<<<(unsigned)std::min<size_t>((npad + 255) / 256, 65535), 256, 0, st>>>(
                d_keys.p, npad, d_act.p, d_cnt.p, na);
            ctx->launches++;
            TRY_S(cudaGetLastError());
            if (bitonic_sort(ctx, d_keys.p, npad) != DVS_OK) return fail(DVS_ERR_CUDA);
            k_mash_compact<<<na, 256, 0, st>>>(d_keys.p, npad, d_slot_rec.p, sketch_size, sk->stride, sk->data.p,
                                               sk->lens.p, d_nd.p);
            ctx->launches++;
            TRY_S(cudaGetLastError());
        }
        TRY_S(cudaMemcpyAsync(h_nd.data(), d_nd.p, na * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        TRY_S(cudaStreamSynchronize(st));
        std::vector<uint32_t> again;
        for (uint32_t a = 0; a < na; ++a) {
            const uint32_t r = act[a];
            if ((uint64_t)h_nd[a] < std::min<uint64_t>(sketch_size, sk->stride) && thresh[r] != 0xFFFFFFFFu) {
                // not enough distinct hashes under the threshold: widen it (x8) and redo this record
                uint64_t t = (uint64_t)thresh[r] * 8 + 7;
                thresh[r] = (uint32_t)std::min<uint64_t>(t, 0xFFFFFFFFu);
                cap[r] = (uint32_t)std::min<uint64_t>((uint64_t)cap[r] * 8, nk[r]);
                if (thresh[r] == 0xFFFFFFFFu) cap[r] = (uint32_t)std::min<uint64_t>(nk[r], 0xFFFFFFFFu);
                again.push_back(r);
            }
        }
        act.swap(again);
    }
#undef TRY_S
    *out = sk;
    return DVS_OK;
}

int dvs_sketches_from_host(dvs_ctx* ctx, const uint32_t* sketches, uint32_t stride, const uint32_t* lens,
                           uint32_t nrec, dvs_sketches** out) {
    if (!ctx || !out || stride == 0 || (nrec && (!sketches || !lens))) {
        set_error("dvs_sketches_from_host: bad argument");
        return DVS_ERR_ARG;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    auto* sk = new dvs_sketches();
    sk->device = ctx->device;
    sk->nrec = nrec;
    sk->stride = stride;
    int rc = sk->data.alloc((size_t)nrec * stride);
    if (rc == DVS_OK) rc = sk->lens.alloc(nrec);
    if (rc == DVS_OK && nrec) {
        cudaError_t e = cudaMemcpyAsync(sk->data.p, sketches, (size_t)nrec * stride * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(sk->lens.p, lens, nrec * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_error("sketch upload failed: %s", cudaGetErrorString(e));
            rc = DVS_ERR_CUDA;
        }
    }
    if (rc != DVS_OK) {
        dvs_sketches_free(sk);
        return rc;
    }
    *out = sk;
    return DVS_OK;
}

uint32_t dvs_sketches_nrec(const dvs_sketches* sk) { return sk->nrec; }
uint32_t dvs_sketches_stride(const dvs_sketches* sk) { return sk->stride; }

int dvs_sketches_download(dvs_ctx* ctx, const dvs_sketches* sk, uint32_t* sketches, uint32_t* lens) {
    DVS_CUDA_TRY(dvs::enter(ctx));
    if (sk->nrec == 0) return DVS_OK;
    if (sketches)
        DVS_CUDA_TRY(cudaMemcpyAsync(sketches, sk->data.p, (size_t)sk->nrec * sk->stride * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (lens) DVS_CUDA_TRY(cudaMemcpyAsync(lens, sk->lens.p, sk->nrec * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DVS_OK;
}

void dvs_sketches_free(dvs_sketches* sk) {
    if (!sk) return;
    cudaSetDevice(sk->device);
    delete sk;
}

int dvs_mash_distances(dvs_ctx* ctx, const dvs_sketches* sk, int k, uint64_t sketch_size, uint32_t row_begin,
                       uint32_t row_end, double* dist, uint32_t* inter, uint32_t* uni) {
    if (!ctx || !sk || !dist || row_begin > row_end || row_end > sk->nrec) {
        set_error("dvs_mash_distances: bad argument");
        return DVS_ERR_ARG;
    }
    const size_t nrows = row_end - row_begin, n = sk->nrec;
    if (nrows == 0 || n == 0) return DVS_OK;
    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    DevBuf<double> d_dist;
    DevBuf<uint32_t> d_inter, d_uni;
    DevBuf<int> d_err;
    DVS_TRY(d_dist.alloc(nrows * n));
    if (inter) DVS_TRY(d_inter.alloc(nrows * n));
    if (uni) DVS_TRY(d_uni.alloc(nrows * n));
    DVS_TRY(d_err.alloc(1));
    DVS_CUDA_TRY(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
    const size_t total = nrows * n;
    PhaseTimer pt(ctx, DVS_PHASE_MASH_PAIRS);
    k_mash_pairs<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(sk->data.p, sk->lens.p, sk->stride, (uint32_t)n, k,
                                                                   sketch_size, row_begin, row_end, d_dist.p,
                                                                   inter ? d_inter.p : nullptr, uni ? d_uni.p : nullptr,
                                                                   d_err.p);
    pt.stop();
    DVS_LAUNCHED(ctx);
    int h_err = 0;
    DVS_CUDA_TRY(cudaMemcpyAsync(dist, d_dist.p, total * sizeof(double), cudaMemcpyDefault, st));
    if (inter) DVS_CUDA_TRY(cudaMemcpyAsync(inter, d_inter.p, total * 4, cudaMemcpyDefault, st));
    if (uni) DVS_CUDA_TRY(cudaMemcpyAsync(uni, d_uni.p, total * 4, cudaMemcpyDefault, st));
    DVS_CUDA_TRY(cudaMemcpyAsync(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    DVS_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_err) {
        set_error("division by zero");  // both sketches empty: ZeroDivisionError in distance.py:283
        return DVS_ERR_VALUE;
    }
    return DVS_OK;
}

int dvs_sketches_allgather(dvs_ctx* ctx, dvs_comm* c, const dvs_sketches* sk, const uint32_t* nrec_per_rank,
                           uint32_t stride_all, dvs_sketches** out) {
    if (!ctx || !c || !c->connected || !sk || !nrec_per_rank || !out || stride_all < sk->stride ||
        nrec_per_rank[c->rank] != sk->nrec) {
        set_error("dvs_sketches_allgather: bad argument (stride_all must be the largest stride of all ranks)");
        return DVS_ERR_ARG;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    uint64_t total = 0;
    std::vector<uint64_t> bytes(c->world), lbytes(c->world);
    for (int r = 0; r < c->world; ++r) {
        total += nrec_per_rank[r];
        bytes[r] = (uint64_t)nrec_per_rank[r] * stride_all * 4;
        lbytes[r] = (uint64_t)nrec_per_rank[r] * 4;
    }
    auto* all = new dvs_sketches();
    all->device = ctx->device;
    all->nrec = (uint32_t)total;
    all->stride = stride_all;
    int rc = all->data.alloc((size_t)total * stride_all);
    if (rc == DVS_OK) rc = all->lens.alloc(total);
    DevBuf<uint32_t> wide;  // this rank's sketches at the common stride
    const uint32_t* src = sk->data.p;
    if (rc == DVS_OK && sk->stride != stride_all && sk->nrec) {
        rc = wide.alloc((size_t)sk->nrec * stride_all);
        cudaError_t e = cudaSuccess;
        if (rc == DVS_OK) e = cudaMemsetAsync(wide.p, 0, (size_t)sk->nrec * stride_all * 4, st);
        if (rc == DVS_OK && e == cudaSuccess)
            e = cudaMemcpy2DAsync(wide.p, (size_t)stride_all * 4, sk->data.p, (size_t)sk->stride * 4, (size_t)sk->stride * 4,
                                  sk->nrec, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) {
            set_error("dvs_sketches_allgather: %s", cudaGetErrorString(e));
            rc = DVS_ERR_CUDA;
        }
        src = wide.p;
    }
    if (rc == DVS_OK) rc = dvs_comm_allgatherv(ctx, c, src, bytes.data(), all->data.p);
    if (rc == DVS_OK) rc = dvs_comm_allgatherv(ctx, c, sk->lens.p, lbytes.data(), all->lens.p);
    if (rc != DVS_OK) {
        dvs_sketches_free(all);
        return rc;
    }
    *out = all;
    return DVS_OK;
}

int dvs_mash_distances_sharded(dvs_ctx* ctx, dvs_comm* c, const dvs_sketches* sk_all, int k, uint64_t sketch_size,
                               double* dist) {
    if (!ctx || !c || !c->connected || !sk_all || !dist) {
        set_error("dvs_mash_distances_sharded: bad argument / communicator not connected");
        return DVS_ERR_ARG;
    }
    const size_t n = sk_all->nrec;
    if (n == 0) return DVS_OK;
    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    uint64_t hoff = 0;
    DVS_TRY(comm_heap_alloc(c, n * n * sizeof(double), &hoff));
    int rc = DVS_OK;
    int h_err = 0;
    cudaError_t e = cudaMemsetAsync(c->window + kCommSelUpdOff, 0, sizeof(int), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->window + hoff, 0, n * n * sizeof(double), st);  // diagonal
    if (rc == DVS_OK && e == cudaSuccess) rc = comm_barrier(ctx, c);  // every matrix is zeroed and free of old readers
    if (rc == DVS_OK && e == cudaSuccess) {
        const uint64_t npairs = (uint64_t)n * (n - 1) / 2;
        const uint64_t mine = npairs > (uint64_t)c->rank ? (npairs - c->rank + c->world - 1) / c->world : 0;
        MashOuts outs;
        memset(&outs, 0, sizeof outs);
        for (int r = 0; r < c->world; ++r) {
            outs.p[r] = reinterpret_cast<double*>(c->peer[(c->rank + r) % c->world] + hoff);
            outs.err[r] = reinterpret_cast<int*>(c->peer[(c->rank + r) % c->world] + kCommSelUpdOff);
        }
        outs.count = c->world;
        PhaseTimer pt(ctx, DVS_PHASE_MASH_PAIRS);
        if (mine) {
            k_mash_pairs_tri<<<(unsigned)((mine + 127) / 128), 128, 0, st>>>(sk_all->data.p, sk_all->lens.p, sk_all->stride,
                                                                            (uint32_t)n, k, sketch_size, (uint64_t)c->rank,
                                                                            (uint64_t)c->world, npairs, outs);
            ctx->launches++;
            e = cudaGetLastError();
        }
        pt.stop();
        if (e == cudaSuccess) rc = comm_barrier(ctx, c);
        if (e == cudaSuccess && rc == DVS_OK)
            e = cudaMemcpyAsync(dist, c->window + hoff, n * n * sizeof(double), cudaMemcpyDefault, st);
        if (e == cudaSuccess && rc == DVS_OK)
            e = cudaMemcpyAsync(&h_err, c->window + kCommSelUpdOff, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && rc == DVS_OK) rc = comm_check_error(ctx, c, "dvs_mash_distances_sharded");
    }
    comm_heap_free(c, hoff);
    if (e != cudaSuccess) {
        set_error("dvs_mash_distances_sharded: %s", cudaGetErrorString(e));
        return DVS_ERR_CUDA;
    }
    if (rc == DVS_OK && h_err) {  // a pair of empty sketches (seen by any rank): ZeroDivisionError in distance.py:283
        set_error("division by zero");
        return DVS_ERR_VALUE;
    }
    return rc;
}

int dvs_mash_sketch_host(dvs_ctx* ctx, const uint8_t* seq, uint64_t len, int k, uint64_t sketch_size, int num_states,
                         int canonical, uint32_t* out, uint64_t cap, uint64_t* out_len) {
    uint64_t offsets[2] = {0, len};
    dvs_seqset* s = nullptr;
    DVS_TRY(dvs_seqset_upload(ctx, seq, offsets, 1, &s));
    dvs_sketches* sk = nullptr;
    int rc = dvs_mash_sketch(ctx, s, k, sketch_size, num_states, canonical, &sk);
    if (rc == DVS_OK) {
        uint32_t n = 0;
        cudaError_t e = cudaMemcpyAsync(&n, sk->lens.p, 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess && n > cap) {
            set_error("dvs_mash_sketch_host: output capacity %llu < sketch length %u", (unsigned long long)cap, n);
            rc = DVS_ERR_ARG;
        } else if (e == cudaSuccess && n) {
            e = cudaMemcpyAsync(out, sk->data.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        }
        if (e != cudaSuccess) {
            set_error("sketch download failed: %s", cudaGetErrorString(e));
            rc = DVS_ERR_CUDA;
        }
        if (rc == DVS_OK) *out_len = n;
    }
    dvs_sketches_free(sk);
    dvs_seqset_free(s);
    return rc;
}

}  // extern "C"
