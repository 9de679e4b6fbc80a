// Synthetic genome generator (bench + tests), bit-identical on host and device.
//
// The reference has no generator; its benchmarks use real genomes (WoL / REFSOIL), which cannot
// be downloaded here.  iid-uniform bases would make every JSD a near-tie, so records come in
// `nfam` families: each family has an ancestor drawn from its own order-2 Markov chain (random
// transition table => distinct k-mer spectrum and GC content) and each record is a prefix of
// its family's ancestor with per-base substitutions at a record-specific rate from
// {0.1%, 1%, 5%, 20%}.  Invalid bytes (value 4) appear in runs of 1..100 at ~1e-4 of all bases
// to exercise the invalid-byte logic of count_kmers (src/record.rs:57-67).
//
// Everything is integer arithmetic on counter-based splitmix64 hashes keyed by
// (seed, stream, a, b), so any byte can be produced independently on either side.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace dvs {

constexpr uint32_t kSynSeg = 256;      // ancestor Markov chains restart every kSynSeg bases
constexpr uint32_t kSynBlock = 1024;   // one candidate invalid run per block of bases

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t syn_hash(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    return splitmix64(splitmix64(splitmix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL)) ^ a) ^ (b * 0xA24BAED4963EE407ULL));
}

// per-record parameters (host and device)
struct SynRec {
    uint64_t len;
    uint32_t fam;
    uint32_t sub_thresh;  // substitution if u32 hash < sub_thresh
};

__host__ __device__ inline SynRec syn_record(uint64_t seed, uint32_t r, uint32_t nfam, uint64_t mean_len) {
    SynRec s;
    const uint64_t lo = mean_len - mean_len / 4;
    const uint64_t span = mean_len / 2 + 1;
    s.len = lo + syn_hash(seed, 1, r, 0) % span;
    s.fam = (uint32_t)(syn_hash(seed, 2, r, 0) % nfam);
    const uint32_t rates[4] = {4294967u, 42949673u, 214748365u, 858993459u};  // 0.1%,1%,5%,20% of 2^32
    s.sub_thresh = rates[syn_hash(seed, 3, r, 0) & 3];
    return s;
}

__host__ __device__ inline uint64_t syn_max_len(uint64_t mean_len) { return mean_len - mean_len / 4 + mean_len / 2; }

// transition table of family f: cum[ctx][0..2] cumulative u32 thresholds (4th is implicit 2^32)
__host__ __device__ inline void syn_table(uint64_t seed, uint32_t fam, uint32_t* cum /*16*3*/) {
    for (uint32_t c = 0; c < 16; ++c) {
        uint64_t w[4], tot = 0;
        for (uint32_t j = 0; j < 4; ++j) {
            uint64_t h = syn_hash(seed, 4, fam, c * 4 + j);
            // sum of two uniforms ~ Gamma(2)-shaped weight => rows are Dirichlet(2,2,2,2)-like
            // (SURVEY.md §8d), i.e. conditional base probabilities mostly within 0.08..0.5 as in
            // real microbial genomes
            w[j] = (h & 255) + ((h >> 8) & 255) + 2;
            tot += w[j];
        }
        uint64_t acc = 0;
        for (uint32_t j = 0; j < 3; ++j) {
            acc += w[j];
            cum[c * 3 + j] = (uint32_t)((acc << 32) / tot);
        }
    }
}

__host__ __device__ __forceinline__ uint32_t syn_draw(const uint32_t* cum, uint32_t ctx, uint32_t u) {
    const uint32_t* c = cum + ctx * 3;
    return (u >= c[0]) + (u >= c[1]) + (u >= c[2]);
}

// one ancestor segment [seg*kSynSeg, ...) of family fam into out (n bases)
__host__ __device__ inline void syn_ancestor_segment(uint64_t seed, uint32_t fam, const uint32_t* cum, uint64_t seg,
                                                     uint8_t* out, uint32_t n) {
    uint64_t h0 = syn_hash(seed, 5, fam, seg);
    uint32_t p2 = (uint32_t)(h0 & 3), p1 = (uint32_t)((h0 >> 2) & 3);
    const uint64_t base = seg * kSynSeg;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t u = (uint32_t)(syn_hash(seed, 6, fam, base + i) >> 32);
        uint32_t b = syn_draw(cum, p2 * 4 + p1, u);
        out[i] = (uint8_t)b;
        p2 = p1;
        p1 = b;
    }
}

// member byte at position p given the ancestor base
__host__ __device__ __forceinline__ uint8_t syn_member_byte(uint64_t seed, uint32_t r, uint32_t sub_thresh, uint64_t p,
                                                            uint8_t anc) {
    uint64_t hb = syn_hash(seed, 7, r, p / kSynBlock);
    if ((hb & 0xFFFF) < 131) {  // ~0.002 of blocks carry one invalid run (mean length 50.5)
        uint32_t start = (uint32_t)((hb >> 16) & (kSynBlock - 1));
        uint32_t rl = 1 + (uint32_t)((hb >> 26) % 100);
        uint32_t off = (uint32_t)(p % kSynBlock);
        if (off >= start && off < start + rl) return 4;
    }
    uint64_t h = syn_hash(seed, 8, r, p);
    if ((uint32_t)h < sub_thresh) return (uint8_t)((anc + 1 + ((h >> 32) % 3)) & 3);
    return anc;
}

__global__ void k_syn_ancestors(uint64_t seed, uint32_t nfam, uint64_t anc_len, const uint32_t* __restrict__ tables,
                                uint8_t* __restrict__ anc) {
    const uint64_t nseg = (anc_len + kSynSeg - 1) / kSynSeg;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nseg * nfam) return;
    const uint32_t fam = (uint32_t)(gid / nseg);
    const uint64_t seg = gid % nseg;
    const uint32_t n = (uint32_t)min((uint64_t)kSynSeg, anc_len - seg * kSynSeg);
    uint32_t cum[48];
    for (int i = 0; i < 48; ++i) cum[i] = tables[fam * 48 + i];
    uint8_t buf[kSynSeg];
    syn_ancestor_segment(seed, fam, cum, seg, buf, n);
    uint8_t* dst = anc + (size_t)fam * anc_len + seg * kSynSeg;
    for (uint32_t i = 0; i < n; ++i) dst[i] = buf[i];
}

// one thread per output byte position group of 16 (records located by binary search on offsets)
// records [first, first + nrec) of the set: local record r is record first + r of the generator
__global__ void k_syn_members(uint64_t seed, uint32_t nrec, uint32_t nfam, uint64_t mean_len, uint64_t anc_len,
                              const uint64_t* __restrict__ offsets, const uint8_t* __restrict__ anc,
                              uint8_t* __restrict__ out, uint64_t total, uint32_t first) {
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (g >= total) return;
    // record containing byte g
    uint32_t lo = 0, hi = nrec;  // offsets[lo] <= g < offsets[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= g) lo = mid; else hi = mid;
    }
    uint32_t r = lo;
    SynRec sr = syn_record(seed, first + r, nfam, mean_len);
    uint64_t rend = offsets[r + 1];
    for (int i = 0; i < 16; ++i) {
        uint64_t q = g + i;
        if (q >= total) break;
        while (q >= rend) {  // crossed into the next (non-empty) record
            ++r;
            sr = syn_record(seed, first + r, nfam, mean_len);
            rend = offsets[r + 1];
        }
        const uint64_t p = q - offsets[r];
        out[q] = syn_member_byte(seed, first + r, sr.sub_thresh, p, anc[(size_t)sr.fam * anc_len + p]);
    }
}

}  // namespace dvs

using namespace dvs;

extern "C" int dvs_seqset_alloc_internal(dvs_ctx* ctx, const uint64_t* offsets, uint32_t nrec, dvs_seqset** out);

extern "C" {

int dvs_synth_lengths(uint64_t seed, uint32_t nrec, uint64_t mean_len, uint64_t* lens_out) {
    if (mean_len < 4) {
        set_error("dvs_synth: mean_len must be >= 4");
        return DVS_ERR_ARG;
    }
    for (uint32_t r = 0; r < nrec; ++r) lens_out[r] = syn_record(seed, r, 1, mean_len).len;
    return DVS_OK;
}

int dvs_synth_host(uint64_t seed, uint32_t nrec, uint32_t nfam, uint64_t mean_len, uint32_t first, uint32_t count,
                   uint8_t* seqs_out, uint64_t* offsets_out) {
    if (mean_len < 4 || nfam == 0 || first + (uint64_t)count > nrec) {
        set_error("dvs_synth_host: bad argument");
        return DVS_ERR_ARG;
    }
    // ancestors are generated lazily per family, only as long as the longest requesting record
    std::vector<std::vector<uint8_t>> anc(nfam);
    std::vector<uint64_t> need(nfam, 0);
    std::vector<SynRec> recs(count);
    for (uint32_t i = 0; i < count; ++i) {
        recs[i] = syn_record(seed, first + i, nfam, mean_len);
        need[recs[i].fam] = std::max(need[recs[i].fam], recs[i].len);
    }
    for (uint32_t f = 0; f < nfam; ++f) {
        if (!need[f]) continue;
        uint32_t cum[48];
        syn_table(seed, f, cum);
        const uint64_t nseg = (need[f] + kSynSeg - 1) / kSynSeg;
        anc[f].resize(nseg * kSynSeg);
        for (uint64_t s = 0; s < nseg; ++s) syn_ancestor_segment(seed, f, cum, s, anc[f].data() + s * kSynSeg, kSynSeg);
    }
    uint64_t off = 0;
    offsets_out[0] = 0;
    for (uint32_t i = 0; i < count; ++i) {
        const SynRec& sr = recs[i];
        if (seqs_out)
            for (uint64_t p = 0; p < sr.len; ++p)
                seqs_out[off + p] = syn_member_byte(seed, first + i, sr.sub_thresh, p, anc[sr.fam][p]);
        off += sr.len;
        offsets_out[i + 1] = off;
    }
    return DVS_OK;
}

int dvs_seqset_synth(dvs_ctx* ctx, uint64_t seed, uint32_t nrec, uint32_t nfam, uint64_t mean_len, dvs_seqset** out) {
    return dvs_seqset_synth_range(ctx, seed, 0, nrec, nfam, mean_len, out);
}

int dvs_seqset_synth_range(dvs_ctx* ctx, uint64_t seed, uint32_t first, uint32_t nrec, uint32_t nfam, uint64_t mean_len,
                           dvs_seqset** out) {
    if (!ctx || !out || mean_len < 4 || nfam == 0 || (uint64_t)first + nrec > 0xFFFFFFFFull) {
        set_error("dvs_seqset_synth: bad argument");
        return DVS_ERR_ARG;
    }
    std::vector<uint64_t> offsets(nrec + 1, 0);
    for (uint32_t r = 0; r < nrec; ++r) offsets[r + 1] = offsets[r] + syn_record(seed, first + r, nfam, mean_len).len;
    dvs_seqset* s = nullptr;
    DVS_TRY(dvs_seqset_alloc_internal(ctx, offsets.data(), nrec, &s));
    auto fail = [&](const char* what, cudaError_t e) {
        set_error("%s failed: %s", what, cudaGetErrorString(e));
        dvs_seqset_free(s);
        return DVS_ERR_CUDA;
    };
    const uint64_t anc_len = syn_max_len(mean_len);
    std::vector<uint32_t> tables((size_t)nfam * 48);
    for (uint32_t f = 0; f < nfam; ++f) syn_table(seed, f, tables.data() + (size_t)f * 48);
    DevBuf<uint32_t> d_tables;
    DevBuf<uint8_t> d_anc;
    if (d_tables.alloc(tables.size()) != DVS_OK || d_anc.alloc((size_t)nfam * anc_len) != DVS_OK) {
        dvs_seqset_free(s);
        return DVS_ERR_CUDA;
    }
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaMemcpyAsync(d_tables.p, tables.data(), tables.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail("table upload", e);
    const uint64_t nseg = (anc_len + kSynSeg - 1) / kSynSeg;
    const uint64_t nthreads = nseg * nfam;
    k_syn_ancestors<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(seed, nfam, anc_len, d_tables.p, d_anc.p);
    ctx->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_syn_ancestors", e);
    if (s->total) {
        const uint64_t groups = (s->total + 15) / 16;
        k_syn_members<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(seed, nrec, nfam, mean_len, anc_len, s->offsets.p,
                                                                        d_anc.p, s->data(), s->total, first);
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_syn_members", e);
    }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail("synth", e);
    *out = s;
    return DVS_OK;
}

}  // extern "C"
