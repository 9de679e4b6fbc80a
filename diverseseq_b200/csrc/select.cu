// SummedRecords on the device: JSD / delta-JSD state and the order-preserving nmost / max
// selection loops.  Replaces /root/reference/src/records.rs:10-524.
//
// Exactness: every entropy the reference evaluates is evaluated here in the reference's own
// sequential order (entropy.cuh), and every elementwise update of summed_kfreqs /
// summed_entropies uses the same operations in the same order (explicit __d*_rn, no FMA), so
// the state is bitwise the reference's and every `>` / `<` decision is the reference's.
//
// Parallel shape: the reference streams candidates one at a time (records.rs:331-340).  Here a
// WINDOW of candidates following the cursor is scored concurrently (one CTA per candidate)
// against the current state; the FIRST position whose predicate `delta_jsd > total_jsd + EPS`
// holds is the one the reference would accept next (everything before it was rejected against
// the same state; everything after it is stale and is re-scored after the update).  A min-index
// reduction (atomicMin) finds it; the state update then runs as three small kernels.
#include <limits.h>
#include <stddef.h>
#include <stdlib.h>

#include <algorithm>
#include <limits>

#include <cooperative_groups.h>

#include "common.cuh"
#include "entropy.cuh"

namespace cg = cooperative_groups;

namespace dvs {

constexpr unsigned kNone = 0xFFFFFFFFu;

struct SelScal {
    double E;          // summed_entropies
    double total_jsd;
    double mean, stdv, cov;  // mean/std/cov of member delta_jsd
    double panic_total;      // offending total for a state-level entropy panic
    unsigned n;              // size
    unsigned lowest;         // lowest_index
    unsigned first_true;     // scan result: first position whose predicate holds
    unsigned first_panic;    // scan result: first position whose evaluation panics
    unsigned panic;          // 0 none; 1 = entropy sum check failed while updating the state
    unsigned ticket;         // last-block-done counter
    // ---- bounded-error fast path (see "fast path" below) ----
    double total_bound;      // |total_jsd - exact total_jsd| <= total_bound (0 when exact)
    unsigned first_unsure;   // scan result: first position the fast scan could not decide
    unsigned state_unsure;   // fast update could not certify lowest_index / the sum checks
    unsigned exact;          // 1 when total_jsd / mdelta / lowest come from the exact kernel
    unsigned pad0_;
    double std_bound;        // fast path: |stdv - reference std_delta_jsd| <= std_bound  (0 when exact)
    double cov_bound;        // fast path: same for cov (infinite when the mean is too close to 0)
    // ---- device-driven rounds (k_sel_scan_dev / k_sel_round_dev) ----
    unsigned cursor;         // next position to examine
    unsigned window;         // candidates scored per round
    unsigned window_max;
    unsigned num;            // number of positions
    unsigned halt;           // 1: the host must resolve an undecided candidate / state exactly
    unsigned accepts;        // candidates that replaced the lowest record so far
    unsigned which;          // which copy of the double-buffered S / member list is current
};

struct SelState {
    // S and the member list are double buffered: the fused fast replace+update kernel reads one copy
    // and writes the other, so it needs no grid-wide barrier between "update S" and "use S"
    DevBuf<double> Sbuf[2];          // summed_kfreqs [dim]
    DevBuf<unsigned> membuf[2];      // row indices in Vec order [cap]
    int which = 0;
    DevBuf<double> mdelta;     // delta_jsd per member [cap]
    DevBuf<double> mbound;     // fast path: error bound of each member's value [cap]
    DevBuf<SelScal> sc;
    unsigned cap = 0;
    double* S() { return Sbuf[which].p; }
    unsigned* members() { return membuf[which].p; }
    double* S_other() { return Sbuf[which ^ 1].p; }
    unsigned* members_other() { return membuf[which ^ 1].p; }
    void flip() { which ^= 1; }
    int alloc(uint64_t dim, unsigned capacity) {
        cap = capacity;
        for (int b = 0; b < 2; ++b) {
            DVS_TRY(Sbuf[b].alloc(dim));
            DVS_TRY(membuf[b].alloc(capacity));
        }
        DVS_TRY(mdelta.alloc(capacity));
        DVS_TRY(mbound.alloc(capacity));
        DVS_TRY(sc.alloc(1));
        return DVS_OK;
    }
};

// S[i] = ((0 + f_m0[i]) + f_m1[i]) + ... (+ f_extra[i]);  E likewise  (records.rs:36-42, :129-133)
__global__ void k_sel_sum(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim,
                          const unsigned* __restrict__ members, unsigned n, int extra_row, double* __restrict__ S,
                          unsigned* __restrict__ members_out, SelScal* sc) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) {
        double s = 0.0;
        for (unsigned j = 0; j < n; ++j) s = __dadd_rn(s, F[(size_t)members[j] * dim + i]);
        if (extra_row >= 0) s = __dadd_rn(s, F[(size_t)extra_row * dim + i]);
        S[i] = s;
    }
    if (i == 0) {
        double e = 0.0;
        for (unsigned j = 0; j < n; ++j) e = __dadd_rn(e, H[members[j]]);
        if (extra_row >= 0) e = __dadd_rn(e, H[extra_row]);
        unsigned nn = n;
        if (members_out != members)
            for (unsigned j = 0; j < n; ++j) members_out[j] = members[j];
        if (extra_row >= 0) members_out[nn++] = (unsigned)extra_row;
        sc->E = e;
        sc->n = nn;
        sc->panic = 0;
        sc->panic_total = 0.0;
        sc->ticket = 0;
        sc->first_true = kNone;
        sc->first_panic = kNone;
        sc->first_unsure = kNone;
        sc->state_unsure = 0;
        sc->total_bound = 0.0;
        sc->exact = 0;
    }
}

// replace_lowest (records.rs:94-135): drop_lowest + push on summed_kfreqs, elementwise; the LAST
// CTA to finish then updates the scalars and the member Vec: E -= H_low; E += H_c;
// Vec::remove(lowest); push(c).
__global__ void k_sel_replace_vec(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim,
                                  unsigned* __restrict__ members, uint8_t* __restrict__ is_member, SelScal* sc,
                                  unsigned cand_row, double* __restrict__ S) {
    __shared__ unsigned s_last;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned low = sc->lowest;
    const unsigned low_row = members[low];
    if (i < dim) {
        double s = __dsub_rn(S[i], F[(size_t)low_row * dim + i]);
        if (s <= kEps) s = 0.0;
        S[i] = __dadd_rn(s, F[(size_t)cand_row * dim + i]);
    }
    __syncthreads();  // every read of members[]/lowest by this CTA is done
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&sc->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {  // block-uniform
        const unsigned n = sc->n;
        // Vec::remove(low) + push(cand): every thread moves a strided share, reads before writes
        unsigned keep[8];
        const unsigned per = (n + blockDim.x - 1) / blockDim.x;
        if (per <= 8) {
            for (unsigned q = 0; q < per; ++q) {
                const unsigned j = threadIdx.x + q * blockDim.x;
                keep[q] = (j >= low && j + 1 < n) ? members[j + 1] : 0u;
            }
            __syncthreads();
            for (unsigned q = 0; q < per; ++q) {
                const unsigned j = threadIdx.x + q * blockDim.x;
                if (j >= low && j + 1 < n) members[j] = keep[q];
            }
        } else if (threadIdx.x == 0) {
            for (unsigned j = low; j + 1 < n; ++j) members[j] = members[j + 1];
        }
        if (threadIdx.x == 0) {
            double e = __dsub_rn(sc->E, H[low_row]);
            sc->E = __dadd_rn(e, H[cand_row]);
            members[n - 1] = cand_row;
            if (is_member) {
                is_member[low_row] = 0;
                is_member[cand_row] = 1;
            }
            sc->ticket = 0;
        }
    }
}

// total_jsd and get_lowest_record_index in one launch (records.rs:136-146, 220-252): CTA j < n
// evaluates member j's leave-one-out entropy, CTA n evaluates H(S/n); none of these entropies
// depends on another, only the final scalar arithmetic does, so the last CTA to finish forms
// total_jsd = H(S/n) - E/n, delta_j = total_jsd - (H(m_j) - (E - H_j)/(n-1)), the strict-<
// argmin and the mean/std/cov statistics (:153-172) in the reference's order.
__global__ void __launch_bounds__(kEntThreads)
k_sel_update(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim, const double* __restrict__ S,
             const unsigned* __restrict__ members, double* __restrict__ mdelta, SelScal* sc) {
    extern __shared__ __align__(16) double ent_smem[];
    __shared__ unsigned s_last;
    const unsigned j = blockIdx.x, n = sc->n;
    const double nd = (double)n;
    if (j == n) {
        EntropyResult h = block_entropy_exact(dim, [&](uint64_t i) { return __ddiv_rn(S[i], nd); }, ent_smem);
        if (threadIdx.x == 0) {
            sc->total_jsd = __dsub_rn(h.e, __ddiv_rn(sc->E, nd));
            if (entropy_total_bad(h.t, dim) && atomicCAS(&sc->panic, 0u, 1u) == 0u) sc->panic_total = h.t;
        }
    } else {
        const unsigned row = members[j];
        const double div = __dsub_rn(nd, 1.0);
        const double* f = F + (size_t)row * dim;
        EntropyResult h = block_entropy_exact(
            dim,
            [&](uint64_t i) {
                double m = __ddiv_rn(__dsub_rn(S[i], f[i]), div);
                return (m <= kEps) ? 0.0 : m;  // updated_mean_freqs clamp, records.rs:281-284
            },
            ent_smem);
        if (threadIdx.x == 0) {
            const double mean_entropy = __ddiv_rn(__dsub_rn(sc->E, H[row]), div);
            mdelta[j] = __dsub_rn(h.e, mean_entropy);  // jsd without record j; delta formed below
            // the reference panics at the first member whose sum check fails; any failure is fatal
            if (entropy_total_bad(h.t, dim) && atomicCAS(&sc->panic, 0u, 1u) == 0u) sc->panic_total = h.t;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&sc->ticket, 1u) == n) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {  // block-uniform: the last CTA finishes get_lowest_record_index + the statistics
        __threadfence();
        // the sums below are sequential (reference order); stage the n values in shared memory first so
        // the single summing thread does not pay one L2 round trip per member
        constexpr unsigned kStage = (unsigned)(kEntSmemBytes / sizeof(double));
        const bool staged = n <= kStage;
        if (staged) {
            for (unsigned t = threadIdx.x; t < n; t += blockDim.x) ent_smem[t] = __ldcg(&mdelta[t]);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            volatile double* mdv = mdelta;
            const double total_jsd = __ldcg(&sc->total_jsd);
            double mn = 1e6;
            unsigned low = 0;
            double sum = 0.0;
            for (unsigned t = 0; t < n; ++t) {
                const double d = __dsub_rn(total_jsd, staged ? ent_smem[t] : mdv[t]);
                if (staged) ent_smem[t] = d; else mdv[t] = d;
                if (d < mn) {
                    mn = d;
                    low = t;
                }
                sum = __dadd_rn(sum, d);
            }
            const double mean = __ddiv_rn(sum, nd);
            double ss = 0.0;
            for (unsigned t = 0; t < n; ++t) {
                const double d = __dsub_rn(staged ? ent_smem[t] : mdv[t], mean);
                ss = __dadd_rn(ss, __dmul_rn(d, d));
            }
            const double sd = __dsqrt_rn(__ddiv_rn(ss, __dsub_rn(nd, 1.0)));
            sc->lowest = low;
            sc->mean = mean;
            sc->stdv = sd;
            sc->cov = __ddiv_rn(sd, mean);
            sc->ticket = 0;
            sc->first_true = kNone;
            sc->first_panic = kNone;
            sc->first_unsure = kNone;
            sc->total_bound = 0.0;
            sc->std_bound = 0.0;
            sc->cov_bound = 0.0;
            sc->state_unsure = 0;
            sc->exact = 1;
        }
        if (staged) {
            __syncthreads();
            for (unsigned t = threadIdx.x; t < n; t += blockDim.x) mdelta[t] = ent_smem[t];
        }
    }
}

// increases_jsd for a window of candidates (records.rs:70-92); one CTA per position
__global__ void __launch_bounds__(kEntThreads)
k_sel_scan(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim, const double* __restrict__ S,
           const unsigned* __restrict__ members, SelScal* sc, const double* __restrict__ candF,
           const double* __restrict__ candH, const uint8_t* __restrict__ cand_valid,
           const uint8_t* __restrict__ is_member, const unsigned* __restrict__ order, unsigned pos0,
           double* __restrict__ delta_out) {
    extern __shared__ __align__(16) double ent_smem[];
    const unsigned pos = pos0 + blockIdx.x;
    const unsigned row = order ? order[pos] : pos;
    if (cand_valid && !cand_valid[row]) return;     // Err("No valid k-mers") -> skipped
    if (is_member && is_member[row]) return;        // seqid already in the set -> false
    const unsigned n = sc->n;
    const double nd = (double)n;
    const unsigned low_row = members[sc->lowest];
    const double* fl = F + (size_t)low_row * dim;
    const double* fc = candF + (size_t)row * dim;
    EntropyResult h = block_entropy_exact(
        dim, [&](uint64_t i) { return __ddiv_rn(__dadd_rn(__dsub_rn(S[i], fl[i]), fc[i]), nd); }, ent_smem);
    if (threadIdx.x == 0) {
        const double mean_entropy = __ddiv_rn(__dadd_rn(__dsub_rn(sc->E, H[low_row]), candH[row]), nd);
        const double d = __dsub_rn(h.e, mean_entropy);
        if (delta_out) delta_out[blockIdx.x] = d;
        if (entropy_total_bad(h.t, dim)) atomicMin(&sc->first_panic, pos);
        if (d > __dadd_rn(sc->total_jsd, kEps)) atomicMin(&sc->first_true, pos);
    }
}

// ------------------------------------------------------------------------------- fast path ----
// The exact kernels above cost one dependent FP64 add per element (~30 us per 4^6-element
// entropy).  Most decisions are nowhere near a tie, so they are first attempted with a PARALLEL
// evaluation whose distance from the reference's value is rigorously bounded:
//   * the frequencies m_i are formed with the same IEEE operations, so they are identical;
//   * each term -m*log2(m) differs from the reference's by <= 3 ulp (two <1-ulp log2s, one product);
//   * a pairwise (tree) sum of D terms is within (ceil(log2 D)+3) u A of the real sum A' of the
//     computed terms, the reference's sequential sum within (D-1) u A        (u = 2^-53, A = sum|term|);
// so |e_fast - e_ref| <= (D + 64) * 1.2e-16 * A =: bound.  A decision is taken from the fast value
// only when it holds for every value in [fast - bound, fast + bound]; otherwise the position is
// reported as "unsure" and the exact kernel decides.  The reference's sum-to-one check
// |t_ref - 1| <= D*EPS is certified from the tree sum t (|t - T| <= 16u, |t_ref - T| <= (D-1)u):
// it cannot fail when |t - 1| <= 0.45*D*EPS; otherwise: unsure.  Selected sets, their order and all
// reported numbers are therefore still the exact path's (the final state is always re-evaluated
// exactly); only the work is reduced.
constexpr int kFastThreads = 512;

struct FastSum {
    double e, t, a;  // entropy, total, sum |term|
    int bad;         // a negative / NaN frequency was seen (reference yields NaN): cannot be bounded
};

// elements [lo, hi) of a vector (the SM-replicated selection kernel splits one vector over several CTAs)
template <class Elem>
__device__ FastSum block_entropy_span(uint64_t lo, uint64_t dim, Elem elem) {
    __shared__ double s_part[3][kFastThreads / 32];
    __shared__ int s_bad[kFastThreads / 32];
    double e0 = 0.0, e1 = 0.0, t0 = 0.0, t1 = 0.0, a0 = 0.0, a1 = 0.0;
    int bad = 0;
    // 8 elements per thread per pass: all frequencies (global loads + divides) are formed before the
    // first log2 so one memory latency is exposed per pass, not one per element
    constexpr int kBatch = 8;
    for (uint64_t base = lo; base < dim; base += (uint64_t)kBatch * kFastThreads) {
        double x[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            const uint64_t i = base + threadIdx.x + (uint64_t)q * kFastThreads;
            x[q] = i < dim ? elem(i) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < kBatch; q += 2) {
            const double x0 = x[q], x1 = x[q + 1];
            if (!(x0 == 0.0)) {
                bad |= !(x0 > 0.0);
                const double tm = __dmul_rn(-x0, log2(x0));
                e0 += tm; a0 += fabs(tm); t0 += x0;
            }
            if (!(x1 == 0.0)) {
                bad |= !(x1 > 0.0);
                const double tm = __dmul_rn(-x1, log2(x1));
                e1 += tm; a1 += fabs(tm); t1 += x1;
            }
        }
    }
    double e = e0 + e1, t = t0 + t1, a = a0 + a1;
    for (int o = 16; o; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        t += __shfl_xor_sync(0xffffffffu, t, o);
        a += __shfl_xor_sync(0xffffffffu, a, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_part[0][w] = e; s_part[1][w] = t; s_part[2][w] = a; s_bad[w] = bad;
    }
    __syncthreads();
    FastSum r{0.0, 0.0, 0.0, 0};
    for (int q = 0; q < kFastThreads / 32; ++q) {
        r.e += s_part[0][q]; r.t += s_part[1][q]; r.a += s_part[2][q]; r.bad |= s_bad[q];
    }
    __syncthreads();
    return r;
}
template <class Elem>
__device__ __forceinline__ FastSum block_entropy_fast(uint64_t dim, Elem elem) {
    return block_entropy_span(0, dim, elem);
}

// ---- the same sums with instruction-level parallelism (SM-replicated rounds) ----
// block_entropy_span evaluates one element at a time: CUDA's ddiv and log2 are ~50 dependent FP64
// instructions behind their own branches, and with 4 warps per scheduler at ~14 cycles per dependent
// FP64 instruction the FP64 pipe idles most of the time (measured 0.47 us per element per thread).
// Here every step is branch-free straight-line code over 8 elements per thread, so 8 chains interleave:
//   * x / b by div_exact (entropy.cuh): reciprocal + two fused-remainder steps, = __ddiv_rn bit for bit;
//   * log2 by the table-driven path of the glibc restatement (dvs_log2_main): the reference's own bits,
//     so a term -m*log2(m) is now IDENTICAL to the reference's, not merely within 3 ulp.
// Inputs that the glibc algorithm sends down its other paths (m within ~4 % of 1, subnormals) or that
// make the reference's value NaN raise `bad`, which the callers already treat as "undecided".
// m is usable by the table path (and was divided exactly) iff 2^-900 <= m < 0x1.ea4afp-1: one unsigned
// compare on its high word; negative, NaN, inf, zero, subnormal and "near 1 or above" all fall outside.
// (Below 2^-900 the fused remainders of div_exact could underflow; k-mer frequencies are >= ~1e-10.)
__device__ __forceinline__ bool fast_term_ok(double m) {
    return (uint32_t)__double2hiint(m) - 0x07b00000u < 0x3feea4afu - 0x07b00000u;
}

template <bool CLAMP, int BATCH, class Num>
__device__ FastSum block_entropy_ilp(unsigned lo, unsigned hi, Num num, const FastDiv dv,
                                     const double2* __restrict__ ltab) {
    __shared__ double s_part[3][kFastThreads / 32];
    __shared__ int s_bad[kFastThreads / 32];
    double e0 = 0.0, e1 = 0.0, t0 = 0.0, t1 = 0.0, a0 = 0.0, a1 = 0.0;
    bool bad = false;
    for (unsigned base = lo; base < hi; base += BATCH * kFastThreads) {
        double x[BATCH], l[BATCH];
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const unsigned i = base + threadIdx.x + (unsigned)q * kFastThreads;
            x[q] = i < hi ? num(i) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            double m = div_exact(x[q], dv);
            if (CLAMP) m = (m <= kEps) ? 0.0 : m;
            x[q] = m;
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            int ignored = 0;
            l[q] = dvs_log2_main(x[q], ltab, ignored);
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const bool ok = fast_term_ok(x[q]);
            bad |= !ok && !(x[q] == 0.0);
            const double tm = ok ? __dmul_rn(-x[q], l[q]) : 0.0;
            if (q & 1) {
                e1 += tm; a1 += fabs(tm); t1 += x[q];
            } else {
                e0 += tm; a0 += fabs(tm); t0 += x[q];
            }
        }
    }
    double e = e0 + e1, t = t0 + t1, a = a0 + a1;
    int badi = bad ? 1 : 0;
    for (int o = 16; o; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        t += __shfl_xor_sync(0xffffffffu, t, o);
        a += __shfl_xor_sync(0xffffffffu, a, o);
        badi |= __shfl_xor_sync(0xffffffffu, badi, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_part[0][w] = e; s_part[1][w] = t; s_part[2][w] = a; s_bad[w] = badi;
    }
    __syncthreads();
    FastSum r{0.0, 0.0, 0.0, 0};
    for (int q = 0; q < kFastThreads / 32; ++q) {
        r.e += s_part[0][q]; r.t += s_part[1][q]; r.a += s_part[2][q]; r.bad |= s_bad[q];
    }
    __syncthreads();
    return r;
}

// The same evaluation for the kernels that keep their state in global memory (host-driven and
// device-driven rounds, the cooperative kernel, the batched grow attempts — the k >= 7 paths): m_i =
// num(i) / divisor, optionally clamped (<= EPS -> 0).  Stages the log2 table itself; divisors beyond the
// range the exact division was verified for, and vectors longer than 2^30, take the one-element-at-a-time
// form with __ddiv_rn (same values either way).
template <bool CLAMP, class Num>
__device__ FastSum block_entropy_div(uint64_t dim, Num num, double divisor) {
    if (dim > (1ull << 30) || !(divisor >= 1.0 && divisor <= 4096.0)) {  // block-uniform
        return block_entropy_fast(dim, [&](uint64_t i) {
            const double m = __ddiv_rn(num(i), divisor);
            return (CLAMP && m <= kEps) ? 0.0 : m;
        });
    }
    __shared__ double2 s_ltab_div[64];
    dvs_log2_stage_table(s_ltab_div);
    __syncthreads();
    return block_entropy_ilp<CLAMP, 8>(0u, (unsigned)dim, num, make_fast_div(divisor), s_ltab_div);
}

// Each thread first sums dim/kFastThreads elements sequentially, then the partials are tree-summed:
// |sum_fast - real sum| <= (dim/kFastThreads + 12) u A, the reference's sequential sum is within
// (dim - 1) u A, and the per-term differences add 6 u A; 1.2e-16 > u = 2^-53 absorbs second-order terms.
// `depth`: additional sequential additions on top of the tree (partials of a vector split over CTAs)
__device__ __forceinline__ double fast_slack(uint64_t dim, double depth = 0.0) {
    return (double)(dim / kFastThreads) + 12.0 + depth;
}
__device__ __forceinline__ double fast_bound(uint64_t dim, double a, double extra, double depth = 0.0) {
    return ((double)dim + fast_slack(dim, depth) + 16.0) * 1.2e-16 * (a + fabs(extra) + 1.0);
}
// reference check: |t_ref - 1| <= dim*EPS = 2 dim u.  |t_ref - T| <= (dim-1) u, |t - T| <= slack u (T ~ 1),
// so the check cannot fail when |t - 1| <= (dim + 1 - slack) u; for tiny dim this is never certified
// and the (then trivially cheap) exact kernel decides.
__device__ __forceinline__ bool fast_total_ok(uint64_t dim, double t, double depth = 0.0) {
    const double lim = ((double)dim + 1.0 - fast_slack(dim, depth) - 2.0) * 1.1102230246251565e-16;
    return lim > 0.0 && fabs(t - 1.0) <= lim;
}

// Block-cooperative end of a fast update (run by every thread of the LAST CTA to finish):
// delta_j = total - jsd_j, argmin (lowest index wins ties) and the certainty test
// "member `low` is smaller than every other member for all admissible errors".  Returns 1 when the
// argmin could not be certified.  (A single thread walking n global values costs n L2 latencies.)
// GLOBAL: the arrays live in global memory and were written by other CTAs (read through L2); otherwise
// they are this CTA's own shared-memory copies.
template <bool GLOBAL = true>
__device__ __forceinline__ unsigned finalize_fast_block(double* mdelta, const double* mbound, unsigned n, double total,
                                                        double total_bound, unsigned* low_out) {
    auto ld = [](const double* p) { return GLOBAL ? __ldcg(p) : *p; };
    __shared__ double s_mn[kFastThreads / 32], s_mb[kFastThreads / 32];
    __shared__ unsigned s_ix[kFastThreads / 32];
    __shared__ double s_best, s_bestb;
    __shared__ unsigned s_besti;
    double mn = 1e300, mb = 0.0;
    unsigned ix = kNone;
    for (unsigned t = threadIdx.x; t < n; t += blockDim.x) {
        const double d = total - ld(&mdelta[t]);
        mdelta[t] = d;
        if (d < mn || (d == mn && t < ix)) {
            mn = d;
            mb = ld(&mbound[t]);
            ix = t;
        }
    }
    for (int o = 16; o; o >>= 1) {
        const double omn = __shfl_xor_sync(0xffffffffu, mn, o), omb = __shfl_xor_sync(0xffffffffu, mb, o);
        const unsigned oix = __shfl_xor_sync(0xffffffffu, ix, o);
        if (omn < mn || (omn == mn && oix < ix)) {
            mn = omn; mb = omb; ix = oix;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_mn[threadIdx.x >> 5] = mn; s_mb[threadIdx.x >> 5] = mb; s_ix[threadIdx.x >> 5] = ix;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned w = 1; w < blockDim.x / 32; ++w)
            if (s_mn[w] < mn || (s_mn[w] == mn && s_ix[w] < ix)) {
                mn = s_mn[w]; mb = s_mb[w]; ix = s_ix[w];
            }
        s_best = mn; s_bestb = mb; s_besti = (ix == kNone) ? 0u : ix;
    }
    __syncthreads();
    mn = s_best; mb = s_bestb;
    const unsigned low = s_besti;
    int unsure = 0;
    for (unsigned t = threadIdx.x; t < n; t += blockDim.x)
        if (t != low && !(mn + mb + 2.0 * kEps < mdelta[t] - ld(&mbound[t]))) unsure = 1;
    if (!(mn + mb + total_bound < 1e6)) unsure = 1;  // the reference's `min_delta_jsd = 1e6` initial value
    unsure = __syncthreads_or(unsure);
    *low_out = low;
    return (unsigned)unsure;
}

// block-wide sum / max of one double per thread (all threads get the result)
__device__ __forceinline__ double block_sum_fast(double v) {
    __shared__ double s_red[kFastThreads / 32];
    __shared__ double s_out;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (unsigned w = 0; w < blockDim.x / 32; ++w) t += s_red[w];
        s_out = t;
    }
    __syncthreads();
    const double r = s_out;
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_max_fast(double v) {
    __shared__ double s_red[kFastThreads / 32];
    __shared__ double s_out;
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = s_red[0];
        for (unsigned w = 1; w < blockDim.x / 32; ++w) t = fmax(t, s_red[w]);
        s_out = t;
    }
    __syncthreads();
    const double r = s_out;
    __syncthreads();
    return r;
}

// mean / std / cov of the (approximate) member deltas with rigorous distance bounds to the values
// the reference computes (src/records.rs:153-172).  delta_j is within mbound_j (+ the common
// total_bound, which cancels in the deviations) of the reference's; std is 1-Lipschitz in the
// deviations scaled by 1/sqrt(n-1), the sequential sums add <= (n+8) u relative error.
__device__ __forceinline__ void stats_fast_block(const double* mdelta, const double* mbound, unsigned n,
                                                 double total_bound, SelScal* sc) {
    double sd = 0.0, bmx = 0.0, amx = 0.0;
    for (unsigned t = threadIdx.x; t < n; t += blockDim.x) {
        const double d = mdelta[t];
        sd += d;
        bmx = fmax(bmx, __ldcg(&mbound[t]));
        amx = fmax(amx, fabs(d));
    }
    const double nd = (double)n;
    const double mean = block_sum_fast(sd) / nd;
    bmx = block_max_fast(bmx);
    amx = block_max_fast(amx);
    double ss = 0.0;
    for (unsigned t = threadIdx.x; t < n; t += blockDim.x) {
        const double d = mdelta[t] - mean;
        ss += d * d;
    }
    const double sdev = sqrt(block_sum_fast(ss) / (nd - 1.0));
    if (threadIdx.x == 0) {
        const double u = 1.2e-16, scale = amx + fabs(mean);
        const double b_mean = bmx + total_bound + (nd + 8.0) * u * scale;
        const double b_std = 1.5 * (bmx + 4.0 * u * scale) + (nd + 8.0) * u * sdev;
        const double cov = sdev / mean;
        double b_cov = 1e300;
        if (fabs(mean) > 2.0 * b_mean) b_cov = (b_std + fabs(cov) * b_mean) / (fabs(mean) - b_mean) + 8.0 * u * fabs(cov);
        sc->mean = mean;
        sc->stdv = sdev;
        sc->cov = cov;
        sc->std_bound = b_std;
        sc->cov_bound = b_cov;
    }
}

// The mutable selection state (S, member list, is_member, scalars) is read with ld.global.cg inside the
// bodies below: the persistent kernel keeps CTAs alive across rounds, where a line cached in L1 during
// an earlier round would be stale.  Rows of F / H / valid / order never change and use the default path.
struct ScanScal {
    double E, total_jsd, total_bound;
    unsigned n, low_row;
};

// fast increases_jsd of the candidate at position `pos`; first_true / first_unsure by atomicMin
__device__ __forceinline__ void scan_fast_body(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim,
                                               const double* __restrict__ S, SelScal* sc, const ScanScal& q,
                                               const uint8_t* __restrict__ valid,
                                               const uint8_t* __restrict__ is_member,
                                               const unsigned* __restrict__ order, unsigned pos) {
    const unsigned row = order[pos];
    if (!valid[row] || __ldcg(is_member + row)) return;
    const double nd = (double)q.n;
    const double* fl = F + (size_t)q.low_row * dim;
    const double* fc = F + (size_t)row * dim;
    FastSum h = block_entropy_div<false>(
        dim, [&](uint64_t i) { return __dadd_rn(__dsub_rn(__ldcg(S + i), fl[i]), fc[i]); }, nd);
    if (threadIdx.x == 0) {
        const double mean_entropy = __ddiv_rn(__dadd_rn(__dsub_rn(q.E, H[q.low_row]), H[row]), nd);
        const double d = h.e - mean_entropy;
        const double b = fast_bound(dim, h.a, mean_entropy);
        const double thr = q.total_jsd + kEps, tb = q.total_bound + 4.0 * kEps;
        if (h.bad || !fast_total_ok(dim, h.t) || !(d == d)) {
            atomicMin(&sc->first_unsure, pos);
        } else if (d - b > thr + tb) {
            atomicMin(&sc->first_true, pos);
        } else if (!(d + b < thr - tb)) {
            atomicMin(&sc->first_unsure, pos);
        }
    }
}

__device__ __forceinline__ ScanScal load_scan_scal(const SelScal* sc, const unsigned* members) {
    ScanScal q;
    q.E = __ldcg(&sc->E);
    q.total_jsd = __ldcg(&sc->total_jsd);
    q.total_bound = __ldcg(&sc->total_bound);
    q.n = __ldcg(&sc->n);
    q.low_row = __ldcg(members + __ldcg(&sc->lowest));
    return q;
}

// host-driven window: one CTA per position pos0 + blockIdx.x
__global__ void __launch_bounds__(kFastThreads)
k_sel_scan_fast(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim, const double* __restrict__ S,
                const unsigned* __restrict__ members, SelScal* sc, const uint8_t* __restrict__ valid,
                const uint8_t* __restrict__ is_member, const unsigned* __restrict__ order, unsigned pos0) {
    scan_fast_body(F, H, dim, S, sc, load_scan_scal(sc, members), valid, is_member, order, pos0 + blockIdx.x);
}

// device-driven window: cursor / window / current buffer come from the scalar block, so the host can
// enqueue many rounds back to back without reading anything back
__global__ void __launch_bounds__(kFastThreads)
k_sel_scan_dev(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim, const double* __restrict__ S0,
               const double* __restrict__ S1, const unsigned* __restrict__ M0, const unsigned* __restrict__ M1,
               SelScal* sc, const uint8_t* __restrict__ valid, const uint8_t* __restrict__ is_member,
               const unsigned* __restrict__ order) {
    if (sc->halt) return;
    const unsigned cursor = sc->cursor, num = sc->num;
    if (cursor >= num) return;
    const unsigned count = min(sc->window, num - cursor);
    if (blockIdx.x >= count) return;
    const unsigned w = sc->which;
    scan_fast_body(F, H, dim, w ? S1 : S0, sc, load_scan_scal(sc, w ? M1 : M0), valid, is_member, order,
                   cursor + blockIdx.x);
}

// fast total_jsd + get_lowest_record_index: same shape as k_sel_update.  Leaves approximate
// total_jsd (with total_bound) and mdelta, and lowest_index only if it is certain.
__global__ void __launch_bounds__(kFastThreads)
k_sel_update_fast(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim,
                  const double* __restrict__ S, const unsigned* __restrict__ members, double* __restrict__ mdelta,
                  double* __restrict__ mbound, SelScal* sc) {
    __shared__ unsigned s_last;
    const unsigned j = blockIdx.x, n = sc->n;
    const double nd = (double)n;
    if (j == n) {
        FastSum h = block_entropy_div<false>(dim, [&](uint64_t i) { return S[i]; }, nd);
        if (threadIdx.x == 0) {
            const double me = __ddiv_rn(sc->E, nd);
            sc->total_jsd = h.e - me;
            sc->total_bound = fast_bound(dim, h.a, me);
            if (h.bad || !fast_total_ok(dim, h.t)) atomicExch(&sc->state_unsure, 1u);
        }
    } else {
        const unsigned row = members[j];
        const double div = __dsub_rn(nd, 1.0);
        const double* f = F + (size_t)row * dim;
        FastSum h = block_entropy_div<true>(dim, [&](uint64_t i) { return __dsub_rn(S[i], f[i]); }, div);
        if (threadIdx.x == 0) {
            const double mean_entropy = __ddiv_rn(__dsub_rn(sc->E, H[row]), div);
            mdelta[j] = h.e - mean_entropy;
            mbound[j] = fast_bound(dim, h.a, mean_entropy);
            if (h.bad || !fast_total_ok(dim, h.t)) atomicExch(&sc->state_unsure, 1u);
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&sc->ticket, 1u) == n) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {  // block-uniform
        __threadfence();
        const double total = __ldcg(&sc->total_jsd), tb = __ldcg(&sc->total_bound);
        unsigned low = 0;
        const unsigned unsure = finalize_fast_block(mdelta, mbound, n, total, tb, &low);
        stats_fast_block(mdelta, mbound, n, tb, sc);
        if (threadIdx.x == 0) {
            sc->lowest = low;
            if (unsure) atomicExch(&sc->state_unsure, 1u);
            sc->exact = 0;
            sc->ticket = 0;
            sc->first_true = kNone;
            sc->first_panic = kNone;
            sc->first_unsure = kNone;
        }
    }
}

// replace_lowest + total_jsd + get_lowest_record_index in ONE launch (fast path).  Every CTA forms
// the updated sums S'[i] = clamp(S[i] - f_low[i]) + f_c[i] on the fly from the OLD buffers (same
// operations as k_sel_replace_vec, so S' is bitwise the reference's); CTA n also stores S' and the
// last CTA to finish stores the new member list / scalars into the other buffer set.
// dev_pos != kNone: device-driven round; the last CTA also advances cursor / window / accepts / which.
__device__ __forceinline__ void replace_update_fast_body(
    const double* __restrict__ F, const double* __restrict__ H, uint64_t dim, const double* __restrict__ S_in,
    double* __restrict__ S_out, const unsigned* __restrict__ m_in, unsigned* __restrict__ m_out,
    uint8_t* __restrict__ is_member, double* __restrict__ mdelta, double* __restrict__ mbound, SelScal* sc,
    unsigned cand_row, unsigned dev_pos, unsigned dev_cursor, unsigned j, unsigned n, unsigned low, double E_old) {
    __shared__ unsigned s_last;
    const double nd = (double)n;
    const unsigned low_row = __ldcg(m_in + low);
    const double* fl = F + (size_t)low_row * dim;
    const double* fc = F + (size_t)cand_row * dim;
    const double E_new = __dadd_rn(__dsub_rn(E_old, H[low_row]), H[cand_row]);  // records.rs:101,129
    auto s_new = [&](uint64_t i) {
        double s = __dsub_rn(__ldcg(S_in + i), fl[i]);
        if (s <= kEps) s = 0.0;
        return __dadd_rn(s, fc[i]);
    };
    if (j == n) {
        FastSum h = block_entropy_div<false>(dim, [&](uint64_t i) {
            const double s = s_new(i);
            S_out[i] = s;
            return s;
        }, nd);
        if (threadIdx.x == 0) {
            const double me = __ddiv_rn(E_new, nd);
            sc->total_jsd = h.e - me;
            sc->total_bound = fast_bound(dim, h.a, me);
            if (h.bad || !fast_total_ok(dim, h.t)) atomicExch(&sc->state_unsure, 1u);
        }
    } else {
        // member j of the list after Vec::remove(low) + push(cand)
        const unsigned row = j < low ? __ldcg(m_in + j) : (j + 1 < n ? __ldcg(m_in + j + 1) : cand_row);
        const double div = __dsub_rn(nd, 1.0);
        const double* f = F + (size_t)row * dim;
        FastSum h = block_entropy_div<true>(dim, [&](uint64_t i) { return __dsub_rn(s_new(i), f[i]); }, div);
        if (threadIdx.x == 0) {
            const double mean_entropy = __ddiv_rn(__dsub_rn(E_new, H[row]), div);
            mdelta[j] = h.e - mean_entropy;
            mbound[j] = fast_bound(dim, h.a, mean_entropy);
            if (h.bad || !fast_total_ok(dim, h.t)) atomicExch(&sc->state_unsure, 1u);
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&sc->ticket, 1u) == n) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {  // block-uniform: the whole last CTA finishes the round
        __threadfence();
        for (unsigned t = threadIdx.x; t < n; t += blockDim.x)
            m_out[t] = t < low ? __ldcg(m_in + t) : (t + 1 < n ? __ldcg(m_in + t + 1) : cand_row);
        const double total = __ldcg(&sc->total_jsd), tb = __ldcg(&sc->total_bound);
        unsigned lo2 = 0;
        const unsigned unsure = finalize_fast_block(mdelta, mbound, n, total, tb, &lo2);
        if (dev_pos == kNone) stats_fast_block(mdelta, mbound, n, tb, sc);  // only max-mode callers read them
        if (threadIdx.x == 0) {
            is_member[low_row] = 0;
            is_member[cand_row] = 1;
            sc->E = E_new;
            sc->lowest = lo2;
            if (unsure) atomicExch(&sc->state_unsure, 1u);
            sc->exact = 0;
            sc->ticket = 0;
            sc->first_true = kNone;
            sc->first_panic = kNone;
            sc->first_unsure = kNone;
            if (dev_pos != kNone) {
                sc->window = max(64u, min(__ldcg(&sc->window_max), 2u * (dev_pos - dev_cursor + 1u)));
                sc->cursor = dev_pos + 1u;
                sc->accepts = __ldcg(&sc->accepts) + 1u;
                sc->which = __ldcg(&sc->which) ^ 1u;
            }
        }
    }
}

__global__ void __launch_bounds__(kFastThreads)
k_sel_replace_update_fast(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim,
                          const double* __restrict__ S_in, double* __restrict__ S_out,
                          const unsigned* __restrict__ m_in, unsigned* __restrict__ m_out,
                          uint8_t* __restrict__ is_member, double* __restrict__ mdelta, double* __restrict__ mbound,
                          SelScal* sc, unsigned cand_row) {
    replace_update_fast_body(F, H, dim, S_in, S_out, m_in, m_out, is_member, mdelta, mbound, sc, cand_row, kNone, 0u,
                             blockIdx.x, sc->n, sc->lowest, sc->E);
}

// One device-driven round after k_sel_scan_dev: every CTA takes the same decision from the scalar
// block (only the last CTA to finish modifies it): accept the first certain candidate (fused
// replace + update), advance past an empty window, or halt for the host when the first
// interesting candidate / the state could not be decided within the error bound.
__global__ void __launch_bounds__(kFastThreads)
k_sel_round_dev(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim, double* __restrict__ S0,
                double* __restrict__ S1, unsigned* __restrict__ M0, unsigned* __restrict__ M1,
                uint8_t* __restrict__ is_member, double* __restrict__ mdelta, double* __restrict__ mbound, SelScal* sc,
                const unsigned* __restrict__ order) {
    __shared__ unsigned s_last2;
    if (sc->halt) return;
    const unsigned cursor = sc->cursor, num = sc->num;
    if (cursor >= num) return;
    const unsigned window = sc->window, count = min(window, num - cursor);
    const unsigned ft = sc->first_true, fu = sc->first_unsure, su = sc->state_unsure, n = sc->n;
    if (!su && !(fu < ft) && ft != kNone) {
        const unsigned w = sc->which;
        replace_update_fast_body(F, H, dim, w ? S1 : S0, w ? S0 : S1, w ? M1 : M0, w ? M0 : M1, is_member, mdelta,
                                 mbound, sc, order[ft], ft, cursor, blockIdx.x, n, sc->lowest, sc->E);
        return;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last2 = (atomicAdd(&sc->ticket, 1u) == n) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last2 && threadIdx.x == 0) {
        sc->ticket = 0;
        if (su || fu < ft) {
            sc->halt = 1u;
        } else {  // empty window
            sc->cursor = cursor + count;
            sc->window = min(window * 2u, sc->window_max);
        }
    }
}

// All device-driven rounds in ONE cooperative launch: the CTAs stay resident, a round is
//   scan (one candidate per CTA and pass) | grid barrier | accept: fused replace + update over n+1 member
//   slots, or advance / halt | grid barrier
// with the same scalar-block protocol as k_sel_scan_dev / k_sel_round_dev, so the host loop and the exact
// fallbacks are unchanged.  Against the two-launches-per-round form this removes the launch and block
// scheduling latency of ~2 x 350 dependent kernels per nmost run and loads the round's scalars once per
// CTA instead of through chains of dependent global loads.
struct RoundScal {
    ScanScal q;
    unsigned lowest, which, cursor, count, window, stop, ft, fu, su;
};

__global__ void __launch_bounds__(kFastThreads)
k_sel_persist(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim, double* S0, double* S1,
              unsigned* M0, unsigned* M1, uint8_t* is_member, double* mdelta, double* mbound, SelScal* sc,
              const uint8_t* __restrict__ valid, const unsigned* __restrict__ order, unsigned max_rounds,
              unsigned long long* trace) {
    cg::grid_group grid = cg::this_grid();
    __shared__ RoundScal rs;
    // DVS_SELECT_TRACE: CTA 0 stamps %globaltimer at the four phase boundaries of the first rounds
    auto stamp = [&](unsigned round, int slot) {
        if (trace && blockIdx.x == 0 && threadIdx.x == 0 && round < 512) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            trace[round * 4 + slot] = t;
        }
    };
    for (unsigned round = 0; round < max_rounds; ++round) {
        stamp(round, 0);
        if (threadIdx.x == 0) {
            const unsigned cursor = __ldcg(&sc->cursor), num = __ldcg(&sc->num), window = __ldcg(&sc->window);
            rs.stop = __ldcg(&sc->halt) || cursor >= num;
            rs.cursor = cursor;
            rs.window = window;
            rs.count = cursor < num ? min(window, num - cursor) : 0u;
            rs.which = __ldcg(&sc->which);
            rs.lowest = __ldcg(&sc->lowest);
            rs.q = load_scan_scal(sc, rs.which ? M1 : M0);
        }
        __syncthreads();
        if (rs.stop) break;  // grid-uniform: every CTA read the same scalar block
        const unsigned which = rs.which, cursor = rs.cursor, count = rs.count, n = rs.q.n;
        const ScanScal q = rs.q;
        for (unsigned c = blockIdx.x; c < count; c += gridDim.x)
            scan_fast_body(F, H, dim, which ? S1 : S0, sc, q, valid, is_member, order, cursor + c);
        stamp(round, 1);
        grid.sync();
        stamp(round, 2);
        if (threadIdx.x == 0) {
            rs.ft = __ldcg(&sc->first_true);
            rs.fu = __ldcg(&sc->first_unsure);
            rs.su = __ldcg(&sc->state_unsure);
        }
        __syncthreads();
        const unsigned ft = rs.ft, fu = rs.fu, su = rs.su;
        if (!su && !(fu < ft) && ft != kNone) {
            const unsigned cand_row = order[ft];
            for (unsigned j = blockIdx.x; j <= n; j += gridDim.x)
                replace_update_fast_body(F, H, dim, which ? S1 : S0, which ? S0 : S1, which ? M1 : M0, which ? M0 : M1,
                                         is_member, mdelta, mbound, sc, cand_row, ft, cursor, j, n, rs.lowest, q.E);
        } else if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (su || fu < ft) {
                sc->halt = 1u;
            } else {  // empty window
                sc->cursor = cursor + count;
                sc->window = min(rs.window * 2u, __ldcg(&sc->window_max));
            }
        }
        stamp(round, 3);
        grid.sync();
    }
}

// ------------------------------------------------------------------ batched grow attempts (max) ----
// While a `max` selection is below max_size every candidate that increases the JSD costs the reference a
// `clone()` + `push` and a comparison of std / cov (records.rs:434-451), and the candidate is DISCARDED
// when the statistic does not improve — the state is unchanged.  On the benchmark set almost every
// candidate is of that kind (stdev 5..10 over 10.5k genomes: 4 adoptions), and one host-driven attempt
// costs ~210 us.  As with the windowed scan, a window of candidates is therefore evaluated concurrently
// against the same state: for each candidate the scan predicate and, for the grown set, H(S'/(n+1)) and the
// n+1 leave-one-out entropies (one CTA each, same operations as k_sel_sum + k_sel_update_fast), then the
// bounded statistic.  A candidate is skipped only if it CERTAINLY does not increase the JSD or CERTAINLY
// does not improve the statistic; the first candidate that is anything else is handed to the existing
// host path (which adopts it, or decides it exactly).  Decisions are unchanged, only certain discards
// are taken in bulk.
__global__ void __launch_bounds__(kFastThreads)
k_grow_eval(const double* __restrict__ F, uint64_t dim, const double* __restrict__ S_cur,
            const unsigned* __restrict__ members, const SelScal* __restrict__ sc_cur,
            const double* __restrict__ S_fresh, const uint8_t* __restrict__ valid,
            const uint8_t* __restrict__ is_member, const unsigned* __restrict__ order, unsigned cursor,
            FastSum* __restrict__ parts) {
    const unsigned c = blockIdx.y, t = blockIdx.x, n = sc_cur->n;
    const unsigned row = order[cursor + c];
    if (!valid[row] || is_member[row]) return;
    const double* fc = F + (size_t)row * dim;
    const double nd = (double)n;
    FastSum h;
    if (t == 0) {  // increases_jsd against the current state (records.rs:70-92)
        const double* fl = F + (size_t)members[sc_cur->lowest] * dim;
        h = block_entropy_div<false>(dim, [&](uint64_t i) { return __dadd_rn(__dsub_rn(S_cur[i], fl[i]), fc[i]); }, nd);
    } else if (t == 1) {  // total of the grown set: clone() re-sums in member order, push adds the candidate
        const double nd1 = __dadd_rn(nd, 1.0);
        h = block_entropy_div<false>(dim, [&](uint64_t i) { return __dadd_rn(S_fresh[i], fc[i]); }, nd1);
    } else {  // leave-one-out of member j of the grown set (j == n: the candidate itself)
        const unsigned j = t - 2;
        const double* f = j < n ? F + (size_t)members[j] * dim : fc;
        h = block_entropy_div<true>(dim, [&](uint64_t i) { return __dsub_rn(__dadd_rn(S_fresh[i], fc[i]), f[i]); }, nd);
    }
    if (threadIdx.x == 0) parts[(size_t)c * (n + 3) + t] = h;
}

__global__ void __launch_bounds__(kFastThreads)
k_grow_decide(const double* __restrict__ H, uint64_t dim, const unsigned* __restrict__ members,
              const SelScal* __restrict__ sc_cur, const SelScal* __restrict__ sc_fresh,
              const uint8_t* __restrict__ valid, const uint8_t* __restrict__ is_member,
              const unsigned* __restrict__ order, unsigned cursor, const FastSum* __restrict__ parts,
              double* __restrict__ md, double* __restrict__ mb, unsigned cap, SelScal* __restrict__ scratch,
              unsigned* __restrict__ first_interesting, int use_cov) {
    __shared__ int s_dec;
    const unsigned c = blockIdx.x, pos = cursor + c, n = sc_cur->n;
    const unsigned row = order[pos];
    if (!valid[row] || is_member[row]) return;  // skipped silently, like the scan
    const FastSum* P = parts + (size_t)c * (n + 3);
    const double nd = (double)n;
    if (threadIdx.x == 0) {
        const FastSum h = P[0];
        const unsigned low_row = members[sc_cur->lowest];
        const double me = __ddiv_rn(__dadd_rn(__dsub_rn(sc_cur->E, H[low_row]), H[row]), nd);
        const double d = h.e - me, b = fast_bound(dim, h.a, me);
        const double thr = sc_cur->total_jsd + kEps, tb = sc_cur->total_bound + 4.0 * kEps;
        int dec = 2;  // 0: certainly not increasing, 1: certainly increasing, 2: undecided
        if (!sc_cur->state_unsure && !h.bad && fast_total_ok(dim, h.t) && d == d) {
            if (d - b > thr + tb) dec = 1;
            else if (d + b < thr - tb) dec = 0;
        }
        s_dec = dec;
    }
    __syncthreads();
    const int dec = s_dec;
    if (dec == 0) return;
    if (dec == 2) {
        if (threadIdx.x == 0) atomicMin(first_interesting, pos);
        return;
    }
    // the grown set's total, member deltas and statistic (same forms as k_sel_update_fast)
    const double nd1 = __dadd_rn(nd, 1.0);
    const double E_try = __dadd_rn(sc_fresh->E, H[row]);
    const FastSum tot = P[1];
    const double me_t = __ddiv_rn(E_try, nd1);
    const double total_try = tot.e - me_t, tbound = fast_bound(dim, tot.a, me_t);
    int unsure = (tot.bad || !fast_total_ok(dim, tot.t)) ? 1 : 0;
    double* mdc = md + (size_t)c * cap;
    double* mbc = mb + (size_t)c * cap;
    for (unsigned j = threadIdx.x; j <= n; j += blockDim.x) {
        const FastSum hj = P[2 + j];
        const double Hj = j < n ? H[members[j]] : H[row];
        const double me = __ddiv_rn(__dsub_rn(E_try, Hj), nd);
        mdc[j] = total_try - (hj.e - me);
        mbc[j] = fast_bound(dim, hj.a, me);
        if (hj.bad || !fast_total_ok(dim, hj.t)) unsure = 1;
    }
    unsure = __syncthreads_or(unsure);
    if (unsure) {
        if (threadIdx.x == 0) atomicMin(first_interesting, pos);
        return;
    }
    stats_fast_block(mdc, mbc, n + 1, tbound, scratch + c);
    __syncthreads();
    if (threadIdx.x == 0) {
        const SelScal* g = scratch + c;
        const double sa = use_cov ? sc_cur->cov : sc_cur->stdv;
        const double ba = sc_cur->exact ? 0.0 : (use_cov ? sc_cur->cov_bound : sc_cur->std_bound);
        const double sb = use_cov ? g->cov : g->stdv, bb = use_cov ? g->cov_bound : g->std_bound;
        if (!(sb + bb < sa - ba)) atomicMin(first_interesting, pos);  // not a certain discard (NaN included)
    }
}

// ---------------------------------------------------------------- SM-replicated selection rounds ----
// k_sel_persist still pays ~10 dependent L2 round trips per round (scalar block, member list, ticket,
// last-CTA tail) and two cooperative-groups barriers.  For vectors that fit in shared memory (dim <= 4096,
// i.e. k <= 6) the whole selection state is instead REPLICATED in every SM:
//   shared memory: S, member rows and their entropies, per-member delta / bound, and the (row, valid,
//   entropy) of the next 512 positions of `order` (their rows are prefetched into L2 when the chunk is
//   staged);   registers: E, total_jsd, lowest, cursor, window.  (The global is_member map is only
//   brought up to date when the kernel ends.)
// Per round only 32-byte partial sums cross the L2, and they carry their own arrival flag, so there is no
// separate grid barrier:
//   scan    CTA b scores slice p = b % P of candidate c = b / P of the window (P = 4, 2 or 1 CTAs per
//           candidate, so a short window still uses every SM) and publishes {e, t, a, bad | tag} as two
//           self-validating 128-bit stores (tag = number of the exchange); CTAs without a candidate
//           publish an empty slot;
//   decide  the leader (CTA 0) polls all G slots until their tags match (one one-way latency after the
//           last writer), combines the partials in order, finds the first certain acceptance / first
//           undecided candidate and broadcasts them in one tagged 128-bit store that every other CTA polls;
//   accept  CTA j computes the leave-one-out entropy of member j (CTA n: H(S'/n)) and publishes its partial,
//           every CTA forms S' = clamp(S - f_lowest) + f_cand and the new member list in shared memory;
//   final   the leader gathers the n + 1 partials, forms the member deltas and the certified argmin and
//           broadcasts {total_jsd, total_bound, lowest, unsure}.
// (Letting every CTA poll every slot and decide redundantly was measured at 2.2 us per exchange against
// 1.3 us for the leader form — 148 x 148 pollers — tools/microbench/gridsync_bench.cu.)
// Slot reuse is safe without resets: the leader only broadcasts scan decision x after EVERY CTA has
// published its slot of exchange x, i.e. after every CTA has consumed all earlier broadcasts, and two
// update exchanges are always separated by a scan exchange; scan slots are double buffered.
// Arithmetic, bounds and the halt protocol are those of the kernels above (the partial sums only add
// P - 1 sequential additions, covered by `depth`; `a` travels as a float rounded UP, which only widens a
// bound), so decisions are identical; CTA 0 writes the state back to global memory when the rounds end or
// halt for the host.
constexpr unsigned kSmMaxDim = 4096, kSmMaxN = 1024, kSmMaxGrid = 256, kSmChunk = kFastThreads;
constexpr double kSmDepth = 4.0;

struct __align__(16) SmPart {
    unsigned long long w[4];  // {e, bad<<32 | tag}, {t, float_ru(a)<<32 | tag}
};

struct SmShared {
    double S[kSmMaxDim];
    double mH[2][kSmMaxN + 1];
    double md[kSmMaxN + 1];
    double mb[kSmMaxN + 1];
    unsigned members[2][kSmMaxN + 1];
    double cH[kSmChunk];
    unsigned crow[kSmChunk];
    unsigned char cvalid[kSmChunk];
    double pe[kSmMaxGrid], pt[kSmMaxGrid], pa[kSmMaxGrid];
    unsigned char pbad[kSmMaxGrid], wskip[kSmMaxGrid];
    unsigned ft, fu, unsure;
    double2 ltab[64];  // glibc log2 table {1/c, log2 c}
};

__device__ __forceinline__ void sm_st128(void* p, unsigned long long lo, unsigned long long hi) {
    asm volatile("{ .reg .b128 v; mov.b128 v, {%1, %2}; st.relaxed.gpu.global.b128 [%0], v; }" ::"l"(p), "l"(lo), "l"(hi)
                 : "memory");
}
__device__ __forceinline__ void sm_ld128(const void* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{ .reg .b128 v; ld.relaxed.gpu.global.b128 v, [%2]; mov.b128 {%0, %1}, v; }"
                 : "=l"(lo), "=l"(hi)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void sm_publish(SmPart* slot, const FastSum& h, unsigned tag) {
    sm_st128(&slot->w[0], (unsigned long long)__double_as_longlong(h.e), ((unsigned long long)(h.bad ? 1u : 0u) << 32) | tag);
    sm_st128(&slot->w[2], (unsigned long long)__double_as_longlong(h.t),
             ((unsigned long long)__float_as_uint(__double2float_ru(h.a)) << 32) | tag);
}
__device__ __forceinline__ FastSum sm_gather(const SmPart* slot, unsigned tag) {
    unsigned long long w0, w1, w2, w3;
    do sm_ld128(&slot->w[0], w0, w1); while ((unsigned)w1 != tag);
    do sm_ld128(&slot->w[2], w2, w3); while ((unsigned)w3 != tag);
    return FastSum{__longlong_as_double((long long)w0), __longlong_as_double((long long)w2),
                   (double)__uint_as_float((unsigned)(w3 >> 32)), (int)(w1 >> 32)};
}

// finalize_fast_block on this CTA's shared-memory copies, with the cross-warp step done redundantly by
// every thread (one barrier less, no serial tail on thread 0)
__device__ __forceinline__ unsigned sm_finalize(double* md, const double* mb_, unsigned n, double total,
                                                double total_bound, unsigned* low_out) {
    __shared__ double s_mn[kFastThreads / 32], s_mb[kFastThreads / 32];
    __shared__ unsigned s_ix[kFastThreads / 32];
    double mn = 1e300, mb = 0.0;
    unsigned ix = kNone;
    for (unsigned t = threadIdx.x; t < n; t += blockDim.x) {
        const double d = total - md[t];
        md[t] = d;
        if (d < mn || (d == mn && t < ix)) {
            mn = d; mb = mb_[t]; ix = t;
        }
    }
    for (int o = 16; o; o >>= 1) {
        const double omn = __shfl_xor_sync(0xffffffffu, mn, o), omb = __shfl_xor_sync(0xffffffffu, mb, o);
        const unsigned oix = __shfl_xor_sync(0xffffffffu, ix, o);
        if (omn < mn || (omn == mn && oix < ix)) {
            mn = omn; mb = omb; ix = oix;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_mn[threadIdx.x >> 5] = mn; s_mb[threadIdx.x >> 5] = mb; s_ix[threadIdx.x >> 5] = ix;
    }
    __syncthreads();
    mn = s_mn[0]; mb = s_mb[0]; ix = s_ix[0];
#pragma unroll
    for (unsigned w = 1; w < kFastThreads / 32; ++w) {
        const double wmn = s_mn[w];
        const unsigned wix = s_ix[w];
        if (wmn < mn || (wmn == mn && wix < ix)) {
            mn = wmn; mb = s_mb[w]; ix = wix;
        }
    }
    const unsigned low = (ix == kNone) ? 0u : ix;
    int unsure = 0;
    for (unsigned t = threadIdx.x; t < n; t += blockDim.x)
        if (t != low && !(mn + mb + 2.0 * kEps < md[t] - mb_[t])) unsure = 1;
    if (!(mn + mb + total_bound < 1e6)) unsure = 1;  // the reference's `min_delta_jsd = 1e6` initial value
    unsure = __syncthreads_or(unsure);  // (also fences s_mn / s_mb / s_ix for the next call)
    *low_out = low;
    return (unsigned)unsure;
}

constexpr int kSmTraceSlots = 8;

__global__ void __launch_bounds__(kFastThreads, 1)
k_sel_persist_sm(const double* __restrict__ F, const double* __restrict__ H, unsigned dim, double* S_glob,
                 unsigned* M_glob, uint8_t* is_member, double* mdelta_g, double* mbound_g, SelScal* sc,
                 const uint8_t* __restrict__ valid, const unsigned* __restrict__ order, SmPart* spart, SmPart* upart,
                 SmPart* dpart, unsigned long long* trace, int trace_all) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    SmShared& sm = *reinterpret_cast<SmShared*>(sm_raw);
    const unsigned tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
    unsigned tr_round = 0;
    // DVS_SELECT_TRACE: CTA 0's timeline of the first 256 rounds; DVS_SELECT_TRACE_ALL: every CTA's
    auto stamp = [&](int slot) {
        if (trace && tid == 0 && tr_round < 256 && (b == 0 || trace_all)) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (trace_all)
                trace[((size_t)tr_round * kSmTraceSlots + slot) * kSmMaxGrid + b] = t;
            else
                trace[tr_round * kSmTraceSlots + slot] = t;
        }
    };

    // ---- every CTA loads the state the host / the previous kernels left in global memory ----
    const unsigned n = sc->n, num = sc->num;
    const double nd = (double)n, div = __dsub_rn(nd, 1.0);
    const FastDiv div_n = make_fast_div(nd), div_n1 = make_fast_div(div);
    double E = sc->E, total_jsd = sc->total_jsd, total_bound = sc->total_bound;
    unsigned lowest = sc->lowest, cursor = sc->cursor, window = sc->window, accepts = sc->accepts;
    unsigned state_unsure = sc->state_unsure, halt = state_unsure ? 1u : 0u, mw = 0;
    bool touched = false;  // an acceptance happened in this launch: md / mb / total are this kernel's
    dvs_log2_stage_table(sm.ltab);
    for (unsigned i = tid; i < dim; i += kFastThreads) sm.S[i] = S_glob[i];
    for (unsigned j = tid; j < n; j += kFastThreads) {
        const unsigned r = M_glob[j];
        sm.members[0][j] = r;
        sm.mH[0][j] = H[r];
    }
    __syncthreads();
    const unsigned wmin = max(1u, G / 4u);
    // loop-invariant factors of fast_bound / fast_total_ok (same expressions, same rounding)
    const double kb0 = ((double)dim + fast_slack(dim) + 16.0) * 1.2e-16;
    const double kb4 = ((double)dim + fast_slack(dim, kSmDepth) + 16.0) * 1.2e-16;
    const double lim0 = ((double)dim + 1.0 - fast_slack(dim) - 2.0) * 1.1102230246251565e-16;
    const double lim4 = ((double)dim + 1.0 - fast_slack(dim, kSmDepth) - 2.0) * 1.1102230246251565e-16;
    auto total_ok = [](double t, double lim) { return lim > 0.0 && fabs(t - 1.0) <= lim; };
    unsigned xs = 0, xu = 0;       // scan / update exchanges so far (= tags)
    unsigned cbase = 0, cend = 0;  // positions [cbase, cend) of `order` are staged in shared memory

    while (!halt && cursor < num) {
        stamp(0);
        window = max(1u, min(window, G));
        const unsigned P = (dim >= 2048u && window * 4u <= G) ? 4u : ((dim >= 2048u && window * 2u <= G) ? 2u : 1u);
        const unsigned count = min(window, num - cursor);
        if (cursor < cbase || cursor + count > cend) {  // stage the next chunk of positions (CTA-uniform)
            __syncthreads();
            cbase = cursor;
            cend = min(num, cbase + kSmChunk);
            if (cbase + tid < cend) {
                const unsigned row = order[cbase + tid];
                sm.crow[tid] = row;
                sm.cvalid[tid] = valid[row];
                sm.cH[tid] = H[row];
            }
            // ... and pull their rows towards L2, one 128-byte line per prefetch, rows dealt over the CTAs
            for (unsigned r = b; r < cend - cbase; r += G) {
                const double* fr = F + (size_t)order[cbase + r] * dim;
                for (unsigned l = tid * 16u; l < dim; l += kFastThreads * 16u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(fr + l));
            }
            __syncthreads();
        }
        const unsigned coff = cursor - cbase;
        if (tid == 0) {
            sm.ft = kNone;
            sm.fu = kNone;
        }
        if (tid < count) sm.wskip[tid] = sm.cvalid[coff + tid] ? 0 : 1;
        // ---- scan: slice p of candidate c ----
        ++xs;
        SmPart* const sbuf = spart + (xs & 1u) * kSmMaxGrid;
        const unsigned c = b / P, p = b % P;
        const double* fl = F + (size_t)sm.members[mw][lowest] * dim;
        if (c < count && sm.cvalid[coff + c]) {  // CTA-uniform (a member's score is published but never read)
            const double* fc = F + (size_t)sm.crow[coff + c] * dim;
            const unsigned lo = (unsigned)(((uint64_t)dim * p) / P), hi = (unsigned)(((uint64_t)dim * (p + 1)) / P);
            auto num = [&](unsigned i) { return __dadd_rn(__dsub_rn(sm.S[i], fl[i]), fc[i]); };
            const FastSum h = P == 4u   ? block_entropy_ilp<false, 2>(lo, hi, num, div_n, sm.ltab)
                              : P == 2u ? block_entropy_ilp<false, 4>(lo, hi, num, div_n, sm.ltab)
                                        : block_entropy_ilp<false, 8>(lo, hi, num, div_n, sm.ltab);
            if (tid == 0) sm_publish(sbuf + b, h, xs);
        } else if (tid == 0) {
            // every CTA publishes in every scan exchange: the leader's wait for all G slots is what bounds
            // how far any CTA can lag, i.e. what makes the reuse of all exchange slots safe
            sm_publish(sbuf + b, FastSum{0.0, 0.0, 0.0, 0}, xs);
        }
        stamp(1);
        // ---- gather + decision by the leader (CTA 0), broadcast of {first_true, first_unsure} ----
        // (an all-to-all gather — every CTA polling every slot — costs 2.2 us per exchange on 148 SMs, the
        // leader form 1.3 us: tools/microbench/gridsync_bench.cu)
        SmPart* const dslot = dpart + (xs & 1u);
        if (b == 0) {
            // a candidate that is already a member is skipped (records.rs:76-78): the replicated member list
            // is the ground truth, searched while the partials are still in flight (thread j holds member j
            // and walks the window)
            __syncthreads();  // wskip initialised (a CTA without a candidate has not passed a barrier yet)
            for (unsigned j = tid; j < n; j += kFastThreads) {
                const unsigned r = sm.members[mw][j];
                for (unsigned cc = 0; cc < count; ++cc)
                    if (sm.crow[coff + cc] == r) sm.wskip[cc] = 1;
            }
            __syncthreads();
            if (tid < G) {
                const FastSum g = sm_gather(sbuf + tid, xs);
                if (tid < count * P && !sm.wskip[tid / P]) {
                    sm.pe[tid] = g.e; sm.pt[tid] = g.t; sm.pa[tid] = g.a; sm.pbad[tid] = (unsigned char)g.bad;
                }
            }
            __syncthreads();
            stamp(2);
            if (tid < count && !sm.wskip[tid]) {
                FastSum h{sm.pe[tid * P], sm.pt[tid * P], sm.pa[tid * P], sm.pbad[tid * P]};
                for (unsigned q = 1; q < P; ++q) {
                    h.e += sm.pe[tid * P + q]; h.t += sm.pt[tid * P + q]; h.a += sm.pa[tid * P + q];
                    h.bad |= sm.pbad[tid * P + q];
                }
                const unsigned pos = cursor + tid;
                const double mean_entropy =
                    div_exact(__dadd_rn(__dsub_rn(E, sm.mH[mw][lowest]), sm.cH[coff + tid]), div_n);
                const double d = h.e - mean_entropy;
                const double bd = (P > 1u ? kb4 : kb0) * (h.a + fabs(mean_entropy) + 1.0);
                const double thr = total_jsd + kEps, tb = total_bound + 4.0 * kEps;
                if (h.bad || !total_ok(h.t, P > 1u ? lim4 : lim0) || !(d == d)) {
                    atomicMin(&sm.fu, pos);
                } else if (d - bd > thr + tb) {
                    atomicMin(&sm.ft, pos);
                } else if (!(d + bd < thr - tb)) {
                    atomicMin(&sm.fu, pos);
                }
            }
            __syncthreads();
            if (tid == 0) sm_st128(&dslot->w[0], ((unsigned long long)sm.fu << 32) | sm.ft, xs);
        } else {
            if (tid == 0) {
                unsigned long long w0, w1;
                do sm_ld128(&dslot->w[0], w0, w1); while ((unsigned)w1 != xs);
                sm.ft = (unsigned)w0;
                sm.fu = (unsigned)(w0 >> 32);
            }
            stamp(2);
            __syncthreads();
        }
        const unsigned ft = sm.ft, fu = sm.fu;
        __syncthreads();
        stamp(3);
        if (fu < ft) {  // the first interesting candidate is undecided: the host resolves it exactly
            halt = 1;
            break;
        }
        if (ft == kNone) {  // empty window
            cursor += count;
            window = min(window * 2u, G);
            stamp(4); stamp(5); stamp(6);
            ++tr_round;
            continue;
        }
        // ---- accept: replace_lowest + leave-one-out update ----
        const unsigned cand = sm.crow[ft - cbase];
        const double Hc = sm.cH[ft - cbase];
        const double E_new = __dadd_rn(__dsub_rn(E, sm.mH[mw][lowest]), Hc);  // records.rs:101,129
        const double* fc = F + (size_t)cand * dim;
        auto member_after = [&](unsigned j) {  // Vec::remove(lowest) + push(cand)
            return j < lowest ? sm.members[mw][j] : (j + 1 < n ? sm.members[mw][j + 1] : cand);
        };
        auto s_new = [&](unsigned i) {
            double s = __dsub_rn(sm.S[i], fl[i]);
            if (s <= kEps) s = 0.0;
            return __dadd_rn(s, fc[i]);
        };
        ++xu;
        bool have_S = false;
        for (unsigned j = b; j <= n; j += G) {
            FastSum h;
            if (j == n) {
                h = block_entropy_ilp<false, 8>(0u, dim, [&](unsigned i) {
                    const double s = have_S ? sm.S[i] : s_new(i);
                    if (!have_S) sm.S[i] = s;
                    return s;
                }, div_n, sm.ltab);
            } else {
                const double* f = F + (size_t)member_after(j) * dim;
                h = block_entropy_ilp<true, 8>(0u, dim, [&](unsigned i) {
                    const double s = have_S ? sm.S[i] : s_new(i);
                    if (!have_S) sm.S[i] = s;
                    return __dsub_rn(s, f[i]);
                }, div_n1, sm.ltab);
            }
            if (tid == 0) sm_publish(upart + j, h, xu);
            have_S = true;
        }
        if (!have_S)
            for (unsigned i = tid; i < dim; i += kFastThreads) sm.S[i] = s_new(i);
        stamp(4);
        // ---- every CTA: new member list; the leader: deltas + certified argmin, broadcast ----
        for (unsigned t = tid; t < n; t += kFastThreads) {
            sm.members[mw ^ 1][t] = member_after(t);
            sm.mH[mw ^ 1][t] = t < lowest ? sm.mH[mw][t] : (t + 1 < n ? sm.mH[mw][t + 1] : Hc);
        }
        if (tid == 0) sm.unsure = 0;
        __syncthreads();
        mw ^= 1;
        E = E_new;
        unsigned lo2 = 0, unsure = 0;
        SmPart* const fslot = dpart + 2;
        if (b == 0) {
            int uns = 0;
            for (unsigned t = tid; t <= n; t += kFastThreads) {
                const FastSum g = sm_gather(upart + t, xu);
                if (t == n) {
                    sm.pe[0] = g.e; sm.pa[0] = g.a;  // (scan staging is free again)
                } else {
                    const double mean_entropy = div_exact(__dsub_rn(E, sm.mH[mw][t]), div_n1);
                    sm.md[t] = g.e - mean_entropy;
                    sm.mb[t] = kb0 * (g.a + fabs(mean_entropy) + 1.0);
                }
                if (g.bad || !total_ok(g.t, lim0)) uns = 1;
            }
            if (uns) sm.unsure = 1;  // benign race: every writer stores 1
            __syncthreads();
            stamp(5);
            const double me = div_exact(E, div_n);
            total_jsd = sm.pe[0] - me;
            total_bound = kb0 * (sm.pa[0] + fabs(me) + 1.0);
            unsure = sm_finalize(sm.md, sm.mb, n, total_jsd, total_bound, &lo2);
            unsure |= sm.unsure;
            if (tid == 0) {
                sm_st128(&fslot->w[0], (unsigned long long)__double_as_longlong(total_jsd), ((unsigned long long)lo2 << 32) | xu);
                sm_st128(&fslot->w[2], (unsigned long long)__double_as_longlong(total_bound),
                         ((unsigned long long)unsure << 32) | xu);
            }
        } else {
            if (tid == 0) {
                unsigned long long w0, w1, w2, w3;
                do sm_ld128(&fslot->w[0], w0, w1); while ((unsigned)w1 != xu);
                do sm_ld128(&fslot->w[2], w2, w3); while ((unsigned)w3 != xu);
                sm.pe[0] = __longlong_as_double((long long)w0);
                sm.pa[0] = __longlong_as_double((long long)w2);
                sm.ft = (unsigned)(w1 >> 32);
                sm.unsure = (unsigned)(w3 >> 32);
            }
            __syncthreads();
            stamp(5);
            total_jsd = sm.pe[0];
            total_bound = sm.pa[0];
            lo2 = sm.ft;
            unsure = sm.unsure;
            __syncthreads();
        }
        lowest = lo2;
        touched = true;
        window = max(wmin, min(G, 2u * (ft - cursor + 1u)));
        cursor = ft + 1u;
        ++accepts;
        stamp(6);
        ++tr_round;
        if (unsure) {  // the argmin / a sum check could not be certified: the host redoes the update exactly
            state_unsure = 1;
            halt = 1;
            break;
        }
    }
    stamp(0);

    // ---- CTA 0 hands the state back in the layout the other kernels and the host loop use ----
    if (b == 0) {
        __syncthreads();
        if (touched) {
            for (unsigned i = tid; i < dim; i += kFastThreads) S_glob[i] = sm.S[i];
            for (unsigned j = tid; j < n; j += kFastThreads) is_member[M_glob[j]] = 0;  // the set at entry
            __syncthreads();
            for (unsigned j = tid; j < n; j += kFastThreads) {
                const unsigned r = sm.members[mw][j];
                M_glob[j] = r;
                is_member[r] = 1;
                mdelta_g[j] = sm.md[j];
                mbound_g[j] = sm.mb[j];
            }
        }
        if (tid == 0) {
            if (touched) {
                sc->E = E;
                sc->total_jsd = total_jsd;
                sc->total_bound = total_bound;
                sc->lowest = lowest;
                sc->exact = 0;
            }
            sc->state_unsure = state_unsure;
            sc->ticket = 0;
            sc->first_true = kNone;
            sc->first_panic = kNone;
            sc->first_unsure = kNone;
            sc->cursor = cursor;
            sc->window = window;
            sc->accepts = accepts;
            sc->halt = halt;
        }
    }
}

// test hook: the two building blocks of block_entropy_ilp on arbitrary operands
__global__ void k_debug_fast_terms(const double* a, const double* b, double* m_out, double* l_out, int* sp_out, uint64_t n) {
    __shared__ double2 ltab[64];
    dvs_log2_stage_table(ltab);
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = div_exact(a[i], make_fast_div(b[i]));
    int sp = 0;
    const double l = dvs_log2_main(m, ltab, sp);
    m_out[i] = m;
    l_out[i] = l;
    sp_out[i] = sp;
}

__global__ void k_sel_set_dev(SelScal* sc, unsigned cursor, unsigned window, unsigned window_max, unsigned num,
                              unsigned accepts, unsigned which) {
    sc->cursor = cursor;
    sc->window = window;
    sc->window_max = window_max;
    sc->num = num;
    sc->accepts = accepts;
    sc->which = which;
    sc->halt = 0;
}

__global__ void k_set_members(uint8_t* is_member, const unsigned* members, unsigned n, uint8_t v) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_member[members[i]] = v;
}

static int panic_error(const SelScal& h) {
    set_error("cannot calculate entropy as frequency vector total %.17g!=1.0", h.panic_total);
    return DVS_ERR_VALUE;
}

struct Selector {
    dvs_ctx* ctx;
    const dvs_kfreqs* f;
    uint64_t dim;
    cudaStream_t st;
    SelScal* h_sc;  // pinned

    int read(SelState& s) {
        DVS_CUDA_TRY(cudaMemcpyAsync(h_sc, s.sc.p, sizeof(SelScal), cudaMemcpyDeviceToHost, st));
        DVS_CUDA_TRY(cudaStreamSynchronize(st));
        return DVS_OK;
    }
    unsigned vec_grid() const { return (unsigned)((dim + 255) / 256); }
    int reset_scan(SelState& s) {
        static_assert(offsetof(SelScal, first_panic) == offsetof(SelScal, first_true) + sizeof(unsigned), "layout");
        DVS_CUDA_TRY(cudaMemsetAsync((char*)s.sc.p + offsetof(SelScal, first_true), 0xFF, 2 * sizeof(unsigned), st));
        DVS_CUDA_TRY(cudaMemsetAsync((char*)s.sc.p + offsetof(SelScal, first_unsure), 0xFF, sizeof(unsigned), st));
        return DVS_OK;
    }
    int update_fast(SelState& s, unsigned n) {
        k_sel_update_fast<<<n + 1, kFastThreads, 0, st>>>(f->freqs.p, f->entropy.p, dim, s.S(), s.members(),
                                                          s.mdelta.p, s.mbound.p, s.sc.p);
        DVS_LAUNCHED(ctx);
        return DVS_OK;
    }
    int replace_fast(SelState& s, unsigned n, unsigned cand_row, uint8_t* is_member) {
        k_sel_replace_update_fast<<<n + 1, kFastThreads, 0, st>>>(f->freqs.p, f->entropy.p, dim, s.S(), s.S_other(),
                                                                  s.members(), s.members_other(), is_member,
                                                                  s.mdelta.p, s.mbound.p, s.sc.p, cand_row);
        DVS_LAUNCHED(ctx);
        s.flip();
        return DVS_OK;
    }
    // (first_true / first_unsure are re-armed by every update kernel and stay armed after an empty
    // window; the caller resets them explicitly after an exact single-candidate scan)
    int scan_fast(SelState& s, const uint8_t* is_member, const unsigned* d_order, unsigned pos0, unsigned count) {
        k_sel_scan_fast<<<count, kFastThreads, 0, st>>>(f->freqs.p, f->entropy.p, dim, s.S(), s.members(), s.sc.p,
                                                         f->valid.p, is_member, d_order, pos0);
        DVS_LAUNCHED(ctx);
        return DVS_OK;
    }
    // one candidate row of another kfreqs object, no order / validity / membership filters
    int scan_raw(SelState& s, const double* candF, const double* candH, unsigned row, double* delta_out) {
        DVS_TRY(reset_scan(s));
        k_sel_scan<<<1, kEntThreads, kEntSmemBytes, st>>>(f->freqs.p, f->entropy.p, dim, s.S(), s.members(), s.sc.p,
                                                           candF, candH, nullptr, nullptr, nullptr, row, delta_out);
        DVS_LAUNCHED(ctx);
        return DVS_OK;
    }

    // SummedRecords::new over `members` (+ optional pushed row): sums, then total_jsd + member deltas
    int build(SelState& s, const unsigned* d_members, unsigned n, int extra_row) {
        k_sel_sum<<<vec_grid(), 256, 0, st>>>(f->freqs.p, f->entropy.p, dim, d_members, n, extra_row, s.S(),
                                              s.members(), s.sc.p);
        DVS_LAUNCHED(ctx);
        return update(s, n + (extra_row >= 0 ? 1u : 0u));
    }
    // fast SummedRecords::new (+ optional pushed row): exact sums, bounded-error total / deltas / stats
    int build_fast(SelState& s, const unsigned* d_members, unsigned n, int extra_row) {
        k_sel_sum<<<vec_grid(), 256, 0, st>>>(f->freqs.p, f->entropy.p, dim, d_members, n, extra_row, s.S(),
                                              s.members(), s.sc.p);
        DVS_LAUNCHED(ctx);
        return update_fast(s, n + (extra_row >= 0 ? 1u : 0u));
    }
    int update(SelState& s, unsigned n) {
        k_sel_update<<<n + 1, kEntThreads, kEntSmemBytes, st>>>(f->freqs.p, f->entropy.p, dim, s.S(), s.members(),
                                                                 s.mdelta.p, s.sc.p);
        DVS_LAUNCHED(ctx);
        return DVS_OK;
    }
    int replace(SelState& s, unsigned n, unsigned cand_row, uint8_t* is_member) {
        k_sel_replace_vec<<<vec_grid(), 256, 0, st>>>(f->freqs.p, f->entropy.p, dim, s.members(), is_member,
                                                      s.sc.p, cand_row, s.S());
        DVS_LAUNCHED(ctx);
        return update(s, n);
    }
    int scan(SelState& s, const dvs_kfreqs* q, const uint8_t* is_member, const unsigned* d_order, unsigned pos0,
             unsigned count, double* delta_out) {
        DVS_TRY(reset_scan(s));  // re-arm the min-index reduction
        k_sel_scan<<<count, kEntThreads, kEntSmemBytes, st>>>(f->freqs.p, f->entropy.p, dim, s.S(), s.members(),
                                                               s.sc.p, q->freqs.p, q->entropy.p, q->valid.p,
                                                               is_member, d_order, pos0, delta_out);
        DVS_LAUNCHED(ctx);
        return DVS_OK;
    }
};

}  // namespace dvs

using namespace dvs;

struct dvs_summed {
    int device = 0;
    const dvs_kfreqs* f = nullptr;
    SelState st;
    unsigned n = 0;
};

extern "C" {

int dvs_select(dvs_ctx* ctx, const dvs_kfreqs* f, const uint32_t* order, uint32_t num, int mode, uint32_t min_size,
               uint32_t max_size, uint32_t* sel_idx, double* sel_delta, double* stats5, uint32_t* size_out) {
    if (!ctx || !f || (!order && num) || !size_out) {
        set_error("dvs_select: NULL argument");
        return DVS_ERR_ARG;
    }
    if (mode < DVS_MODE_NMOST || mode > DVS_MODE_MAX_COV) {
        set_error("dvs_select: bad mode %d", mode);
        return DVS_ERR_ARG;
    }
    ctx->last_accepts = 0;
    if (num < min_size) {  // records.rs:323-325, :404-410
        set_error("The number of sequences %u is < n %u", num, min_size);
        return DVS_ERR_VALUE;
    }
    if (mode == DVS_MODE_NMOST) max_size = min_size;
    if (!(num > max_size)) max_size = num;  // records.rs:412-416
    for (uint32_t i = 0; i < num; ++i)
        if (order[i] >= f->nrec) {
            set_error("dvs_select: order[%u]=%u out of range (nrec=%u)", i, order[i], f->nrec);
            return DVS_ERR_ARG;
        }
    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    const uint64_t dim = f->dim;

    // validity / panic flags of the records, needed on the host to form the initial set
    std::vector<uint8_t> valid(f->nrec), err(f->nrec);
    std::vector<double> err_total(f->nrec);
    if (f->nrec) {
        DVS_CUDA_TRY(cudaMemcpyAsync(valid.data(), f->valid.p, f->nrec, cudaMemcpyDeviceToHost, st));
        DVS_CUDA_TRY(cudaMemcpyAsync(err.data(), f->err.p, f->nrec, cudaMemcpyDeviceToHost, st));
        DVS_CUDA_TRY(cudaMemcpyAsync(err_total.data(), f->err_total.p, f->nrec * sizeof(double),
                                     cudaMemcpyDeviceToHost, st));
        DVS_CUDA_TRY(cudaStreamSynchronize(st));
    }
    // every record in `order` goes through KmerSeq::new -> entropy(); a failing sum check panics
    for (uint32_t i = 0; i < num; ++i)
        if (valid[order[i]] && err[order[i]]) {
            set_error("cannot calculate entropy as frequency vector total %.17g!=1.0", err_total[order[i]]);
            return DVS_ERR_VALUE;
        }

    std::vector<unsigned> init;  // records.rs:299-305: first min_size records, failures dropped
    for (uint32_t i = 0; i < min_size; ++i)
        if (valid[order[i]]) init.push_back(order[i]);
    if (init.empty()) {
        set_error("records cannot be empty");  // records.rs:28-30
        return DVS_ERR_VALUE;
    }
    if (init.size() <= 1) {
        set_error("must have > 1 KmerSeq");  // records.rs:227-230
        return DVS_ERR_VALUE;
    }
    {   // duplicates in the initial set: the reference's Vec keeps both rows; keep that behaviour
    }

    const unsigned cap = std::max<unsigned>(std::max(min_size, max_size), (unsigned)init.size()) + 1;
    SelState A, B;
    DVS_TRY(A.alloc(dim, cap));
    const bool grow_mode = (mode != DVS_MODE_NMOST);
    if (grow_mode) DVS_TRY(B.alloc(dim, cap));
    DevBuf<uint8_t> is_member;
    DevBuf<unsigned> d_order, d_init;
    DVS_TRY(is_member.alloc(f->nrec));
    DVS_TRY(d_order.alloc(num));
    DVS_TRY(d_init.alloc(init.size()));
    DVS_CUDA_TRY(cudaMemsetAsync(is_member.p, 0, f->nrec, st));
    DVS_CUDA_TRY(cudaMemcpyAsync(d_order.p, order, num * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    DVS_CUDA_TRY(cudaMemcpyAsync(d_init.p, init.data(), init.size() * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    k_set_members<<<(unsigned)((init.size() + 255) / 256), 256, 0, st>>>(is_member.p, d_init.p, (unsigned)init.size(), 1);
    DVS_LAUNCHED(ctx);

    PhaseTimer pt(ctx, DVS_PHASE_SELECT);
    Selector sel{ctx, f, dim, st, (SelScal*)ctx->pinned};
    SelState* cur = &A;
    SelState* alt = &B;
    unsigned n = (unsigned)init.size();
    DVS_TRY(sel.build(*cur, d_init.p, n, -1));

    const unsigned window_max = std::max(64u, (unsigned)ctx->sm_count * 12u);
    unsigned cursor = min_size;
    unsigned accepts = 0;
    unsigned window = 64;  // adaptive: grows while windows come back empty, shrinks after a hit
    // DVS_SELECT_EXACT_ONLY=1 disables the bounded-error fast path (tests compare both)
    const char* exact_env = getenv("DVS_SELECT_EXACT_ONLY");
    const bool use_fast = !(exact_env && exact_env[0] == '1');
    unsigned exact_evals = 0;
    const unsigned window_max_dev = std::max(64u, (unsigned)ctx->sm_count * 4u);  // grid of a device-driven scan
    const char* dev_env = getenv("DVS_SELECT_HOST_LOOP");
    const bool use_dev = use_fast && !(dev_env && dev_env[0] == '1');
    constexpr int kRoundsPerBatch = 32;
    // DVS_SELECT_PERSIST=0 keeps the two-launches-per-round form, =1 the global-state persistent kernel
    // (A/B measurements, fallbacks); default: SM-replicated rounds when the state fits in shared memory
    const char* per_env = getenv("DVS_SELECT_PERSIST");
    int coop = 0, per_sm = 0, per_sm2 = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sel_persist, kFastThreads, 0) != cudaSuccess)
        per_sm = 0;
    const bool use_persist = use_dev && coop && per_sm > 0 && !(per_env && per_env[0] == '0');
    const unsigned persist_grid = (unsigned)ctx->sm_count * (unsigned)std::min(per_sm, 2);
    bool sm_ok = use_persist && !(per_env && per_env[0] == '1') && dim <= kSmMaxDim;
    if (sm_ok) {
        sm_ok = cudaFuncSetAttribute(k_sel_persist_sm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(SmShared)) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_sel_persist_sm, kFastThreads,
                                                              sizeof(SmShared)) == cudaSuccess &&
                per_sm2 > 0;
        (void)cudaGetLastError();
    }
    const unsigned sm_grid = std::min<unsigned>((unsigned)ctx->sm_count, kSmMaxGrid);
    DevBuf<SmPart> d_spart, d_upart, d_dpart;
    if (sm_ok) {
        DVS_TRY(d_spart.alloc(2 * kSmMaxGrid));
        DVS_TRY(d_upart.alloc(kSmMaxN + 1));
        DVS_TRY(d_dpart.alloc(3));  // the leader's broadcasts: scan decision (double buffered), update result
    }
    // batched grow attempts (see k_grow_eval): scratch for a window of candidates; DVS_SELECT_GROW_BATCH=0
    // keeps one host-driven attempt per candidate
    const char* gb_env = getenv("DVS_SELECT_GROW_BATCH");
    const bool use_grow_batch = use_fast && grow_mode && !(gb_env && gb_env[0] == '0');
    constexpr unsigned kGrowWindowMax = 64;
    SelState fresh;  // clone(): the members re-summed in order (S, E), refreshed whenever the set changes
    DevBuf<FastSum> d_gparts;
    DevBuf<double> d_gmd, d_gmb;
    DevBuf<SelScal> d_gscratch;
    DevBuf<unsigned> d_ginteresting;
    bool fresh_valid = false;
    unsigned grow_window = 16;
    if (use_grow_batch && n < max_size) {
        DVS_TRY(fresh.alloc(dim, cap));
        DVS_TRY(d_gparts.alloc((size_t)kGrowWindowMax * (cap + 3)));
        DVS_TRY(d_gmd.alloc((size_t)kGrowWindowMax * cap));
        DVS_TRY(d_gmb.alloc((size_t)kGrowWindowMax * cap));
        DVS_TRY(d_gscratch.alloc(kGrowWindowMax));
        DVS_TRY(d_ginteresting.alloc(1));
    }
    unsigned fresh_n = 0;
    const SelState* fresh_of = nullptr;
    int fresh_which = -1;
    bool single_candidate = false;  // the grow filter has already isolated the candidate at `cursor`
    // while the set grows at almost every candidate (e.g. cov right after the start) the filter skips
    // nothing: after two such windows in a row it is bypassed for 1, 2, 4, 8 candidates
    unsigned filter_streak = 0, filter_bypass = 0;
    while (cursor < num) {
        if (use_grow_batch && n < max_size && !single_candidate && filter_bypass > 0) {
            --filter_bypass;
            single_candidate = true;  // straight to the host path for this candidate
            continue;
        }
        if (use_grow_batch && n < max_size && !single_candidate) {
            if (!fresh_valid || fresh_n != n || fresh_of != cur || fresh_which != cur->which) {
                k_sel_sum<<<sel.vec_grid(), 256, 0, st>>>(f->freqs.p, f->entropy.p, dim, cur->members(), n, -1, fresh.S(),
                                                          fresh.members(), fresh.sc.p);
                DVS_LAUNCHED(ctx);
                fresh_valid = true;
                fresh_n = n;
                fresh_of = cur;
                fresh_which = cur->which;
            }
            const unsigned W = std::min(grow_window, num - cursor);
            DVS_CUDA_TRY(cudaMemsetAsync(d_ginteresting.p, 0xFF, sizeof(unsigned), st));
            k_grow_eval<<<dim3(n + 3, W), kFastThreads, 0, st>>>(f->freqs.p, dim, cur->S(), cur->members(), cur->sc.p,
                                                                 fresh.S(), f->valid.p, is_member.p, d_order.p, cursor,
                                                                 d_gparts.p);
            DVS_LAUNCHED(ctx);
            k_grow_decide<<<W, kFastThreads, 0, st>>>(f->entropy.p, dim, cur->members(), cur->sc.p, fresh.sc.p, f->valid.p,
                                                      is_member.p, d_order.p, cursor, d_gparts.p, d_gmd.p, d_gmb.p, cap,
                                                      d_gscratch.p, d_ginteresting.p, mode == DVS_MODE_MAX_COV ? 1 : 0);
            DVS_LAUNCHED(ctx);
            unsigned* h_fi = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ctx->pinned) + sizeof(SelScal));
            DVS_CUDA_TRY(cudaMemcpyAsync(h_fi, d_ginteresting.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            DVS_CUDA_TRY(cudaStreamSynchronize(st));
            const unsigned fi = *h_fi;
            if (fi == kNone) {  // the whole window was certainly rejected
                cursor += W;
                grow_window = std::min(grow_window * 2u, kGrowWindowMax);
                continue;
            }
            grow_window = std::max(4u, std::min(kGrowWindowMax, 2u * (fi - cursor + 1u)));
            if (fi == cursor) {
                if (++filter_streak >= 2) filter_bypass = 1u << std::min(filter_streak - 2u, 3u);
            } else {
                filter_streak = 0;
            }
            cursor = fi;  // everything before it is certainly rejected; the host path below takes this one
            single_candidate = true;
            continue;
        }
        if (use_persist && (!grow_mode || n == max_size)) {
            // every remaining round in one cooperative launch (until done, or halted for the host)
            const bool use_sm = sm_ok && n <= kSmMaxN;
            const unsigned grid = use_sm ? sm_grid : persist_grid;
            k_sel_set_dev<<<1, 1, 0, st>>>(cur->sc.p, cursor, std::min(window, grid), grid, num, accepts,
                                           (unsigned)cur->which);
            DVS_LAUNCHED(ctx);
            const double* a_F = f->freqs.p;
            const double* a_H = f->entropy.p;
            uint64_t a_dim = dim;
            double *a_S0 = cur->Sbuf[0].p, *a_S1 = cur->Sbuf[1].p;
            unsigned *a_M0 = cur->membuf[0].p, *a_M1 = cur->membuf[1].p;
            uint8_t* a_mem = is_member.p;
            double *a_md = cur->mdelta.p, *a_mb = cur->mbound.p;
            SelScal* a_sc = cur->sc.p;
            const uint8_t* a_valid = f->valid.p;
            const unsigned* a_order = d_order.p;
            unsigned a_rounds = 0x7fffffffu;
            unsigned long long* a_trace = nullptr;
            DevBuf<unsigned long long> d_trace;
            const char* tr_env = getenv("DVS_SELECT_TRACE");
            const char* tra_env = getenv("DVS_SELECT_TRACE_ALL");  // raw u64 [256 rounds][8 slots][256 CTAs]
            int a_trace_all = (use_sm && tra_env && tra_env[0]) ? 1 : 0;
            const size_t trace_len = a_trace_all ? (size_t)256 * kSmTraceSlots * kSmMaxGrid : 2048;
            if (a_trace_all) tr_env = nullptr;
            if ((tr_env && tr_env[0]) || a_trace_all) {
                DVS_TRY(d_trace.alloc(trace_len));
                DVS_CUDA_TRY(cudaMemsetAsync(d_trace.p, 0, trace_len * sizeof(unsigned long long), st));
                a_trace = d_trace.p;
            }
            if (use_sm) {
                unsigned a_dim32 = (unsigned)dim;
                double* a_S = cur->S();
                unsigned* a_M = cur->members();
                SmPart *a_sp = d_spart.p, *a_up = d_upart.p, *a_dp = d_dpart.p;
                DVS_CUDA_TRY(cudaMemsetAsync(d_dpart.p, 0, 3 * sizeof(SmPart), st));
                // exchange tags restart at 1 in every launch
                DVS_CUDA_TRY(cudaMemsetAsync(d_spart.p, 0, 2 * kSmMaxGrid * sizeof(SmPart), st));
                DVS_CUDA_TRY(cudaMemsetAsync(d_upart.p, 0, (kSmMaxN + 1) * sizeof(SmPart), st));
                void* args[] = {&a_F, &a_H, &a_dim32, &a_S, &a_M, &a_mem, &a_md, &a_mb, &a_sc, &a_valid, &a_order,
                                &a_sp, &a_up, &a_dp, &a_trace, &a_trace_all};
                DVS_CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_sel_persist_sm, dim3(grid), dim3(kFastThreads), args,
                                                         sizeof(SmShared), st));
            } else {
                void* args[] = {&a_F, &a_H, &a_dim, &a_S0, &a_S1, &a_M0, &a_M1, &a_mem, &a_md, &a_mb, &a_sc, &a_valid,
                                &a_order, &a_rounds, &a_trace};
                DVS_CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_sel_persist, dim3(grid), dim3(kFastThreads), args, 0, st));
            }
            ctx->launches++;
            DVS_TRY(sel.read(*cur));
            if (a_trace && a_trace_all) {
                std::vector<unsigned long long> tr(trace_len);
                DVS_CUDA_TRY(cudaMemcpy(tr.data(), a_trace, trace_len * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                if (FILE* fp = fopen(tra_env, "wb")) {
                    fwrite(tr.data(), sizeof(unsigned long long), trace_len, fp);
                    fclose(fp);
                }
            } else if (a_trace) {  // phase timeline of CTA 0 to the file named by DVS_SELECT_TRACE
                std::vector<unsigned long long> tr(2048);
                DVS_CUDA_TRY(cudaMemcpy(tr.data(), a_trace, 2048 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                if (FILE* fp = fopen(tr_env, "a")) {
                    const int slots = use_sm ? kSmTraceSlots : 4, used = use_sm ? 7 : 4;
                    fprintf(fp, use_sm ? "# round scan_ns gather_ns decide_ns update_ns gather2_ns finalize_ns next_ns (CTA 0)\n"
                                       : "# round scan_ns wait1_ns update_ns wait2_ns (CTA 0; wait = grid barrier incl. the slowest CTA)\n");
                    for (int r = 0; (r + 1) * slots < 2048 && tr[slots * (r + 1)]; ++r) {
                        fprintf(fp, "%d", r);
                        for (int q = 0; q + 1 < used; ++q) fprintf(fp, " %llu", tr[slots * r + q + 1] - tr[slots * r + q]);
                        fprintf(fp, " %llu\n", tr[slots * (r + 1)] - tr[slots * r + used - 1]);
                    }
                    fclose(fp);
                }
            }
            const SelScal hd = *sel.h_sc;
            if (hd.panic) return panic_error(hd);
            cursor = hd.cursor;
            window = hd.window;
            accepts = hd.accepts;
            cur->which = (int)hd.which;
            if (!hd.halt) continue;  // finished
            DVS_TRY(sel.reset_scan(*cur));
            if (cursor >= num) break;
        } else if (use_dev && (!grow_mode || n == max_size)) {
            // Device-driven rounds: scan + decide + replace/update are enqueued kRoundsPerBatch times
            // without any read-back; the kernels carry cursor / window / accepts in the scalar block
            // and stop doing work once they halt (undecided within the error bound) or finish.
            k_sel_set_dev<<<1, 1, 0, st>>>(cur->sc.p, cursor, std::min(window, window_max_dev), window_max_dev, num,
                                           accepts, (unsigned)cur->which);
            DVS_LAUNCHED(ctx);
            for (int r = 0; r < kRoundsPerBatch; ++r) {
                k_sel_scan_dev<<<window_max_dev, kFastThreads, 0, st>>>(
                    f->freqs.p, f->entropy.p, dim, cur->Sbuf[0].p, cur->Sbuf[1].p, cur->membuf[0].p, cur->membuf[1].p,
                    cur->sc.p, f->valid.p, is_member.p, d_order.p);
                DVS_LAUNCHED(ctx);
                k_sel_round_dev<<<n + 1, kFastThreads, 0, st>>>(
                    f->freqs.p, f->entropy.p, dim, cur->Sbuf[0].p, cur->Sbuf[1].p, cur->membuf[0].p, cur->membuf[1].p,
                    is_member.p, cur->mdelta.p, cur->mbound.p, cur->sc.p, d_order.p);
                DVS_LAUNCHED(ctx);
            }
            DVS_TRY(sel.read(*cur));
            const SelScal hd = *sel.h_sc;
            if (hd.panic) return panic_error(hd);
            cursor = hd.cursor;
            window = hd.window;
            accepts = hd.accepts;
            cur->which = (int)hd.which;
            if (!hd.halt) continue;  // finished, or simply out of enqueued rounds
            // halted: one host-driven iteration below resolves the undecided state / candidate exactly
            DVS_TRY(sel.reset_scan(*cur));
            if (cursor >= num) break;
        }
        const unsigned count = single_candidate ? 1u : std::min(window, num - cursor);
        single_candidate = false;
        SelScal h;
        unsigned pos = kNone;
        if (use_fast) {
            DVS_TRY(sel.scan_fast(*cur, is_member.p, d_order.p, cursor, count));
            DVS_TRY(sel.read(*cur));
            h = *sel.h_sc;
            if (h.panic) return panic_error(h);
            if (h.state_unsure) {
                // the last fast update could not certify lowest_index / a sum check: redo it exactly and
                // re-score the window against the now certain state
                DVS_TRY(sel.update(*cur, n));
                ++exact_evals;
                continue;
            }
            if (h.first_unsure < h.first_true) {
                // an undecided candidate comes before the first certain acceptance: decide it exactly
                const unsigned pu = h.first_unsure;
                if (!h.exact) DVS_TRY(sel.update(*cur, n));
                DVS_TRY(sel.scan(*cur, f, is_member.p, d_order.p, pu, 1, nullptr));
                DVS_TRY(sel.read(*cur));
                h = *sel.h_sc;
                ++exact_evals;
                if (h.panic) return panic_error(h);
                if (h.first_panic == pu) {
                    set_error("cannot calculate entropy as frequency vector total !=1.0 (candidate at position %u)", pu);
                    return DVS_ERR_VALUE;
                }
                if (h.first_true != pu) {  // rejected by the exact evaluation: carry on after it
                    DVS_TRY(sel.reset_scan(*cur));
                    cursor = pu + 1;
                    continue;
                }
                pos = pu;
            } else if (h.first_true != kNone) {
                pos = h.first_true;
            }
        } else {
            DVS_TRY(sel.scan(*cur, f, is_member.p, d_order.p, cursor, count, nullptr));
            DVS_TRY(sel.read(*cur));
            h = *sel.h_sc;
            if (h.panic) return panic_error(h);
            if (h.first_panic != kNone && h.first_panic <= h.first_true) {
                set_error("cannot calculate entropy as frequency vector total !=1.0 (candidate at position %u)",
                          h.first_panic);
                return DVS_ERR_VALUE;
            }
            pos = h.first_true;
        }
        if (pos == kNone) {
            cursor += count;
            window = std::min(window * 2, window_max);
            continue;
        }
        const unsigned row = order[pos];
        window = std::max(64u, std::min(window_max, 2 * (pos - cursor + 1)));
        cursor = pos + 1;
        if (!grow_mode || n == max_size) {  // replace_lowest, records.rs:111-118
            DVS_TRY(use_fast ? sel.replace_fast(*cur, n, row, is_member.p) : sel.replace(*cur, n, row, is_member.p));
            ++accepts;
            continue;
        }
        // records.rs:434-451: nw = clone(); nw.push(rec); keep whichever has the larger statistic.
        // First try with the bounded-error kernels: the comparison is decided from them only when it
        // holds for every admissible error; otherwise both states are evaluated by the exact kernel.
        const bool use_cov = (mode == DVS_MODE_MAX_COV);
        int decided = -1;  // -1 undecided, 0 discard candidate, 1 adopt the grown set
        if (use_fast && !h.state_unsure) {
            DVS_TRY(sel.build_fast(*alt, cur->members(), n, (int)row));
            DVS_TRY(sel.read(*alt));
            const SelScal hf = *sel.h_sc;
            if (hf.panic) return panic_error(hf);
            if (!hf.state_unsure) {
                const double sa_f = use_cov ? h.cov : h.stdv, ba = h.exact ? 0.0 : (use_cov ? h.cov_bound : h.std_bound);
                const double sb_f = use_cov ? hf.cov : hf.stdv, bb = use_cov ? hf.cov_bound : hf.std_bound;
                if (sb_f - bb > sa_f + ba) decided = 1;
                else if (sb_f + bb < sa_f - ba) decided = 0;
            }
        }
        if (decided == 1) {
            std::swap(cur, alt);
            ++n;
            ++accepts;
            uint8_t one = 1;
            DVS_CUDA_TRY(cudaMemcpyAsync(is_member.p + row, &one, 1, cudaMemcpyHostToDevice, st));
            DVS_CUDA_TRY(cudaStreamSynchronize(st));
            continue;
        }
        if (decided == 0) {
            DVS_TRY(sel.reset_scan(*cur));
            continue;
        }
        if (use_fast) ++exact_evals;
        if (!h.exact) {
            DVS_TRY(sel.update(*cur, n));
            DVS_TRY(sel.read(*cur));
            h = *sel.h_sc;
            if (h.panic) return panic_error(h);
        }
        DVS_TRY(sel.build(*alt, cur->members(), n, (int)row));
        DVS_TRY(sel.read(*alt));
        const SelScal hb = *sel.h_sc;
        if (hb.panic) return panic_error(hb);
        const double sa = use_cov ? h.cov : h.stdv;
        const double sb = use_cov ? hb.cov : hb.stdv;
        if (sb > sa) {
            std::swap(cur, alt);
            ++n;
            ++accepts;
            uint8_t one = 1;
            DVS_CUDA_TRY(cudaMemcpyAsync(is_member.p + row, &one, 1, cudaMemcpyHostToDevice, st));
            DVS_CUDA_TRY(cudaStreamSynchronize(st));
        } else {
            DVS_TRY(sel.reset_scan(*cur));  // candidate discarded: re-arm the min-index reduction of `cur`
        }
    }
    DVS_TRY(sel.read(*cur));
    if (!sel.h_sc->exact || sel.h_sc->state_unsure) DVS_TRY(sel.update(*cur, n));  // reported numbers are exact
    ctx->last_exact_evals = exact_evals;
    DVS_TRY(sel.read(*cur));
    const SelScal h = *sel.h_sc;
    if (h.panic) return panic_error(h);
    if (sel_idx) DVS_CUDA_TRY(cudaMemcpyAsync(sel_idx, cur->members(), n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    if (sel_delta) DVS_CUDA_TRY(cudaMemcpyAsync(sel_delta, cur->mdelta.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    DVS_CUDA_TRY(cudaStreamSynchronize(st));
    if (stats5) {
        stats5[0] = h.total_jsd;
        stats5[1] = h.mean;
        stats5[2] = h.stdv;
        stats5[3] = h.cov;
        stats5[4] = h.E;
    }
    *size_out = n;
    ctx->last_accepts = accepts;
    return DVS_OK;
}

int dvs_debug_fast_terms(dvs_ctx* ctx, const double* a, const double* b, double* m, double* l, int32_t* special,
                         uint64_t n) {
    if (n == 0) return DVS_OK;
    DVS_CUDA_TRY(dvs::enter(ctx));
    DevBuf<double> da, db, dm, dl;
    DevBuf<int> ds;
    DVS_TRY(da.alloc(n));
    DVS_TRY(db.alloc(n));
    DVS_TRY(dm.alloc(n));
    DVS_TRY(dl.alloc(n));
    DVS_TRY(ds.alloc(n));
    DVS_CUDA_TRY(cudaMemcpyAsync(da.p, a, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DVS_CUDA_TRY(cudaMemcpyAsync(db.p, b, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_debug_fast_terms<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(da.p, db.p, dm.p, dl.p, ds.p, n);
    DVS_LAUNCHED(ctx);
    DVS_CUDA_TRY(cudaMemcpyAsync(m, dm.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DVS_CUDA_TRY(cudaMemcpyAsync(l, dl.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DVS_CUDA_TRY(cudaMemcpyAsync(special, ds.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DVS_OK;
}

uint32_t dvs_select_last_accepts(dvs_ctx* ctx) { return ctx->last_accepts; }
uint32_t dvs_select_last_exact_evals(dvs_ctx* ctx) { return ctx->last_exact_evals; }

int dvs_summed_create(dvs_ctx* ctx, const dvs_kfreqs* f, const uint32_t* members, uint32_t n, dvs_summed** out) {
    if (!ctx || !f || !out || (!members && n)) {
        set_error("dvs_summed_create: NULL argument");
        return DVS_ERR_ARG;
    }
    if (n == 0) {
        set_error("records cannot be empty");
        return DVS_ERR_VALUE;
    }
    if (n <= 1) {
        set_error("must have > 1 KmerSeq");
        return DVS_ERR_VALUE;
    }
    for (uint32_t i = 0; i < n; ++i)
        if (members[i] >= f->nrec) {
            set_error("dvs_summed_create: member %u out of range", members[i]);
            return DVS_ERR_ARG;
        }
    DVS_CUDA_TRY(dvs::enter(ctx));
    auto* s = new dvs_summed();
    s->device = ctx->device;
    s->f = f;
    s->n = n;
    int rc = s->st.alloc(f->dim, n + 1);
    DevBuf<unsigned> d_m;
    if (rc == DVS_OK) rc = d_m.alloc(n);
    if (rc == DVS_OK) {
        cudaError_t e = cudaMemcpyAsync(d_m.p, members, n * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            set_error("member upload failed: %s", cudaGetErrorString(e));
            rc = DVS_ERR_CUDA;
        }
    }
    Selector sel{ctx, f, f->dim, ctx->stream, (SelScal*)ctx->pinned};
    if (rc == DVS_OK) rc = sel.build(s->st, d_m.p, n, -1);
    if (rc == DVS_OK) rc = sel.read(s->st);
    if (rc == DVS_OK && sel.h_sc->panic) rc = panic_error(*sel.h_sc);
    if (rc != DVS_OK) {
        dvs_summed_free(s);
        return rc;
    }
    *out = s;
    return DVS_OK;
}

int dvs_summed_delta_jsd(dvs_ctx* ctx, dvs_summed* s, const dvs_kfreqs* q, uint32_t q_row, int is_member,
                         double* out) {
    if (!ctx || !s || !q || !out || q_row >= q->nrec || q->dim != s->f->dim) {
        set_error("dvs_summed_delta_jsd: bad argument");
        return DVS_ERR_ARG;
    }
    if (is_member) {  // records.rs:71-73
        *out = 0.0;
        return DVS_OK;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    Selector sel{ctx, s->f, s->f->dim, ctx->stream, (SelScal*)ctx->pinned};
    DevBuf<double> d_out;
    DVS_TRY(d_out.alloc(1));
    // a window of one candidate at position == row (order == nullptr); validity is the caller's business
    DVS_TRY(sel.scan_raw(s->st, q->freqs.p, q->entropy.p, q_row, d_out.p));
    DVS_CUDA_TRY(cudaMemcpyAsync(out, d_out.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DVS_TRY(sel.read(s->st));
    const SelScal h = *sel.h_sc;
    if (h.first_panic != kNone) {
        set_error("cannot calculate entropy as frequency vector total !=1.0");
        return DVS_ERR_VALUE;
    }
    return DVS_OK;
}

int dvs_summed_delta_jsd_batch(dvs_ctx* ctx, dvs_summed* s, const dvs_kfreqs* q, const uint8_t* is_member_or_null,
                               double* out) {
    if (!ctx || !s || !q || !out || q->dim != s->f->dim) {
        set_error("dvs_summed_delta_jsd_batch: bad argument");
        return DVS_ERR_ARG;
    }
    const uint32_t nq = q->nrec;
    if (nq == 0) return DVS_OK;
    DVS_CUDA_TRY(dvs::enter(ctx));
    Selector sel{ctx, s->f, s->f->dim, ctx->stream, (SelScal*)ctx->pinned};
    DevBuf<double> d_out;
    DevBuf<uint8_t> d_mem;
    DVS_TRY(d_out.alloc(nq));
    DVS_TRY(d_mem.alloc(nq));
    std::vector<uint8_t> h_mem(nq, 0), h_valid(nq);
    if (is_member_or_null)
        for (uint32_t i = 0; i < nq; ++i) h_mem[i] = is_member_or_null[i] ? 1 : 0;
    DVS_CUDA_TRY(cudaMemcpyAsync(d_mem.p, h_mem.data(), nq, cudaMemcpyHostToDevice, ctx->stream));
    DVS_CUDA_TRY(cudaMemsetAsync(d_out.p, 0, nq * sizeof(double), ctx->stream));
    DVS_TRY(sel.reset_scan(s->st));
    // one CTA per query row; position == row (order == nullptr); invalid / member rows are skipped
    k_sel_scan<<<nq, kEntThreads, kEntSmemBytes, ctx->stream>>>(s->f->freqs.p, s->f->entropy.p, s->f->dim, s->st.S(),
                                                                 s->st.members(), s->st.sc.p, q->freqs.p,
                                                                 q->entropy.p, q->valid.p, d_mem.p, nullptr, 0u,
                                                                 d_out.p);
    DVS_LAUNCHED(ctx);
    DVS_CUDA_TRY(cudaMemcpyAsync(out, d_out.p, nq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DVS_CUDA_TRY(cudaMemcpyAsync(h_valid.data(), q->valid.p, nq, cudaMemcpyDeviceToHost, ctx->stream));
    DVS_TRY(sel.read(s->st));
    const SelScal h = *sel.h_sc;
    DVS_TRY(sel.reset_scan(s->st));
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h.first_panic != kNone) {
        set_error("cannot calculate entropy as frequency vector total !=1.0 (query %u)", h.first_panic);
        return DVS_ERR_VALUE;
    }
    for (uint32_t i = 0; i < nq; ++i)
        if (!h_valid[i]) out[i] = std::numeric_limits<double>::quiet_NaN();
    return DVS_OK;
}

int dvs_summed_result(dvs_ctx* ctx, dvs_summed* s, uint32_t* sel_idx, double* sel_delta, double* stats5,
                      uint32_t* size_out, uint32_t* lowest_out) {
    DVS_CUDA_TRY(dvs::enter(ctx));
    Selector sel{ctx, s->f, s->f->dim, ctx->stream, (SelScal*)ctx->pinned};
    DVS_TRY(sel.read(s->st));
    const SelScal h = *sel.h_sc;
    if (sel_idx)
        DVS_CUDA_TRY(cudaMemcpyAsync(sel_idx, s->st.members(), s->n * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    if (sel_delta)
        DVS_CUDA_TRY(cudaMemcpyAsync(sel_delta, s->st.mdelta.p, s->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats5) {
        stats5[0] = h.total_jsd;
        stats5[1] = h.mean;
        stats5[2] = h.stdv;
        stats5[3] = h.cov;
        stats5[4] = h.E;
    }
    if (size_out) *size_out = s->n;
    if (lowest_out) *lowest_out = h.lowest;
    return DVS_OK;
}

void dvs_summed_free(dvs_summed* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    delete s;
}

}  // extern "C"
