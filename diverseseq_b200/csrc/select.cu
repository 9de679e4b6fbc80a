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
#include <string.h>

#include <algorithm>
#include <limits>

#include <cooperative_groups.h>

#include "comm.cuh"
#include "common.cuh"
#include "entropy.cuh"
#include "fused.cuh"

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
    unsigned limit;          // trailing rounds: positions below it had been published when the kernel returned
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

#include "select_shard.cuh"
#include "select_fast.cuh"
#include "select_grow.cuh"
#include "select_sm.cuh"

// waits (bounded: ~150 us) until `want` counting CTAs are resident, or the counting is over
__global__ void k_wait_resident(const unsigned* resident, unsigned want, const unsigned* ready, unsigned num) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        if (*reinterpret_cast<const volatile unsigned*>(resident) >= want) break;
        if (*reinterpret_cast<const volatile unsigned*>(ready) >= num) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 150000ull) break;
        __nanosleep(500);
    }
}

constexpr auto k_sel_persist_sm_full = k_sel_persist_sm_t<kFastThreads, 1, kSmMaxN, kSmChunk>;
// the trailing form: 128 threads x <= 64 registers and ~66 KB of shared memory fit beside one counting CTA
// (1,024 threads x 56 registers, 144 KB) on every SM
template <int NT>
struct SlimKernel {
    static constexpr auto fn = k_sel_persist_sm_t<NT, (1024 + NT - 1) / NT, kSmSlimMaxN, kSmSlimChunk>;  // (<= 64 registers)
};

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
    if (h.panic == 2u) {
        set_error("dvs_select_sharded: a peer GPU did not answer an exchange within 20 s (a rank is missing or died)");
        return DVS_ERR_CUDA;
    }
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

// comm == nullptr: one GPU.  Otherwise every rank calls this with the same rows (a kfreqs made by
// dvs_count_kmers_sharded / dvs_kfreqs_allgather), the same order and the same arguments: the state and
// every host-side decision are replicated (identical inputs, identical kernels), while the windows of
// candidates are scored candidate-sharded with an all-reduce(min) of the first interesting position.
static int select_core(dvs_ctx* ctx, dvs_comm* comm, const dvs_kfreqs* f, const uint32_t* order, uint32_t num, int mode,
                       uint32_t min_size, uint32_t max_size, uint32_t* sel_idx, double* sel_delta, double* stats5,
                       uint32_t* size_out, const TrailArgs* trail = nullptr) {
    if (!ctx || !f || (!order && num) || !size_out) {
        set_error("dvs_select: NULL argument");
        return DVS_ERR_ARG;
    }
    if (mode < DVS_MODE_NMOST || mode > DVS_MODE_MAX_COV) {
        set_error("dvs_select: bad mode %d", mode);
        return DVS_ERR_ARG;
    }
    ctx->last_accepts = 0;
    ctx->last_trail_accepts = 0;
    ctx->last_trail_launches = 0;
    ctx->last_trail_sms = 0xFFFFFFFFu;
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
    const unsigned world = comm ? (unsigned)comm->world : 1u, rank = comm ? (unsigned)comm->rank : 0u;
    ShardArgs shard;
    memset(&shard, 0, sizeof shard);
    shard.rank = (int)rank;
    shard.world = (int)world;
    for (unsigned r = 0; r < world && comm; ++r) shard.xbase[r] = comm->peer[r];
    if (world > 1) DVS_TRY(comm_barrier(ctx, comm));  // every rank is here: exchange slots of earlier calls are dead
    // trailing mode (dvs_count_select): the rows are still being produced on another stream; positions below
    // `trail_limit` have been published.  Start once the first chunk (>= min_size positions) is there.
    unsigned trail_limit = num;
    if (trail) {
        DVS_CUDA_TRY(cudaStreamWaitEvent(st, trail->first_ready, 0));
        trail_limit = std::min(trail->limit0, num);
        if (mode != DVS_MODE_NMOST || trail_limit < min_size) {  // nothing to gain / not enough rows yet: wait for all
            DVS_CUDA_TRY(cudaStreamWaitEvent(st, trail->count_done, 0));
            trail_limit = num;
        }
    }

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
    // (trailing mode: the records published so far; all of them once the counting has finished, below)
    for (uint32_t i = 0; i < trail_limit; ++i)
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

    const bool grow_mode = (mode != DVS_MODE_NMOST);
    // max modes with max_size below the initial size (e.g. max_size < min_size, which only the reference's
    // CLI rejects, cli.py:311): `size == max_size` never holds, so the set keeps growing as in
    // records.rs:427-451 - every buffer is then sized for all `num` records
    const bool unbounded = grow_mode && max_size < (uint32_t)init.size();
    const unsigned cap = (unbounded ? num : std::max<unsigned>(std::max(min_size, max_size), (unsigned)init.size())) + 1;
    SelState A, B;
    DVS_TRY(A.alloc(dim, cap));
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
    // DVS_SELECT_GRID: CTAs of the persistent kernels (tests run several ranks on ONE GPU: their kernels must be
    // co-resident, so each takes a share of the SMs)
    const char* grid_env = getenv("DVS_SELECT_GRID");
    const unsigned grid_cap = grid_env && atoi(grid_env) > 0 ? (unsigned)atoi(grid_env) : 0xFFFFFFFFu;
    const char* per_env = getenv("DVS_SELECT_PERSIST");
    int coop = 0, per_sm = 0, per_sm2 = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sel_persist, kFastThreads, 0) != cudaSuccess)
        per_sm = 0;
    const bool use_persist = use_dev && coop && per_sm > 0 && !(per_env && per_env[0] == '0');
    const unsigned persist_grid = std::min((unsigned)ctx->sm_count * (unsigned)std::min(per_sm, 2), std::max(1u, std::min(grid_cap, 0x7FFFFFFFu)));
    bool sm_ok = use_persist && !(per_env && per_env[0] == '1') && dim <= kSmMaxDim;
    if (sm_ok) {
        sm_ok = cudaFuncSetAttribute(k_sel_persist_sm_full, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(SmShared)) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_sel_persist_sm_full, kFastThreads,
                                                              sizeof(SmShared)) == cudaSuccess &&
                per_sm2 > 0;
        (void)cudaGetLastError();
    }
    const unsigned sm_grid = std::min(std::min<unsigned>((unsigned)ctx->sm_count, kSmMaxGrid), grid_cap);
    const int slim_threads = 128 * (trail_shape() + 1);
    const auto k_slim = slim_threads == 384 ? SlimKernel<384>::fn : (slim_threads == 256 ? SlimKernel<256>::fn : SlimKernel<128>::fn);
    bool slim_ok = false;
    if (trail && sm_ok) {
        slim_ok = cudaFuncSetAttribute(k_slim, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(SmSharedSlim)) == cudaSuccess &&
                  cudaFuncSetAttribute(k_slim, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared) == cudaSuccess;
        (void)cudaGetLastError();
    }
    DevBuf<SlicePart> d_slparts;  // k_sel_persist: partial sums of sliced vectors and their arrival tickets
    DevBuf<unsigned> d_slticks;
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
    const unsigned kGrowWindowMax = 64 * world;
    SelState fresh;  // clone(): the members re-summed in order (S, E), refreshed whenever the set changes
    DevBuf<FastSum> d_gparts;
    DevBuf<double> d_gmd, d_gmb;
    DevBuf<SelScal> d_gscratch;
    DevBuf<unsigned> d_ginteresting;
    bool fresh_valid = false;
    unsigned grow_window = 16;
    if (use_grow_batch && n != max_size) {
        DVS_TRY(fresh.alloc(dim, cap));
        DVS_TRY(d_gparts.alloc((size_t)kGrowWindowMax * (cap + 3)));
        DVS_TRY(d_gmd.alloc((size_t)kGrowWindowMax * cap));
        DVS_TRY(d_gmb.alloc((size_t)kGrowWindowMax * cap));
        DVS_TRY(d_gscratch.alloc(kGrowWindowMax));  // (indexed by the window offset)
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
        if (use_grow_batch && n != max_size && !single_candidate && filter_bypass > 0) {
            --filter_bypass;
            single_candidate = true;  // straight to the host path for this candidate
            continue;
        }
        if (use_grow_batch && n != max_size && !single_candidate) {
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
            // candidate-sharded: this GPU evaluates the window candidates c with (cursor + c) % world == rank
            const unsigned gc0 = (rank + world - cursor % world) % world;
            const unsigned Wloc = W > gc0 ? (W - gc0 + world - 1) / world : 0;
            if (Wloc) {
                k_grow_eval<<<dim3(n + 3, Wloc), kFastThreads, 0, st>>>(f->freqs.p, dim, cur->S(), cur->members(), cur->sc.p,
                                                                        fresh.S(), f->valid.p, is_member.p, d_order.p,
                                                                        cursor, d_gparts.p, gc0, world);
                DVS_LAUNCHED(ctx);
                k_grow_decide<<<Wloc, kFastThreads, 0, st>>>(f->entropy.p, dim, cur->members(), cur->sc.p, fresh.sc.p,
                                                             f->valid.p, is_member.p, d_order.p, cursor, d_gparts.p, d_gmd.p,
                                                             d_gmb.p, cap, d_gscratch.p, d_ginteresting.p,
                                                             mode == DVS_MODE_MAX_COV ? 1 : 0, gc0, world);
                DVS_LAUNCHED(ctx);
            }
            if (world > 1) DVS_TRY(comm_min_u32(ctx, comm, d_ginteresting.p, d_ginteresting.p));
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
        if (trail && trail_limit < num && !use_persist) {  // only the persistent kernel knows how to trail
            DVS_CUDA_TRY(cudaStreamWaitEvent(st, trail->count_done, 0));
            trail_limit = num;
        }
        if (use_persist && (!grow_mode || n == max_size)) {
            // every remaining round in one cooperative launch (until done, or halted for the host)
            const bool use_sm = sm_ok && n <= kSmMaxN;
            // trailing rounds need the slim SM-replicated kernel; anything else waits for the counting to finish
            const bool dedicated = trail_shape() == 3;  // the selection owns whole SMs (fused.cuh)
            const bool trailing = trail && trail_limit < num && use_sm && world == 1 &&
                                  (dedicated || (slim_ok && n <= kSmSlimMaxN));
            if (trail && trail_limit < num && !trailing) {
                DVS_CUDA_TRY(cudaStreamWaitEvent(st, trail->count_done, 0));
                trail_limit = num;
            }
            const unsigned grid = (trailing && dedicated) ? std::min(sm_grid, trail_sms(ctx->sm_count))
                                                           : (use_sm ? sm_grid : persist_grid);
            k_sel_set_dev<<<1, 1, 0, st>>>(cur->sc.p, cursor, std::min(window, grid * world), grid * world, num, accepts,
                                           (unsigned)cur->which);
            if (comm) {
                shard.tag_base = (unsigned)((++comm->sel_tag_base) << 20);
                DVS_TRY(comm_host_rendezvous(ctx, comm));  // (ranks sharing one GPU only)
            }
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
            unsigned a_rounds = world > 1 ? 0xFFFF0u : 0x7fffffffu;  // (exchange tags of one launch are 20 bits)
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
                const unsigned* a_ready = trailing ? trail->d_ready : nullptr;
                unsigned a_limit0 = trail_limit;
                void* args[] = {&a_F, &a_H, &a_dim32, &a_S, &a_M, &a_mem, &a_md, &a_mb, &a_sc, &a_valid, &a_order,
                                &a_sp, &a_up, &a_dp, &a_trace, &a_trace_all, &shard, &a_ready, &a_limit0};
                if (trailing && dedicated) {
                    if (!ctx->stream_hi) {
                        int lo_p = 0, hi_p = 0;
                        DVS_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
                        DVS_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->stream_hi, cudaStreamNonBlocking, hi_p));
                        DVS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_hi_in, cudaEventDisableTiming));
                        DVS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_hi_out, cudaEventDisableTiming));
                    }
                    // 512 threads x 128 registers: one CTA fills an SM's register file, so these CTAs take whole SMs -
                    // the ones the counting CTAs of the launch in flight leave when it ends (the higher stream priority
                    // puts them ahead of the next counting launch's CTAs)
                    DVS_CUDA_TRY(cudaEventRecord(ctx->ev_hi_in, st));
                    DVS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream_hi, ctx->ev_hi_in, 0));
                    k_sel_persist_sm_full<<<grid, kFastThreads, sizeof(SmShared), ctx->stream_hi>>>(
                        a_F, a_H, a_dim32, a_S, a_M, a_mem, a_md, a_mb, a_sc, a_valid, a_order, a_sp, a_up, a_dp, a_trace,
                        a_trace_all, shard, a_ready, a_limit0);
                    DVS_CUDA_TRY(cudaGetLastError());
                    DVS_CUDA_TRY(cudaEventRecord(ctx->ev_hi_out, ctx->stream_hi));
                    DVS_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_hi_out, 0));
                } else if (trailing) {
                    // a plain launch, made when every SM holds a counting CTA: beside it exactly one of these CTAs
                    // fits (registers), so they land one per SM.  Launched between two counting launches they
                    // would pack three to an SM and keep the counting off a third of the GPU.
                    k_wait_resident<<<1, 1, 0, st>>>(trail->d_resident, std::min<unsigned>(grid, (unsigned)ctx->sm_count),
                                                     trail->d_ready, num);
                    DVS_LAUNCHED(ctx);
                }
                if (trailing && dedicated) {
                } else if (trailing)
                    k_slim<<<grid, slim_threads, sizeof(SmSharedSlim), st>>>(
                        a_F, a_H, a_dim32, a_S, a_M, a_mem, a_md, a_mb, a_sc, a_valid, a_order, a_sp, a_up, a_dp, a_trace,
                        a_trace_all, shard, a_ready, a_limit0);
                else
                    DVS_CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_sel_persist_sm_full, dim3(grid), dim3(kFastThreads),
                                                             args, sizeof(SmShared), st));
            } else {
                // DVS_SELECT_SLICES (A/B measurements): 0 = one CTA per candidate / member slot, 1 = candidates of a short
                // window are cut into slices (default), 2 = member slots of the update too (measured slower: a slice
                // repeats the slot's set-up and pays a fence + ticket, 55 us against 33 us per update at k = 8, n = 100)
                const char* sl_env = getenv("DVS_SELECT_SLICES");
                unsigned a_slice_mode = sl_env ? (unsigned)atoi(sl_env) : 1u;
                if (!d_slparts.p && a_slice_mode) {
                    const size_t slots = std::max<size_t>(grid, (size_t)cap + 1);
                    DVS_TRY(d_slparts.alloc(slots * kMaxSlices));
                    DVS_TRY(d_slticks.alloc(slots));
                    DVS_CUDA_TRY(cudaMemsetAsync(d_slticks.p, 0, slots * sizeof(unsigned), st));
                }
                SlicePart* a_parts = d_slparts.p;
                unsigned* a_ticks = d_slticks.p;
                void* args[] = {&a_F, &a_H, &a_dim, &a_S0, &a_S1, &a_M0, &a_M1, &a_mem, &a_md, &a_mb, &a_sc, &a_valid,
                                &a_order, &a_rounds, &a_trace, &shard, &a_parts, &a_ticks, &a_slice_mode};
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
            if (trailing) {
                ctx->last_trail_accepts += hd.accepts - accepts;
                ++ctx->last_trail_launches;
                std::vector<unsigned> smid(grid);
                DVS_CUDA_TRY(cudaMemcpyAsync(smid.data(), trail->d_ready + kTrailSmidOff, grid * sizeof(unsigned),
                                             cudaMemcpyDeviceToHost, st));
                DVS_CUDA_TRY(cudaStreamSynchronize(st));
                std::sort(smid.begin(), smid.end());
                const unsigned distinct = (unsigned)(std::unique(smid.begin(), smid.end()) - smid.begin());
                ctx->last_trail_sms = std::min(ctx->last_trail_sms, distinct);
            }
            cursor = hd.cursor;
            window = hd.window;
            accepts = hd.accepts;
            cur->which = (int)hd.which;
            if (trailing) trail_limit = std::max(trail_limit, std::min(hd.limit, num));
            if (!hd.halt) continue;  // finished (or, trailing, everything is published: the stand-alone kernel goes on)
            DVS_TRY(sel.reset_scan(*cur));
            if (cursor >= num) break;
        } else if (use_dev && world == 1 && (!grow_mode || n == max_size)) {
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
        if (cursor >= trail_limit) {  // (trailing) nothing published beyond the cursor: wait for all of it
            DVS_CUDA_TRY(cudaStreamWaitEvent(st, trail->count_done, 0));
            trail_limit = num;
        }
        const unsigned count = single_candidate ? 1u : std::min(window, trail_limit - cursor);
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
    if (trail) {  // the records that were not published when the selection started: the reference's panic check
        DVS_CUDA_TRY(cudaStreamWaitEvent(st, trail->count_done, 0));
        DVS_CUDA_TRY(cudaMemcpyAsync(valid.data(), f->valid.p, f->nrec, cudaMemcpyDeviceToHost, st));
        DVS_CUDA_TRY(cudaMemcpyAsync(err.data(), f->err.p, f->nrec, cudaMemcpyDeviceToHost, st));
        DVS_CUDA_TRY(cudaMemcpyAsync(err_total.data(), f->err_total.p, f->nrec * sizeof(double), cudaMemcpyDeviceToHost, st));
        DVS_CUDA_TRY(cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < num; ++i)
            if (valid[order[i]] && err[order[i]]) {
                set_error("cannot calculate entropy as frequency vector total %.17g!=1.0", err_total[order[i]]);
                return DVS_ERR_VALUE;
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
    if (world > 1) DVS_TRY(comm_check_error(ctx, comm, "dvs_select_sharded"));
    return DVS_OK;
}

}  // extern "C"
namespace dvs {
int select_with_trail(dvs_ctx* ctx, const dvs_kfreqs* f, const uint32_t* order, uint32_t num, int mode, uint32_t min_size,
                      uint32_t max_size, uint32_t* sel_idx, double* sel_delta, double* stats5, uint32_t* size_out,
                      const TrailArgs* trail) {
    return select_core(ctx, nullptr, f, order, num, mode, min_size, max_size, sel_idx, sel_delta, stats5, size_out, trail);
}
}  // namespace dvs
extern "C" {

int dvs_select(dvs_ctx* ctx, const dvs_kfreqs* f, const uint32_t* order, uint32_t num, int mode, uint32_t min_size,
               uint32_t max_size, uint32_t* sel_idx, double* sel_delta, double* stats5, uint32_t* size_out) {
    return select_core(ctx, nullptr, f, order, num, mode, min_size, max_size, sel_idx, sel_delta, stats5, size_out);
}

int dvs_select_sharded(dvs_ctx* ctx, dvs_comm* comm, const dvs_kfreqs* f_all, const uint32_t* order, uint32_t num,
                       int mode, uint32_t min_size, uint32_t max_size, uint32_t* sel_idx, double* sel_delta,
                       double* stats5, uint32_t* size_out) {
    if (!comm || !comm->connected) {
        set_error("dvs_select_sharded: the communicator is not connected");
        return DVS_ERR_ARG;
    }
    return select_core(ctx, comm->world > 1 ? comm : nullptr, f_all, order, num, mode, min_size, max_size, sel_idx,
                       sel_delta, stats5, size_out);
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
uint32_t dvs_select_last_trail_accepts(dvs_ctx* ctx) { return ctx->last_trail_accepts; }
uint32_t dvs_select_last_trail_launches(dvs_ctx* ctx) { return ctx->last_trail_launches; }
uint32_t dvs_select_last_trail_sms(dvs_ctx* ctx) { return ctx->last_trail_launches ? ctx->last_trail_sms : 0; }

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
