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

#include <algorithm>

#include "common.cuh"
#include "entropy.cuh"

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
};

struct SelState {
    DevBuf<double> S;          // summed_kfreqs [dim]
    DevBuf<unsigned> members;  // row indices in Vec order [cap]
    DevBuf<double> mdelta;     // delta_jsd per member [cap]
    DevBuf<SelScal> sc;
    unsigned cap = 0;
    int alloc(uint64_t dim, unsigned capacity) {
        cap = capacity;
        DVS_TRY(S.alloc(dim));
        DVS_TRY(members.alloc(capacity));
        DVS_TRY(mdelta.alloc(capacity));
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
    if (s_last && threadIdx.x == 0) {
        const unsigned n = sc->n;
        double e = __dsub_rn(sc->E, H[low_row]);
        sc->E = __dadd_rn(e, H[cand_row]);
        for (unsigned j = low; j + 1 < n; ++j) members[j] = members[j + 1];
        members[n - 1] = cand_row;
        if (is_member) {
            is_member[low_row] = 0;
            is_member[cand_row] = 1;
        }
        sc->ticket = 0;
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
    extern __shared__ double ent_smem[];
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
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        volatile double* md = mdelta;
        const double total_jsd = *(volatile double*)&sc->total_jsd;
        double mn = 1e6;
        unsigned low = 0;
        double sum = 0.0;
        for (unsigned t = 0; t < n; ++t) {
            const double d = __dsub_rn(total_jsd, md[t]);
            md[t] = d;
            if (d < mn) {
                mn = d;
                low = t;
            }
            sum = __dadd_rn(sum, d);
        }
        const double mean = __ddiv_rn(sum, nd);
        double ss = 0.0;
        for (unsigned t = 0; t < n; ++t) {
            const double d = __dsub_rn(md[t], mean);
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
    }
}

// increases_jsd for a window of candidates (records.rs:70-92); one CTA per position
__global__ void __launch_bounds__(kEntThreads)
k_sel_scan(const double* __restrict__ F, const double* __restrict__ H, uint64_t dim, const double* __restrict__ S,
           const unsigned* __restrict__ members, SelScal* sc, const double* __restrict__ candF,
           const double* __restrict__ candH, const uint8_t* __restrict__ cand_valid,
           const uint8_t* __restrict__ is_member, const unsigned* __restrict__ order, unsigned pos0,
           double* __restrict__ delta_out) {
    extern __shared__ double ent_smem[];
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
        return DVS_OK;
    }
    // one candidate row of another kfreqs object, no order / validity / membership filters
    int scan_raw(SelState& s, const double* candF, const double* candH, unsigned row, double* delta_out) {
        DVS_TRY(reset_scan(s));
        k_sel_scan<<<1, kEntThreads, kEntSmemBytes, st>>>(f->freqs.p, f->entropy.p, dim, s.S.p, s.members.p, s.sc.p,
                                                           candF, candH, nullptr, nullptr, nullptr, row, delta_out);
        DVS_LAUNCHED(ctx);
        return DVS_OK;
    }

    // SummedRecords::new over `members` (+ optional pushed row): sums, then total_jsd + member deltas
    int build(SelState& s, const unsigned* d_members, unsigned n, int extra_row) {
        k_sel_sum<<<vec_grid(), 256, 0, st>>>(f->freqs.p, f->entropy.p, dim, d_members, n, extra_row, s.S.p,
                                              s.members.p, s.sc.p);
        DVS_LAUNCHED(ctx);
        return update(s, n + (extra_row >= 0 ? 1u : 0u));
    }
    int update(SelState& s, unsigned n) {
        k_sel_update<<<n + 1, kEntThreads, kEntSmemBytes, st>>>(f->freqs.p, f->entropy.p, dim, s.S.p, s.members.p,
                                                                 s.mdelta.p, s.sc.p);
        DVS_LAUNCHED(ctx);
        return DVS_OK;
    }
    int replace(SelState& s, unsigned n, unsigned cand_row, uint8_t* is_member) {
        k_sel_replace_vec<<<vec_grid(), 256, 0, st>>>(f->freqs.p, f->entropy.p, dim, s.members.p, is_member,
                                                      s.sc.p, cand_row, s.S.p);
        DVS_LAUNCHED(ctx);
        return update(s, n);
    }
    int scan(SelState& s, const dvs_kfreqs* q, const uint8_t* is_member, const unsigned* d_order, unsigned pos0,
             unsigned count, double* delta_out) {
        DVS_TRY(reset_scan(s));  // re-arm the min-index reduction
        k_sel_scan<<<count, kEntThreads, kEntSmemBytes, st>>>(f->freqs.p, f->entropy.p, dim, s.S.p, s.members.p,
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
    while (cursor < num) {
        const unsigned count = std::min(window, num - cursor);
        DVS_TRY(sel.scan(*cur, f, is_member.p, d_order.p, cursor, count, nullptr));
        DVS_TRY(sel.read(*cur));
        const SelScal h = *sel.h_sc;
        if (h.panic) return panic_error(h);
        if (h.first_panic != kNone && h.first_panic <= h.first_true) {
            set_error("cannot calculate entropy as frequency vector total !=1.0 (candidate at position %u)",
                      h.first_panic);
            return DVS_ERR_VALUE;
        }
        if (h.first_true == kNone) {
            cursor += count;
            window = std::min(window * 2, window_max);
            continue;
        }
        const unsigned pos = h.first_true;
        const unsigned row = order[pos];
        window = std::max(64u, std::min(window_max, 2 * (pos - cursor + 1)));
        cursor = pos + 1;
        if (!grow_mode || n == max_size) {
            DVS_TRY(sel.replace(*cur, n, row, is_member.p));  // replace_lowest, records.rs:111-118
            ++accepts;
            continue;
        }
        // records.rs:434-451: nw = clone(); nw.push(rec); keep whichever has the larger statistic
        DVS_TRY(sel.build(*alt, cur->members.p, n, (int)row));
        DVS_TRY(sel.read(*alt));
        const SelScal hb = *sel.h_sc;
        if (hb.panic) return panic_error(hb);
        const double sa = (mode == DVS_MODE_MAX_COV) ? h.cov : h.stdv;
        const double sb = (mode == DVS_MODE_MAX_COV) ? hb.cov : hb.stdv;
        if (sb > sa) {
            std::swap(cur, alt);
            ++n;
            ++accepts;
            uint8_t one = 1;
            DVS_CUDA_TRY(cudaMemcpyAsync(is_member.p + row, &one, 1, cudaMemcpyHostToDevice, st));
            DVS_CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    DVS_TRY(sel.read(*cur));
    const SelScal h = *sel.h_sc;
    if (h.panic) return panic_error(h);
    if (sel_idx) DVS_CUDA_TRY(cudaMemcpyAsync(sel_idx, cur->members.p, n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
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

uint32_t dvs_select_last_accepts(dvs_ctx* ctx) { return ctx->last_accepts; }

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

int dvs_summed_result(dvs_ctx* ctx, dvs_summed* s, uint32_t* sel_idx, double* sel_delta, double* stats5,
                      uint32_t* size_out, uint32_t* lowest_out) {
    DVS_CUDA_TRY(dvs::enter(ctx));
    Selector sel{ctx, s->f, s->f->dim, ctx->stream, (SelScal*)ctx->pinned};
    DVS_TRY(sel.read(s->st));
    const SelScal h = *sel.h_sc;
    if (sel_idx)
        DVS_CUDA_TRY(cudaMemcpyAsync(sel_idx, s->st.members.p, s->n * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
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
