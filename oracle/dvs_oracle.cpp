// dvs_oracle — CPU restatement of the diverse-seq hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (diverseseq_b200/) never does.
//
// The reference's hot path is Rust (+ a little Python) and cannot be built in this
// image (no cargo/rustc, no cogent3), so this file restates its arithmetic in C++17,
// function by function, with the reference file:line each one follows.  Build with
// `-O2 -ffp-contract=off` (Rust never contracts a*b+c) against glibc (`std::log2` is
// what Rust's `f64::log2` calls on Linux).  Parity is PINNED: tests/test_oracle_golden.py
// checks every exact known-answer value in the reference's own Rust unit tests
// (src/record.rs:276-351, src/records.rs:602-621,676-685,726-740, src/distance.rs:186-191).
// Unpinned by the reference (no exact test exists upstream): murmur hash values, sketch
// contents, mash intersections, `max` membership — see DESIGN.md §oracle.
//
// All f64 sums are left-to-right sequential exactly as the Rust iterators evaluate them.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace {

constexpr double EPS = std::numeric_limits<double>::epsilon();  // f64::EPSILON

thread_local std::string g_err;

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------- record.rs ----

// src/record.rs:10-15  coord_conversion_coeffs
std::vector<size_t> coord_conversion_coeffs(size_t num_states, size_t k) {
    std::vector<size_t> c(k);
    for (size_t j = 0; j < k; ++j) {
        size_t e = k - 1 - j, v = 1;
        for (size_t t = 0; t < e; ++t) v *= num_states;
        c[j] = v;
    }
    return c;
}

// src/record.rs:18-29  kmer_to_index
size_t kmer_to_index(const uint8_t* kmer, size_t k, size_t num_states, const std::vector<size_t>& coeffs,
                     size_t max_index) {
    size_t index = 0;
    for (size_t i = 0; i < k; ++i) {
        if ((size_t)kmer[i] >= num_states) {
            index = max_index;
            break;
        }
        index += coeffs[i] * (size_t)kmer[i];
    }
    return index;
}

// src/record.rs:31-39  count_monomers
void count_monomers(const uint8_t* seq, size_t len, size_t num_states, uint64_t* counts) {
    for (size_t i = 0; i < num_states; ++i) counts[i] = 0;
    for (size_t i = 0; i < len; ++i)
        if (seq[i] < (uint8_t)num_states) counts[seq[i]] += 1;
}

// src/record.rs:41-84  count_kmers — literal transliteration incl. skip_until logic
void count_kmers(const uint8_t* seq, size_t len, size_t num_states, size_t k, uint64_t* counts) {
    auto coeffs = coord_conversion_coeffs(num_states, k);
    size_t size = 1;
    for (size_t t = 0; t < k; ++t) size *= num_states;
    for (size_t i = 0; i < size; ++i) counts[i] = 0;
    size_t skip_until = 0;
    for (size_t i = 0; i < std::min(k, len); ++i)
        if ((size_t)seq[i] >= num_states) skip_until = i + 1;
    int64_t index = -1;
    uint8_t nstates = (uint8_t)num_states;
    int64_t biggest_coeff = (int64_t)coeffs[0];
    if (len < k) return;  // seq.windows(k) is empty
    for (size_t i = 0; i + k <= len; ++i) {
        uint8_t gained = seq[i + k - 1];
        if (gained >= nstates) {
            index = -1;
            skip_until = i + k;
        }
        if (i < skip_until) continue;
        if (index < 0) {
            index = (int64_t)kmer_to_index(seq + i, k, num_states, coeffs, size - 1);
        } else {
            int64_t dropped = (int64_t)seq[i - 1];
            index = (index - dropped * biggest_coeff) * (int64_t)num_states + (int64_t)gained;
        }
        if (index < 0) continue;
        counts[(size_t)index] += 1;
    }
}

// src/record.rs:124-131  SeqRecord::to_kcounts
void to_kcounts(const uint8_t* seq, size_t len, size_t num_states, size_t k, uint64_t* counts) {
    if (k == 0) throw Panic("k cannot be 0");
    if (k == 1)
        count_monomers(seq, len, num_states, counts);
    else
        count_kmers(seq, len, num_states, k, counts);
}

// src/record.rs:86-106  entropy
double entropy(const double* f, size_t n) {
    if (n == 0) throw Panic("cannot calculate entropy as frequency vector empty");
    double e = 0.0, total = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double x = f[i];
        if (x == 0.0) continue;
        e += -x * std::log2(x);
        total += x;
    }
    double tol = (double)n * EPS;
    if (std::fabs(total - 1.0) > tol) {
        char buf[128];
        snprintf(buf, sizeof buf, "cannot calculate entropy as frequency vector total %.17g!=1.0", total);
        throw Panic(buf);
    }
    return e;
}

size_t ipow(size_t b, size_t e) {
    size_t v = 1;
    for (size_t t = 0; t < e; ++t) v *= b;
    return v;
}

// src/record.rs:145-188  KmerSeq
struct KmerSeq {
    int64_t id;  // stands in for seqid (identity only)
    std::vector<double> kfreqs;
    double entropy;
    mutable double delta_jsd = 0.0;
};

// src/record.rs:133-141  SeqRecord::to_kmerseq  (nullopt == Err("No valid k-mers ..."))
std::optional<KmerSeq> to_kmerseq(int64_t id, const uint8_t* seq, size_t len, size_t num_states, size_t k) {
    size_t d = (k == 1) ? num_states : ipow(num_states, k);
    std::vector<uint64_t> c(d);
    to_kcounts(seq, len, num_states, k, c.data());
    uint64_t tot = 0;
    for (auto v : c) tot += v;
    double total = (double)tot;
    if (total == 0.0) return std::nullopt;
    KmerSeq ks;
    ks.id = id;
    ks.kfreqs.resize(d);
    for (size_t i = 0; i < d; ++i) ks.kfreqs[i] = (double)c[i] / total;
    ks.entropy = entropy(ks.kfreqs.data(), d);  // KmerSeq::new, record.rs:157-168
    return ks;
}

// --------------------------------------------------------------- records.rs ----

// src/records.rs:276-286  updated_mean_freqs
void updated_mean_freqs(std::vector<double>& dest, const std::vector<double>& total, const std::vector<double>& rec,
                        double div) {
    for (size_t i = 0; i < dest.size(); ++i) {
        dest[i] = (total[i] - rec[i]) / div;
        if (dest[i] <= EPS) dest[i] = 0.0;
    }
}

// src/records.rs:220-252  get_lowest_record_index
uint32_t get_lowest_record_index(const std::vector<KmerSeq>& records, size_t num_kmers,
                                 const std::vector<double>& summed_kfreqs, double summed_entropies,
                                 double total_jsd) {
    double div = (double)records.size() - 1.0;
    if (div <= 0.0) throw Panic("must have > 1 KmerSeq");
    double min_delta = 1e6;
    uint32_t lowest = 0;
    std::vector<double> mean(num_kmers, 0.0);
    for (size_t i = 0; i < records.size(); ++i) {
        const KmerSeq& r = records[i];
        double mean_entropy = (summed_entropies - r.entropy) / div;
        updated_mean_freqs(mean, summed_kfreqs, r.kfreqs, div);
        double entropy_of_mean = entropy(mean.data(), num_kmers);
        double jsd = entropy_of_mean - mean_entropy;
        r.delta_jsd = total_jsd - jsd;
        if (r.delta_jsd < min_delta) {
            min_delta = r.delta_jsd;
            lowest = (uint32_t)i;
        }
    }
    return lowest;
}

// src/records.rs:10-217  SummedRecords
struct SummedRecords {
    std::vector<KmerSeq> records;
    uint32_t size = 0;
    std::vector<double> summed_kfreqs;
    double summed_entropies = 0.0;
    double total_jsd = 0.0;
    uint32_t lowest_index = 0;
    std::unordered_set<int64_t> seqids;

    // :27-68
    explicit SummedRecords(std::vector<KmerSeq> recs) : records(std::move(recs)) {
        if (records.empty()) throw Panic("records cannot be empty");
        size = (uint32_t)records.size();
        size_t num_kmers = records[0].kfreqs.size();
        summed_kfreqs.assign(num_kmers, 0.0);
        for (auto& r : records) {
            if (r.kfreqs.size() != num_kmers) throw Panic("length mismatch for add_vectors");
            for (size_t j = 0; j < num_kmers; ++j) summed_kfreqs[j] += r.kfreqs[j];
            summed_entropies += r.entropy;
        }
        std::vector<double> mean(num_kmers);
        for (size_t j = 0; j < num_kmers; ++j) mean[j] = summed_kfreqs[j] / (double)size;
        total_jsd = entropy(mean.data(), num_kmers) - summed_entropies / (double)size;
        for (auto& r : records) seqids.insert(r.id);
        lowest_index = get_lowest_record_index(records, num_kmers, summed_kfreqs, summed_entropies, total_jsd);
    }

    // :70-84
    double delta_jsd(const KmerSeq& rec) const {
        if (seqids.count(rec.id)) return 0.0;
        const KmerSeq& low = records[lowest_index];
        size_t n = low.kfreqs.size();
        std::vector<double> mean(n, 0.0);
        double mean_entropy = (summed_entropies - low.entropy + rec.entropy) / (double)size;
        for (size_t i = 0; i < n; ++i)
            mean[i] = (summed_kfreqs[i] - low.kfreqs[i] + rec.kfreqs[i]) / (double)size;
        double entropy_of_mean = entropy(mean.data(), n);
        return entropy_of_mean - mean_entropy;
    }

    // :86-92
    bool increases_jsd(const KmerSeq& rec) const {
        if (seqids.count(rec.id)) return false;
        double jsd = delta_jsd(rec);
        return jsd > total_jsd + EPS;
    }

    // :94-109
    void drop_lowest() {
        KmerSeq old = std::move(records[lowest_index]);
        records.erase(records.begin() + lowest_index);
        seqids.erase(old.id);
        summed_entropies -= old.entropy;
        for (size_t i = 0; i < old.kfreqs.size(); ++i) {
            summed_kfreqs[i] -= old.kfreqs[i];
            if (summed_kfreqs[i] <= EPS) summed_kfreqs[i] = 0.0;
        }
    }

    // :120-147
    void push(KmerSeq rec) {
        if (seqids.count(rec.id)) return;
        size_t num_kmers = records[0].kfreqs.size();
        seqids.insert(rec.id);
        summed_entropies += rec.entropy;
        for (size_t i = 0; i < num_kmers; ++i) summed_kfreqs[i] += rec.kfreqs[i];
        records.push_back(std::move(rec));
        size = (uint32_t)records.size();
        std::vector<double> mean(num_kmers);
        for (size_t j = 0; j < num_kmers; ++j) mean[j] = summed_kfreqs[j] / (double)size;
        double mean_entropy = entropy(mean.data(), num_kmers);
        total_jsd = mean_entropy - summed_entropies / (double)size;
        lowest_index = get_lowest_record_index(records, num_kmers, summed_kfreqs, summed_entropies, total_jsd);
    }

    // :111-118
    void replace_lowest(KmerSeq rec) {
        if (seqids.count(rec.id)) return;
        drop_lowest();
        push(std::move(rec));
    }

    // :153-172
    double mean_delta_jsd() const {
        double s = 0.0;
        for (auto& r : records) s += r.delta_jsd;
        return s / (double)size;
    }
    double std_delta_jsd() const {
        double mean = mean_delta_jsd();
        double sum = 0.0;
        for (auto& r : records) {
            double d = r.delta_jsd - mean;
            sum += d * d;  // powi(2)
        }
        return std::sqrt(sum / ((double)size - 1.0));
    }
    double cov_delta_jsd() const { return std_delta_jsd() / mean_delta_jsd(); }

    // :182-189  clone() == rebuild from scratch with delta_jsd reset
    SummedRecords clone() const {
        std::vector<KmerSeq> recs;
        recs.reserve(records.size());
        for (auto& r : records) {
            KmerSeq c{r.id, r.kfreqs, r.entropy};
            c.delta_jsd = 0.0;
            recs.push_back(std::move(c));
        }
        return SummedRecords(std::move(recs));
    }
};

using RowProvider = std::function<std::optional<KmerSeq>(size_t)>;  // position in order -> row or Err

// src/records.rs:311-342  select_nmost_divergent  (and :363-382 with a never-failing provider)
SummedRecords select_nmost(size_t num, size_t n, const RowProvider& get, std::vector<int64_t>* trace) {
    if (num < n) {
        char buf[96];
        snprintf(buf, sizeof buf, "The number of sequences %zu is < n %zu", num, n);
        throw Panic(buf);
    }
    std::vector<KmerSeq> init;
    for (size_t i = 0; i < n; ++i) {
        auto r = get(i);
        if (r) init.push_back(std::move(*r));
    }
    SummedRecords summed(std::move(init));
    for (size_t i = n; i < num; ++i) {
        auto r = get(i);
        if (!r) continue;
        if (summed.increases_jsd(*r)) {
            if (trace) trace->push_back((int64_t)i);
            summed.replace_lowest(std::move(*r));
        }
    }
    return summed;
}

// src/records.rs:390-454  select_max_divergent  (and :456-507)
SummedRecords select_max(size_t num, size_t min_size, size_t max_size, bool use_cov, const RowProvider& get,
                         std::vector<int64_t>* trace) {
    if (num < min_size) {
        char buf[96];
        snprintf(buf, sizeof buf, "The number of sequences %zu is < n %zu", num, min_size);
        throw Panic(buf);
    }
    if (!(num > max_size)) max_size = num;
    std::vector<KmerSeq> init;
    for (size_t i = 0; i < min_size; ++i) {
        auto r = get(i);
        if (r) init.push_back(std::move(*r));
    }
    SummedRecords summed(std::move(init));
    for (size_t i = min_size; i < num; ++i) {
        auto r = get(i);
        if (!r) continue;
        if (!summed.increases_jsd(*r)) continue;
        if (summed.size == (uint32_t)max_size) {
            if (trace) trace->push_back((int64_t)i);
            summed.replace_lowest(std::move(*r));
            continue;
        }
        SummedRecords nw = summed.clone();
        nw.push(std::move(*r));
        bool better = use_cov ? (nw.cov_delta_jsd() > summed.cov_delta_jsd())
                              : (nw.std_delta_jsd() > summed.std_delta_jsd());
        if (better) {
            if (trace) trace->push_back(-(int64_t)i - 1);  // negative == grew
            summed = std::move(nw);
        }
    }
    return summed;
}

// -------------------------------------------------------------- distance.rs ----

// src/distance.rs:21-49  murmurhash3_32 (non-standard: every byte is a 32-bit block)
uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
uint32_t murmurhash3_32(const uint8_t* data, size_t len, uint32_t seed) {
    if (seed == 0) seed = 0x9747B28Cu;
    uint32_t h = seed ^ (uint32_t)len;
    for (size_t i = 0; i < len; ++i) {
        uint32_t k = data[i];
        k *= 0xCC9E2D51u;
        k = rotl32(k, 15);
        k *= 0x1B873593u;
        h ^= k;
        h = rotl32(h, 13);
        h = h * 5u + 0xE6546B64u;
    }
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

// src/distance.rs:65-87  hash_kmer
uint32_t hash_kmer(const uint8_t* kmer, size_t k, bool canonical) {
    if (canonical) {
        std::vector<uint8_t> rev(k);  // :17-19 reverse_complement
        for (size_t i = 0; i < k; ++i) rev[k - 1 - i] = (uint8_t)((kmer[i] + 2) % 4);
        for (size_t i = 0; i < k; ++i) {
            if (kmer[i] < rev[i]) break;
            if (kmer[i] > rev[i]) return murmurhash3_32(rev.data(), k, 0);
        }
    }
    return murmurhash3_32(kmer, k, 0);
}

// src/distance.rs:101-134  get_kmer_hashes
std::vector<uint32_t> get_kmer_hashes(const uint8_t* seq, size_t len, size_t k, uint8_t num_states, bool canonical) {
    std::vector<uint32_t> out;
    if (len < k) return out;
    out.reserve(len - k + 1);
    size_t skip_until = 0;
    for (size_t i = 0; i < k; ++i)
        if (seq[i] >= num_states) skip_until = i + 1;
    for (size_t i = 0; i + k <= len; ++i) {
        if (seq[i + k - 1] >= num_states) skip_until = i + k;
        if (i < skip_until) continue;
        out.push_back(hash_kmer(seq + i, k, canonical));
    }
    return out;
}

// src/distance.rs:151-182  mash_sketch  (set -> max-heap bottom-s -> ascending)
std::vector<uint32_t> mash_sketch(const uint8_t* seq, size_t len, size_t k, size_t sketch_size, uint8_t num_states,
                                  bool canonical) {
    auto hashes = get_kmer_hashes(seq, len, k, num_states, canonical);
    std::unordered_set<uint32_t> uniq(hashes.begin(), hashes.end());
    std::priority_queue<uint32_t> heap;
    for (uint32_t h : uniq) {
        if (heap.size() < sketch_size) {
            heap.push(h);
        } else if (!heap.empty() && h < heap.top()) {
            heap.pop();
            heap.push(h);
        }
    }
    std::vector<uint32_t> result;
    result.reserve(heap.size());
    while (!heap.empty()) {
        result.push_back(heap.top());
        heap.pop();
    }
    std::sort(result.begin(), result.end());
    return result;
}

// diverse_seq/distance.py:230-291  mash_distance
double mash_distance(const uint32_t* a, size_t la, const uint32_t* b, size_t lb, int k, uint64_t sketch_size,
                     uint64_t* inter_out, uint64_t* union_out) {
    uint64_t inter = 0, uni = 0;
    size_t i = 0, j = 0;
    while (uni < sketch_size && i < la && j < lb) {
        uint32_t l = a[i], r = b[j];
        if (l < r)
            ++i;
        else if (r < l)
            ++j;
        else {
            ++i;
            ++j;
            ++inter;
        }
        ++uni;
    }
    if (uni < sketch_size) {
        if (i < la) uni += la - i;
        if (j < lb) uni += lb - j;
        uni = std::min<uint64_t>(uni, sketch_size);
    }
    if (inter_out) *inter_out = inter;
    if (union_out) *union_out = uni;
    // Python raises ZeroDivisionError when union_size == 0 (both sketches empty)
    if (uni == 0) throw Panic("division by zero");
    double jaccard = (double)inter / (double)uni;
    if (inter == uni) return 0.0;
    if (inter == 0) return 1.0;
    double distance = -std::log(2 * jaccard / (1.0 + jaccard)) / (double)k;
    if (distance > 1) distance = 1.0;
    return distance;
}

template <class F>
void parallel_for(size_t n, int threads, F&& f) {
    if (threads <= 1 || n < 2) {
        for (size_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    std::atomic<bool> failed{false};
    std::string msg;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            try {
                for (;;) {
                    size_t i = next.fetch_add(1);
                    if (i >= n || failed.load()) break;
                    f(i);
                }
            } catch (const std::exception& e) {
                if (!failed.exchange(true)) msg = e.what();
            }
        });
    for (auto& th : pool) th.join();
    if (failed) throw Panic(msg);
}

template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

void write_result(const SummedRecords& s, int64_t* sel_ids, double* sel_delta, double* sel_freqs, double* stats,
                  uint64_t* size_out) {
    size_t d = s.records[0].kfreqs.size();
    for (size_t i = 0; i < s.records.size(); ++i) {
        if (sel_ids) sel_ids[i] = s.records[i].id;
        if (sel_delta) sel_delta[i] = s.records[i].delta_jsd;
        if (sel_freqs) memcpy(sel_freqs + i * d, s.records[i].kfreqs.data(), d * sizeof(double));
    }
    if (stats) {
        stats[0] = s.total_jsd;
        stats[1] = s.mean_delta_jsd();
        stats[2] = s.std_delta_jsd();
        stats[3] = s.cov_delta_jsd();
        stats[4] = s.summed_entropies;
    }
    if (size_out) *size_out = s.records.size();
}

}  // namespace

// ------------------------------------------------------------------- C API ----
extern "C" {

const char* dvso_last_error() { return g_err.c_str(); }

uint64_t dvso_kmer_to_index(const uint8_t* kmer, uint64_t k, uint64_t num_states, uint64_t max_index) {
    auto c = coord_conversion_coeffs(num_states, k);
    return kmer_to_index(kmer, k, num_states, c, max_index);
}

// counts must hold num_states^k (k>=2) or num_states (k==1) entries
int dvso_kcounts(const uint8_t* seq, uint64_t len, int k, int num_states, uint64_t* counts) {
    return guarded([&] { to_kcounts(seq, len, (size_t)num_states, (size_t)k, counts); });
}

double dvso_entropy(const double* f, uint64_t n, int* err) {
    double v = 0.0;
    int rc = guarded([&] { v = entropy(f, n); });
    if (err) *err = rc;
    return v;
}

// returns 0 ok, 2 = Err("No valid k-mers"), 1 = panic
int dvso_kmerseq(const uint8_t* seq, uint64_t len, int k, int num_states, double* kfreqs, double* entropy_out) {
    int code = 0;
    int rc = guarded([&] {
        auto ks = to_kmerseq(0, seq, len, (size_t)num_states, (size_t)k);
        if (!ks) {
            code = 2;
            return;
        }
        memcpy(kfreqs, ks->kfreqs.data(), ks->kfreqs.size() * sizeof(double));
        *entropy_out = ks->entropy;
    });
    return rc ? rc : code;
}

// LazySeq.get_kfreqs (src/record.rs:256-261): no zero check -> NaN when no valid k-mers
int dvso_kfreqs_unchecked(const uint8_t* seq, uint64_t len, int k, int num_states, double* kfreqs) {
    return guarded([&] {
        size_t d = (k == 1) ? (size_t)num_states : ipow(num_states, k);
        std::vector<uint64_t> c(d);
        to_kcounts(seq, len, (size_t)num_states, (size_t)k, c.data());
        uint64_t tot = 0;
        for (auto v : c) tot += v;
        double total = (double)tot;
        for (size_t i = 0; i < d; ++i) kfreqs[i] = (double)c[i] / total;
    });
}

// Batch count -> rows.  valid[r]: 1 ok, 0 no valid k-mers.  Any of counts/freqs may be null.
int dvso_count_batch(const uint8_t* seqs, const uint64_t* offsets, uint64_t nrec, int k, int num_states,
                     uint64_t* counts, double* freqs, double* entropies, uint8_t* valid, int threads) {
    return guarded([&] {
        size_t d = (k == 1) ? (size_t)num_states : ipow(num_states, k);
        parallel_for(nrec, threads, [&](size_t r) {
            std::vector<uint64_t> c(d);
            to_kcounts(seqs + offsets[r], offsets[r + 1] - offsets[r], (size_t)num_states, (size_t)k, c.data());
            if (counts) memcpy(counts + r * d, c.data(), d * sizeof(uint64_t));
            uint64_t tot = 0;
            for (auto v : c) tot += v;
            double total = (double)tot;
            if (total == 0.0) {
                if (valid) valid[r] = 0;
                if (entropies) entropies[r] = 0.0;
                if (freqs)
                    for (size_t i = 0; i < d; ++i) freqs[r * d + i] = 0.0;
                return;
            }
            std::vector<double> local;
            double* f = freqs ? freqs + r * d : (local.resize(d), local.data());
            for (size_t i = 0; i < d; ++i) f[i] = (double)c[i] / total;
            double h = entropy(f, d);
            if (entropies) entropies[r] = h;
            if (valid) valid[r] = 1;
        });
    });
}

// Selection over precomputed rows.  `order[i]` = row index examined at position i; the
// row index is also the record identity (stands in for seqid).  valid[row]==0 marks a
// record whose to_kmerseq returned Err (skipped silently, records.rs:302,333).
// recompute_entropy != 0 reproduces KmerSeq::new on stored kfreqs (final_*, records.rs:353).
// mode 0 = nmost (n = min_size), 1 = max/stdev, 2 = max/cov.
// trace (optional, cap trace_cap): positions whose candidate changed the set
// (negative = -(pos)-1 for a max-mode growth).
int dvso_select_rows(const double* rows, const double* entropies, const uint8_t* valid, uint64_t d,
                     const uint64_t* order, uint64_t num, int mode, uint64_t min_size, uint64_t max_size,
                     int recompute_entropy, int64_t* sel_ids, double* sel_delta, double* sel_freqs, double* stats,
                     uint64_t* size_out, int64_t* trace, uint64_t trace_cap, uint64_t* trace_len) {
    return guarded([&] {
        RowProvider get = [&](size_t pos) -> std::optional<KmerSeq> {
            uint64_t r = order[pos];
            if (valid && !valid[r]) return std::nullopt;
            KmerSeq ks;
            ks.id = (int64_t)r;
            ks.kfreqs.assign(rows + r * d, rows + (r + 1) * d);
            ks.entropy = recompute_entropy ? entropy(ks.kfreqs.data(), d) : entropies[r];
            return ks;
        };
        std::vector<int64_t> tr;
        SummedRecords s = (mode == 0) ? select_nmost(num, min_size, get, &tr)
                                      : select_max(num, min_size, max_size, mode == 2, get, &tr);
        write_result(s, sel_ids, sel_delta, sel_freqs, stats, size_out);
        if (trace_len) *trace_len = tr.size();
        if (trace)
            for (size_t i = 0; i < std::min<size_t>(tr.size(), trace_cap); ++i) trace[i] = tr[i];
    });
}

// Selection straight from sequences, streaming one record at a time like the reference
// (rows of non-members are not kept).  Same conventions as dvso_select_rows.
int dvso_select_seqs(const uint8_t* seqs, const uint64_t* offsets, const uint64_t* order, uint64_t num, int k,
                     int num_states, int mode, uint64_t min_size, uint64_t max_size, int64_t* sel_ids,
                     double* sel_delta, double* sel_freqs, double* stats, uint64_t* size_out, int64_t* trace,
                     uint64_t trace_cap, uint64_t* trace_len) {
    return guarded([&] {
        RowProvider get = [&](size_t pos) -> std::optional<KmerSeq> {
            uint64_t r = order[pos];
            return to_kmerseq((int64_t)r, seqs + offsets[r], offsets[r + 1] - offsets[r], (size_t)num_states,
                              (size_t)k);
        };
        std::vector<int64_t> tr;
        SummedRecords s = (mode == 0) ? select_nmost(num, min_size, get, &tr)
                                      : select_max(num, min_size, max_size, mode == 2, get, &tr);
        write_result(s, sel_ids, sel_delta, sel_freqs, stats, size_out);
        if (trace_len) *trace_len = tr.size();
        if (trace)
            for (size_t i = 0; i < std::min<size_t>(tr.size(), trace_cap); ++i) trace[i] = tr[i];
    });
}

// make_summed_records (records.rs:509-524) + SummedRecordsWrapper (records_py.rs:90-125)
struct DvsoSummed {
    std::unique_ptr<SummedRecords> s;
    int k, num_states;
};

void* dvso_summed_create(const uint8_t* seqs, const uint64_t* offsets, uint64_t nrec, int k, int num_states) {
    DvsoSummed* h = nullptr;
    guarded([&] {
        std::vector<KmerSeq> recs;
        for (uint64_t r = 0; r < nrec; ++r) {
            auto ks = to_kmerseq((int64_t)r, seqs + offsets[r], offsets[r + 1] - offsets[r], (size_t)num_states,
                                 (size_t)k);
            if (ks) recs.push_back(std::move(*ks));
        }
        h = new DvsoSummed{std::make_unique<SummedRecords>(std::move(recs)), k, num_states};
    });
    return h;
}

void dvso_summed_free(void* h) { delete (DvsoSummed*)h; }

// id >= 0 names a record already given to create (same seqid); id < 0 is a new seqid.
// returns 0 ok, 2 no valid k-mers, 1 panic
int dvso_summed_delta_jsd(void* hv, int64_t id, const uint8_t* seq, uint64_t len, double* out) {
    auto* h = (DvsoSummed*)hv;
    int code = 0;
    int rc = guarded([&] {
        auto ks = to_kmerseq(id, seq, len, (size_t)h->num_states, (size_t)h->k);
        if (!ks) {
            code = 2;
            return;
        }
        *out = h->s->delta_jsd(*ks);
    });
    return rc ? rc : code;
}

int dvso_summed_result(void* hv, int64_t* sel_ids, double* sel_delta, double* sel_freqs, double* stats,
                       uint64_t* size_out) {
    auto* h = (DvsoSummed*)hv;
    return guarded([&] { write_result(*h->s, sel_ids, sel_delta, sel_freqs, stats, size_out); });
}

uint64_t dvso_summed_size(void* hv) { return ((DvsoSummed*)hv)->s->records.size(); }
uint32_t dvso_summed_lowest(void* hv) { return ((DvsoSummed*)hv)->s->lowest_index; }

// ---- mash ----
uint32_t dvso_murmurhash3_32(const uint8_t* data, uint64_t len, uint32_t seed) {
    return murmurhash3_32(data, len, seed);
}
void dvso_reverse_complement(const uint8_t* kmer, uint64_t k, uint8_t* out) {
    for (uint64_t i = 0; i < k; ++i) out[k - 1 - i] = (uint8_t)((kmer[i] + 2) % 4);
}
uint32_t dvso_hash_kmer(const uint8_t* kmer, uint64_t k, int canonical) { return hash_kmer(kmer, k, canonical != 0); }

// out must hold min(sketch_size, max(len-k+1,0)) entries; returns count via *out_len
int dvso_mash_sketch(const uint8_t* seq, uint64_t len, int k, uint64_t sketch_size, int num_states, int canonical,
                     uint32_t* out, uint64_t* out_len) {
    return guarded([&] {
        auto s = mash_sketch(seq, len, (size_t)k, (size_t)sketch_size, (uint8_t)num_states, canonical != 0);
        memcpy(out, s.data(), s.size() * sizeof(uint32_t));
        *out_len = s.size();
    });
}

// sketches: [nrec][stride] u32 ascending, lens[nrec]
int dvso_mash_sketch_batch(const uint8_t* seqs, const uint64_t* offsets, uint64_t nrec, int k, uint64_t sketch_size,
                           int num_states, int canonical, uint32_t* sketches, uint64_t stride, uint32_t* lens,
                           int threads) {
    return guarded([&] {
        parallel_for(nrec, threads, [&](size_t r) {
            auto s = mash_sketch(seqs + offsets[r], offsets[r + 1] - offsets[r], (size_t)k, (size_t)sketch_size,
                                 (uint8_t)num_states, canonical != 0);
            if (s.size() > stride) throw Panic("sketch stride too small");
            memcpy(sketches + r * stride, s.data(), s.size() * sizeof(uint32_t));
            lens[r] = (uint32_t)s.size();
        });
    });
}

double dvso_mash_distance(const uint32_t* a, uint64_t la, const uint32_t* b, uint64_t lb, int k,
                          uint64_t sketch_size, uint64_t* inter, uint64_t* uni, int* err) {
    double d = 0.0;
    int rc = guarded([&] { d = mash_distance(a, la, b, lb, k, sketch_size, inter, uni); });
    if (err) *err = rc;
    return d;
}

// diverse_seq/distance.py:163-175: lower triangle i>j, mirrored, zero diagonal
int dvso_mash_matrix(const uint32_t* sketches, uint64_t stride, const uint32_t* lens, uint64_t nrec, int k,
                     uint64_t sketch_size, double* dist, uint32_t* inter, uint32_t* uni, int threads) {
    return guarded([&] {
        for (uint64_t i = 0; i < nrec; ++i) dist[i * nrec + i] = 0.0;
        parallel_for(nrec, threads, [&](size_t i) {
            for (size_t j = 0; j < i; ++j) {
                uint64_t x, u;
                double d = mash_distance(sketches + i * stride, lens[i], sketches + j * stride, lens[j], k,
                                         sketch_size, &x, &u);
                dist[i * nrec + j] = d;
                dist[j * nrec + i] = d;
                if (inter) inter[i * nrec + j] = inter[j * nrec + i] = (uint32_t)x;
                if (uni) uni[i * nrec + j] = uni[j * nrec + i] = (uint32_t)u;
            }
        });
    });
}

// diverse_seq/distance.py:335-336  np.linalg.norm(f1 - f2); matrix per :318-332
int dvso_euclid_matrix(const double* rows, uint64_t nrec, uint64_t d, double* dist, int threads) {
    return guarded([&] {
        for (uint64_t i = 0; i < nrec; ++i) dist[i * nrec + i] = 0.0;
        parallel_for(nrec, threads, [&](size_t i) {
            const double* a = rows + i * d;
            for (size_t j = 0; j < i; ++j) {
                const double* b = rows + j * d;
                double s = 0.0;
                for (size_t t = 0; t < d; ++t) {
                    double x = a[t] - b[t];
                    s += x * x;
                }
                double v = std::sqrt(s);
                dist[i * nrec + j] = v;
                dist[j * nrec + i] = v;
            }
        });
    });
}

int dvso_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
