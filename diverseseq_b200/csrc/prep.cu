// `dvs prep` encode on the device: FASTA text -> one index-encoded record per file.
//
// Reference behaviour restated:
//   diverse_seq/io.py:30-34   converter_fasta: a-z -> A-Z, delete b"\n\r\t- "
//   diverse_seq/io.py:47-57   cogent3 iter_fasta_records(path, converter): the file is split at every
//                             '>' byte; a piece without '\n' is dropped; otherwise the first line is
//                             the label and the rest, run through the converter, the sequence
//   diverse_seq/io.py:95-104  one record per FILE: b"-".join(seqs) -> str2arr
//   diverse_seq/util.py:32-45 str2arr: most_degen_alphabet().to_indices (bytes.translate with an
//                             identity default: a byte outside the alphabet keeps its own value)
//
// As a byte-level state machine (state H = inside a label line, B = inside a sequence body; a file
// starts in H because the text before the first '>' is also cut at its first '\n'):
//   H: '\n' -> B, and emit one separator code unless this is the file's first label end; else nothing
//   B: '>'  -> H;  delete-set byte -> nothing;  any other byte c -> emit table[upper(c)]
// The state before a byte only depends on the last '>' or '\n' before it, so the text is cut into
// fixed chunks, one warp per chunk:
//   k_prep<false>  per chunk and for both possible entry states: emitted bytes, label ends, exit state
//   k_prep_carry   one thread per file chains its chunks: entry state, label ends so far, output offset
//   k_prep<true>   re-reads the chunk, translates + compacts through a per-warp shared-memory stage and
//                  writes the record bytes with aligned 16-byte stores
// Algorithmic traffic: text read once + record bytes written once (≈2 B/base); this two-pass form reads
// the text twice (≈3 B/base).
#include <algorithm>
#include <cstring>

#include "common.cuh"

using dvs::DevBuf;
using dvs::set_error;

extern "C" int dvs_seqset_alloc_internal(dvs_ctx* ctx, const uint64_t* offsets, uint32_t nrec, dvs_seqset** out);

namespace {

constexpr uint32_t kPrepChunk = 32 * 1024;  // bytes of text per warp work item
constexpr int kPrepThreads = 512;
constexpr int kPrepWarps = kPrepThreads / 32;
constexpr int kStageBytes = 16 + 512 + 16;  // pending (<16) + one 4-row step + slack

// table entry: low byte = output code, flags above
constexpr uint32_t F_DEL = 0x100, F_GT = 0x200, F_NL = 0x400;

struct PrepChunk {
    uint64_t begin;  // absolute byte offset in the text
    uint32_t len;
    uint32_t file;
};

struct PrepSum {  // per chunk, for entry state H ([0]) and B ([1])
    uint32_t emit[2];  // emitted sequence bytes (separators not included)
    uint32_t ends[2];  // label ends
    uint32_t exit_state[2];
};

struct PrepCarry {
    uint64_t out_off;     // output offset of the chunk inside its record
    uint32_t ends_before;  // label ends before the chunk (saturating)
    uint32_t state;        // entry state: 1 = H, 0 = B
};

struct PrepParams {
    uint32_t sep_code;
    uint32_t swar;     // 1: the SWAR fast classification below is valid for this table
    uint32_t letters;  // the four canonical letters (upper case), byte s = the letter with ((c >> 1) & 3) == s
    uint32_t codes;    // their codes, same byte order
};

struct RowOut {
    uint32_t out4;   // lane: bytes that produce output (sequence or separator)
    uint32_t sep4;   // lane: subset of out4 that are separators
    uint32_t state;  // warp: state after the row
    uint32_t ends;   // warp: label ends in the row
    uint32_t emit;   // warp: sequence bytes emitted in the row
};

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// 0x80 in every byte of x that is non-zero
__device__ __forceinline__ uint32_t nz_flags(uint32_t x) {
    return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// per-byte flags (0x80 each) <-> 4-bit masks
__device__ __forceinline__ uint32_t flags_to_mask4(uint32_t f) { return (((f >> 7) * 0x01020408u) >> 24) & 0xFu; }
__device__ __forceinline__ uint32_t mask4_to_flags(uint32_t m) { return ((m * 0x00204081u) & 0x01010101u) << 7; }

// Classes of the four bytes of `w` as 0x80-per-byte flags: deleted and '\n' (nlf, also deleted), plus
// `other` = bytes that need the 256-entry table.  SWAR form: with CODES, canonical letters of either
// case are recognised and translated by two PRMT lookups keyed by bits 1-2 of the byte; without, only
// bytes < 0x40 can be deleted or special at all.  '\n' is found by a SWAR compare.  Branch-free.
template <bool CODES, bool SWAR>
__device__ __forceinline__ void classify(uint32_t w, const PrepParams& p, uint32_t& codes, uint32_t& nlf,
                                         uint32_t& other) {
    codes = 0;
    if (!SWAR) {
        nlf = 0;
        other = 0x80808080u;
        return;
    }
    nlf = ~nz_flags(w ^ 0x0A0A0A0Au) & 0x80808080u;
    if (CODES) {
        const uint32_t u = w & 0xDFDFDFDFu;
        const uint32_t x = (u >> 1) & 0x03030303u;
        uint32_t sel = (x & 0x00030003u) | ((x >> 4) & 0x00300030u);
        sel = (sel | (sel >> 8)) & 0x3333u;
        other = nz_flags(__byte_perm(p.letters, 0, sel) ^ u) & ~nlf;
        codes = __byte_perm(p.codes, 0, sel);
    } else {
        other = ~w & (~w << 1) & 0x80808080u & ~nlf;  // bytes < 0x40 are the only deleted / special ones
    }
}

// table lookups for the bytes flagged in `other` (rare: labels, ambiguity codes, CR, blanks)
template <bool CODES>
__device__ __forceinline__ void classify_table(uint32_t w, uint32_t other, const uint16_t* tab, uint32_t& codes,
                                               uint32_t& delf, uint32_t& gtf, uint32_t& nlf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (other & (0x80u << (8 * i))) {
            const uint32_t e = tab[(w >> (8 * i)) & 0xFFu];
            if (CODES) codes = (codes & ~(0xFFu << (8 * i))) | ((e & 0xFFu) << (8 * i));
            delf |= ((e >> 8) & 1u) << (8 * i + 7);
            gtf |= ((e >> 9) & 1u) << (8 * i + 7);
            nlf |= ((e >> 10) & 1u) << (8 * i + 7);
        }
    }
}

// the label/body state machine over one 128-byte row (lane = 4 bytes), general case
__device__ __forceinline__ RowOut prep_row(uint32_t gt4, uint32_t nl4, uint32_t del4, uint32_t state,
                                           uint32_t ends_before) {
    RowOut r;
    r.sep4 = 0;
    r.ends = 0;
    r.state = state;
    const uint32_t any_gt = __ballot_sync(0xffffffffu, gt4 != 0);
    const uint32_t any_nl = __ballot_sync(0xffffffffu, nl4 != 0);
    if (state == 0 && any_gt == 0) {  // body all the way
        r.out4 = ~del4 & 0xFu;
    } else if (state == 1 && any_nl == 0) {  // label all the way
        r.out4 = 0;
    } else {
        const uint32_t sp4 = gt4 | nl4;
        const uint32_t m = any_gt | any_nl;
        const uint32_t last_is_gt = sp4 ? ((gt4 >> (31 - __clz(sp4))) & 1u) : 0u;
        const uint32_t prev = m & lanemask_lt();
        const uint32_t from = __shfl_sync(0xffffffffu, last_is_gt, prev ? (31 - __clz(prev)) : 0);
        uint32_t s = prev ? from : state;
        uint32_t emit4 = 0, he4 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t bit = 1u << i;
            if (s) {
                if (nl4 & bit) {
                    he4 |= bit;
                    s = 0;
                }
            } else if (gt4 & bit) {
                s = 1;
            } else if (!(del4 & bit)) {
                emit4 |= bit;
            }
        }
        const uint32_t nhe = __popc(he4);  // 0..2
        const uint32_t b0 = __ballot_sync(0xffffffffu, nhe & 1u), b1 = __ballot_sync(0xffffffffu, nhe & 2u);
        const uint32_t lt = lanemask_lt();
        const uint32_t before = ends_before + __popc(b0 & lt) + 2 * __popc(b1 & lt);
        r.ends = __popc(b0) + 2 * __popc(b1);
        r.sep4 = he4;
        if (before == 0 && he4) r.sep4 = he4 & (he4 - 1);  // the file's first label end joins nothing
        r.out4 = emit4 | r.sep4;
        r.state = __shfl_sync(0xffffffffu, last_is_gt, 31 - __clz(m));  // m != 0 on this path
    }
    // warp totals of emitted sequence bytes (0..4 per lane)
    const uint32_t ne = __popc(r.out4 & ~r.sep4);
    const uint32_t e0 = __ballot_sync(0xffffffffu, ne & 1u), e1 = __ballot_sync(0xffffffffu, ne & 2u),
                   e2 = __ballot_sync(0xffffffffu, ne & 4u);
    r.emit = __popc(e0) + 2 * __popc(e1) + 4 * __popc(e2);
    return r;
}

template <bool WRITE, bool SWAR>
__global__ void __launch_bounds__(kPrepThreads)
k_prep(const uint8_t* __restrict__ text, const PrepChunk* __restrict__ chunks, uint32_t nchunks,
       const uint16_t* __restrict__ table, PrepParams prm, PrepSum* __restrict__ sums,
       const PrepCarry* __restrict__ carry, const uint64_t* __restrict__ rec_offsets, uint8_t* __restrict__ out) {
    __shared__ uint16_t tab[256];
    __shared__ __align__(16) uint8_t stage_all[WRITE ? kPrepWarps * kStageBytes : 16];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = table[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * kPrepWarps;
    uint8_t* stage = stage_all + (WRITE ? warp * kStageBytes : 0);

    for (uint32_t ci = blockIdx.x * kPrepWarps + warp; ci < nchunks; ci += nwarps) {
        const PrepChunk ch = chunks[ci];
        const uint64_t row0 = ch.begin & ~127ull;
        const uint32_t rel_begin = (uint32_t)(ch.begin - row0), rel_end = rel_begin + ch.len;  // bytes from row0
        const uint32_t nsteps = (rel_end + 511) >> 9;                                       // 512-byte steps
        const uint32_t s_lo = rel_begin ? 1u : 0u, s_hi = rel_end >> 9;  // steps [s_lo, s_hi) lie fully inside
        const uint8_t* base = text + row0 + 4 * lane;
        // WRITE: the real entry state in [0]; else both hypotheses ([0]: entered in H, [1]: entered in B)
        uint32_t st[2] = {1u, 0u}, emit[2] = {0, 0}, ends[2] = {0, 0};
        uint32_t lane_emit[2] = {0, 0};  // per-lane partial counts of the fast path (summed at the end)
        uint64_t gbase = 0;
        uint32_t fill = 0, head_skip = 0;
        if (WRITE) {
            const PrepCarry cy = carry[ci];
            st[0] = cy.state;
            ends[0] = cy.ends_before;
            const uint64_t o0 = rec_offsets[ch.file] + cy.out_off;
            gbase = o0 & ~15ull;
            fill = head_skip = (uint32_t)(o0 & 15);
        }
        uint32_t w[4], wn[4];
        auto load_step = [&](uint32_t s, uint32_t* dst) {
            const uint8_t* q0 = base + ((uint64_t)s << 9);
            if (s >= s_lo && s < s_hi) {  // warp-uniform: no guards
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = __ldg(reinterpret_cast<const uint32_t*>(q0 + 128 * j));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t q = (s << 9) + 128 * j + 4 * lane;  // offset of the lane's word from row0
                    dst[j] = (s < nsteps && q + 4 > rel_begin && q < rel_end)
                                 ? __ldg(reinterpret_cast<const uint32_t*>(q0 + 128 * j))
                                 : 0u;
                }
            }
        };
        load_step(0, w);
        for (uint32_t s = 0; s < nsteps; ++s) {
            load_step(s + 1, wn);  // next step in flight while this one is processed
            uint32_t codes[4], delf[4], gtf[4], nlf[4], other[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                classify<WRITE, SWAR>(w[j], prm, codes[j], nlf[j], other[j]);
                delf[j] = nlf[j];
                gtf[j] = 0;
            }
            if (other[0] | other[1] | other[2] | other[3]) {
#pragma unroll
                for (int j = 0; j < 4; ++j) classify_table<WRITE>(w[j], other[j], tab, codes[j], delf[j], gtf[j], nlf[j]);
            }
            if (!(s >= s_lo && s < s_hi)) {  // warp-uniform: bytes outside [begin, end) count as deleted
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t q = (s << 9) + 128 * j + 4 * lane;
                    uint32_t valid = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (q + i >= rel_begin && q + i < rel_end) valid |= 0x80u << (8 * i);
                    delf[j] = (delf[j] & valid) | (~valid & 0x80808080u);
                    gtf[j] &= valid;
                    nlf[j] &= valid;
                }
            }
            const uint32_t any_gt = __ballot_sync(0xffffffffu, (gtf[0] | gtf[1] | gtf[2] | gtf[3]) != 0);
            const uint32_t any_nl = __ballot_sync(0xffffffffu, (nlf[0] | nlf[1] | nlf[2] | nlf[3]) != 0);
            if (WRITE) {
                uint32_t outf[4];  // bytes of the step that produce output
                if (st[0] == 0 && !any_gt) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) outf[j] = ~delf[j] & 0x80808080u;
                } else if (st[0] == 1 && !any_nl) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) outf[j] = 0;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const RowOut r = prep_row(flags_to_mask4(gtf[j]), flags_to_mask4(nlf[j]), flags_to_mask4(delf[j]),
                                                  st[0], ends[0]);
                        st[0] = r.state;
                        ends[0] = min(ends[0] + r.ends, 0x7fffffffu);
                        outf[j] = mask4_to_flags(r.out4);
                        const uint32_t sepf = mask4_to_flags(r.sep4);
                        const uint32_t sm = (sepf >> 7) * 0xFFu;  // 0xFF in separator bytes
                        codes[j] = (codes[j] & ~sm) | ((prm.sep_code * 0x01010101u) & sm);
                    }
                }
                // positions: one packed scan over (row, lane) order, 8 bits per row (<= 128 each)
                const uint32_t c = __popc(outf[0]) | (__popc(outf[1]) << 8) | (__popc(outf[2]) << 16) | (__popc(outf[3]) << 24);
                uint32_t incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += v;
                }
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t excl = incl - c;
                uint32_t rowbase = fill;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint8_t* dst = stage + rowbase + ((excl >> (8 * j)) & 0xFFu);
                    if (outf[j] == 0x80808080u) {  // the common word: nothing deleted
                        dst[0] = (uint8_t)codes[j];
                        dst[1] = (uint8_t)(codes[j] >> 8);
                        dst[2] = (uint8_t)(codes[j] >> 16);
                        dst[3] = (uint8_t)(codes[j] >> 24);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (outf[j] & (0x80u << (8 * i))) *dst++ = (uint8_t)(codes[j] >> (8 * i));
                    }
                    rowbase += (tot >> (8 * j)) & 0xFFu;
                }
                fill = rowbase;
                __syncwarp();
                const uint32_t nfull = fill >> 4;
                for (uint32_t q = lane; q < nfull; q += 32) {
                    if (q == 0 && head_skip) {  // the bytes before head_skip belong to the previous chunk
                        for (uint32_t b = head_skip; b < 16; ++b) out[gbase + b] = stage[b];
                    } else {
                        *reinterpret_cast<uint4*>(out + gbase + 16ull * q) = *reinterpret_cast<const uint4*>(stage + 16 * q);
                    }
                }
                if (nfull) {
                    const uint32_t rem = fill & 15;
                    const uint8_t keep = lane < rem ? stage[16 * nfull + lane] : (uint8_t)0;
                    __syncwarp();
                    if (lane < rem) stage[lane] = keep;
                    gbase += 16ull * nfull;
                    fill = rem;
                    head_skip = 0;
                }
                __syncwarp();
            } else {
                const bool conv = st[0] == st[1];
                const uint32_t kept = __popc(~delf[0] & 0x80808080u) + __popc(~delf[1] & 0x80808080u) +
                                      __popc(~delf[2] & 0x80808080u) + __popc(~delf[3] & 0x80808080u);
                if (conv && st[0] == 0 && !any_gt) {
                    lane_emit[0] += kept;
                    lane_emit[1] += kept;
                } else if (conv && st[0] == 1 && !any_nl) {
                } else if (!conv && !any_gt && !any_nl) {
                    lane_emit[1] += kept;  // entered in B: body so far; entered in H: still in the label
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t g4 = flags_to_mask4(gtf[j]), n4 = flags_to_mask4(nlf[j]), d4 = flags_to_mask4(delf[j]);
                        if (st[0] == st[1]) {
                            const RowOut r = prep_row(g4, n4, d4, st[0], 1);
                            st[0] = st[1] = r.state;
                            emit[0] += r.emit;
                            emit[1] += r.emit;
                            ends[0] += r.ends;
                            ends[1] += r.ends;
                        } else {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const RowOut r = prep_row(g4, n4, d4, st[h], 1);
                                st[h] = r.state;
                                emit[h] += r.emit;
                                ends[h] += r.ends;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = wn[j];
        }
        if (WRITE) {
            if (lane < fill && lane >= head_skip) out[gbase + lane] = stage[lane];
            __syncwarp();
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int d = 16; d; d >>= 1) lane_emit[h] += __shfl_xor_sync(0xffffffffu, lane_emit[h], d);
            if (lane == 0) {
                PrepSum sm;
                for (int h = 0; h < 2; ++h) {
                    sm.emit[h] = emit[h] + lane_emit[h];
                    sm.ends[h] = ends[h];
                    sm.exit_state[h] = st[h];
                }
                sums[ci] = sm;
            }
        }
    }
}

// one thread per file: chain the chunk summaries (entry state, label ends so far, output offsets)
__global__ void k_prep_carry(const PrepSum* __restrict__ sums, const uint32_t* __restrict__ first_chunk,
                             uint32_t nfiles, PrepCarry* __restrict__ carry, uint64_t* __restrict__ rec_len) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfiles) return;
    uint32_t state = 1, ends = 0;
    uint64_t off = 0;
    for (uint32_t c = first_chunk[f]; c < first_chunk[f + 1]; ++c) {
        const PrepSum s = sums[c];
        const int h = state ? 0 : 1;
        carry[c] = PrepCarry{off, ends, state};
        const uint32_t seps = s.ends[h] - ((ends == 0 && s.ends[h] > 0) ? 1u : 0u);
        off += (uint64_t)s.emit[h] + seps;
        ends = min(ends + s.ends[h], 0x7fffffffu);
        state = s.exit_state[h];
    }
    rec_len[f] = off;
}

}  // namespace

extern "C" {

int dvs_prep_fasta(dvs_ctx* ctx, const uint8_t* text, const uint64_t* file_offsets, uint32_t nfiles,
                   const char* alphabet, const char* delete_chars, int sep_char, int text_on_device,
                   dvs_seqset** out) {
    if (!ctx || !file_offsets || !out || (!text && file_offsets[nfiles] > 0)) {
        set_error("dvs_prep_fasta: NULL argument");
        return DVS_ERR_ARG;
    }
    for (uint32_t f = 0; f < nfiles; ++f)
        if (file_offsets[f + 1] < file_offsets[f]) {
            set_error("dvs_prep_fasta: file_offsets must be non-decreasing (file %u)", f);
            return DVS_ERR_ARG;
        }
    if (text_on_device && ((uintptr_t)text & 3)) {
        set_error("dvs_prep_fasta: device text must be 4-byte aligned");
        return DVS_ERR_ARG;
    }
    if (!alphabet) alphabet = DVS_DNA_ALPHABET;
    if (!delete_chars) delete_chars = "\n\r\t- ";
    if (sep_char < 0) sep_char = '-';
    const size_t na = strlen(alphabet);
    if (na == 0 || na > 255) {
        set_error("dvs_prep_fasta: alphabet must have 1..255 characters");
        return DVS_ERR_ARG;
    }
    // translate table: upper-case, then alphabet position, identity for bytes outside the alphabet
    uint8_t code[256];
    for (int b = 0; b < 256; ++b) code[b] = (uint8_t)b;
    for (size_t i = 0; i < na; ++i) code[(uint8_t)alphabet[i]] = (uint8_t)i;
    uint16_t table[256];
    for (int b = 0; b < 256; ++b) {
        const int up = (b >= 'a' && b <= 'z') ? b - 32 : b;
        table[b] = code[up];
    }
    for (const char* d = delete_chars; *d; ++d) table[(uint8_t)*d] |= F_DEL;
    table[(uint8_t)'\n'] |= F_NL;
    table[(uint8_t)'>'] |= F_GT;
    PrepParams prm{};
    prm.sep_code = code[(uint8_t)sep_char];
    {
        // SWAR fast path: four upper-case letters with distinct ((c >> 1) & 3), kept (not deleted), every
        // deleted byte and '>' below 0x40, and '\n' deleted.  Otherwise every byte takes the table.
        bool ok = na >= 4 && (table[(uint8_t)'\n'] & F_DEL);
        uint32_t seen = 0;
        for (int i = 0; i < 4 && ok; ++i) {
            const uint8_t c = (uint8_t)alphabet[i];
            ok = c >= 'A' && c <= 'Z' && !(table[c] & F_DEL) && !(table[c | 0x20] & F_DEL) && !(seen & (1u << ((c >> 1) & 3)));
            seen |= 1u << ((c >> 1) & 3);
            prm.letters |= (uint32_t)c << (8 * ((c >> 1) & 3));
            prm.codes |= (uint32_t)code[c] << (8 * ((c >> 1) & 3));
        }
        for (int b = 0x40; b < 256 && ok; ++b) ok = !(table[b] & (F_DEL | F_GT | F_NL));
        prm.swar = ok ? 1u : 0u;
    }

    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    const uint64_t total = file_offsets[nfiles];

    std::vector<PrepChunk> chunks;
    std::vector<uint32_t> first_chunk(nfiles + 1, 0);
    for (uint32_t f = 0; f < nfiles; ++f) {
        first_chunk[f] = (uint32_t)chunks.size();
        for (uint64_t b = file_offsets[f]; b < file_offsets[f + 1]; b += kPrepChunk)
            chunks.push_back({b, (uint32_t)std::min<uint64_t>(kPrepChunk, file_offsets[f + 1] - b), f});
        if (chunks.size() > 0xFFFFFFF0ull) {
            set_error("dvs_prep_fasta: text too large");
            return DVS_ERR_ARG;
        }
    }
    first_chunk[nfiles] = (uint32_t)chunks.size();
    const uint32_t nchunks = (uint32_t)chunks.size();

    DevBuf<uint8_t> d_text;
    const uint8_t* dtext = text;
    if (!text_on_device && total) {
        DVS_TRY(d_text.alloc(total + 256));
        PhaseTimer up(ctx, DVS_PHASE_UPLOAD);
        DVS_CUDA_TRY(cudaMemcpyAsync(d_text.p, text, total, cudaMemcpyHostToDevice, st));
        up.stop();
        dtext = d_text.p;
    }
    DevBuf<PrepChunk> d_chunks;
    DevBuf<uint32_t> d_first;
    DevBuf<PrepSum> d_sums;
    DevBuf<PrepCarry> d_carry;
    DevBuf<uint64_t> d_len;
    DevBuf<uint16_t> d_table;
    DVS_TRY(d_chunks.alloc(nchunks));
    DVS_TRY(d_first.alloc(nfiles + 1));
    DVS_TRY(d_sums.alloc(nchunks));
    DVS_TRY(d_carry.alloc(nchunks));
    DVS_TRY(d_len.alloc(nfiles));
    DVS_TRY(d_table.alloc(256));
    if (nchunks)
        DVS_CUDA_TRY(cudaMemcpyAsync(d_chunks.p, chunks.data(), (size_t)nchunks * sizeof(PrepChunk),
                                     cudaMemcpyHostToDevice, st));
    DVS_CUDA_TRY(cudaMemcpyAsync(d_first.p, first_chunk.data(), (nfiles + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    DVS_CUDA_TRY(cudaMemcpyAsync(d_table.p, table, sizeof(table), cudaMemcpyHostToDevice, st));

    const uint32_t grid = (uint32_t)std::max<uint64_t>(
        1, std::min<uint64_t>((uint64_t)ctx->sm_count * 4, (nchunks + kPrepWarps - 1) / kPrepWarps));
    std::vector<uint64_t> rec_offsets(nfiles + 1, 0);
    PhaseTimer pt(ctx, DVS_PHASE_PREP);
    if (nfiles) {
        if (nchunks) {
            if (prm.swar)
                k_prep<false, true><<<grid, kPrepThreads, 0, st>>>(dtext, d_chunks.p, nchunks, d_table.p, prm, d_sums.p,
                                                                   nullptr, nullptr, nullptr);
            else
                k_prep<false, false><<<grid, kPrepThreads, 0, st>>>(dtext, d_chunks.p, nchunks, d_table.p, prm, d_sums.p,
                                                                    nullptr, nullptr, nullptr);
            DVS_LAUNCHED(ctx);
        }
        k_prep_carry<<<(nfiles + 127) / 128, 128, 0, st>>>(d_sums.p, d_first.p, nfiles, d_carry.p, d_len.p);
        DVS_LAUNCHED(ctx);
        std::vector<uint64_t> len(nfiles);
        DVS_CUDA_TRY(cudaMemcpyAsync(len.data(), d_len.p, nfiles * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        DVS_CUDA_TRY(cudaStreamSynchronize(st));
        for (uint32_t f = 0; f < nfiles; ++f) rec_offsets[f + 1] = rec_offsets[f] + len[f];
    }
    dvs_seqset* s = nullptr;
    DVS_TRY(dvs_seqset_alloc_internal(ctx, rec_offsets.data(), nfiles, &s));
    if (nchunks && s->total) {
        if (prm.swar)
            k_prep<true, true><<<grid, kPrepThreads, 0, st>>>(dtext, d_chunks.p, nchunks, d_table.p, prm, nullptr,
                                                              d_carry.p, s->offsets.p, s->data());
        else
            k_prep<true, false><<<grid, kPrepThreads, 0, st>>>(dtext, d_chunks.p, nchunks, d_table.p, prm, nullptr,
                                                               d_carry.p, s->offsets.p, s->data());
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e), __FILE__, __LINE__);
            delete s;
            return DVS_ERR_CUDA;
        }
    }
    pt.stop();
    // the caller's text (possibly pageable) and our local tables must outlive the copies
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("dvs_prep_fasta failed: %s", cudaGetErrorString(e));
        delete s;
        return DVS_ERR_CUDA;
    }
    *out = s;
    return DVS_OK;
}

}  // extern "C"
