"""ctypes front-end for the CPU oracle (oracle/libdvs_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs.  The product package
``diverseseq_b200`` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libdvs_oracle.so"

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> pathlib.Path:
    if force or not _LIB_PATH.exists():
        subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []), check=True,
                       capture_output=True)
    return _LIB_PATH


def _load() -> C.CDLL:
    build()
    lib = C.CDLL(str(_LIB_PATH))
    lib.dvso_last_error.restype = C.c_char_p
    lib.dvso_kmer_to_index.restype = C.c_uint64
    lib.dvso_kmer_to_index.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.dvso_kcounts.argtypes = [_u8p, C.c_uint64, C.c_int, C.c_int, _u64p]
    lib.dvso_entropy.restype = C.c_double
    lib.dvso_entropy.argtypes = [_f64p, C.c_uint64, C.POINTER(C.c_int)]
    lib.dvso_kmerseq.argtypes = [_u8p, C.c_uint64, C.c_int, C.c_int, _f64p, C.POINTER(C.c_double)]
    lib.dvso_kfreqs_unchecked.argtypes = [_u8p, C.c_uint64, C.c_int, C.c_int, _f64p]
    lib.dvso_count_batch.argtypes = [_u8p, _u64p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int]
    lib.dvso_select_rows.argtypes = [_f64p, C.c_void_p, C.c_void_p, C.c_uint64, _u64p, C.c_uint64, C.c_int,
                                     C.c_uint64, C.c_uint64, C.c_int, _i64p, _f64p, C.c_void_p, _f64p,
                                     C.POINTER(C.c_uint64), _i64p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.dvso_select_seqs.argtypes = [_u8p, _u64p, _u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                     C.c_uint64, _i64p, _f64p, C.c_void_p, _f64p, C.POINTER(C.c_uint64), _i64p,
                                     C.c_uint64, C.POINTER(C.c_uint64)]
    lib.dvso_summed_create.restype = C.c_void_p
    lib.dvso_summed_create.argtypes = [_u8p, _u64p, C.c_uint64, C.c_int, C.c_int]
    lib.dvso_summed_free.argtypes = [C.c_void_p]
    lib.dvso_summed_delta_jsd.argtypes = [C.c_void_p, C.c_int64, _u8p, C.c_uint64, C.POINTER(C.c_double)]
    lib.dvso_summed_result.argtypes = [C.c_void_p, _i64p, _f64p, C.c_void_p, _f64p, C.POINTER(C.c_uint64)]
    lib.dvso_summed_size.restype = C.c_uint64
    lib.dvso_summed_size.argtypes = [C.c_void_p]
    lib.dvso_summed_lowest.restype = C.c_uint32
    lib.dvso_summed_lowest.argtypes = [C.c_void_p]
    lib.dvso_murmurhash3_32.restype = C.c_uint32
    lib.dvso_murmurhash3_32.argtypes = [_u8p, C.c_uint64, C.c_uint32]
    lib.dvso_reverse_complement.argtypes = [_u8p, C.c_uint64, _u8p]
    lib.dvso_hash_kmer.restype = C.c_uint32
    lib.dvso_hash_kmer.argtypes = [_u8p, C.c_uint64, C.c_int]
    lib.dvso_mash_sketch.argtypes = [_u8p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int, _u32p,
                                     C.POINTER(C.c_uint64)]
    lib.dvso_mash_sketch_batch.argtypes = [_u8p, _u64p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                           _u32p, C.c_uint64, _u32p, C.c_int]
    lib.dvso_mash_distance.restype = C.c_double
    lib.dvso_mash_distance.argtypes = [_u32p, C.c_uint64, _u32p, C.c_uint64, C.c_int, C.c_uint64,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    lib.dvso_mash_matrix.argtypes = [_u32p, C.c_uint64, _u32p, C.c_uint64, C.c_int, C.c_uint64, _f64p,
                                     C.c_void_p, C.c_void_p, C.c_int]
    lib.dvso_euclid_matrix.argtypes = [_f64p, C.c_uint64, C.c_uint64, _f64p, C.c_int]
    lib.dvso_hardware_threads.restype = C.c_int
    lib.dvso_log2_port.restype = C.c_double
    lib.dvso_log2_port.argtypes = [C.c_double]
    lib.dvso_log2_libm.restype = C.c_double
    lib.dvso_log2_libm.argtypes = [C.c_double]
    lib.dvso_log2_port_mismatches.restype = C.c_uint64
    lib.dvso_log2_port_mismatches.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    lib.dvso_log2_samples.argtypes = [C.c_uint64, C.c_uint64, _f64p, _f64p]
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


class OraclePanic(ValueError):
    """A Rust panic in the reference (surfaces to Python as ValueError, src/lib.rs:36-57)."""


def _check(rc: int) -> None:
    if rc == 1:
        raise OraclePanic(lib().dvso_last_error().decode())


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray))
                                else np.asarray(a, dtype=np.uint8))


def _vptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_kmers(k: int, num_states: int = 4) -> int:
    return num_states if k == 1 else num_states ** k


def hardware_threads() -> int:
    return lib().dvso_hardware_threads()


def concat(seqs) -> tuple[np.ndarray, np.ndarray]:
    """list of uint8 arrays -> (flat bytes, offsets[n+1])"""
    arrs = [_u8(s) for s in seqs]
    offsets = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offsets[1:] = np.cumsum([len(a) for a in arrs], dtype=np.uint64)
    flat = np.concatenate(arrs) if arrs and offsets[-1] else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(flat), offsets


def kmer_to_index(kmer, num_states: int, max_index: int) -> int:
    a = _u8(kmer)
    return int(lib().dvso_kmer_to_index(a, len(a), num_states, max_index))


def kcounts(seq, k: int, num_states: int = 4) -> np.ndarray:
    a = _u8(seq)
    out = np.zeros(num_kmers(k, num_states), dtype=np.uint64)
    _check(lib().dvso_kcounts(a, len(a), k, num_states, out))
    return out


def entropy(freqs) -> float:
    f = np.ascontiguousarray(freqs, dtype=np.float64)
    err = C.c_int(0)
    v = lib().dvso_entropy(f, len(f), C.byref(err))
    _check(err.value)
    return float(v)


def kmerseq(seq, k: int, num_states: int = 4):
    """SeqRecord::to_kmerseq -> (kfreqs, entropy) or None for Err("No valid k-mers")."""
    a = _u8(seq)
    f = np.zeros(num_kmers(k, num_states), dtype=np.float64)
    h = C.c_double(0.0)
    rc = lib().dvso_kmerseq(a, len(a), k, num_states, f, C.byref(h))
    _check(rc)
    return None if rc == 2 else (f, h.value)


def kfreqs_unchecked(seq, k: int, num_states: int = 4) -> np.ndarray:
    a = _u8(seq)
    f = np.zeros(num_kmers(k, num_states), dtype=np.float64)
    _check(lib().dvso_kfreqs_unchecked(a, len(a), k, num_states, f))
    return f


def count_batch(flat, offsets, k: int, num_states: int = 4, threads: int = 1, want_counts=True,
                want_freqs=True):
    flat = _u8(flat)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    d = num_kmers(k, num_states)
    counts = np.zeros((n, d), dtype=np.uint64) if want_counts else None
    freqs = np.zeros((n, d), dtype=np.float64) if want_freqs else None
    ent = np.zeros(n, dtype=np.float64)
    valid = np.zeros(n, dtype=np.uint8)
    if len(flat) == 0:
        flat = np.zeros(1, dtype=np.uint8)
    _check(lib().dvso_count_batch(flat, offsets, n, k, num_states, _vptr(counts), _vptr(freqs), _vptr(ent),
                                  _vptr(valid), threads))
    return counts, freqs, ent, valid


class Selection:
    __slots__ = ("ids", "delta_jsd", "kfreqs", "total_jsd", "mean_delta_jsd", "std_delta_jsd",
                 "cov_delta_jsd", "summed_entropies", "size", "trace")

    def __repr__(self):
        return f"Selection(size={self.size}, ids={self.ids.tolist()}, total_jsd={self.total_jsd!r})"


def _mk_selection(ids, delta, freqs, stats, size, trace):
    s = Selection()
    n = int(size)
    s.ids, s.delta_jsd = ids[:n].copy(), delta[:n].copy()
    s.kfreqs = None if freqs is None else freqs[:n].copy()
    s.total_jsd, s.mean_delta_jsd, s.std_delta_jsd, s.cov_delta_jsd, s.summed_entropies = map(float, stats)
    s.size, s.trace = n, trace
    return s


_MODES = {"nmost": 0, "stdev": 1, "cov": 2}


def select_rows(rows, entropies, order, mode: str, min_size: int, max_size: int = 0, valid=None,
                recompute_entropy: bool = False, want_freqs: bool = False) -> Selection:
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    d = rows.shape[1]
    order = np.ascontiguousarray(order, dtype=np.uint64)
    num = len(order)
    ent = None if entropies is None else np.ascontiguousarray(entropies, dtype=np.float64)
    val = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    cap = max(num, max_size, min_size, 1)
    ids = np.zeros(cap, dtype=np.int64)
    delta = np.zeros(cap, dtype=np.float64)
    freqs = np.zeros((cap, d), dtype=np.float64) if want_freqs else None
    stats = np.zeros(5, dtype=np.float64)
    size, tlen = C.c_uint64(0), C.c_uint64(0)
    trace = np.zeros(max(num, 1), dtype=np.int64)
    _check(lib().dvso_select_rows(rows, _vptr(ent), _vptr(val), d, order, num, _MODES[mode], min_size,
                                  max_size, int(recompute_entropy), ids, delta, _vptr(freqs), stats,
                                  C.byref(size), trace, len(trace), C.byref(tlen)))
    return _mk_selection(ids, delta, freqs, stats, size.value, trace[: tlen.value].copy())


def select_seqs(flat, offsets, order, k: int, mode: str, min_size: int, max_size: int = 0,
                num_states: int = 4, want_freqs: bool = False) -> Selection:
    flat = _u8(flat)
    if len(flat) == 0:
        flat = np.zeros(1, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    order = np.ascontiguousarray(order, dtype=np.uint64)
    num = len(order)
    d = num_kmers(k, num_states)
    cap = max(num, max_size, min_size, 1)
    ids = np.zeros(cap, dtype=np.int64)
    delta = np.zeros(cap, dtype=np.float64)
    freqs = np.zeros((cap, d), dtype=np.float64) if want_freqs else None
    stats = np.zeros(5, dtype=np.float64)
    size, tlen = C.c_uint64(0), C.c_uint64(0)
    trace = np.zeros(max(num, 1), dtype=np.int64)
    _check(lib().dvso_select_seqs(flat, offsets, order, num, k, num_states, _MODES[mode], min_size, max_size,
                                  ids, delta, _vptr(freqs), stats, C.byref(size), trace, len(trace),
                                  C.byref(tlen)))
    return _mk_selection(ids, delta, freqs, stats, size.value, trace[: tlen.value].copy())


class Summed:
    """make_summed_records + SummedRecordsWrapper (src/records.rs:509, src/records_py.rs:90)."""

    def __init__(self, seqs, k: int, num_states: int = 4):
        flat, offsets = concat(seqs)
        if len(flat) == 0:
            flat = np.zeros(1, dtype=np.uint8)
        self._d = num_kmers(k, num_states)
        self._h = lib().dvso_summed_create(flat, offsets, len(offsets) - 1, k, num_states)
        if not self._h:
            raise OraclePanic(lib().dvso_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dvso_summed_free(self._h)
            self._h = None

    def delta_jsd(self, seq, ident: int = -1):
        a = _u8(seq)
        if len(a) == 0:
            a = np.zeros(1, dtype=np.uint8)
            n = 0
        else:
            n = len(a)
        out = C.c_double(0.0)
        rc = lib().dvso_summed_delta_jsd(self._h, ident, a, n, C.byref(out))
        _check(rc)
        if rc == 2:
            raise ValueError("No valid k-mers")
        return out.value

    @property
    def lowest_index(self) -> int:
        return int(lib().dvso_summed_lowest(self._h))

    def result(self, want_freqs: bool = False) -> Selection:
        n = int(lib().dvso_summed_size(self._h))
        ids = np.zeros(n, dtype=np.int64)
        delta = np.zeros(n, dtype=np.float64)
        freqs = np.zeros((n, self._d), dtype=np.float64) if want_freqs else None
        stats = np.zeros(5, dtype=np.float64)
        size = C.c_uint64(0)
        _check(lib().dvso_summed_result(self._h, ids, delta, _vptr(freqs), stats, C.byref(size)))
        return _mk_selection(ids, delta, freqs, stats, size.value, None)


def murmurhash3_32(data, seed: int = 0) -> int:
    a = _u8(data)
    if len(a) == 0:
        return int(lib().dvso_murmurhash3_32(np.zeros(1, dtype=np.uint8), 0, seed))
    return int(lib().dvso_murmurhash3_32(a, len(a), seed))


def reverse_complement(kmer) -> np.ndarray:
    a = _u8(kmer)
    out = np.zeros_like(a)
    lib().dvso_reverse_complement(a, len(a), out)
    return out


def hash_kmer(kmer, canonical: bool = False) -> int:
    a = _u8(kmer)
    return int(lib().dvso_hash_kmer(a, len(a), int(canonical)))


def mash_sketch(seq, k: int, sketch_size: int, num_states: int = 4, canonical: bool = False) -> np.ndarray:
    a = _u8(seq)
    cap = max(min(sketch_size, max(len(a) - k + 1, 0)), 1)
    out = np.zeros(cap, dtype=np.uint32)
    n = C.c_uint64(0)
    if len(a) == 0:
        a = np.zeros(1, dtype=np.uint8)
        ln = 0
    else:
        ln = len(a)
    _check(lib().dvso_mash_sketch(a, ln, k, sketch_size, num_states, int(canonical), out, C.byref(n)))
    return out[: n.value].copy()


def mash_sketch_batch(flat, offsets, k: int, sketch_size: int, num_states: int = 4, canonical: bool = False,
                      threads: int = 1):
    flat = _u8(flat)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    maxlen = int(np.max(np.diff(offsets.astype(np.int64)))) if n else 0
    stride = max(min(sketch_size, max(maxlen - k + 1, 0)), 1)
    sk = np.zeros((n, stride), dtype=np.uint32)
    lens = np.zeros(n, dtype=np.uint32)
    _check(lib().dvso_mash_sketch_batch(flat, offsets, n, k, sketch_size, num_states, int(canonical), sk, stride,
                                        lens, threads))
    return sk, lens


def mash_distance(a, b, k: int, sketch_size: int):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    x, u, err = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
    pa = a if len(a) else np.zeros(1, dtype=np.uint32)
    pb = b if len(b) else np.zeros(1, dtype=np.uint32)
    d = lib().dvso_mash_distance(pa, len(a), pb, len(b), k, sketch_size, C.byref(x), C.byref(u), C.byref(err))
    if err.value:
        raise ZeroDivisionError(lib().dvso_last_error().decode())
    return float(d), int(x.value), int(u.value)


def mash_matrix(sketches, lens, k: int, sketch_size: int, threads: int = 1):
    sk = np.ascontiguousarray(sketches, dtype=np.uint32)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    n = len(lens)
    dist = np.zeros((n, n), dtype=np.float64)
    inter = np.zeros((n, n), dtype=np.uint32)
    uni = np.zeros((n, n), dtype=np.uint32)
    _check(lib().dvso_mash_matrix(sk, sk.shape[1], lens, n, k, sketch_size, dist, _vptr(inter), _vptr(uni),
                                  threads))
    return dist, inter, uni


def euclid_matrix(rows, threads: int = 1) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n, d = rows.shape
    dist = np.zeros((n, n), dtype=np.float64)
    _check(lib().dvso_euclid_matrix(rows, n, d, dist, threads))
    return dist


def log2_port(x: float) -> float:
    return float(lib().dvso_log2_port(x))


def log2_libm(x: float) -> float:
    return float(lib().dvso_log2_libm(x))


def log2_port_mismatches(seed: int, n_per_thread: int, threads: int = 1):
    bad = C.c_double(0.0)
    m = lib().dvso_log2_port_mismatches(seed, n_per_thread, threads, C.byref(bad))
    return int(m), bad.value


def log2_samples(seed: int, n: int):
    xs = np.zeros(n, dtype=np.float64)
    ys = np.zeros(n, dtype=np.float64)
    lib().dvso_log2_samples(seed, n, xs, ys)
    return xs, ys


# ---- `dvs prep` encode ------------------------------------------------
# PARITY UNPINNED: cogent3 (the parser and the alphabet) is not installed here and the reference holds
# no golden vector for the encode step, so this restates the published behaviour of
#   diverse_seq/io.py:30-34   converter_fasta = convert_alphabet(a-z -> A-Z, delete=b"\n\r\t- ")
#   diverse_seq/io.py:47-57   cogent3.parse.fasta.iter_fasta_records(path, converter)
#   diverse_seq/io.py:95-104  dvs_load_seqs.main: b"-".join(seqs) -> str2arr
#   diverse_seq/util.py:32-45 str2arr: most_degen_alphabet().to_indices
# with bytes.split / bytes.translate, the way the parser itself is written.  The codes of the four
# canonical bases (T,C,A,G -> 0..3, src/distance.rs:6-8) are pinned by the reference's own tests; the
# order of the codes >= 4 follows cogent3's degenerate gapped DNA alphabet as remembered and does not
# influence any hot-path result (every code >= num_states is equally invalid).
DNA_ALPHABET = "TCAG-NRYWSKMBDHV?"
_FASTA_DELETE = b"\n\r\t- "


def _converter_fasta(body: bytes, delete: bytes = _FASTA_DELETE) -> bytes:
    import string

    table = bytes.maketrans(string.ascii_lowercase.encode(), string.ascii_uppercase.encode())
    return body.translate(table, delete=delete)


def iter_fasta_records(data: bytes, delete: bytes = _FASTA_DELETE):
    """(label, converted sequence bytes) of every '>'-delimited piece that has a label line"""
    for piece in data.split(b">"):
        if not piece:
            continue
        eol = piece.find(b"\n")
        if eol == -1:
            continue
        yield piece[:eol].strip().decode("utf8", errors="replace"), _converter_fasta(piece[eol + 1:], delete)


def prep_fasta(data: bytes, alphabet: str = DNA_ALPHABET, delete: bytes = _FASTA_DELETE) -> np.ndarray:
    """index-encoded record of one FASTA file (all its sequences joined with '-')"""
    joined = b"-".join(seq for _, seq in iter_fasta_records(data, delete))
    chars = alphabet.encode()
    table = bytes.maketrans(chars, bytes(range(len(chars))))  # bytes outside the alphabet map to themselves
    return np.frombuffer(joined.translate(table), dtype=np.uint8).copy()


# ---- average-linkage tree ----------------------------------------------
# Restatement of what diverse_seq/cluster.py:191-237 (make_cluster_tree) gets from
# sklearn.cluster.AgglomerativeClustering(metric="precomputed", linkage="average").children_:
# scipy's nn_chain + stable sort + label() on the upper triangle.  PINNED: tests/test_cluster.py checks
# it against scikit-learn itself (installed here and on the GPU box; it is the reference's own
# dependency), ties included.
def linkage_average(dist: np.ndarray):
    """(children (n-1,2) int, heights (n-1,), counts (n-1,)) of a symmetric n x n distance matrix"""
    d = np.array(dist, dtype=np.float64)
    n = d.shape[0]
    iu = np.triu_indices(n, 1)
    w = np.zeros((n, n))
    w[iu] = d[iu]
    w = w + w.T  # exact: one of the two addends is 0.0
    size = np.ones(n, dtype=np.int64)
    chain, merges = [], []
    for _ in range(n - 1):
        if not chain:
            chain.append(int(np.flatnonzero(size > 0)[0]))
        while True:
            x = chain[-1]
            if len(chain) > 1:
                y = chain[-2]
                cur = w[x, y]
            else:
                y, cur = -1, np.inf
            row = np.where((size > 0) & (np.arange(n) != x), w[x], np.inf)
            i = int(np.argmin(row))  # first index of the minimum
            if row[i] < cur:
                cur, y = row[i], i
            if len(chain) > 1 and y == chain[-2]:
                break
            chain.append(y)
        chain.pop()
        chain.pop()
        if x > y:
            x, y = y, x
        nx, ny = int(size[x]), int(size[y])
        merges.append((x, y, float(cur), nx + ny))
        size[x] = 0
        size[y] = nx + ny
        live = (size > 0) & (np.arange(n) != y)
        new = (nx * w[live, x] + ny * w[live, y]) / (nx + ny)
        w[live, y] = new
        w[y, live] = new
    order = sorted(range(n - 1), key=lambda r: merges[r][2])  # sorted() is stable
    parent = list(range(2 * n - 1))
    csize = [1] * (2 * n - 1)

    def find(a):
        while parent[a] != a:
            a = parent[a]
        return a

    children = np.zeros((n - 1, 2), dtype=np.int64)
    heights = np.zeros(n - 1)
    counts = np.zeros(n - 1, dtype=np.int64)
    nxt = n
    for r, idx in enumerate(order):
        x, y, h, _ = merges[idx]
        xr, yr = find(x), find(y)
        children[r] = (min(xr, yr), max(xr, yr))
        parent[xr] = parent[yr] = nxt
        csize[nxt] = csize[xr] + csize[yr]
        heights[r], counts[r] = h, csize[nxt]
        nxt += 1
    return children, heights, counts
