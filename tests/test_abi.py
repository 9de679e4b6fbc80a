"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/dvs_b200.h declares (no compute calls without a GPU), and the host-side mirror of the
reference interface keeps the reference's names / defaults."""
import inspect
import pathlib
import re

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _header_functions():
    text = (ROOT / "include" / "dvs_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dvs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from diverseseq_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dvs_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert b"sm_100a" in lib.dvs_version()


def test_no_cpu_fallback_without_gpu():
    from diverseseq_b200 import _lib
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Context(0)
    from diverseseq_b200 import _dvs
    st = _dvs.make_zarr_store()
    for n, s in (("a", b"\x00\x01\x02\x03"), ("b", b"\x00\x00\x02\x03"), ("c", b"\x01\x01\x02\x03")):
        st.write(n, s)
    with pytest.raises(RuntimeError):
        _dvs.nmost_divergent(st, n=2, k=1)  # the product path fails loudly instead of computing on the CPU


def test_product_never_imports_oracle():
    for p in (ROOT / "diverseseq_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".h"}:
            assert "oracle" not in p.read_text().replace("no CPU oracle", "").replace("CPU oracle", "").lower() \
                or p.name == "__init__.py", p


def test_dvs_surface_matches_reference_signatures():
    from diverseseq_b200 import _dvs
    sig = lambda f: str(inspect.signature(f))
    assert sig(_dvs.nmost_divergent) == "(store, n: 'int', k: 'int', num_states: 'int' = 4, seqids=None) -> 'SummedRecordsResult'"
    assert list(inspect.signature(_dvs.max_divergent).parameters) == \
        ["store", "min_size", "max_size", "k", "num_states", "seqids", "stat"]  # src/lib.rs:105-115
    assert inspect.signature(_dvs.max_divergent).parameters["stat"].default == "stdev"
    assert list(inspect.signature(_dvs.final_nmost).parameters) == ["records", "n"]
    assert list(inspect.signature(_dvs.final_max).parameters) == ["records", "min_size", "max_size", "stat"]
    assert list(inspect.signature(_dvs.mash_sketch).parameters) == \
        ["seq_array", "k", "sketch_size", "num_states", "mash_canonical"]  # src/distance.rs:136-137
    assert list(inspect.signature(_dvs.get_delta_jsd_calculator).parameters) == ["seqids_seqs", "k", "num_states"]
    r = _dvs.SummedRecordsResult()
    assert set(r.__getstate__()) == {"total_jsd", "records", "mean_delta_jsd", "std_delta_jsd", "cov_delta_jsd",
                                     "size", "k", "num_states"}  # src/records_py.rs:57-64
    import pickle
    r.records = [("x", [0.5, 0.5], 0.1)]
    r.size = 1
    q = pickle.loads(pickle.dumps(r))
    assert q.records == r.records and q.record_names == ["x"]


def test_inmemory_store_semantics():
    from diverseseq_b200 import _dvs
    st = _dvs.make_zarr_store()
    st.write("a", bytes([0, 1, 2, 3]), {"source": "t"})
    st.write("b", bytes([0, 1, 2, 3]))
    st.write("c", np.array([3, 3], dtype=np.uint8).tobytes())
    st.write("a", bytes([1, 1]))  # existing seqid is skipped (src/zarr_io.rs:217-219)
    assert len(st) == 3 and st.num_unique() == 2 and st.get_seqids() == ["a", "b", "c"]
    assert st.unique_seqids == ["a", "c"] and st.read("b") == bytes([0, 1, 2, 3])
    assert st.read_metadata("a") == {"source": "t"}
    with pytest.raises(ValueError):
        st.write("empty", b"")  # tests/test_zarr_store.py:19-22
    with pytest.raises(RuntimeError):
        st.read("nope")
    lz = st.get_lazyseqs(4)
    assert [x.seqid for x in lz] == ["a", "b", "c"] and lz[0].get_seq() == bytes([0, 1, 2, 3])


def test_synth_host_is_deterministic_and_structured():
    from diverseseq_b200 import _lib
    a, oa = _lib.synth_host(7, 40, 4, 20000)
    b, ob = _lib.synth_host(7, 40, 4, 20000)
    assert np.array_equal(a, b) and np.array_equal(oa, ob)
    lens = np.diff(oa.astype(np.int64))
    assert lens.min() >= 15000 and lens.max() <= 25000
    assert a.max() <= 4


def test_host_pack_for_packed_upload():
    """transfer encoding of the packed upload path (csrc/hostpack.cpp): 4 bases per byte + exceptions"""
    import ctypes as C
    from diverseseq_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(4)
    for n in (0, 1, 3, 4, 5, 31, 32, 33, 1000, 100_003):
        src = rng.integers(0, 4, size=n, dtype=np.uint8)
        bad = rng.random(n) < 0.01
        src[bad] = rng.integers(4, 256, size=int(bad.sum()), dtype=np.uint8)
        packed = np.zeros(max((n + 3) // 4, 1), dtype=np.uint8)
        pos = np.zeros(max(n, 1), dtype=np.uint32)
        val = np.zeros(max(n, 1), dtype=np.uint8)
        nexc = C.c_uint32(0)
        lib.dvs_debug_pack_host(_lib.ptr(src) if n else None, n, _lib.ptr(packed), _lib.ptr(pos), _lib.ptr(val),
                                pos.size, C.byref(nexc))
        assert nexc.value == int(bad.sum())
        clean = np.where(src > 3, 0, src)
        pad = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
        pad[:n] = clean
        q = pad.reshape(-1, 4)
        expect = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
        assert np.array_equal(packed[: expect.size], expect)
        order = np.argsort(pos[: nexc.value])
        assert np.array_equal(pos[: nexc.value][order], np.nonzero(bad)[0].astype(np.uint32))
        assert np.array_equal(val[: nexc.value][order], src[bad])
