"""CPU tests of the `.dvseqsz` store module: layout, postcard sidecar, zstd frames,
dedup and the reference's store semantics (tests/test_zarr_store.py upstream)."""
import json

import numpy as np
import pytest

from diverseseq_b200 import _dvs, dvseqsz


def test_postcard_roundtrip():
    m = {"seq1": "0123456789abcdef", "a-longer-seqid" * 12: "fedcba9876543210", "": "0" * 16}
    buf = dvseqsz.encode_seqid_to_hash(m)
    assert buf[0] == 3 and dvseqsz.decode_seqid_to_hash(buf) == m
    # a 168-byte seqid needs a two-byte varint length
    assert dvseqsz._varint(168) == bytes([0xA8, 0x01]) and dvseqsz._varint_read(bytes([0xA8, 0x01]), 0) == (168, 2)
    md = {"source": "brca1-dataset:Human", "moltype": "dna"}
    assert dvseqsz.decode_str_map(dvseqsz.encode_str_map(md)) == md


def test_zstd_frame_has_content_size_and_checksum():
    rng = np.random.default_rng(0)
    data = rng.integers(0, 4, size=100_000, dtype=np.uint8)
    frame = dvseqsz.zstd_compress(data, 3, True)
    assert frame[:4] == bytes([0x28, 0xB5, 0x2F, 0xFD])  # zstd magic
    assert frame[4] & 0x04  # frame header descriptor: content checksum flag (codec checksum=true upstream)
    assert dvseqsz.zstd_content_size(frame) == data.size and len(frame) < data.size // 3
    out = np.empty(data.size, dtype=np.uint8)
    assert dvseqsz.zstd_decompress_into(frame, out) == data.size and np.array_equal(out, data)
    bad = bytearray(frame)
    bad[len(bad) // 2] ^= 0xFF
    with pytest.raises(RuntimeError):
        dvseqsz.zstd_decompress_into(bytes(bad), out)


def test_store_layout_roundtrip_and_dedup(tmp_path, brca1):
    path = tmp_path / "brca1.dvseqsz"
    st = _dvs.make_zarr_store(str(path), mode="w")
    for name in ("Human", "Chimpanzee", "Dugong"):
        st.write(name, brca1[name].tobytes(), {"source": f"brca1-dataset:{name}"})
    st.write("HumanCopy", brca1["Human"].tobytes())          # same content -> same array
    st.write("Human", brca1["Dugong"].tobytes())              # existing seqid is skipped (zarr_io.rs:217-219)
    with pytest.raises(ValueError):
        st.write("empty", b"")                               # tests/test_zarr_store.py:19-22
    assert len(st) == 4 and st.num_unique() == 3 and "Human" in st and "nope" not in st
    # on-disk layout
    assert (path / ".seqid_to_hash.bin").exists() and (path / "seqdata" / "zarr.json").exists()
    import xxhash
    hexd = xxhash.xxh3_64_hexdigest(brca1["Human"].tobytes())
    meta = json.loads((path / "seqdata" / hexd / "zarr.json").read_text())
    assert meta["zarr_format"] == 3 and meta["shape"] == [brca1["Human"].size] and meta["data_type"] == "uint8"
    assert meta["chunk_grid"]["configuration"]["chunk_shape"] == meta["shape"]
    assert [c["name"] for c in meta["codecs"]] == ["bytes", "zstd"]
    assert (path / "seqdata" / hexd / "c" / "0").exists()
    assert len(list((path / "seqdata").iterdir())) == 4  # 3 arrays + zarr.json
    # reopen read-only
    ro = _dvs.make_zarr_store(str(path))
    assert ro.get_seqids() == ["Human", "Chimpanzee", "Dugong", "HumanCopy"]
    assert ro.unique_seqids == ["Human", "Chimpanzee", "Dugong"]
    assert _dvs.get_seqids_from_store(str(path)) == ro.get_seqids()
    assert ro.read("HumanCopy") == brca1["Human"].tobytes() and ro.read("Dugong") == brca1["Dugong"].tobytes()
    assert ro.read_metadata("Chimpanzee") == {"source": "brca1-dataset:Chimpanzee"}
    assert ro.get_lazyseq("Human", 4).get_seq() == brca1["Human"].tobytes()
    with pytest.raises(RuntimeError):
        ro.read("missing")
    with pytest.raises(ValueError):
        ro.write("x", b"\x00\x01")
    with pytest.raises(FileNotFoundError):
        _dvs.make_zarr_store(str(tmp_path / "absent.dvseqsz"))


def test_all_fill_value_record_without_chunk_file_reads_as_zeros(tmp_path):
    """zarrs omits a chunk made only of the fill value (0 = 'T'): the reader must return zeros, not raise"""
    path = tmp_path / "t.dvseqsz"
    st = _dvs.make_zarr_store(str(path), mode="w")
    st.write("polyT", bytes(1000))
    st.write("other", bytes([0, 1, 2, 3] * 10))
    import xxhash
    hexd = xxhash.xxh3_64_hexdigest(bytes(1000))
    (path / "seqdata" / hexd / "c" / "0").unlink()  # what a reference-written store looks like for this record
    ro = _dvs.make_zarr_store(str(path))
    assert ro.read("polyT") == bytes(1000) and ro.read("other") == bytes([0, 1, 2, 3] * 10)
    # a raw (codec-less) frame of the wrong length is an error, not a silent short read
    meta_p = path / "seqdata" / xxhash.xxh3_64_hexdigest(bytes([0, 1, 2, 3] * 10)) / "zarr.json"
    meta = json.loads(meta_p.read_text())
    meta["codecs"] = [c for c in meta["codecs"] if c["name"] != "zstd"]
    meta_p.write_text(json.dumps(meta))
    with pytest.raises(RuntimeError):
        _dvs.make_zarr_store(str(path)).read("other")
