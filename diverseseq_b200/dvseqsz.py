"""Reader (and spec-conformant writer) for the reference's `.dvseqsz` stores, feeding the GPU.

SURVEY.md §8(f) rank 1 — the data format on the input side of the hot path.  Layout written by
the reference through the `zarrs` crate (/root/reference/src/zarr_io.rs):

    <name>.dvseqsz/
      .seqid_to_hash.bin          postcard Vec<(String, [u8;16])>: varint(n), then per entry
                                  varint(len) + utf8 seqid + 16 raw bytes = ASCII hex of xxh3_64(data)
                                  (:57-60, :121-190, :221-224)
      seqdata/zarr.json           Zarr v3 group (:81-93)
      seqdata/<16-hex>/zarr.json  Zarr v3 array: shape [L], ONE chunk [L], uint8, fill 0,
                                  codecs bytes + zstd(level 3, checksum) (:237-258),
                                  attributes {"metadata": [postcard bytes of map<String,String>]}
      seqdata/<16-hex>/c/0        the single zstd frame

One array per unique sequence content; several seqids may share a digest (:217-235).
zstd comes from the system libzstd through ctypes (decompression releases the GIL, so records are
decoded on a thread pool while earlier batches are already being uploaded and counted).

Byte compatibility of the WRITER with zarrs' exact JSON spelling cannot be verified here (no
reference-written store and no zarrs in this image); the reader accepts any Zarr v3 array with a
single zstd (or uncompressed) chunk, the writer emits spec-conformant Zarr v3 and is used by the
round-trip tests.  Reading a store written by the reference itself is untested ("parity unpinned").
"""
from __future__ import annotations

import ctypes as C
import json
import os
import pathlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

__all__ = ["DvseqszStore", "load_seqset"]

_zstd = None


def _libzstd():
    global _zstd
    if _zstd is None:
        lib = C.CDLL("libzstd.so.1")
        lib.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        lib.ZSTD_decompress.restype = C.c_size_t
        lib.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ZSTD_isError.restype = C.c_uint
        lib.ZSTD_isError.argtypes = [C.c_size_t]
        lib.ZSTD_getErrorName.restype = C.c_char_p
        lib.ZSTD_getErrorName.argtypes = [C.c_size_t]
        lib.ZSTD_compressBound.restype = C.c_size_t
        lib.ZSTD_compressBound.argtypes = [C.c_size_t]
        lib.ZSTD_createCCtx.restype = C.c_void_p
        lib.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        lib.ZSTD_CCtx_setParameter.restype = C.c_size_t
        lib.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.ZSTD_compress2.restype = C.c_size_t
        lib.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        _zstd = lib
    return _zstd


_ZSTD_c_compressionLevel, _ZSTD_c_checksumFlag = 100, 201
_CONTENTSIZE_UNKNOWN, _CONTENTSIZE_ERROR = 2 ** 64 - 1, 2 ** 64 - 2


def zstd_decompress_into(frame: bytes, out: np.ndarray) -> int:
    """decode one zstd frame into `out` (uint8, C-contiguous); returns the number of bytes written"""
    lib = _libzstd()
    src = np.frombuffer(frame, dtype=np.uint8)
    n = lib.ZSTD_decompress(out.ctypes.data_as(C.c_void_p), out.size, src.ctypes.data_as(C.c_void_p), src.size)
    if lib.ZSTD_isError(n):
        raise RuntimeError(f"zstd: {lib.ZSTD_getErrorName(n).decode()}")
    return int(n)


def zstd_content_size(frame: bytes) -> int:
    lib = _libzstd()
    src = np.frombuffer(frame, dtype=np.uint8)
    n = lib.ZSTD_getFrameContentSize(src.ctypes.data_as(C.c_void_p), src.size)
    if n in (_CONTENTSIZE_UNKNOWN, _CONTENTSIZE_ERROR):
        raise RuntimeError("zstd frame without a content size")
    return int(n)


def zstd_compress(data: np.ndarray, level: int = 3, checksum: bool = True) -> bytes:
    lib = _libzstd()
    cctx = lib.ZSTD_createCCtx()
    try:
        lib.ZSTD_CCtx_setParameter(cctx, _ZSTD_c_compressionLevel, level)
        lib.ZSTD_CCtx_setParameter(cctx, _ZSTD_c_checksumFlag, int(checksum))
        cap = lib.ZSTD_compressBound(data.size)
        out = np.empty(cap, dtype=np.uint8)
        n = lib.ZSTD_compress2(cctx, out.ctypes.data_as(C.c_void_p), cap, data.ctypes.data_as(C.c_void_p), data.size)
        if lib.ZSTD_isError(n):
            raise RuntimeError(f"zstd: {lib.ZSTD_getErrorName(n).decode()}")
        return out[:n].tobytes()
    finally:
        lib.ZSTD_freeCCtx(cctx)


# ---- postcard (the subset used by the store: varint, str, [u8;16], map<str,str>) ---------------
def _varint_read(buf: bytes, pos: int) -> tuple[int, int]:
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _str(s: str) -> bytes:
    b = s.encode()
    return _varint(len(b)) + b


def decode_seqid_to_hash(buf: bytes) -> dict[str, str]:
    n, pos = _varint_read(buf, 0)
    out: dict[str, str] = {}
    for _ in range(n):
        ln, pos = _varint_read(buf, pos)
        seqid = buf[pos:pos + ln].decode()
        pos += ln
        out[seqid] = buf[pos:pos + 16].decode("ascii")
        pos += 16
    return out


def encode_seqid_to_hash(m: dict[str, str]) -> bytes:
    out = bytearray(_varint(len(m)))
    for seqid, hexd in m.items():
        out += _str(seqid) + hexd.encode("ascii")
    return bytes(out)


def decode_str_map(buf: bytes) -> dict[str, str]:
    n, pos = _varint_read(buf, 0)
    out = {}
    for _ in range(n):
        ln, pos = _varint_read(buf, pos)
        k = buf[pos:pos + ln].decode()
        pos += ln
        ln, pos = _varint_read(buf, pos)
        out[k] = buf[pos:pos + ln].decode()
        pos += ln
    return out


def encode_str_map(m: dict[str, str]) -> bytes:
    out = bytearray(_varint(len(m)))
    for k, v in m.items():
        out += _str(k) + _str(v)
    return bytes(out)


class DvseqszStore:
    """Directory store with the surface of the reference's ZarrStoreWrapper (src/zarr_py.rs:9-247)."""

    ROOT = "seqdata"
    SIDECAR = ".seqid_to_hash.bin"

    def __init__(self, path: str | os.PathLike, mode: str = "r"):
        self.path = pathlib.Path(path)
        self._mode = mode
        if mode == "r" and not self.path.is_dir():
            raise FileNotFoundError(str(self.path))
        if mode != "r":
            (self.path / self.ROOT).mkdir(parents=True, exist_ok=True)
            gmeta = self.path / self.ROOT / "zarr.json"
            if not gmeta.exists():
                gmeta.write_text(json.dumps({"zarr_format": 3, "node_type": "group", "attributes": {}}))
        side = self.path / self.SIDECAR
        self._seqid_to_hash: dict[str, str] = decode_seqid_to_hash(side.read_bytes()) if side.exists() else {}

    # -- reference surface --
    def __repr__(self) -> str:
        return f"ZarrStoreWrapper(source={self.path}, n={len(self)})"

    def __contains__(self, key: str) -> bool:
        return key in self._seqid_to_hash

    def __len__(self) -> int:
        return len(self._seqid_to_hash)

    @property
    def source(self) -> str:
        return str(self.path)

    def get_seqids(self) -> list[str]:
        return list(self._seqid_to_hash)

    @property
    def unique_seqids(self) -> list[str]:
        seen, out = set(), []
        for sid, h in self._seqid_to_hash.items():
            if h not in seen:
                seen.add(h)
                out.append(sid)
        return out

    def num_unique(self) -> int:
        return len(set(self._seqid_to_hash.values()))

    def _array_dir(self, seqid: str) -> pathlib.Path:
        try:
            return self.path / self.ROOT / self._seqid_to_hash[seqid]
        except KeyError:
            raise RuntimeError(f"Failed to create add {seqid}") from None

    def _meta(self, seqid: str) -> dict:
        return json.loads((self._array_dir(seqid) / "zarr.json").read_text())

    def length(self, seqid: str) -> int:
        return int(self._meta(seqid)["shape"][0])

    def read_into(self, seqid: str, out: np.ndarray) -> int:
        """decode the record into `out` (may be a slice of a pinned staging buffer)"""
        d = self._array_dir(seqid)
        meta = json.loads((d / "zarr.json").read_text())
        if meta.get("data_type") != "uint8" or len(meta["shape"]) != 1:
            raise RuntimeError(f"{seqid}: not a 1-D uint8 array")
        n = int(meta["shape"][0])
        if list(meta["chunk_grid"]["configuration"]["chunk_shape"]) != [n]:
            raise RuntimeError(f"{seqid}: expected a single chunk")
        sep = meta.get("chunk_key_encoding", {}).get("configuration", {}).get("separator", "/")
        chunk = (d / "c" / "0") if sep == "/" else (d / "c.0")
        if not chunk.exists():
            # zarrs does not store a chunk that consists of the fill value only (a record of code 0 = 'T')
            out[:n] = int(meta.get("fill_value", 0))
            return n
        frame = chunk.read_bytes()
        codecs = [c["name"] for c in meta.get("codecs", [])]
        if "zstd" in codecs:
            got = zstd_decompress_into(frame, out[:n])
        else:
            got = len(frame)
            if got == n:
                out[:n] = np.frombuffer(frame, dtype=np.uint8)
        if got != n:
            raise RuntimeError(f"{seqid}: decoded {got} bytes, expected {n}")
        return n

    def _array(self, seqid: str) -> np.ndarray:
        out = np.empty(self.length(seqid), dtype=np.uint8)
        self.read_into(seqid, out)
        return out

    def read(self, seqid: str) -> bytes:
        return self._array(seqid).tobytes()

    def read_metadata(self, seqid: str) -> dict:
        md = self._meta(seqid).get("attributes", {}).get("metadata")
        return decode_str_map(bytes(md)) if md is not None else {}

    def get_lazyseq(self, seqid: str, num_states: int):
        from ._dvs import LazySeq
        return LazySeq(seqid, self, num_states)

    def get_lazyseqs(self, num_states: int):
        return [self.get_lazyseq(s, num_states) for s in self.get_seqids()]

    def write_log(self, unique_id: str, data: str) -> None:
        return None

    def write_citations(self, data) -> None:
        return None

    # -- writer (add_uint8_array, src/zarr_io.rs:211-282) --
    def write(self, seqid: str, seq, metadata: dict | None = None) -> None:
        import xxhash

        if self._mode == "r":
            raise ValueError(f"Failed to create add {seqid}")
        data = np.frombuffer(bytes(seq), dtype=np.uint8) if not isinstance(seq, np.ndarray) else \
            np.ascontiguousarray(seq, dtype=np.uint8)
        if data.size == 0:
            raise ValueError(f"Failed to create add {seqid}")
        if seqid in self._seqid_to_hash:
            return
        hexd = xxhash.xxh3_64_hexdigest(data.tobytes())
        known = hexd in self._seqid_to_hash.values()
        self._seqid_to_hash[seqid] = hexd
        if not known:
            d = self.path / self.ROOT / hexd
            (d / "c").mkdir(parents=True, exist_ok=True)
            (d / "c" / "0").write_bytes(zstd_compress(data, 3, True))
            n = int(data.size)
            meta = {"zarr_format": 3, "node_type": "array", "shape": [n], "data_type": "uint8",
                    "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [n]}},
                    "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
                    "fill_value": 0,
                    "codecs": [{"name": "bytes"}, {"name": "zstd", "configuration": {"level": 3, "checksum": True}}],
                    "attributes": {"metadata": list(encode_str_map(metadata or {"source": "unknown"}))}}
            (d / "zarr.json").write_text(json.dumps(meta))
        self.save_metadata()

    def save_metadata(self) -> None:
        tmp = self.path / (self.SIDECAR + ".tmp")  # atomic rename like src/zarr_io.rs:121-190
        tmp.write_bytes(encode_seqid_to_hash(self._seqid_to_hash))
        os.replace(tmp, self.path / self.SIDECAR)


def load_seqset(ctx, store: DvseqszStore, seqids, threads: int | None = None, pinned: bool = True):
    """Decode `seqids` (distinct) from the store straight into one (pinned) staging buffer on a thread
    pool and upload it as a device-resident SeqSet.  Returns (SeqSet, offsets)."""
    from . import _lib

    lens = [store.length(s) for s in seqids]
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens, dtype=np.uint64)
    total = int(offsets[-1])
    host = None
    if pinned:
        try:
            host = _lib.pinned_array(max(total, 1))
        except Exception:
            host = None
    if host is None:
        host = np.empty(max(total, 1), dtype=np.uint8)

    def decode(i: int) -> None:
        store.read_into(seqids[i], host[int(offsets[i]):int(offsets[i + 1])])

    with ThreadPoolExecutor(max_workers=threads or min(32, (os.cpu_count() or 4))) as pool:
        list(pool.map(decode, range(len(seqids))))
    return _lib.SeqSet.upload(ctx, host[:total], offsets), offsets
