"""`dvs prep` encode: FASTA text -> index-encoded records.

CPU part: the oracle restatement (oracle/oracle.py::prep_fasta, split/translate like the cogent3 parser
the reference calls, diverse_seq/io.py:30-57,95-104) on hand-checked cases.  GPU part: dvs_prep_fasta
(csrc/prep.cu, a chunked state machine) against that oracle, bit-exact, on edge cases, fuzzed text
that puts labels across chunk boundaries, and a wrapped/lower-cased rendering of the brca1 fixture.
Parity of the oracle itself is UNPINNED for this row (no cogent3 here, no golden vector upstream).
"""
import numpy as np
import pytest

from oracle import oracle

T, C, A, G, GAP, N = 0, 1, 2, 3, 4, 5


def test_oracle_prep_known_answers():
    assert oracle.prep_fasta(b">s1 human\nACGT\nacgt\n").tolist() == [A, C, G, T, A, C, G, T]
    # several sequences in one file are joined by '-', gaps/whitespace inside a sequence are deleted
    assert oracle.prep_fasta(b">a\nAC-GT \r\n>b\nTT\tN\n").tolist() == [A, C, G, T, GAP, T, T, N]
    # bytes outside the alphabet keep their (upper-cased) value, as bytes.translate does
    assert oracle.prep_fasta(b">a\nAxC*\n").tolist() == [A, ord("X"), C, ord("*")]
    # a piece without a newline is dropped; text before the first '>' is a piece of its own
    assert oracle.prep_fasta(b"\n>a\nAC>b>c\nGG").tolist() == [GAP, A, C, GAP, G, G]
    assert oracle.prep_fasta(b"").size == 0
    assert oracle.prep_fasta(b">only a label").size == 0
    assert oracle.prep_fasta(b">x\n").size == 0
    assert oracle.prep_fasta(b"ACGT").size == 0  # no newline at all: no label line, nothing yielded
    assert oracle.prep_fasta(b"label\nACGT").tolist() == [A, C, G, T]


def test_oracle_prep_matches_hot_path_alphabet():
    # T,C,A,G -> 0..3 is what src/distance.rs:6-8 and tests/test_util.py:9-17 pin; everything else > 3
    enc = oracle.prep_fasta(b">a\nTCAGNRY?-X\n")
    assert enc[:4].tolist() == [0, 1, 2, 3]
    assert (enc[4:] > 3).all()


# ------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from diverseseq_b200 import _lib

    return _lib.Context(0)


def _check(ctx, blobs, alphabet=None, delete=None):
    from diverseseq_b200 import _lib

    text, offsets = _lib.concat([np.frombuffer(b, dtype=np.uint8) for b in blobs])
    ss = _lib.SeqSet.prep_fasta(ctx, text, offsets, alphabet=alphabet, delete_chars=delete)
    off = ss.offsets()
    flat = ss.download()
    assert ss.nrec == len(blobs)
    for i, b in enumerate(blobs):
        want = oracle.prep_fasta(b, alphabet or oracle.DNA_ALPHABET, delete or b"\n\r\t- ")
        got = flat[int(off[i]):int(off[i + 1])]
        assert got.size == want.size, (i, got.size, want.size, b[:80])
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, (i, bad[:5], got[bad[:5]], want[bad[:5]])
    return ss


@gpu
def test_prep_edge_cases(ctx):
    _check(ctx, [b">s1 human\nACGT\nacgt\n", b">a\nAC-GT \r\n>b\nTT\tN\n", b">a\nAxC*\n", b"\n>a\nAC>b>c\nGG",
                 b"", b">only a label", b">x\n", b"ACGT", b"label\nACGT", b">a\r\nAC\r\nGT\r\n", b">\n>\n>\nA",
                 b">>>>\n\n\n>", b"\n", b">", b">a\nACGT", b"-\n-", bytes(range(256)) * 3, b">a\n" + bytes(range(256))])


@gpu
def test_prep_empty_batch_and_empty_files(ctx):
    from diverseseq_b200 import _lib

    ss = _lib.SeqSet.prep_fasta(ctx, np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert ss.nrec == 0 and ss.total_bases == 0
    ss = _check(ctx, [b"", b"", b""])
    assert ss.total_bases == 0


def _fuzz(rng, n, p_gt, p_nl):
    letters = np.frombuffer(b"ACGTacgtNnRYxX*- \t\r", dtype=np.uint8)
    t = letters[rng.integers(0, letters.size, n)]
    u = rng.random(n)
    t = np.where(u < p_gt, ord(">"), t)
    t = np.where((u >= p_gt) & (u < p_gt + p_nl), ord("\n"), t)
    return t.astype(np.uint8).tobytes()


@gpu
@pytest.mark.parametrize("p_gt,p_nl", [(0.2, 0.2), (0.01, 0.02), (1e-4, 1 / 61), (1e-5, 1e-5), (0.0, 0.0), (0.5, 0.0),
                                        (0.0, 0.5), (1e-5, 0.0)])
def test_prep_fuzz_vs_oracle(ctx, p_gt, p_nl):
    rng = np.random.default_rng(int(p_gt * 1e6) * 7919 + int(p_nl * 1e6) + 5)
    sizes = [0, 1, 3, 4, 5, 127, 128, 129, 511, 513, 4097, 32767, 32768, 32769, 65536, 100_003, 300_001]
    _check(ctx, [_fuzz(rng, n, p_gt, p_nl) for n in sizes])


@gpu
@pytest.mark.parametrize("alphabet,delete", [("ABCD-", None), ("TCAG-N", b"\n\r\t- Xx"), ("TCAG", b"\r\t "),
                                             ("GATC-", b"\n")])
def test_prep_table_only_path(ctx, alphabet, delete):
    """alphabets / delete sets the SWAR classification cannot serve run every byte through the table"""
    rng = np.random.default_rng(17)
    blobs = [_fuzz(rng, n, 0.01, 0.02) for n in (0, 5, 600, 40_000, 70_001)]
    blobs.append(b">a\nABCDabcdXx-\n>b\nTCAGtcag\n")
    _check(ctx, blobs, alphabet, delete)


@gpu
def test_prep_long_labels_across_chunks(ctx):
    # label lines longer than a chunk (32 KB) and bodies without any newline
    rng = np.random.default_rng(11)
    body = _fuzz(rng, 200_000, 0.0, 0.0)
    blobs = [b">" + b"x" * 70_000 + b"\n" + body + b"\n>" + b"y" * 40_000,
             b"z" * 33_000 + b"\n" + body[:1000] + b">" + b"w" * 32_768 + b"\nAC",
             b">" + b"h" * 32_765 + b"\nACGT>" + b"k" * 5 + b"\nGG"]
    _check(ctx, blobs)


def _render(seq, width, rng):
    letters = np.frombuffer(b"TCAGN", dtype=np.uint8)[np.minimum(seq, 4)]
    lower = rng.random(letters.size) < 0.3
    letters = np.where(lower, letters | 0x20, letters).astype(np.uint8)
    lines = [letters[i:i + width].tobytes() for i in range(0, letters.size, width)]
    return b"\n".join(lines) + b"\n"


@gpu
def test_prep_roundtrip_brca1(ctx, brca1):
    """FASTA rendering (wrapped, mixed case, one file per sequence and all-in-one) -> same indices"""
    rng = np.random.default_rng(3)
    names = list(brca1)
    files = [b">" + n.encode() + b" some description\n" + _render(brca1[n], 60, rng) for n in names]
    ss = _check(ctx, files)
    off, flat = ss.offsets(), ss.download()
    for i, n in enumerate(names):
        want = np.where(brca1[n] > 3, N, brca1[n])
        assert np.array_equal(flat[int(off[i]):int(off[i + 1])], want)
    # one multi-FASTA file: a single record, sequences joined by the gap code
    ss = _check(ctx, [b"".join(files)])
    want = np.concatenate([np.concatenate([np.where(brca1[n] > 3, N, brca1[n]), [GAP]]) for n in names])[:-1]
    assert np.array_equal(ss.download(), want)


@gpu
def test_prep_large_matches_oracle_and_counts(ctx):
    """8 genomes of ~3 Mbp as 80-column FASTA; the encoded records feed the counting kernel"""
    from diverseseq_b200 import _lib

    flat, offsets = _lib.synth_host(77, 8, 4, 3_000_000)
    rng = np.random.default_rng(9)
    files = []
    for i in range(8):
        seq = flat[int(offsets[i]):int(offsets[i + 1])]
        cut = seq.size // 3
        files.append(b">contig1\n" + _render(seq[:cut], 80, rng) + b">contig2 x\n" + _render(seq[cut:], 80, rng))
    ss = _check(ctx, files)
    kf = _lib.KFreqs.count(ctx, ss, 6)
    counts = kf.download()[0]
    for i in (0, 7):
        want = oracle.kcounts(oracle.prep_fasta(files[i]), 6)
        assert np.array_equal(counts[i], want)


@gpu
def test_prep_device_text_pointer(ctx):
    import torch

    from diverseseq_b200 import _lib

    blobs = [b">a\nACGTNN\nacg\n", b">b\nTTTT\n>c\nGG\n"]
    text, offsets = _lib.concat([np.frombuffer(b, dtype=np.uint8) for b in blobs])
    d = torch.zeros(text.size + 64, dtype=torch.uint8, device="cuda:0")
    d[:text.size] = torch.from_numpy(text.copy()).cuda()
    torch.cuda.synchronize()
    ss = _lib.SeqSet.prep_fasta(ctx, None, offsets, device_ptr=d.data_ptr())
    want = np.concatenate([oracle.prep_fasta(b) for b in blobs])
    assert np.array_equal(ss.download(), want)


@gpu
def test_prep_custom_alphabet_and_errors(ctx):
    from diverseseq_b200 import _lib

    text = np.frombuffer(b">a\nUCAGT\n", dtype=np.uint8)
    off = np.array([0, text.size], dtype=np.uint64)
    ss = _lib.SeqSet.prep_fasta(ctx, text, off, alphabet="UCAG-")
    assert ss.download().tolist() == [0, 1, 2, 3, ord("T")]
    with pytest.raises((ValueError, TypeError)):
        _lib.SeqSet.prep_fasta(ctx, text, off, alphabet="")
    with pytest.raises((ValueError, TypeError)):
        _lib.SeqSet.prep_fasta(ctx, text, np.array([5, 2], dtype=np.uint64))


@gpu
def test_prep_directory_to_store_and_select(ctx, tmp_path, brca1):
    """`dvs prep` directory mode into a .dvseqsz store, then nmost from the store == from the arrays"""
    from diverseseq_b200 import _dvs, prep

    rng = np.random.default_rng(5)
    seqdir = tmp_path / "seqs"
    seqdir.mkdir()
    names = list(brca1)[:20]
    for n in names:
        (seqdir / f"{n}.fa").write_bytes(b">" + n.encode() + b"\n" + _render(brca1[n], 70, rng))
    store = prep.prep(seqdir, tmp_path / "out", suffix="fa", ctx=ctx)
    assert sorted(store.get_seqids()) == sorted(names)
    for n in names:
        assert np.array_equal(np.frombuffer(store.read(n), dtype=np.uint8), np.where(brca1[n] > 3, N, brca1[n]))
    mem = _dvs.make_zarr_store()
    for n in names:
        mem.write(n, brca1[n].tobytes())
    order = sorted(names)
    a = _dvs.nmost_divergent(store, 5, 4, seqids=order)
    b = _dvs.nmost_divergent(mem, 5, 4, seqids=order)
    assert a.record_names == b.record_names and a.total_jsd == b.total_jsd
    with pytest.raises(FileExistsError):
        prep.prep(seqdir, tmp_path / "out", suffix="fa", ctx=ctx)


@gpu
def test_prep_single_file_mode(ctx, tmp_path, brca1):
    from diverseseq_b200 import prep

    rng = np.random.default_rng(6)
    names = list(brca1)[:7]
    text = b"".join(b">" + n.encode() + b" desc\n" + _render(brca1[n], 60, rng) for n in names)
    text += b">" + names[0].encode() + b" desc\nACGT\n"  # a repeated label: the later sequence wins
    p = tmp_path / "all.fasta"
    p.write_bytes(text)
    labels, ss = prep.encode_records(ctx, p)
    assert labels == [n + " desc" for n in names]
    off, flat = ss.offsets(), ss.download()
    assert flat[int(off[0]):int(off[1])].tolist() == [A, C, G, T]
    for i, n in enumerate(names[1:], start=1):
        assert np.array_equal(flat[int(off[i]):int(off[i + 1])], np.where(brca1[n] > 3, N, brca1[n]))
