#!/usr/bin/env python3
"""Builds tests/golden/brca1.npz from the reference's own fixture.

Source: /root/reference/tests/data/brca1.fasta (55 aligned BRCA1 sequences; identical
to diverse_seq/data/brca1.fa).  Processing mirrors what the reference's tests do before
the hot path sees the data (tests/conftest.py:17-27: `get_dataset("brca1").degap()` then
`numpy.array(seq)`), i.e. gaps removed, then cogent3's DNA `most_degen_alphabet()`
indices: T,C,A,G -> 0,1,2,3 and every other symbol >= 4 (diverse_seq/util.py:33-45,
src/distance.rs:6-8).  The exact index of an ambiguity code is irrelevant to the hot
path (everything >= num_states is "invalid"); cogent3's order "TCAG-?" + IUPAC is
approximated by putting them at 6.. in the order below.

Run here (needs /root/reference); the .npz is committed because /root/reference does
not exist on the GPU box.
"""
import pathlib

import numpy as np

SRC = pathlib.Path("/root/reference/tests/data/brca1.fasta")
OUT = pathlib.Path(__file__).resolve().parent / "brca1.npz"
ALPHABET = "TCAG-?" + "RYMKSWHBVDN"


def main() -> None:
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(ALPHABET):
        lut[ord(ch)] = i
        lut[ord(ch.lower())] = i
    lut[ord("U")] = lut[ord("u")] = 0
    names, seqs, cur = [], [], []
    for line in SRC.read_text().splitlines():
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur))
                cur = []
            names.append(line[1:].strip())
        else:
            cur.append(line.strip())
    seqs.append("".join(cur))
    assert len(names) == len(seqs) == 55
    arrays = []
    for s in seqs:
        s = s.replace("-", "").replace("?", "")  # degap()
        a = lut[np.frombuffer(s.encode(), dtype=np.uint8)]
        assert a.max() < 255, set(s)
        arrays.append(a)
    offsets = np.zeros(len(arrays) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(a) for a in arrays])
    np.savez_compressed(OUT, names=np.array(names), data=np.concatenate(arrays), offsets=offsets)
    print(OUT, len(names), int(offsets[-1]))


if __name__ == "__main__":
    main()
