import os
import pathlib
import sys

# The multi-GPU tests run several ranks as threads on ONE GPU; their kernels wait for each other, so their
# streams must not share a hardware queue (8 by default): a spinning kernel at the head of a shared queue
# would hold back another rank's copies.  Must be set before the CUDA runtime initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def brca1():
    """The reference's brca1 fixture, degapped + encoded (tests/golden/make_brca1_fixture.py)."""
    z = np.load(ROOT / "tests" / "golden" / "brca1.npz")
    names = [str(n) for n in z["names"]]
    data, offsets = z["data"], z["offsets"].astype(np.int64)
    return {n: np.ascontiguousarray(data[offsets[i]:offsets[i + 1]]) for i, n in enumerate(names)}


def random_seqs(rng, nrec, lo, hi, invalid_rate=0.01, num_states=4):
    """ragged uint8 records with some invalid bytes; family structure so JSDs are not all ties"""
    out = []
    base = rng.integers(0, num_states, size=hi, dtype=np.uint8)
    for _ in range(nrec):
        n = int(rng.integers(lo, hi + 1))
        s = base[:n].copy()
        mut = rng.random(n) < rng.choice([0.01, 0.1, 0.5, 1.0])
        s[mut] = rng.integers(0, num_states, size=int(mut.sum()), dtype=np.uint8)
        bad = rng.random(n) < invalid_rate
        s[bad] = rng.integers(num_states, num_states + 3, size=int(bad.sum()), dtype=np.uint8)
        out.append(s)
    return out
