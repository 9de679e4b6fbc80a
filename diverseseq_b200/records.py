"""Host mirror of the `dvs nmost` / `dvs max` command flow around the hot path
(/root/reference/diverse_seq/cli.py:395-483): shuffle the seqids with the seed, run the selection, and - with
`include` - the re-run that adds the user's records to the selected set (cli.py:465-474).  Everything numeric
goes through diverseseq_b200._dvs (the CUDA library); this module only reproduces the order of the calls.
"""
from __future__ import annotations

import numpy as np

from . import _dvs


def shuffled_seqids(seqids, seed: int | None, limit: int | None = None) -> list[str]:
    """cli.py:445-448: numpy default_rng(seed).shuffle on the list of seqids, then the optional limit"""
    seqids = list(seqids)
    rng = np.random.default_rng(seed=seed)
    rng.shuffle(seqids)
    return seqids if limit is None else seqids[:limit]


def include_rerun(store, result, include, k: int, num_states: int = 4):
    """cli.py:465-474: the user's inclusions are appended to the selected names and nmost is run over exactly
    those names with n = their number, i.e. SummedRecords::new over all of them (src/records.rs:299-309; a
    name that is listed twice contributes two rows, as upstream)."""
    names = list(result.record_names) + list(include)
    return _dvs.nmost_divergent(store, len(names), k, num_states=num_states, seqids=names)


def select_nmost(store, n: int, k: int, seed: int | None = None, include=None, limit: int | None = None,
                 num_states: int = 4):
    """`dvs nmost -n N -k K [--include ...]` with numprocs = 1 (cli.py:395-483)"""
    seqids = store.get_seqids() if hasattr(store, "get_seqids") else list(store.unique_seqids)
    if len(seqids) < n:
        raise ValueError(f"Num seqs={len(seqids)} < number={n}. Nothing to do!")
    if include and not set(include) <= set(seqids):
        raise ValueError(f"provided include={include!r} not in the sequence data")
    order = shuffled_seqids(seqids, seed, limit)
    result = _dvs.nmost_divergent(store, n, k, num_states=num_states, seqids=order)
    return include_rerun(store, result, include, k, num_states) if include else result


def select_max(store, min_size: int, max_size: int, k: int, stat: str = "stdev", seed: int | None = None,
               include=None, limit: int | None = None, num_states: int = 4):
    """`dvs max` with numprocs = 1 (cli.py:296-392): the same flow with max_divergent; the inclusion re-run of
    `max` is an nmost over the selected + included names as well (cli.py:373-382)"""
    seqids = store.get_seqids() if hasattr(store, "get_seqids") else list(store.unique_seqids)
    if max_size is not None and min_size > max_size:
        raise ValueError(f"min_size={min_size} cannot be greater than max_size={max_size}")  # cli.py:311
    if len(seqids) < min_size:
        raise ValueError(f"Num seqs={len(seqids)} < min_size={min_size}. Nothing to do!")
    if include and not set(include) <= set(seqids):
        raise ValueError(f"provided include={include!r} not in the sequence data")
    order = shuffled_seqids(seqids, seed, limit)
    result = _dvs.max_divergent(store, min_size, max_size if max_size is not None else len(order), k,
                                num_states=num_states, seqids=order, stat=stat)
    return include_rerun(store, result, include, k, num_states) if include else result
