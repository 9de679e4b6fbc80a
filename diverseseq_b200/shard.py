"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed
for the plumbing (NCCL on GPUs, gloo in the CPU tests).

* counting / sketching: records are independent -> each rank counts its own shard, no collective.
* nmost / max, two modes:
  - "chunked" (default; the reference's own `-np N` semantics, diverse_seq/records.py:206-251): every
    rank selects from its own records, the N x n winning rows are all-gathered (n x 4^k f64 per rank)
    and merged with final_nmost / final_max (src/records.rs:363-382,456-507) - per-GPU work is
    constant, i.e. true weak scaling;
  - "union": the per-rank frequency rows (N x 4^k f64) are all-gathered once and every rank replays
    the single-pass selection on the full row set (numprocs=1 semantics over all records; replicated
    state, identical results on every rank, no per-step exchange).
* distance matrices: rows of the symmetric matrix are block-partitioned; each rank computes its
  row block against all columns and the blocks are all-gathered for the (CPU) clustering.

The functions here only touch torch tensors / python ints so they run unchanged under gloo.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous block partition [begin, end) of n_items over world ranks (sizes differ by <= 1)"""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def global_order(seed: int, n_total: int) -> np.ndarray:
    """the shuffled examination order, identical on every rank (cli.py:445-446 uses default_rng(seed))"""
    return np.random.default_rng(seed).permutation(n_total).astype(np.uint32)


def all_gather_concat(t, group=None):
    """all-gather equally shaped tensors and concatenate along dim 0 in rank order"""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group) if t.is_cuda else \
        dist.all_gather(list(out.chunk(world, dim=0)), t.contiguous(), group=group)
    return out


def max_over_ranks(value: float, device=None, group=None) -> float:
    """timing rule: a multi-GPU duration is the max over ranks"""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, device=None, group=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


class DeviceArray:
    """zero-copy torch view of a raw device pointer owned by libdvs_b200 (via __cuda_array_interface__)"""

    def __init__(self, ptr: int, shape: tuple, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def kfreqs_as_tensors(kf, device):
    """(rows [nrec, dim] f64, entropies [nrec] f64, valid [nrec] u8) torch views of a KFreqs"""
    import torch

    rows_p, ent_p, val_p = kf.device_ptrs()
    n, d = kf.nrec, kf.dim
    rows = torch.as_tensor(DeviceArray(rows_p, (n, d), "<f8"), device=device)
    ent = torch.as_tensor(DeviceArray(ent_p, (n,), "<f8"), device=device)
    valid = torch.as_tensor(DeviceArray(val_p, (n,), "|u1"), device=device)
    return rows, ent, valid


def all_gather_kfreqs(ctx, kf, device, group=None):
    """every rank ends up with a KFreqs holding the rows of all ranks, in rank order"""
    import torch

    from . import _lib

    rows, ent, valid = kfreqs_as_tensors(kf, device)
    ctx.sync()  # the rows were produced on the library's stream
    g_rows = all_gather_concat(rows, group)
    g_ent = all_gather_concat(ent, group)
    g_valid = all_gather_concat(valid, group)
    torch.cuda.synchronize(device)
    return _lib.KFreqs.from_device(ctx, g_rows.data_ptr(), g_ent.data_ptr(), g_valid.data_ptr(),
                                   g_rows.shape[0], g_rows.shape[1])


def chunked_select(ctx, kf, local_order, mode: int, min_size: int, max_size: int, device, group=None):
    """The reference's multi-process selection with one chunk per GPU: select locally, all-gather the
    winners' rows, merge with final_nmost / final_max on every rank (identical results everywhere).

    Returns (global ids, delta_jsd, stats5); a global id is rank * kf.nrec + local row."""
    import torch
    import torch.distributed as dist

    from . import _lib

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    idx, _delta, _stats = kf.select(local_order, mode, min_size, max_size)
    cap = max(int(min_size), int(max_size))
    won = kf.take_rows(idx)
    rows, _ent, _valid = kfreqs_as_tensors(won, device)
    pad = torch.zeros((cap, kf.dim), dtype=torch.float64, device=device)
    pad[: rows.shape[0]] = rows
    ids = torch.full((cap,), -1, dtype=torch.int64, device=device)
    ids[: len(idx)] = torch.from_numpy(idx.astype(np.int64) + rank * kf.nrec).to(device)
    ctx.sync()
    g_rows = all_gather_concat(pad, group)
    g_ids = all_gather_concat(ids, group)
    torch.cuda.synchronize(device)
    keep = torch.nonzero(g_ids >= 0).flatten()
    g_rows = g_rows[keep].contiguous()
    g_ids = g_ids[keep].cpu().numpy()
    # final_*: entropies recomputed from the stored rows (KmerSeq::new), examined in concatenation order
    merged = _lib.KFreqs.from_device(ctx, g_rows.data_ptr(), None, None, g_rows.shape[0], g_rows.shape[1])
    order = np.arange(g_rows.shape[0], dtype=np.uint32)
    midx, mdelta, mstats = merged.select(order, mode, min_size, max_size)
    return g_ids[midx], mdelta, mstats


def row_blocks(n: int, world: int, align: int = 128) -> list[tuple[int, int]]:
    """block partition of the rows of an n x n matrix; block starts are multiples of `align` (the
    Euclidean kernel mirrors tiles only when its row range is tile aligned)"""
    per = -(-n // world)
    per = -(-per // align) * align
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def gather_row_blocks(local_rows: np.ndarray, n: int, device, group=None) -> np.ndarray:
    """all-gather the per-rank row blocks of an n x n matrix (ctree's distance matrix) -> full matrix on
    every rank (`ncclAllGather` of row blocks before the CPU clustering)"""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    blocks = row_blocks(n, world)
    per = max(e - b for b, e in blocks)
    pad = torch.zeros((per, n), dtype=torch.float64, device=device)
    b, e = blocks[rank]
    if e > b:
        pad[: e - b] = torch.from_numpy(np.ascontiguousarray(local_rows)).to(device)
    full = all_gather_concat(pad, group)
    out = np.empty((n, n), dtype=np.float64)
    for r, (b, e) in enumerate(blocks):
        if e > b:
            out[b:e] = full[r * per: r * per + (e - b)].cpu().numpy()
    return out


def sharded_mash_distances(ctx, seqset_local, k: int, sketch_size: int, num_states: int, canonical: bool, device,
                           group=None):
    """ctree mash on N GPUs: every rank sketches its own records, the sketches are all-gathered
    (records are rank-ordered), every rank computes its row block of the matrix, blocks are gathered."""
    import torch
    import torch.distributed as dist

    from . import _lib

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sk = _lib.Sketches.sketch(ctx, seqset_local, k, sketch_size, num_states, canonical)
    data, lens = sk.download()
    stride = torch.tensor([data.shape[1]], dtype=torch.int64, device=device)
    dist.all_reduce(stride, op=dist.ReduceOp.MAX, group=group)
    wide = np.zeros((data.shape[0], int(stride.item())), dtype=np.uint32)
    wide[:, : data.shape[1]] = data
    g_data = all_gather_concat(torch.from_numpy(wide.view(np.int32)).to(device), group).cpu().numpy().view(np.uint32)
    g_lens = all_gather_concat(torch.from_numpy(lens.view(np.int32)).to(device), group).cpu().numpy().view(np.uint32)
    allsk = _lib.Sketches.from_host(ctx, g_data, g_lens)
    n = allsk.nrec
    b, e = row_blocks(n, world)[rank]
    local = allsk.distances(k, sketch_size, b, e) if e > b else np.zeros((0, n))
    return gather_row_blocks(local, n, device, group)


def sharded_euclidean(ctx, kf_local, device, group=None):
    """ctree Euclidean on N GPUs: all-gather the frequency rows, row block per rank, gather blocks"""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    allf = all_gather_kfreqs(ctx, kf_local, device, group)
    n = allf.nrec
    b, e = row_blocks(n, world)[rank]
    local = allf.euclidean(b, e) if e > b else np.zeros((0, n))
    return gather_row_blocks(local, n, device, group)
