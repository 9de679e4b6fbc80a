"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in diverseseq_b200/shard.py:
record sharding, rank-ordered all-gather of frequency rows, max-over-ranks timing, and that a
selection replayed on the gathered rows is identical on every rank and equal to the
single-process result (the CUDA selection itself is covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diverseseq_b200 import _lib, shard
        from oracle import oracle as orc

        nrec_total, k = 40, 3
        flat, off = _lib.synth_host(11, nrec_total, 4, 3000)
        b, e = shard.shard_bounds(nrec_total, world, rank)
        sub_off = (off[b:e + 1] - off[b]).astype(np.uint64)
        sub = flat[int(off[b]):int(off[e])]
        _, freqs, ent, valid = orc.count_batch(sub, sub_off, k)  # this rank's shard ("prep" is record-sharded)
        g_rows = shard.all_gather_concat(torch.from_numpy(freqs))
        g_ent = shard.all_gather_concat(torch.from_numpy(ent))
        g_valid = shard.all_gather_concat(torch.from_numpy(valid))
        order = shard.global_order(5, nrec_total)
        sel = orc.select_rows(g_rows.numpy(), g_ent.numpy(), order, "nmost", 6, valid=g_valid.numpy())
        t = shard.max_over_ranks(10.0 + rank)
        s = shard.sum_over_ranks(1.5)
        q.put((rank, sel.ids.tolist(), sel.delta_jsd.tolist(), g_rows.numpy().tobytes(), t, s, (b, e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_replicated_selection():
    from diverseseq_b200 import _lib, shard
    from oracle import oracle as orc

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat, off = _lib.synth_host(11, 40, 4, 3000)
    _, freqs, ent, valid = orc.count_batch(flat, off, 3)
    exp = orc.select_rows(freqs, ent, shard.global_order(5, 40), "nmost", 6, valid=valid)
    for rank, ids, delta, rows_bytes, t, s, bounds in res:
        assert ids == exp.ids.tolist() and delta == exp.delta_jsd.tolist()
        assert rows_bytes == freqs.tobytes()  # gathered in rank order == unsharded rows
        assert t == 11.0 and s == 3.0
    assert [r[6] for r in res] == [(0, 20), (20, 40)]


@pytest.mark.parametrize("n,world", [(10, 3), (7, 8), (0, 2), (16, 4)])
def test_shard_bounds_partition(n, world):
    from diverseseq_b200 import shard
    spans = [shard.shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1
