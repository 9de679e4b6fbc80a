"""world_size-2 CPU tests of the multi-GPU host logic (diverseseq_b200/shard.py): the TCP rendezvous that
exchanges window handles / record counts, uneven record sharding (n % world != 0), the position-interleaved
global order, and that a selection over the union of the shards' rows is identical on every rank and equal to
the single-process result.  torch.distributed (gloo) is only bench.py's timing plumbing (max / sum over
ranks); it is exercised here as well.  The CUDA kernels themselves are covered by the -m gpu tests."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import bench
    from diverseseq_b200 import _lib, shard
    from oracle import oracle as orc

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rv = shard.Rendezvous()  # MASTER_PORT + 1
        nrec_total, k = 41, 3  # 41 % 2 != 0: shards of 21 and 20 records
        flat, off = _lib.synth_host(11, nrec_total, 4, 3000)
        b, e = shard.shard_bounds(nrec_total, world, rank)
        sub_off = (off[b:e + 1] - off[b]).astype(np.uint64)
        sub = flat[int(off[b]):int(off[e])]
        _, freqs, ent, valid = orc.count_batch(sub, sub_off, k)  # this rank's shard ("prep" is record-sharded)
        counts = rv.allgather(int(e - b))
        parts = rv.allgather((freqs, ent, valid))
        g_rows = np.concatenate([p[0] for p in parts])
        g_ent = np.concatenate([p[1] for p in parts])
        g_valid = np.concatenate([p[2] for p in parts])
        local_orders = rv.allgather(np.random.default_rng(100 + rank).permutation(e - b))
        order = shard.interleaved_order(local_orders, counts)
        sel = orc.select_rows(g_rows, g_ent, order, "nmost", 6, valid=g_valid)
        rv.barrier()
        t = bench.max_over_ranks(10.0 + rank)
        s = bench.sum_over_ranks(1.5)
        q.put((rank, counts, order.tolist(), sel.ids.tolist(), sel.delta_jsd.tolist(), g_rows.tobytes(), t, s))
        rv.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_rendezvous_uneven_shards_and_replicated_selection():
    from diverseseq_b200 import _lib, shard
    from oracle import oracle as orc

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat, off = _lib.synth_host(11, 41, 4, 3000)
    _, freqs, ent, valid = orc.count_batch(flat, off, 3)
    orders = [np.random.default_rng(100 + r).permutation(n) for r, n in enumerate((21, 20))]
    order = shard.interleaved_order(orders, [21, 20])
    exp = orc.select_rows(freqs, ent, order, "nmost", 6, valid=valid)
    for rank, counts, got_order, ids, delta, rows_bytes, t, s in res:
        assert counts == [21, 20] and got_order == order.tolist()
        assert ids == exp.ids.tolist() and delta == exp.delta_jsd.tolist()
        assert rows_bytes == freqs.tobytes()  # gathered in rank order == unsharded rows
        assert t == 11.0 and s == 3.0


def test_interleaved_order_alternates_ranks_and_covers_every_row():
    from diverseseq_b200 import shard
    npr = [5, 3, 0, 4]
    orders = [np.random.default_rng(r).permutation(n) for r, n in enumerate(npr)]
    order = shard.interleaved_order(orders, npr)
    assert sorted(order.tolist()) == list(range(12))
    base = np.concatenate([[0], np.cumsum(npr)])
    owner = np.searchsorted(base, order, side="right") - 1
    assert owner[:6].tolist() == [0, 1, 3, 0, 1, 3]  # position-cyclic over the ranks that still have records
    for r in range(4):  # each rank's records keep their own order
        assert (order[owner == r] - base[r]).tolist() == orders[r].tolist()


def test_local_group_allgather_between_threads():
    import threading
    from diverseseq_b200 import shard
    g = shard.LocalGroup(3)
    out = [None] * 3

    def w(r):
        m = g.member(r)
        out[r] = (m.allgather(r * 10), m.allgather(str(r)))
        m.barrier()

    ts = [threading.Thread(target=w, args=(r,)) for r in range(3)]
    [t.start() for t in ts]
    [t.join(30) for t in ts]
    assert out == [([0, 10, 20], ["0", "1", "2"])] * 3


@pytest.mark.parametrize("n,world", [(10, 3), (7, 8), (0, 2), (16, 4)])
def test_shard_bounds_partition(n, world):
    from diverseseq_b200 import shard
    spans = [shard.shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1
