"""Multi-GPU paths against the oracle: sharded counting with pipelined row pushes, the
candidate-sharded single-pass selection (numprocs=1 semantics over the union of all ranks' records - NOT the
-np merge), the -np chunk merge, and the ctree matrices with pairs / tiles dealt over the GPUs.

The ranks are host threads of this process, each with its own context, stream and peer window.  With 2+ GPUs
visible every rank takes its own device (peer access over NVLink); on a 1-GPU box all ranks share device 0 and
their persistent kernels are made co-resident by giving each a share of the SMs (DVS_SELECT_GRID) - the
exchange protocol, tags, slots and host control flow are the same code either way.  A second test drives
real processes through the TCP rendezvous and cudaIpc (tests/sharded_worker.py)."""
import os
import pathlib
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from diverseseq_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def device_count() -> int:
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 1
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 1


def run_ranks(world, fn, window_bytes=256 << 20):
    """fn(rank, ctx, comm) on `world` threads; returns the per-rank results"""
    from diverseseq_b200 import _lib, shard

    ndev = device_count()
    devices = list(range(world)) if ndev >= world else [0] * world
    old = os.environ.get("DVS_SELECT_GRID")
    os.environ["DVS_COMM_CHECK"] = "1"  # a timed-out device-side wait is reported by the call it happened in
    if ndev < world:
        os.environ["DVS_SELECT_GRID"] = str(max(8, 148 // world - 2))
    group = shard.LocalGroup(world)
    results, errors = [None] * world, []

    def worker(rank):
        try:
            ctx = _lib.Context(devices[rank])
            comm = shard.connect(ctx, group.member(rank), window_bytes)
            results[rank] = fn(rank, ctx, comm)
            ctx.sync()
            group.member(rank).barrier()
            comm.close()
        except BaseException as exc:  # noqa: BLE001
            errors.append((rank, exc))
            group._bar.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    if old is None:
        os.environ.pop("DVS_SELECT_GRID", None)
    else:
        os.environ["DVS_SELECT_GRID"] = old
    real = [e for e in errors if not isinstance(e[1], threading.BrokenBarrierError)]
    if real or errors:
        raise (real or errors)[0][1]
    return results


def make_shards(lib, seed, nrec_per_rank, nfam, mean_len):
    """per-rank (flat, offsets) of a synthetic set + the union in rank-major order"""
    total = sum(nrec_per_rank)
    shards, first = [], 0
    for n in nrec_per_rank:
        shards.append(lib.synth_host(seed, total, nfam, mean_len, first, n))
        first += n
    flat = np.concatenate([s[0] for s in shards])
    off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(s[1].astype(np.int64)) for s in shards]))]).astype(np.uint64)
    return shards, flat, off


@pytest.mark.parametrize("world,npr", [(2, [37, 52]), (3, [20, 0, 31])])
def test_count_sharded_rows_of_all_ranks_bit_identical(lib, orc, world, npr):
    """uneven shards (n % world != 0, an empty rank): every rank ends with every record's row, entropy, validity"""
    shards, flat, off = make_shards(lib, 77, npr, 5, 30_000)
    _, of, oe, ov = orc.count_batch(flat, off, 6, want_counts=False)

    def fn(rank, ctx, comm):
        from diverseseq_b200 import shard
        ss = lib.SeqSet.upload(ctx, *shards[rank])
        out = []
        for _ in range(2):  # twice: heap blocks and epochs are reused
            kf, nrec = shard.count_sharded(ctx, comm, ss, 6)
            assert nrec == npr and kf.nrec == sum(npr)
            out.append(kf.download(counts=False)[1:])
            kf.close()
        # the same through allgather of rows that already exist
        kl = lib.KFreqs.count(ctx, ss, 6)
        ka = kl.allgather(comm, npr)
        out.append(ka.download(counts=False)[1:])
        ka.close()
        return out

    for res in run_ranks(world, fn):
        for f, e, v in res:
            assert np.array_equal(v, ov) and np.array_equal(e, oe)
            ok = ov.astype(bool)
            assert np.array_equal(f[ok], of[ok])


@pytest.mark.parametrize("world", [2, 4])
def test_select_sharded_k6_equals_single_pass_oracle(lib, orc, world):
    """nmost over the union of all ranks' records with a position-interleaved order: ids, delta_jsd bits and
    total_jsd equal the oracle's numprocs=1 pass; identical on every rank (SM-replicated rounds, k=6)"""
    npr = [260 + 7 * r for r in range(world)]
    shards, flat, off = make_shards(lib, 4242, npr, 9, 20_000)
    _, of, oe, ov = orc.count_batch(flat, off, 6, want_counts=False)
    from diverseseq_b200 import shard
    orders = [np.random.default_rng(10 + r).permutation(npr[r]) for r in range(world)]
    order = shard.interleaved_order(orders, npr)
    assert sorted(order.tolist()) == list(range(sum(npr)))
    exp = orc.select_rows(of, oe, order, "nmost", 24, valid=ov)
    mexp = orc.select_rows(of, oe, order, "cov", 8, 30, valid=ov)
    assert len(exp.trace) > 10

    def fn(rank, ctx, comm):
        kf, _ = shard.count_sharded(ctx, comm, lib.SeqSet.upload(ctx, *shards[rank]), 6)
        a = shard.select_sharded(ctx, comm, kf, order, lib.MODE_NMOST, 24)
        acc = int(ctx._lib.dvs_select_last_accepts(ctx.handle))
        b = shard.select_sharded(ctx, comm, kf, order, lib.MODE_MAX_COV, 8, 30)
        kf.close()
        return a, b, acc

    for a, b, acc in run_ranks(world, fn):
        assert a[0].tolist() == exp.ids.tolist() and np.array_equal(a[1], exp.delta_jsd) and a[2][0] == exp.total_jsd
        assert b[0].tolist() == mexp.ids.tolist() and np.array_equal(b[1], mexp.delta_jsd)
        assert acc == len(exp.trace)


def test_select_sharded_k8_global_state_rounds_and_grow(lib, orc):
    """k=8 rows do not fit shared memory: cooperative global-state rounds with the cross-GPU exchange between
    two grid barriers, and `max` with candidate-sharded grow windows (configs[2]'s path)"""
    npr = [150, 141]
    shards, flat, off = make_shards(lib, 99, npr, 6, 40_000)
    _, of, oe, ov = orc.count_batch(flat, off, 8, want_counts=False)
    order = np.random.default_rng(3).permutation(sum(npr)).astype(np.uint32)
    e_n = orc.select_rows(of, oe, order, "nmost", 12, valid=ov)
    e_s = orc.select_rows(of, oe, order, "stdev", 5, 20, valid=ov)

    def fn(rank, ctx, comm):
        from diverseseq_b200 import shard
        kf, _ = shard.count_sharded(ctx, comm, lib.SeqSet.upload(ctx, *shards[rank]), 8)
        a = shard.select_sharded(ctx, comm, kf, order, lib.MODE_NMOST, 12)
        b = shard.select_sharded(ctx, comm, kf, order, lib.MODE_MAX_STDEV, 5, 20)
        kf.close()
        return a, b

    for a, b in run_ranks(2, fn, window_bytes=512 << 20):
        assert a[0].tolist() == e_n.ids.tolist() and np.array_equal(a[1], e_n.delta_jsd)
        assert b[0].tolist() == e_s.ids.tolist() and np.array_equal(b[1], e_s.delta_jsd) and b[2][2] == e_s.std_delta_jsd


def test_chunked_select_is_the_np_merge(lib, orc):
    """-np N semantics (records.py:206-251): per-chunk selection, then final_nmost over the winners' rows"""
    npr = [120, 131]
    shards, flat, off = make_shards(lib, 5, npr, 6, 4000)
    orders = [np.random.default_rng(40 + r).permutation(npr[r]).astype(np.uint32) for r in range(2)]
    firsts = []
    for r in range(2):
        _, f, e, v = orc.count_batch(*shards[r], 4, want_counts=False)
        firsts.append(orc.select_rows(f, e, orders[r], "nmost", 9, valid=v, want_freqs=True))
    rows = np.concatenate([s.kfreqs for s in firsts])
    merged = orc.select_rows(rows, None, np.arange(len(rows)), "nmost", 9, recompute_entropy=True)
    ids = [(r, int(i)) for r in range(2) for i in firsts[r].ids]
    want = [ids[i] for i in merged.ids]

    def fn(rank, ctx, comm):
        from diverseseq_b200 import shard
        kf = lib.KFreqs.count(ctx, lib.SeqSet.upload(ctx, *shards[rank]), 4)
        return shard.chunked_select(ctx, comm, kf, orders[rank], lib.MODE_NMOST, 9, 9)

    for got, delta, stats in run_ranks(2, fn):
        assert got == want and np.array_equal(delta, merged.delta_jsd) and stats[0] == merged.total_jsd


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_distance_matrices_equal_single_gpu(lib, orc, world):
    npr = [33 + 5 * r for r in range(world)]
    shards, flat, off = make_shards(lib, 808, npr, 4, 9000)
    ctx0 = lib.Context(0)
    ss = lib.SeqSet.upload(ctx0, flat, off)
    full_mash = lib.Sketches.sketch(ctx0, ss, 12, 300, 4, True).distances(12, 300)
    full_eu = lib.KFreqs.count(ctx0, ss, 5).euclidean()
    o_sk, o_lens = orc.mash_sketch_batch(flat, off, 12, 300, canonical=True)
    np.testing.assert_allclose(full_mash, orc.mash_matrix(o_sk, o_lens, 12, 300)[0], rtol=1e-9, atol=0)

    def fn(rank, ctx, comm):
        from diverseseq_b200 import shard
        ssl = lib.SeqSet.upload(ctx, *shards[rank])
        m = shard.sharded_mash_distances(ctx, comm, ssl, 12, 300, 4, True)
        e = shard.sharded_euclidean(ctx, comm, lib.KFreqs.count(ctx, ssl, 5))
        return m, e

    for m, e in run_ranks(world, fn):
        assert np.array_equal(m, full_mash)
        assert np.array_equal(e, full_eu)


def test_processes_over_tcp_rendezvous_and_cuda_ipc(lib, orc):
    """two real processes (cudaIpc-mapped windows, TCP rendezvous): the path torchrun-launched runs take"""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ndev = device_count()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LOCAL_RANK=str(r if ndev >= 2 else 0), DVS_SHARED_GPU="0" if ndev >= 2 else "1")
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "sharded_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "SHARDED_WORKER_OK" in o, o
