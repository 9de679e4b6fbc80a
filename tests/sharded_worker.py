"""One rank of the multi-process sharded test (launched by tests/test_gpu_sharded.py, or by hand under
torchrun: RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT from the environment).  Every rank checks
its own results against the CPU oracle and prints SHARDED_WORKER_OK."""
import os
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from diverseseq_b200 import _lib, shard  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rv = shard.Rendezvous()
    rank, world = rv.rank, rv.world
    shared_gpu = os.environ.get("DVS_SHARED_GPU", "0") == "1"
    ctx = _lib.Context(int(os.environ.get("LOCAL_RANK", "0")))
    comm = shard.connect(ctx, rv, 256 << 20)
    npr = [140 + 9 * r for r in range(world)]
    total = sum(npr)
    first = sum(npr[:rank])
    mine = _lib.synth_host(2026, total, 7, 12_000, first, npr[rank])
    flat, off = _lib.synth_host(2026, total, 7, 12_000)
    _, of, oe, ov = orc.count_batch(flat, off, 6, want_counts=False)
    ss = _lib.SeqSet.upload(ctx, *mine)
    kf, nrec = shard.count_sharded(ctx, comm, ss, 6)
    assert nrec == npr
    _, f, e, v = kf.download(counts=False)
    assert np.array_equal(v, ov) and np.array_equal(e, oe) and np.array_equal(f[ov.astype(bool)], of[ov.astype(bool)])
    # ctree matrices: pairs / tiles dealt over the ranks, results stored into every peer's window
    eu = shard.sharded_euclidean(ctx, comm, kf_all=kf)
    np.testing.assert_allclose(eu, orc.euclid_matrix(of), rtol=1e-9, atol=1e-18)
    mash = shard.sharded_mash_distances(ctx, comm, ss, 10, 200, 4, True)
    o_sk, o_lens = orc.mash_sketch_batch(flat, off, 10, 200, canonical=True)
    np.testing.assert_allclose(mash, orc.mash_matrix(o_sk, o_lens, 10, 200)[0], rtol=1e-9, atol=0)
    if not shared_gpu:
        # persistent selection kernels of different PROCESSES only run concurrently on different GPUs
        order = shard.interleaved_order([np.random.default_rng(r).permutation(npr[r]) for r in range(world)], npr)
        exp = orc.select_rows(of, oe, order, "nmost", 16, valid=ov)
        idx, delta, stats = shard.select_sharded(ctx, comm, kf, order, _lib.MODE_NMOST, 16)
        assert idx.tolist() == exp.ids.tolist() and np.array_equal(delta, exp.delta_jsd) and stats[0] == exp.total_jsd
        mexp = orc.select_rows(of, oe, order, "stdev", 5, 12, valid=ov)
        idx, delta, stats = shard.select_sharded(ctx, comm, kf, order, _lib.MODE_MAX_STDEV, 5, 12)
        assert idx.tolist() == mexp.ids.tolist() and np.array_equal(delta, mexp.delta_jsd)
    kf.close()
    ctx.sync()
    rv.barrier()
    comm.close()
    rv.close()
    print(f"SHARDED_WORKER_OK rank {rank}/{world} shared_gpu={shared_gpu}", flush=True)


if __name__ == "__main__":
    main()
