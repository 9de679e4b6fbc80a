#!/usr/bin/env python3
"""Secondary measurements for BASELINE.json configs[2..4] and the k=12 counting headline
(bench.py covers configs[1], the headline).  Prints one JSON line per config; run on a GPU box:

    python tools/bench_configs.py [--quick] [--only max,mash,euclid,count12]

  max     configs[2]: dvs max (min/max size sweep), k=8, 10.5k genomes
  mash    configs[3]: ctree mash distance k=16 sketch 3000, 1k genomes (sketch Gbp/s, pairs/s)
  euclid  configs[4]: ctree Euclidean k=8, 10.5k genomes (pairs/s, FP64 TFLOP/s)
  cluster ctree tail: average-linkage tree of the 10.5k x 10.5k Euclidean matrix (device) vs scikit-learn (host)
  count12 north-star headline: k-mer counting at k=12 (Gbp/s); dense u32 rows, few records at a time
"""
from __future__ import annotations

import argparse
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
SEED = 20261017


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="1/10 size (smoke)")
    ap.add_argument("--only", default="max,mash,euclid,cluster,count12")
    a = ap.parse_args()
    import os

    from diverseseq_b200 import _lib

    world, rank, local = (int(os.environ.get(v, d)) for v, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    ctx = _lib.Context(local)
    ctx.enable_timing(True)
    which = set(a.only.split(","))
    scale = 10 if a.quick else 1

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    if world > 1:
        # torchrun: tile/row-sharded ctree matrices over N GPUs: records sharded for
        # sketching / counting, all-gather, one row block of the matrix per rank, gather of the blocks
        import torch
        import torch.distributed as dist

        from diverseseq_b200 import shard

        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=device)

        def timed(fn):
            dist.barrier(); torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize(device); dist.barrier()
            return shard.max_over_ranks(time.perf_counter() - t0, device), out

        if "mash" in which:
            nrec, mean_len, k, s = 1000 // scale, 4_000_000 // scale, 16, 3000
            b, e = shard.shard_bounds(nrec, world, rank)
            flat, off = None, None
            ss_all = _lib.SeqSet.synth(ctx, SEED, nrec, 64, mean_len)  # same set on every rank; keep own shard
            offs = ss_all.offsets()
            part = ss_all.download(b, e - b)
            ss = _lib.SeqSet.upload(ctx, part, (offs[b:e + 1] - offs[b]).astype(np.uint64))
            del ss_all
            shard.sharded_mash_distances(ctx, ss, k, s, 4, True, device)
            dt, d = timed(lambda: shard.sharded_mash_distances(ctx, ss, k, s, 4, True, device))
            emit(config="configs[3] ctree mash k=16 s=3000 canonical", n_gpus=world, nrec=nrec, wall_s=dt,
                 pairs=nrec * (nrec - 1) // 2, pairs_per_s_wall=nrec * (nrec - 1) // 2 / dt, mean_dist=float(d.mean()),
                 note="sketch (record-sharded) + all-gather sketches + row-block pairs + gather of row blocks")
        if "euclid" in which:
            nrec, mean_len, k = 10500 // scale, 400_000, 8
            b, e = shard.shard_bounds(nrec, world, rank)
            ss_all = _lib.SeqSet.synth(ctx, SEED, nrec, 64, mean_len)
            offs = ss_all.offsets()
            part = ss_all.download(b, e - b)
            ss = _lib.SeqSet.upload(ctx, part, (offs[b:e + 1] - offs[b]).astype(np.uint64))
            del ss_all
            kf = _lib.KFreqs.count(ctx, ss, k)
            if (e - b) * world != nrec:
                raise SystemExit("euclid multi-GPU bench needs nrec divisible by the world size")
            shard.sharded_euclidean(ctx, kf, device)
            dt, d = timed(lambda: shard.sharded_euclidean(ctx, kf, device))
            npairs = nrec * (nrec - 1) // 2
            emit(config="configs[4] ctree euclidean k=8", n_gpus=world, nrec=nrec, wall_s=dt, pairs=npairs,
                 pairs_per_s_wall=npairs / dt, mean_dist=float(d.mean()),
                 note="all-gather rows + row-block tiles per rank + gather of row blocks (host matrix included)")
        dist.destroy_process_group()
        return 0

    if "max" in which:
        nrec, mean_len, k = 10500 // scale, 4_000_000 // scale, 8
        ss = _lib.SeqSet.synth(ctx, SEED, nrec, 64, mean_len)
        t0 = time.perf_counter()
        kf = _lib.KFreqs.count(ctx, ss, k)
        ctx.sync()
        t_count = time.perf_counter() - t0
        count_ms = ctx.phase_ms(_lib.PHASE_COUNT_KERNEL)
        order = np.random.default_rng(SEED).permutation(nrec).astype(np.uint32)
        sweep = []
        for lo, hi, mode, name in ((5, 10, _lib.MODE_MAX_STDEV, "stdev"), (10, 100, _lib.MODE_MAX_STDEV, "stdev"),
                                   (10, 100, _lib.MODE_MAX_COV, "cov"), (100, 100, _lib.MODE_NMOST, "nmost")):
            t0 = time.perf_counter()
            idx, delta, stats = kf.select(order, mode, lo, hi)
            dt = time.perf_counter() - t0
            sweep.append({"min": lo, "max": hi, "stat": name, "wall_s": dt, "size": int(len(idx)),
                          "accepts": int(ctx._lib.dvs_select_last_accepts(ctx.handle)),
                          "exact_evals": int(ctx._lib.dvs_select_last_exact_evals(ctx.handle))})
        emit(config="configs[2] dvs max k=8", nrec=nrec, gbp=ss.total_bases / 1e9, count_wall_s=t_count,
             count_kernel_ms=count_ms, count_kernel_gbp_per_s=ss.total_bases / count_ms / 1e6, sweep=sweep)
        del kf, ss

    if "mash" in which:
        nrec, mean_len, k, s = 1000 // scale, 4_000_000 // scale, 16, 3000
        ss = _lib.SeqSet.synth(ctx, SEED, nrec, 64, mean_len)
        for _ in range(2):
            t0 = time.perf_counter()
            sk = _lib.Sketches.sketch(ctx, ss, k, s, 4, True)
            ctx.sync()
            t_sk = time.perf_counter() - t0
        sk_ms = ctx.phase_ms(_lib.PHASE_SKETCH)
        for _ in range(2):
            t0 = time.perf_counter()
            dist = sk.distances(k, s)
            t_pairs = time.perf_counter() - t0
        pairs_ms = ctx.phase_ms(_lib.PHASE_MASH_PAIRS)
        npairs = nrec * (nrec - 1) // 2
        emit(config="configs[3] ctree mash k=16 s=3000 canonical", nrec=nrec, gbp=ss.total_bases / 1e9,
             sketch_wall_s=t_sk, sketch_device_ms=sk_ms, sketch_gbp_per_s=ss.total_bases / sk_ms / 1e6,
             pairs=npairs, pairs_kernel_ms=pairs_ms, pairs_per_s_kernel=npairs / pairs_ms * 1e3,
             pairs_wall_s=t_pairs, pairs_per_s_wall=npairs / t_pairs, mean_dist=float(dist.mean()))
        del sk, ss

    if "euclid" in which:
        nrec, mean_len, k = 10500 // scale, 400_000, 8  # rows only depend on nrec x 4^k; shorter genomes suffice
        ss = _lib.SeqSet.synth(ctx, SEED, nrec, 64, mean_len)
        kf = _lib.KFreqs.count(ctx, ss, k)
        for _ in range(2):
            t0 = time.perf_counter()
            d = kf.euclidean()
            t_eu = time.perf_counter() - t0
        eu_ms = ctx.phase_ms(_lib.PHASE_EUCLID)
        npairs = nrec * (nrec - 1) // 2
        dim = 4 ** k
        emit(config="configs[4] ctree euclidean k=8", nrec=nrec, dim=dim, pairs=npairs, kernel_ms=eu_ms,
             pairs_per_s_kernel=npairs / eu_ms * 1e3, wall_s=t_eu, pairs_per_s_wall=npairs / t_eu,
             fp64_tflops_useful=2.0 * dim * npairs / eu_ms / 1e9,  # 2*D flops per pair
             fp64_tflops_issued=3.0 * dim * (npairs + nrec / 2) / eu_ms / 1e9, mean_dist=float(d.mean()))
        del kf, ss

    if "cluster" in which:
        import torch

        nrec, mean_len, k = 10500 // scale, 100_000, 6  # the tree only needs the matrix
        ss = _lib.SeqSet.synth(ctx, SEED, nrec, 64, mean_len)
        kf = _lib.KFreqs.count(ctx, ss, k)
        dmat = torch.empty((nrec, nrec), dtype=torch.float64, device="cuda:0")
        kf.euclidean_into(dmat.data_ptr())
        for _ in range(2):
            t0 = time.perf_counter()
            children, heights, counts = _lib.linkage_average(ctx, n=nrec, device_ptr=dmat.data_ptr())
            t_gpu = time.perf_counter() - t0
        cl_ms = ctx.phase_ms(_lib.PHASE_CLUSTER)
        host = dmat.cpu().numpy()
        from sklearn.cluster import AgglomerativeClustering

        t0 = time.perf_counter()
        ref = AgglomerativeClustering(metric="precomputed", linkage="average").fit(host)
        t_ref = time.perf_counter() - t0
        emit(config="ctree tail: average-linkage tree", nrec=nrec, device_ms=cl_ms, wall_s=t_gpu,
             sklearn_wall_s=t_ref, identical_children=bool(np.array_equal(children, ref.children_)),
             note="sklearn time excludes the 882 MB device->host copy the device path avoids")
        del kf, ss, dmat

    if "count12" in which:
        k = 12
        nrec, mean_len = max(4, 64 // scale), 4_000_000 // scale  # dense u32+f64 rows: 192 MB per record
        ss = _lib.SeqSet.synth(ctx, SEED, nrec, 8, mean_len)
        for _ in range(2):
            kf = _lib.KFreqs.count(ctx, ss, k)
            ms = ctx.phase_ms(_lib.PHASE_COUNT_KERNEL)
            fe = ctx.phase_ms(_lib.PHASE_FREQ_ENTROPY)
            del kf
        emit(config="north-star headline: counting k=12 (dense rows, global RED.ADD)", nrec=nrec,
             gbp=ss.total_bases / 1e9, count_kernel_ms=ms, count_kernel_gbp_per_s=ss.total_bases / ms / 1e6,
             freq_entropy_ms=fe, note="8 GPUs shard records with no collective: aggregate = 8x")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
