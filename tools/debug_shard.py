#!/usr/bin/env python3
"""debug: thread ranks on one GPU, stage by stage with timestamps (python tools/debug_shard.py 20,0,31)"""
import os, sys, threading, time, pathlib
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("DVS_WATCHDOG_MS", "4000")
os.environ["DVS_COMM_CHECK"] = "1"
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
from diverseseq_b200 import _lib, shard

npr = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,0,31").split(",")]
world = len(npr)
group = shard.LocalGroup(world)
T0 = time.time()
def log(rank, msg):
    print(f"[{time.time()-T0:7.3f}] rank {rank}: {msg}", flush=True)
def worker(rank):
    try:
        ctx = _lib.Context(0)
        comm = shard.connect(ctx, group.member(rank), 256 << 20)
        first = sum(npr[:rank])
        flat, off = _lib.synth_host(77, sum(npr), 5, 30_000, first, npr[rank])
        ss = _lib.SeqSet.upload(ctx, flat, off)
        log(rank, "uploaded")
        for it in range(2):
            nrec = comm.rv.allgather(int(ss.nrec))
            log(rank, f"op {it}: calling count_sharded")
            kf = _lib.KFreqs.count_sharded(ctx, comm, ss, 6, nrec)
            log(rank, f"op {it}: returned")
            ctx.sync()
            log(rank, f"op {it}: synced")
            kf.close()
        comm.rv.barrier(); comm.close()
    except BaseException as e:
        log(rank, f"ERROR {type(e).__name__}: {e}")
        group._bar.abort()
ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
[t.start() for t in ts]; [t.join(120) for t in ts]
