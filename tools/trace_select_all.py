import os, sys, numpy as np
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
from diverseseq_b200 import _lib
SEED=20261017
ctx=_lib.Context(0); ctx.enable_timing(True)
ss=_lib.SeqSet.synth(ctx, SEED, 10500, 64, 4_000_000)
kf=_lib.KFreqs.count(ctx, ss, 6)
order=np.random.default_rng(SEED).permutation(10500).astype(np.uint32)
kf.select(order,_lib.MODE_NMOST,100)
os.environ["DVS_SELECT_TRACE_ALL"]="gpurun_out/sel_trace_all.bin"
kf.select(order,_lib.MODE_NMOST,100)
os.environ.pop("DVS_SELECT_TRACE_ALL")
t=np.fromfile("gpurun_out/sel_trace_all.bin",dtype=np.uint64).reshape(256,8,256)[:, :7, :148].astype(np.int64)
t0=t[:,0,:].min(axis=1)  # earliest round start
acc=(t[:,4,0]-t[:,3,0])>200   # accepting rounds
r=np.arange(5,250)[acc[5:250]]
names=["start","scan_pub","gather1","decide","upd_pub","gather2","final"]
rel=t[r]-t0[r][:,None,None]
for i,nm in enumerate(names):
    x=rel[:,i,:]
    print(f"{nm:9s} min {np.median(x.min(axis=1)):7.0f} med {np.median(np.median(x,axis=1)):7.0f} max {np.median(x.max(axis=1)):7.0f}  argmax-cta mode {np.bincount(x.argmax(axis=1)).argmax()} ")
# per-CTA mean lateness of update publish
x=rel[:,4,:]; late=x.mean(axis=0)
print("upd_pub mean by cta (first 110):", np.round(late[:110:6]).tolist())
print("upd_pub mean ctas 101..147:", np.round(late[101:148:6]).tolist())
x=rel[:,1,:]; print("scan_pub mean by cta:", np.round(x.mean(axis=0)[::12]).tolist())
print("round length", np.median(t[r+1,0,0]-t[r,0,0]))
