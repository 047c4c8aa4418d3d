"""Blocking-call timing of one 2^20 G1 MSM over the batched-affine knobs (EXPERIMENTS build: ZK_BA_MINB, ZK_BA_K; levels and the
threshold through zk_ctx_set_opt).  python tools/ba_tune.py [log_n]"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
ctx = zk.Context(0)
bases = zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, sy.random_fr_limbs(n, 1))
b = zk.Bases(ctx, 1, bases, precompute=True)
ds = [torch.from_numpy(sy.random_fr_limbs(n, 2 + k).view(np.int64)).cuda() for k in range(4)]
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(ctx.stream)
def run(tag):
    for k in range(3):
        ref = zk.multiexp_device(b, ds[k % 4].data_ptr(), n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(8):
        zk.multiexp_device(b, ds[k % 4].data_ptr(), n)
    e1.record(stream); torch.cuda.synchronize()
    print("%-40s %.3f ms per MSM" % (tag, e0.elapsed_time(e1) / 8), flush=True)
    return ref
ctx.set_opt(1, 1 << 40)
r0 = run("xyzz only")
ctx.set_opt(1, 0)
for lv in (2,):
    ctx.set_opt(2, lv)
    for minb in ("4", "3"):
        for K in ("16", "32", "48", "64", "96"):
            os.environ["ZK_BA_MINB"] = minb; os.environ["ZK_BA_K"] = K
            r = run("levels=%d minb=%s K=%s" % (lv, minb, K))
            assert r == r0
