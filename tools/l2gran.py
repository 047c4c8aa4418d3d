"""Does cudaLimitMaxL2FetchGranularity change the gather-bound batched-affine passes?  The ncu captures show 160 B of DRAM reads
per 48-B x gather in k_ba_forward (round 1): L2 fills more than the sectors asked for.  Blocking 2^20 MSM + per-kernel times
per setting.  python tools/l2gran.py [log_n]"""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy
rt = ctypes.CDLL("libcudart.so.12")
LIMIT = 0x05                                   # cudaLimitMaxL2FetchGranularity
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
torch.cuda.init(); torch.zeros(1).cuda()
def get():
    v = ctypes.c_size_t(0); rc = rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT); return rc, v.value
print("default limit:", get(), flush=True)
ctx = zk.Context(0)
bases = zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, sy.random_fr_limbs(n, 1))
b = zk.Bases(ctx, 1, bases, precompute=True)
ds = [torch.from_numpy(sy.random_fr_limbs(n, 2 + k).view(np.int64)).cuda() for k in range(4)]
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(ctx.stream)
ctx.set_opt(1, 0)
def run(tag):
    for k in range(3):
        ref = zk.multiexp_device(b, ds[k % 4].data_ptr(), n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(16):
        zk.multiexp_device(b, ds[k % 4].data_ptr(), n)
    e1.record(stream); torch.cuda.synchronize()
    ctx.profile(True)
    for k in range(8):
        zk.multiexp_device(b, ds[k % 4].data_ptr(), n)
    st = ctx.profile_read() if hasattr(ctx, "profile_read") else None
    ctx.profile(False)
    print("%-24s %.3f ms per MSM   stage %s" % (tag, e0.elapsed_time(e1) / 16, st), flush=True)
    return ref
r0 = run("as found")
for g in (32, 64, 128, 32):
    rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(g))
    r = run("granularity %d (rc %d, now %s)" % (g, rc, get()))
    assert r == r0
