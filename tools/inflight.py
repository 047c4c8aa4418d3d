"""Experiment: K MSMs of 2^20 terms with 1 vs 2 in-flight (two contexts/streams, two host threads)."""
import sys, time, threading
import numpy as np
sys.path.insert(0, ".")
import torch
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy
n = 1 << 20
ctxs = [zk.Context(0), zk.Context(0)]
bases_l = zk.scalar_mul_many(ctxs[0], 1, zk.G1_GENERATOR, sy.random_fr_limbs(n, 7))
bases = zk.Bases(ctxs[0], 1, bases_l, window_bits=16, precompute=True)
sets = [torch.from_numpy(sy.random_fr_limbs(n, 100 + k).view(np.int64)).cuda() for k in range(8)]
torch.cuda.synchronize()
import ctypes as C
from zero_chain_b200 import _lib
L = _lib.lib()
def run(ctx, ks, out):
    o = np.zeros(96, np.uint8)
    for k in ks:
        rc = L.zk_msm_device(ctx._h, bases._h, C.c_void_p(sets[k % 8].data_ptr()), n, o.ctypes.data_as(C.c_void_p))
        assert rc == 0
        out[k] = o.tobytes()
K = 20
for inflight in (1, 2, 1, 2):
    res = {}
    for c in ctxs: run(c, [0, 1], res)
    torch.cuda.synchronize()
    t = time.perf_counter()
    if inflight == 1:
        run(ctxs[0], range(K), res)
    else:
        th = [threading.Thread(target=run, args=(ctxs[i], range(i, K, 2), res)) for i in range(2)]
        [x.start() for x in th]; [x.join() for x in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("inflight=%d: %.3f ms/step -> %.1f Mop/s" % (inflight, dt / K * 1e3, n * K / dt / 1e6), flush=True)
    ref = res
# consistency between modes
r1 = {}; run(ctxs[1], range(4), r1); r0 = {}; run(ctxs[0], range(4), r0)
assert all(r0[k] == r1[k] for k in range(4))
print("ok")
