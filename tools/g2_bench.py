import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy
ctx = zk.Context(0)
for logn in (14, 17):
    n = 1 << logn
    bases = zk.scalar_mul_many(ctx, 2, zk.G2_GENERATOR, sy.random_fr_limbs(n, 1))
    b = zk.Bases(ctx, 2, bases, precompute=True)
    d = torch.from_numpy(sy.random_fr_limbs(n, 2).view(np.int64)).cuda()
    torch.cuda.synchronize()
    for _ in range(2): zk.multiexp_device(b, d.data_ptr(), n)
    ts = []
    for _ in range(5):
        s = time.perf_counter(); zk.multiexp_device(b, d.data_ptr(), n); ts.append(time.perf_counter() - s)
    print("G2 msm 2^%d c=%d: %.3f ms -> %.2f Mop/s" % (logn, b.window_bits, sorted(ts)[2] * 1e3, n / sorted(ts)[2] / 1e6), flush=True)
    b.free()
