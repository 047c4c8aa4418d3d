"""Run under `ncu --metrics gpu__time_duration.sum`: a few MSM / NTT invocations for a launch list."""
import sys
import numpy as np
sys.path.insert(0, ".")
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy
import torch
what = sys.argv[1] if len(sys.argv) > 1 else "msm20"
ctx = zk.Context(0)
if what.startswith("msm"):
    logn = int(what[3:])
    n = 1 << logn
    bases = zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, sy.random_fr_limbs(n, 1))
    b = zk.Bases(ctx, 1, bases, window_bits=int(__import__("os").environ.get("ZK_WB", 0)), precompute=len(sys.argv) < 3 or sys.argv[2] != "fresh")
    d = torch.from_numpy(sy.random_fr_limbs(n, 2).view(np.int64)).cuda()
    torch.cuda.synchronize()
    for _ in range(3):
        zk.multiexp_device(b, d.data_ptr(), n)
else:
    logn = int(what[3:])
    n = 1 << logn
    d = torch.from_numpy(sy.random_fr_limbs(n, 2).view(np.int64)).cuda()
    torch.cuda.synchronize()
    from zero_chain_b200 import _lib
    import ctypes as C
    for _ in range(3):
        assert _lib.lib().zk_ntt_fr_device(ctx._h, C.c_void_p(d.data_ptr()), logn, 0) == 0
    ctx.sync()
