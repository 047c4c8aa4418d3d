"""GPU calibration probe (run under gpurun): modmul roofline + first MSM timings. Not a bench line."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy

ctx = zk.Context(0)
res = {}
for field, name in ((0, "fq"), (1, "fr")):
    for blocks, threads in ((148 * 4, 128), (148 * 4, 256), (148 * 8, 128), (148 * 2, 512), (148 * 16, 64)):
        per_s, ms = zk.bench_modmul(ctx, field, blocks, threads, 3000)
        res["%s_b%d_t%d" % (name, blocks, threads)] = per_s
        print("modmul %s blocks=%d threads=%d : %.3e /s  (%.2f ms)" % (name, blocks, threads, per_s, ms), flush=True)

import torch
for logn in (16, 18, 20):
    n = 1 << logn
    t0 = time.time()
    bases = zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, sy.random_fr_limbs(n, 1))
    t1 = time.time()
    for c in (13, 16) if logn < 20 else (14, 16):
        b = zk.Bases(ctx, 1, bases, window_bits=c, precompute=True)
        t2 = time.time()
        scal = sy.random_fr_limbs(n, 2)
        d = torch.from_numpy(scal.view(np.int64)).cuda()
        torch.cuda.synchronize()
        for _ in range(2):
            zk.multiexp_device(b, d.data_ptr(), n)
        ts = []
        for _ in range(5):
            s = time.perf_counter(); zk.multiexp_device(b, d.data_ptr(), n); ts.append(time.perf_counter() - s)
        print("msm 2^%d c=%d tables: gen %.2fs tables %.2fs  best %.3f ms  median %.3f ms -> %.1f Mop/s" %
              (logn, c, t1 - t0, t2 - t1, min(ts) * 1e3, sorted(ts)[2] * 1e3, n / sorted(ts)[2] / 1e6), flush=True)
        res["msm_%d_c%d_ms" % (logn, c)] = sorted(ts)[2] * 1e3
        b.free()
json.dump(res, open("gpurun_out/calibrate.json", "w"), indent=1)
