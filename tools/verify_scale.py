"""Verifier throughput against batch size (device-resident inputs, CUDA events on the library's stream)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import coracle as co
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy
ctx = zk.Context(0)
import os
if os.environ.get("ZK_LANES") is not None:
    ctx.set_opt(zk.Context.OPT_VERIFY_LANES, int(os.environ["ZK_LANES"]))
r1cs = sy.make_r1cs(seed=1, n_constraints=60, n_inputs=23, n_aux=50, a_aux_density=40, b_density=33)
crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=2)
params = zk.Parameters.read(ctx, crs.params_bytes, checked=True)
z = sy.make_witness(r1cs, 1)
a, b, c = sy.evaluate(r1cs, z)
pa = zk.ProvingAssignment(co.ints_to_limbs(a, 4), co.ints_to_limbs(b, 4), co.ints_to_limbs(c, 4), co.ints_to_limbs(z[:23], 4), co.ints_to_limbs(z[23:], 4),
                          *sy.densities(r1cs))
proof = np.frombuffer(zk.create_proof(pa, params, 5, 7), np.uint8)
inp = co.ints_to_limbs(z[1:23], 4).reshape(-1)
pvk = zk.PreparedVerifyingKey.prepare(ctx, crs.params_bytes)
stream = torch.cuda.ExternalStream(ctx.stream)
for n in [int(x) for x in sys.argv[1:]] or [1024, 4096, 8192, 16384, 32768, 65536]:
    dp = torch.from_numpy(np.tile(proof, n)).cuda()
    di = torch.from_numpy(np.tile(inp, n).view(np.int64)).cuda()
    dv = torch.zeros(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    zk.verify_proofs_device(pvk, n, dp.data_ptr(), di.data_ptr(), 22, dv.data_ptr()); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(2):
        zk.verify_proofs_device(pvk, n, dp.data_ptr(), di.data_ptr(), 22, dv.data_ptr())
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    assert bool((dv == 1).all())
    print("n=%6d  %8.2f ms  %9.0f verifications/s" % (n, ms, n / ms * 1e3), flush=True)

# two contexts, half a batch each, enqueued back to back (both calls are asynchronous on their context's stream)
ctx_b = zk.Context(0)
pvk_b = pvk          # the prepared key is read-only and shareable between contexts of one device
for n in [int(x) for x in sys.argv[1:]] or [8192]:
    h = n // 2
    dp = torch.from_numpy(np.tile(proof, n)).cuda()
    di = torch.from_numpy(np.tile(inp, n).view(np.int64)).cuda()
    dv = torch.zeros(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    def both():
        zk._ck(_lib.lib().zk_groth16_verify_batch_device(ctx._h, pvk._h, h, C.c_void_p(dp.data_ptr()), C.c_void_p(di.data_ptr()), 22, C.c_void_p(dv.data_ptr())))
        zk._ck(_lib.lib().zk_groth16_verify_batch_device(ctx_b._h, pvk._h, n - h, C.c_void_p(dp.data_ptr() + 192 * h), C.c_void_p(di.data_ptr() + 32 * 22 * h), 22, C.c_void_p(dv.data_ptr() + h)))
    from zero_chain_b200 import _lib
    import ctypes as C
    both(); ctx.sync(); ctx_b.sync()
    t = time.perf_counter()
    for _ in range(3):
        both()
    ctx.sync(); ctx_b.sync()
    ms = (time.perf_counter() - t) / 3 * 1e3
    assert bool((dv == 1).all())
    print("two contexts: n=%6d  %8.2f ms  %9.0f verifications/s" % (n, ms, n / ms * 1e3), flush=True)
