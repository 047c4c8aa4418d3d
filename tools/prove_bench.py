"""Probe: batched proving of the confidential_transfer-shaped synthetic circuit (GPU) vs the oracle (CPU).
usage: python tools/prove_bench.py [gpu|cpu] [batch]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from zero_chain_b200 import synthetic as sy
mode = sys.argv[1] if len(sys.argv) > 1 else "gpu"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
r1cs = sy.make_r1cs(seed=1, **sy.CONF_SHAPE)
dens = sy.densities(r1cs)
lim = sy.ints_to_limbs


def witness(seed):
    z = sy.make_witness(r1cs, seed)
    a, b, c = sy.evaluate(r1cs, z)
    return (lim(a), lim(b), lim(c), lim(z[:r1cs.n_inputs]), lim(z[r1cs.n_inputs:]))


if mode == "cpu":
    from oracle import coracle as co
    t = time.time(); crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=2); print("crs %.1fs" % (time.time() - t))
    P = co.Params(crs.params_bytes, checked=False)
    w = witness(5)
    for k in range(3):
        t = time.time(); P.prove(*w, *dens, 123 + k, 456); dt = time.time() - t
        print("cpu prove %.3f s  (%d threads) -> %.2f proofs/s" % (dt, co.num_threads(), 1 / dt))
else:
    from zero_chain_b200 import groth16 as zk
    ctx = zk.Context(0)
    if len(sys.argv) > 4:                       # batched-affine bucket rounds: min entries, levels
        ctx.set_opt(zk.Context.OPT_AFFINE_MIN_ENTRIES, int(sys.argv[3])); ctx.set_opt(zk.Context.OPT_AFFINE_LEVELS, int(sys.argv[4]))
    g1 = lambda s: zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, s)
    g2 = lambda s: zk.scalar_mul_many(ctx, 2, zk.G2_GENERATOR, s)
    t = time.time(); crs = sy.make_toy_crs(r1cs, g1, g2, seed=2); print("crs (gpu points) %.1fs" % (time.time() - t), flush=True)
    t = time.time(); params = zk.Parameters.read(ctx, crs.params_bytes, checked=True); print("params load checked %.2fs" % (time.time() - t), flush=True)
    ws = [witness(100 + k) for k in range(8)]
    provers = [zk.ProvingAssignment(*ws[k % 8], *dens) for k in range(batch)]
    rng = sy.SplitMix64(77)
    rs = [rng.fr() for _ in range(batch)]; ss = [rng.fr() for _ in range(batch)]       # full-size blinding scalars, as Fr::rand gives
    # pre-concatenate like create_proof_batch does, but outside the timed region for the device-side number
    for rep in range(3):
        t = time.time(); out = zk.create_proof_batch(provers, params, rs, ss); dt = time.time() - t
        print("gpu batch=%d: %.3f s -> %.1f proofs/s (incl. host concat + H2D)" % (batch, dt, batch / dt), flush=True)
    t = time.time(); one = zk.create_proof(provers[0], params, rs[0], ss[0]); print("single proof latency %.1f ms" % ((time.time() - t) * 1e3))
    assert one == out[:192]
    for rep in range(3):      # steady state after the batch
        t = time.time(); zk.create_proof(provers[0], params, rs[0], ss[0]); print("single proof latency (repeat) %.1f ms" % ((time.time() - t) * 1e3))
