"""Run under `ncu --metrics gpu__time_duration.sum` for a launch list of the verifier (or plain, for wall-clock numbers)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import coracle as co
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = zk.Context(0)
r1cs = sy.make_r1cs(seed=1, n_constraints=60, n_inputs=23, n_aux=50, a_aux_density=40, b_density=33)
crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=2)
params = zk.Parameters.read(ctx, crs.params_bytes, checked=True)
z = sy.make_witness(r1cs, 1)
a, b, c = sy.evaluate(r1cs, z)
pa = zk.ProvingAssignment(co.ints_to_limbs(a, 4), co.ints_to_limbs(b, 4), co.ints_to_limbs(c, 4), co.ints_to_limbs(z[:23], 4), co.ints_to_limbs(z[23:], 4),
                          *sy.densities(r1cs))
proof = zk.create_proof(pa, params, 5, 7)
pvk = zk.PreparedVerifyingKey.prepare(ctx, crs.params_bytes)
proofs = proof * n
inputs = [z[1:23]] * n
for _ in range(3):
    t = time.perf_counter()
    v = zk.verify_proofs(pvk, proofs, inputs)
    dt = time.perf_counter() - t
assert v == [1] * n
print("n=%d  %.2f ms  %.0f verifications/s (python-side input packing included)" % (n, dt * 1e3, n / dt))
