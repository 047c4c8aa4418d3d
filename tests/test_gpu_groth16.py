"""GPU parity tests of the Groth16 path through the C ABI: zk_params_load (Parameters::read),
zk_groth16_prove / _batch (create_proof below synthesis) against (1) the oracle's restatement of
bellman's create_proof and (2) the closed-form trapdoor algebra (oracle/pyref + synthetic.py)."""
import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as pr
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def _witness(r1cs, seed):
    z = sy.make_witness(r1cs, seed)
    a, b, c = sy.evaluate(r1cs, z)
    a_d, bi_d, ba_d = sy.densities(r1cs)
    pa = zk.ProvingAssignment(co.ints_to_limbs(a, 4), co.ints_to_limbs(b, 4), co.ints_to_limbs(c, 4),
                              co.ints_to_limbs(z[:r1cs.n_inputs], 4), co.ints_to_limbs(z[r1cs.n_inputs:], 4), a_d, bi_d, ba_d)
    return z, pa


SHAPES = {
    "tiny": dict(n_constraints=60, n_inputs=4, n_aux=50, a_aux_density=40, b_density=33),
    "mid": dict(n_constraints=1500, n_inputs=23, n_aux=1400, a_aux_density=1000, b_density=800),
}


@pytest.mark.parametrize("shape", ["tiny", "mid"])
def test_prove_matches_oracle_and_closed_form(ctx, shape):
    r1cs = sy.make_r1cs(seed=3, **SHAPES[shape])
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=4)
    params = zk.Parameters.read(ctx, crs.params_bytes, checked=True)
    oparams = co.Params(crs.params_bytes, checked=False)
    assert (params.n_ic, params.n_h, params.n_l, params.n_a, params.n_b_g1) == (oparams.n_ic, oparams.n_h, oparams.n_l, oparams.n_a, oparams.n_b)
    rng = pr.SplitMix64(99)
    for seed in (1, 2):
        z, pa = _witness(r1cs, seed)
        r, s = rng.fr(), rng.fr()
        got = zk.create_proof(pa, params, r, s)
        want = oparams.prove(pa.a, pa.b, pa.c, pa.input_assignment, pa.aux_assignment, pa.a_aux_density, pa.b_input_density, pa.b_aux_density, r, s)
        assert got == want
        A, B, C = sy.expected_proof_scalars(crs, z, r, s)
        assert got == pr.proof_bytes(pr.ec_mul(pr.FQ, pr.G1_GEN, A), pr.ec_mul(pr.FQ2, pr.G2_GEN, B), pr.ec_mul(pr.FQ, pr.G1_GEN, C))
    # r = s = 0 and an all-zero aux-independent corner: still equal to the oracle
    z, pa = _witness(r1cs, 5)
    assert zk.create_proof(pa, params, 0, 0) == oparams.prove(pa.a, pa.b, pa.c, pa.input_assignment, pa.aux_assignment,
                                                              pa.a_aux_density, pa.b_input_density, pa.b_aux_density, 0, 0)
    params.free()


def test_prove_batch_confidential_shape(ctx):
    """The reference circuit's shape (19 974 constraints, 23 inputs, domain 2^15):
    a batch of proofs in one device pass, every proof bit-compared with the oracle.  Eight proofs put the H-query MSM above the
    2^22-entry threshold, so the batched-affine bucket rounds are part of what is compared."""
    r1cs = sy.make_r1cs(seed=1, **sy.CONF_SHAPE)
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=2)
    params = zk.Parameters.read(ctx, crs.params_bytes, checked=False)
    oparams = co.Params(crs.params_bytes, checked=False)
    assert (params.n_h, params.n_l, params.n_a, params.n_b_g1, params.n_b_g2, params.n_ic) == (32767, 19955, 15598, 12402, 12402, 23)
    rng = pr.SplitMix64(1234)
    batch = 8
    provers, zs, rs, ss = [], [], [], []
    for k in range(batch):
        z, pa = _witness(r1cs, 100 + k)
        provers.append(pa); zs.append(z); rs.append(rng.fr()); ss.append(rng.fr())
    got = zk.create_proof_batch(provers, params, rs, ss)
    for k in range(batch):
        pa = provers[k]
        want = oparams.prove(pa.a, pa.b, pa.c, pa.input_assignment, pa.aux_assignment, pa.a_aux_density, pa.b_input_density, pa.b_aux_density, rs[k], ss[k])
        assert got[192 * k:192 * k + 192] == want, k
    A, B, C = sy.expected_proof_scalars(crs, zs[0], rs[0], ss[0])
    assert got[:192] == pr.proof_bytes(pr.ec_mul(pr.FQ, pr.G1_GEN, A), pr.ec_mul(pr.FQ2, pr.G2_GEN, B), pr.ec_mul(pr.FQ, pr.G1_GEN, C))
    assert zk.create_proof(provers[1], params, rs[1], ss[1]) == got[192:384]
    params.free()


def test_params_load_rejects(ctx):
    r1cs = sy.make_r1cs(seed=3, **SHAPES["tiny"])
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=4)
    buf = bytearray(crs.params_bytes)
    with pytest.raises(zk.SynthesisError) as e:
        zk.Parameters.read(ctx, bytes(buf[:-9]))                          # truncated -> IoError
    assert e.value.code == -6
    lay = pr.params_layout(bytes(buf))
    off = lay["l"][0]
    bad = bytearray(buf); bad[off + 95] ^= 1
    zk.Parameters.read(ctx, bytes(bad), checked=False).free()              # unchecked read accepts any field elements
    with pytest.raises(zk.SynthesisError) as e:
        zk.Parameters.read(ctx, bytes(bad), checked=True)                  # NotOnCurve
    assert e.value.code == -7
    bad = bytearray(buf); bad[off:off + 96] = bytes([0x40]) + bytes(95)    # infinity in a query
    with pytest.raises(zk.SynthesisError) as e:
        zk.Parameters.read(ctx, bytes(bad), checked=False)
    assert e.value.code == -5
    # a point on the curve but outside the r-torsion (ec.rs:675-685)
    x = 4
    while True:
        y = pr.FQ.sqrt((x ** 3 + 4) % pr.Q)
        if y is not None and pr.ec_mul(pr.FQ, (x, y), pr.R) is not pr.INF:
            break
        x += 1
    bad = bytearray(buf); bad[off:off + 96] = pr.g1_uncompressed((x, y))
    with pytest.raises(zk.SynthesisError) as e:
        zk.Parameters.read(ctx, bytes(bad), checked=True)
    assert e.value.code == -7
    # witness/CRS shape mismatch -> AssignmentMissing
    params = zk.Parameters.read(ctx, bytes(buf), checked=False)
    z, pa = _witness(r1cs, 1)
    pa.aux_assignment = pa.aux_assignment[:-1]; pa.a_aux_density = pa.a_aux_density[:-1]; pa.b_aux_density = pa.b_aux_density[:-1]
    with pytest.raises(zk.SynthesisError) as e:
        zk.create_proof(pa, params, 1, 2)
    assert e.value.code == -3
    params.free()


def test_gpu_proof_verifies_by_pairing(ctx):
    """The reference's own acceptance check for gen_proof (`check_proof` -> verify_proof, core/proofs/src/confidential.rs:
    208-278): the GPU proof must satisfy e(A,B) = e(alpha,beta) e(sum x_i ic_i, gamma) e(C,delta) under the CRS's verifying
    key.  The pairing is the big-integer oracle pinned by conf_vk.dat (tests/test_oracle_pairing.py)."""
    from tests.test_oracle_pairing import proof_points, vk_from_params
    r1cs = sy.make_r1cs(seed=8, **SHAPES["tiny"])
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=9)
    params = zk.Parameters.read(ctx, crs.params_bytes, checked=True)
    z, pa = _witness(r1cs, 3)
    import random
    proof = zk.create_random_proof(pa, params, random.Random(5))          # r, s drawn like create_random_proof
    vk = vk_from_params(crs.params_bytes)
    assert pr.groth16_verify(vk, proof_points(proof), z[1:r1cs.n_inputs])
    bad = bytearray(proof); bad[100] ^= 4
    try:
        pts = proof_points(bytes(bad))
        assert not pr.groth16_verify(vk, pts, z[1:r1cs.n_inputs])
    except ValueError:
        pass                                                              # not even a curve point any more
    params.free()


def test_prove_anonymous_transfer_shape(ctx):
    """the reference's other KeyContext flavour — the anonymous_transfer circuit shape
    (domain 2^16, |h| 65 535, |l| 50 429, |a| 39 133, |b| 31 257, 105 public inputs) — through the same kernels."""
    r1cs = sy.make_r1cs(seed=11, **sy.ANON_SHAPE)
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=12)
    params = zk.Parameters.read(ctx, crs.params_bytes, checked=False)
    assert (params.n_h, params.n_l, params.n_a, params.n_b_g1, params.n_ic) == (65535, 50429, 39133, 31257, 105)
    oparams = co.Params(crs.params_bytes, checked=False)
    z0, p0 = _witness(r1cs, 21)
    z1, p1 = _witness(r1cs, 22)
    rs, ss = [0x1111, 0x2222], [0x3333, 0x4444]
    got = zk.create_proof_batch([p0, p1], params, rs, ss)
    for k, pa in enumerate((p0, p1)):
        want = oparams.prove(pa.a, pa.b, pa.c, pa.input_assignment, pa.aux_assignment, pa.a_aux_density, pa.b_input_density, pa.b_aux_density, rs[k], ss[k])
        assert got[192 * k:192 * k + 192] == want
    A, B, C = sy.expected_proof_scalars(crs, z0, rs[0], ss[0])
    assert got[:192] == pr.proof_bytes(pr.ec_mul(pr.FQ, pr.G1_GEN, A), pr.ec_mul(pr.FQ2, pr.G2_GEN, B), pr.ec_mul(pr.FQ, pr.G1_GEN, C))
    params.free()


@pytest.mark.parametrize("shape", ["tiny", "mid"])
def test_prove_from_witness_matches_evals_path(ctx, shape):
    """with the fixed constraint system resident on the device, the per-constraint evaluations
    <A_j,z>, <B_j,z>, <C_j,z> (ProvingAssignment::enforce on the host in bellman) are computed by the GPU from the
    assignment alone; proofs must equal the host-evaluated path and the oracle byte for byte."""
    r1cs = sy.make_r1cs(seed=3, **SHAPES[shape])
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=4)
    params = zk.Parameters.read(ctx, crs.params_bytes, checked=False)
    oparams = co.Params(crs.params_bytes, checked=False)
    cs = zk.ConstraintSystem(ctx, r1cs.n_inputs, r1cs.n_aux, r1cs.A, r1cs.B, r1cs.C)
    batch = 3
    ws = [_witness(r1cs, 40 + k) for k in range(batch)]
    rs = [0xAA + k for k in range(batch)]; ss = [0xBB00 + k for k in range(batch)]
    inputs = np.stack([w[1].input_assignment for w in ws]); aux = np.stack([w[1].aux_assignment for w in ws])
    got = zk.create_proof_from_witness_batch(cs, params, batch, inputs, aux, co.ints_to_limbs(rs, 4), co.ints_to_limbs(ss, 4))
    assert got == zk.create_proof_batch([w[1] for w in ws], params, rs, ss)
    pa = ws[1][1]
    assert got[192:384] == oparams.prove(pa.a, pa.b, pa.c, pa.input_assignment, pa.aux_assignment, pa.a_aux_density, pa.b_input_density, pa.b_aux_density, rs[1], ss[1])
    bad = aux.copy(); bad[0, 0, :] = 0xFFFFFFFFFFFFFFFF
    with pytest.raises(zk.SynthesisError):
        zk.create_proof_from_witness_batch(cs, params, batch, inputs, bad, co.ints_to_limbs(rs, 4), co.ints_to_limbs(ss, 4))   # non-canonical
    cs.free(); params.free()
