"""GPU parity tests of the verifier row through the C ABI:
  zk_pvk_prepare / zk_pvk_write   vs the reference's shipped PreparedVerifyingKey files (tests/golden/*_pvk.dat), byte for byte
  zk_pvk_load                     round trip of those files
  zk_pairing_batch                vs the fixture e(alpha, beta) and the oracle's independent pairing; bilinearity at scale
  zk_groth16_verify_batch         vs the oracle's verify_proof on proofs made by the GPU prover under a toy CRS, tampered
                                  proofs / inputs, and every Proof::read rejection class"""
import os

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as pr
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["conf", "anony"])
def test_prepare_verifying_key_matches_reference_file(ctx, name):
    head = open(os.path.join(GOLD, "%s_vk_head.bin" % name), "rb").read()
    want = open(os.path.join(GOLD, "%s_pvk.dat" % name), "rb").read()
    k = zk.PreparedVerifyingKey.prepare(ctx, head)
    assert k.num_inputs == (22 if name == "conf" else 104)
    assert k.write() == want                      # e(alpha, beta), both coefficient tables, ic: identical to the shipped file
    k2 = zk.PreparedVerifyingKey.read(ctx, want)
    assert k2.num_inputs == k.num_inputs and k2.write() == want
    k.free(); k2.free()


def test_pvk_load_rejections(ctx):
    want = open(os.path.join(GOLD, "conf_pvk.dat"), "rb").read()
    with pytest.raises(zk.SynthesisError) as e:
        zk.PreparedVerifyingKey.read(ctx, want[:1000])
    assert e.value.code == -6
    bad = bytearray(want); bad[0:48] = b"\xff" * 48                     # Fq12 coefficient >= q
    with pytest.raises(zk.SynthesisError) as e:
        zk.PreparedVerifyingKey.read(ctx, bytes(bad))
    assert e.value.code == -7
    bad = bytearray(want); bad[-1] ^= 1                                  # last ic point off the curve
    with pytest.raises(zk.SynthesisError) as e:
        zk.PreparedVerifyingKey.read(ctx, bytes(bad))
    assert e.value.code == -7


def test_pairing_kat_and_bilinearity(ctx):
    head = open(os.path.join(GOLD, "conf_vk_head.bin"), "rb").read()
    want = open(os.path.join(GOLD, "conf_pvk.dat"), "rb").read()
    assert zk.pairing(ctx, head[0:96], head[192:384]) == want[:576]       # Engine::pairing(alpha_g1, beta_g2), reference fixture
    rng = pr.SplitMix64(31)
    ks = [(rng.below(pr.R, 4), rng.below(pr.R, 4)) for _ in range(3)]
    g1 = b"".join(pr.g1_uncompressed(pr.ec_mul(pr.FQ, pr.G1_GEN, a)) for a, _ in ks) + pr.g1_uncompressed(pr.INF)
    g2 = b"".join(pr.g2_uncompressed(pr.ec_mul(pr.FQ2, pr.G2_GEN, b)) for _, b in ks) + pr.g2_uncompressed(pr.G2_GEN)
    got = zk.pairing(ctx, g1, g2)
    e = pr.pairing_reference(pr.G1_GEN, pr.G2_GEN)
    for i, (a, b) in enumerate(ks):
        assert got[576 * i:576 * i + 576] == pr.f12_to_tower_bytes(pr._f12_pow(e, a * b % pr.R))
    assert got[576 * 3:] == pr.f12_to_tower_bytes(pr.F12_ONE)            # infinity pairs to one (mod.rs:50-54)
    # size-independent property at a full batch: e(a_i G1, G2) * e(G1, -a_i G2) == 1 is implied by equality of the two sides
    n = 512
    sc = co.ints_to_limbs([rng.below(pr.R, 4) for _ in range(n)], 4)
    p1 = zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, sc)
    p2 = zk.scalar_mul_many(ctx, 2, zk.G2_GENERATOR, sc)
    enc1 = b"".join(co.g1_encode(p1[i], False) for i in range(n))
    enc2 = b"".join(co.g2_encode(p2[i], False) for i in range(n))
    left = zk.pairing(ctx, enc1, pr.g2_uncompressed(pr.G2_GEN) * n)
    right = zk.pairing(ctx, pr.g1_uncompressed(pr.G1_GEN) * n, enc2)
    assert left == right and len(set(left[576 * i:576 * i + 576] for i in range(n))) == n


def _setup(ctx, seed=3):
    shape = dict(n_constraints=60, n_inputs=4, n_aux=50, a_aux_density=40, b_density=33)
    r1cs = sy.make_r1cs(seed=seed, **shape)
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=seed + 1)
    params = zk.Parameters.read(ctx, crs.params_bytes, checked=True)
    return r1cs, crs, params


def _prove(r1cs, params, seed, r, s):
    z = sy.make_witness(r1cs, seed)
    a, b, c = sy.evaluate(r1cs, z)
    pa = zk.ProvingAssignment(co.ints_to_limbs(a, 4), co.ints_to_limbs(b, 4), co.ints_to_limbs(c, 4),
                              co.ints_to_limbs(z[:r1cs.n_inputs], 4), co.ints_to_limbs(z[r1cs.n_inputs:], 4), *sy.densities(r1cs))
    return z, zk.create_proof(pa, params, r, s)


def test_verify_batch_matches_oracle(ctx):
    r1cs, crs, params = _setup(ctx)
    pvk = zk.PreparedVerifyingKey.prepare(ctx, crs.params_bytes)        # a proving-key buffer is accepted as is
    vk = pr.vk_read(crs.params_bytes)
    assert pvk.write() == pr.pvk_write(vk)
    ab = pr.pairing_reference(vk["alpha_g1"], vk["beta_g2"])
    gam, dlt = pr.g2_prepare(pr.ec_neg(pr.FQ2, vk["gamma_g2"])), pr.g2_prepare(pr.ec_neg(pr.FQ2, vk["delta_g2"]))
    rng = pr.SplitMix64(8)
    proofs, inputs = [], []
    for seed in range(1, 5):
        z, proof = _prove(r1cs, params, seed, rng.fr(), rng.fr())
        proofs.append(proof); inputs.append(z[1:4])
    # tampered variants of proof 0: wrong input, C + G, A negated, B replaced by another proof's B, proofs swapped inputs
    A, B, C = pr.proof_read(proofs[0])
    cases = [(proofs[0], [inputs[0][0], inputs[0][1], (inputs[0][2] + 1) % pr.R]),
             (pr.proof_bytes(A, B, pr.ec_add(pr.FQ, C, pr.G1_GEN)), inputs[0]),
             (pr.proof_bytes(pr.ec_neg(pr.FQ, A), B, C), inputs[0]),
             (pr.proof_bytes(A, pr.proof_read(proofs[1])[1], C), inputs[0]),
             (proofs[1], inputs[2]), (proofs[0], [0, 0, 0]), (proofs[0], [pr.R - 1] * 3)]
    all_p = proofs + [c[0] for c in cases]
    all_i = inputs + [c[1] for c in cases]
    got = zk.verify_proofs(pvk, b"".join(all_p), all_i)
    want = [int(pr.verify_prepared(ab, gam, dlt, vk["ic"], pr.proof_read(p), x)) for p, x in zip(all_p, all_i)]
    assert want == [1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
    assert got == want
    # the thread-per-proof kernels (kept as the A/B reference of the lane-parallel ones) give the same verdicts
    ctx.set_opt(zk.Context.OPT_VERIFY_LANES, 0)
    try:
        assert zk.verify_proofs(pvk, b"".join(all_p), all_i) == want
    finally:
        ctx.set_opt(zk.Context.OPT_VERIFY_LANES, 1)
    assert zk.verify_proof(pvk, proofs[0], inputs[0]) is True
    # the same key loaded from its PreparedVerifyingKey::write image gives the same verdicts
    k2 = zk.PreparedVerifyingKey.read(ctx, pvk.write())
    assert zk.verify_proofs(k2, b"".join(all_p), all_i) == want
    # MalformedVerifyingKey (verifier.rs:38-40) and a non-canonical public input
    with pytest.raises(zk.SynthesisError) as e:
        zk.verify_proofs(pvk, proofs[0], [inputs[0][:2]])
    assert e.value.code == -9
    with pytest.raises(zk.SynthesisError) as e:
        zk.verify_proofs(pvk, proofs[0], [[pr.R, 1, 2]])
    assert e.value.code == -8
    assert zk.verify_proofs(pvk, b"", []) == []
    pvk.free(); k2.free(); params.free()


def test_proof_read_rejections_on_device(ctx):
    """Every Proof::read failure class (lib.rs:67-108) gets the verdict the reference's error maps to; a good proof in the
    same batch is unaffected."""
    r1cs, crs, params = _setup(ctx, seed=9)
    pvk = zk.PreparedVerifyingKey.prepare(ctx, crs.params_bytes)
    z, good = _prove(r1cs, params, 1, 5, 6)
    x = 1
    while pr.FQ.sqrt((x ** 3 + 4) % pr.Q) is not None:
        x += 1
    off_curve = bytes([0x80]) + x.to_bytes(47, "big")
    x = 0
    while True:
        y = pr.FQ.sqrt((x ** 3 + 4) % pr.Q)
        if y is not None and pr.ec_mul(pr.FQ, (x, y), pr.R) is not pr.INF:
            break
        x += 1
    off_group = pr.g1_compressed((x, y))
    inf1, inf2 = bytes([0xC0]) + bytes(47), bytes([0xC0]) + bytes(95)
    bad = [
        (bytes([good[0] & 0x7F]) + good[1:], 2),                       # A: compression flag missing
        (inf1 + good[48:], 3),                                         # A = O
        (good[:48] + inf2 + good[144:], 3),                            # B = O
        (good[:144] + inf1, 3),                                        # C = O
        (bytes([0xC0]) + bytes(46) + b"\x01" + good[48:], 2),          # infinity flag with stray bits
        (bytes([0x9F]) + b"\xff" * 47 + good[48:], 2),                 # x >= q
        (off_curve + good[48:], 2), (good[:144] + off_curve, 2),       # no such point
        (off_group + good[48:], 2),                                    # on the curve, outside the subgroup
        (good[:48] + bytes([good[48] & 0x7F]) + good[49:], 2),         # B: compression flag missing
        (inf1 + bytes([good[48] & 0x7F]) + good[49:], 3),              # A = O is reported before B's bad flag (read order)
        (bytes([good[0] ^ 0x20]) + good[1:], 0),                       # A -> -A: well-formed, just false
    ]
    for p, code in bad:
        with_oracle = None
        try:
            pr.proof_read(p)
        except ValueError as e:
            with_oracle = 3 if "PointInfinity" in str(e) else 2
        assert (with_oracle or 0) == (code if code >= 2 else 0)
    got = zk.verify_proofs(pvk, b"".join(p for p, _ in bad) + good, [z[1:4]] * (len(bad) + 1))
    assert got == [c for _, c in bad] + [1]
    with pytest.raises(zk.ZkError):
        zk.verify_proof(pvk, bad[1][0], z[1:4])
    pvk.free(); params.free()


def test_verify_full_batch_round_trip(ctx):
    """Size-independent property at the batch size bench.py uses: prove -> verify accepts all; flipping one public input
    per proof rejects exactly those."""
    r1cs, crs, params = _setup(ctx, seed=13)
    pvk = zk.PreparedVerifyingKey.prepare(ctx, crs.params_bytes)
    rng = pr.SplitMix64(2)
    base = [_prove(r1cs, params, s, rng.fr(), rng.fr()) for s in range(1, 9)]
    n = 1024
    proofs = b"".join(base[i % 8][1] for i in range(n))
    inputs = [list(base[i % 8][0][1:4]) for i in range(n)]
    assert zk.verify_proofs(pvk, proofs, inputs) == [1] * n
    flip = set(range(0, n, 7))
    for i in flip:
        inputs[i][i % 3] = (inputs[i][i % 3] + 1) % pr.R
    want = [0 if i in flip else 1 for i in range(n)]
    assert zk.verify_proofs(pvk, proofs, inputs) == want
    # the slicing path of very large batches (slice: 2^18 proofs): 2^18 + 333 proofs, the tampered pattern continued
    big = (1 << 18) + 333
    reps = (big + n - 1) // n
    got = zk.verify_proofs(pvk, (proofs * reps)[:192 * big], (inputs * reps)[:big])
    assert got == (want * reps)[:big]
    pvk.free(); params.free()


def test_verify_key_without_public_inputs(ctx):
    """ic.len() == 1 (only the constant ONE): the public-input sum is ic[0] itself.  The key is the toy key with its ic cut
    to one element, so no proof can be valid under it; what is checked is that the path agrees with the oracle."""
    r1cs, crs, params = _setup(ctx, seed=17)
    vkb = crs.params_bytes[:864] + (1).to_bytes(4, "big") + crs.params_bytes[868:868 + 96]
    pvk = zk.PreparedVerifyingKey.prepare(ctx, vkb)
    opvk = co.PreparedVerifyingKey.prepare(vkb)
    assert pvk.num_inputs == 0 and pvk.write() == opvk.write()
    z, proof = _prove(r1cs, params, 1, 3, 4)
    got = zk.verify_proofs(pvk, proof * 2, [[], []])
    assert got == opvk.verify_batch(proof * 2, np.zeros(0, np.uint64), 0) == [0, 0]
    with pytest.raises(zk.SynthesisError) as e:
        zk.verify_proofs(pvk, proof, [[1]])
    assert e.value.code == -9
    pvk.free(); params.free()


def test_reference_literal_proof_is_read_by_the_device(ctx):
    """The reference's own 192-byte proof literal (core/primitives/src/proof.rs:89): three points this repository did not make go
    through the device Proof::read (square roots, sign bits, subgroup tests) and the pairing check under the shipped key; the
    verdict must be the oracle's (a proper `false`: read succeeded, the proof belongs to other inputs)."""
    import json
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    raw = bytes.fromhex(json.load(open(os.path.join(gold, "kats.json")))["proof_kat"]["proof_hex"])
    pvk = zk.PreparedVerifyingKey.read(ctx, open(os.path.join(gold, "conf_pvk.dat"), "rb").read())
    opvk = co.PreparedVerifyingKey.read(open(os.path.join(gold, "conf_pvk.dat"), "rb").read())
    inputs = list(range(1, 23))
    bad_flag = bytes([raw[0] & 0x7f]) + raw[1:]
    minus_a = bytes([raw[0] ^ 0x20]) + raw[1:]
    batch = raw + bad_flag + minus_a + raw
    got = zk.verify_proofs(pvk, batch, [inputs] * 4)
    want = opvk.verify_batch(batch, co.ints_to_limbs(inputs * 4, 4), 22)
    assert got == want == [0, 2, 0, 0]
    pvk.free()
