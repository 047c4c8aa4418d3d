"""Verifier-side oracle: the big-integer pairing in oracle/pyref.py is pinned by the reference's
fixture e(alpha_g1, beta_g2) = conf_vk.dat[0:576] and then used as the acceptance check the reference's own prover tests
apply (`check_proof` / verify_proof, core/proofs/src/confidential.rs:208-278, core/bellman-verifier/src/verifier.rs:32-63):
a proof produced by the prover oracle must verify under the CRS's verifying key, and a tampered one must not."""
import json
import os

import pytest

from oracle import coracle as co
from oracle import pyref as pr
from zero_chain_b200 import synthetic as sy

K = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kats.json")))


def test_pairing_matches_reference_fixture():
    kat = K["pairing_kat"]
    alpha = pr.g1_from_uncompressed(bytes.fromhex(kat["alpha_g1_uncompressed"]))
    beta = pr.g2_from_uncompressed(bytes.fromhex(kat["beta_g2_uncompressed"]))
    assert pr.ec_on_curve(pr.FQ, alpha) and pr.ec_on_curve(pr.FQ2, beta)
    assert pr.f12_to_tower_bytes(pr.pairing_reference(alpha, beta)).hex() == kat["alpha_g1_beta_g2_fq12"]


def test_pairing_bilinear_and_nondegenerate():
    e = pr.pairing(pr.G1_GEN, pr.G2_GEN)
    assert e != pr.F12_ONE and pr._f12_pow(e, pr.R) == pr.F12_ONE
    a, b = 0x1234567, 0x89abcdef01
    assert pr.pairing(pr.ec_mul(pr.FQ, pr.G1_GEN, a), pr.ec_mul(pr.FQ2, pr.G2_GEN, b)) == pr._f12_pow(e, a * b)
    assert pr.pairing(pr.INF, pr.G2_GEN) == pr.F12_ONE


def vk_from_params(buf: bytes):
    lay = pr.params_layout(buf)
    off, n = lay["ic"]
    return dict(alpha_g1=pr.g1_from_uncompressed(buf[0:96]), beta_g2=pr.g2_from_uncompressed(buf[192:384]),
                gamma_g2=pr.g2_from_uncompressed(buf[384:576]), delta_g2=pr.g2_from_uncompressed(buf[672:864]),
                ic=[pr.g1_from_uncompressed(buf[off + 96 * i:off + 96 * i + 96]) for i in range(n)])


def proof_points(proof: bytes):
    return (pr.g1_from_compressed(proof[0:48]), pr.g2_from_compressed(proof[48:144]), pr.g1_from_compressed(proof[144:192]))


def test_oracle_proof_verifies_and_tampered_fails():
    shape = dict(n_constraints=60, n_inputs=4, n_aux=50, a_aux_density=40, b_density=33)
    r1cs = sy.make_r1cs(seed=3, **shape)
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=4)
    z = sy.make_witness(r1cs, 1)
    a, b, c = sy.evaluate(r1cs, z)
    dens = sy.densities(r1cs)
    P = co.Params(crs.params_bytes, checked=True)
    proof = P.prove(co.ints_to_limbs(a, 4), co.ints_to_limbs(b, 4), co.ints_to_limbs(c, 4), co.ints_to_limbs(z[:4], 4),
                    co.ints_to_limbs(z[4:], 4), *dens, 0xabcdef, 0x123456789)
    vk = vk_from_params(crs.params_bytes)
    assert pr.groth16_verify(vk, proof_points(proof), z[1:4])
    assert not pr.groth16_verify(vk, proof_points(proof), [z[1], z[2], (z[3] + 1) % pr.R])        # wrong public input
    A, B, C = proof_points(proof)
    assert not pr.groth16_verify(vk, (A, B, pr.ec_add(pr.FQ, C, pr.G1_GEN)), z[1:4])               # tampered C


GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["conf", "anony"])
def test_prepared_verifying_key_matches_shipped_file(name):
    """prepare_verifying_key (verifier.rs:15-30) + PreparedVerifyingKey::write (lib.rs:183-202) restated in the oracle
    reproduce zface/params/{conf,anony}_vk.dat from the VerifyingKey inside the matching proving key — pins g2_prepare's
    line coefficients (scaling included), the Fq12 encoding and the pairing value."""
    vk = pr.vk_read(open(os.path.join(GOLD, "%s_vk_head.bin" % name), "rb").read())
    want = open(os.path.join(GOLD, "%s_pvk.dat" % name), "rb").read()
    assert pr.pvk_write(vk) == want


def test_prepared_verifier_agrees_with_plain_verifier():
    shape = dict(n_constraints=40, n_inputs=3, n_aux=30, a_aux_density=25, b_density=20)
    r1cs = sy.make_r1cs(seed=5, **shape)
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=6)
    z = sy.make_witness(r1cs, 2)
    a, b, c = sy.evaluate(r1cs, z)
    P = co.Params(crs.params_bytes, checked=True)
    proof = P.prove(co.ints_to_limbs(a, 4), co.ints_to_limbs(b, 4), co.ints_to_limbs(c, 4), co.ints_to_limbs(z[:3], 4),
                    co.ints_to_limbs(z[3:], 4), *sy.densities(r1cs), 77, 99)
    vk = pr.vk_read(crs.params_bytes)
    ab = pr.pairing_reference(vk["alpha_g1"], vk["beta_g2"])
    gam, dlt = pr.g2_prepare(pr.ec_neg(pr.FQ2, vk["gamma_g2"])), pr.g2_prepare(pr.ec_neg(pr.FQ2, vk["delta_g2"]))
    pts = pr.proof_read(proof)
    assert pts == proof_points(proof)
    assert pr.verify_prepared(ab, gam, dlt, vk["ic"], pts, z[1:3])
    assert not pr.verify_prepared(ab, gam, dlt, vk["ic"], pts, [z[1], (z[2] + 1) % pr.R])
    with pytest.raises(ValueError, match="MalformedVerifyingKey"):
        pr.verify_prepared(ab, gam, dlt, vk["ic"], pts, z[1:2])


def test_proof_read_rejections():
    """Proof::read (lib.rs:67-108): InvalidData for bad encodings / off-curve / out-of-subgroup points, PointInfinity for O."""
    good = pr.proof_bytes(pr.ec_mul(pr.FQ, pr.G1_GEN, 5), pr.ec_mul(pr.FQ2, pr.G2_GEN, 7), pr.ec_mul(pr.FQ, pr.G1_GEN, 11))
    assert pr.proof_read(good)[0] == pr.ec_mul(pr.FQ, pr.G1_GEN, 5)
    inf1 = bytes([0xC0]) + bytes(47)
    with pytest.raises(ValueError, match="PointInfinity"):
        pr.proof_read(inf1 + good[48:])
    with pytest.raises(ValueError, match="InvalidData"):
        pr.proof_read(bytes([good[0] & 0x7F]) + good[1:])                    # compression bit cleared
    with pytest.raises(ValueError, match="InvalidData"):
        pr.proof_read(bytes([0xC0]) + bytes(46) + b"\x01" + good[48:])       # infinity with stray bits
    with pytest.raises(ValueError, match="InvalidData"):
        pr.proof_read(bytes([0x9F]) + b"\xff" * 47 + good[48:])              # x >= q
    # x with no point on the curve / a curve point outside the r-torsion
    x = 1
    while pr.FQ.sqrt((x ** 3 + 4) % pr.Q) is not None:
        x += 1
    with pytest.raises(ValueError, match="InvalidData"):
        pr.proof_read(bytes([0x80 | (x >> 376)]) + (x & ((1 << 376) - 1)).to_bytes(47, "big") + good[48:])
    x = 0
    while True:
        y = pr.FQ.sqrt((x ** 3 + 4) % pr.Q)
        if y is not None and pr.ec_mul(pr.FQ, (x, y), pr.R) is not pr.INF:
            break
        x += 1
    with pytest.raises(ValueError, match="InvalidData"):
        pr.proof_read(pr.g1_compressed((x, y)) + good[48:])


@pytest.mark.parametrize("name", ["conf", "anony"])
def test_c_oracle_prepared_key_matches_shipped_file(name):
    """The C restatement (oracle/pairing_oracle.inc: tower, G2Prepared, Miller loop, the reference's final-exponentiation
    chain) against the same fixtures; it is the CPU baseline bench.py times and the checker of the big GPU batches."""
    head = open(os.path.join(GOLD, "%s_vk_head.bin" % name), "rb").read()
    want = open(os.path.join(GOLD, "%s_pvk.dat" % name), "rb").read()
    assert co.PreparedVerifyingKey.prepare(head).write() == want
    assert co.PreparedVerifyingKey.read(want).write() == want
    assert co.pairing(head[0:96], head[192:384]) == want[:576]


def test_c_oracle_verifier_agrees_with_pyref():
    shape = dict(n_constraints=40, n_inputs=3, n_aux=30, a_aux_density=25, b_density=20)
    r1cs = sy.make_r1cs(seed=5, **shape)
    crs = sy.make_toy_crs(r1cs, co.g1_fixed_base, co.g2_fixed_base, seed=6)
    P = co.Params(crs.params_bytes, checked=True)
    vk = pr.vk_read(crs.params_bytes)
    k = co.PreparedVerifyingKey.prepare(crs.params_bytes)
    assert k.write() == pr.pvk_write(vk)
    ab = pr.pairing_reference(vk["alpha_g1"], vk["beta_g2"])
    gam, dlt = pr.g2_prepare(pr.ec_neg(pr.FQ2, vk["gamma_g2"])), pr.g2_prepare(pr.ec_neg(pr.FQ2, vk["delta_g2"]))
    proofs, inputs = [], []
    for seed in (1, 2, 3):
        z = sy.make_witness(r1cs, seed)
        a, b, c = sy.evaluate(r1cs, z)
        proofs.append(P.prove(co.ints_to_limbs(a, 4), co.ints_to_limbs(b, 4), co.ints_to_limbs(c, 4), co.ints_to_limbs(z[:3], 4),
                              co.ints_to_limbs(z[3:], 4), *sy.densities(r1cs), 7 * seed, 9 * seed))
        inputs.append(z[1:3])
    A, B, Cc = pr.proof_read(proofs[0])
    proofs += [proofs[0], pr.proof_bytes(A, B, pr.ec_add(pr.FQ, Cc, pr.G1_GEN)), bytes([0xC0]) + bytes(47) + proofs[0][48:],
               bytes([proofs[0][0] & 0x7F]) + proofs[0][1:], proofs[0][:48] + pr.g2_compressed(pr.ec_neg(pr.FQ2, B)) + proofs[0][144:]]
    inputs += [[inputs[0][0], (inputs[0][1] + 1) % pr.R], inputs[0], inputs[0], inputs[0], inputs[0]]
    got = k.verify_batch(b"".join(proofs), co.ints_to_limbs([v for row in inputs for v in row], 4), 2)
    want = []
    for p, x in zip(proofs, inputs):
        try:
            want.append(int(pr.verify_prepared(ab, gam, dlt, vk["ic"], pr.proof_read(p), x)))
        except ValueError as e:
            want.append(3 if "PointInfinity" in str(e) else 2)
    assert got == want == [1, 1, 1, 0, 0, 3, 2, 0]
    with pytest.raises(ValueError, match="MalformedVerifyingKey"):
        k.verify_batch(proofs[0], co.ints_to_limbs(inputs[0][:1], 4), 1)
    # a second sqrt algorithm (pyref: norm method; C: Algorithm 9) must land on the same points
    for kk in (3, 5, 1234567):
        q = pr.ec_mul(pr.FQ2, pr.G2_GEN, kk)
        pp = pr.proof_bytes(pr.ec_mul(pr.FQ, pr.G1_GEN, kk), q, pr.G1_GEN)
        assert k.verify_batch(pp, co.ints_to_limbs(inputs[0], 4), 2) == [0]


def test_reference_literal_proof_reads_and_writes_back():
    """The 192-byte proof the reference's own test holds (core/primitives/src/proof.rs:86-98, test_proof_into_from): Proof::read
    accepts it and Proof::write returns the same bytes — in both oracles; and the SCALE wrapper round-trips it."""
    import json, os
    K = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kats.json")))
    raw = bytes.fromhex(K["proof_kat"]["proof_hex"])
    assert len(raw) == 192
    a, b, c = pr.proof_read(raw)                                   # big-integer oracle: decompression, sign bits, r-torsion of A, B, C
    assert pr.proof_bytes(a, b, c) == raw
    from zero_chain_b200 import groth16 as zk
    w = zk.Proof.from_slice(raw)
    enc = w.encode()
    assert enc[:2] == bytes([0x01, 0x03]) and len(enc) == 194       # compact(192) = (192 << 2) | 1 little-endian
    assert zk.Proof.decode(enc) == w and zk.Proof.decode(enc).as_bytes() == raw and str(w) == "0x" + raw.hex()
    # C oracle: verdict must be a proper boolean (Proof::read succeeded), never InvalidData / PointInfinity
    vk = open(os.path.join(os.path.dirname(__file__), "golden", "conf_vk_head.bin"), "rb").read()
    k = co.PreparedVerifyingKey.prepare(vk)
    inputs = co.ints_to_limbs(list(range(1, 23)), 4)
    assert k.verify_batch(raw, inputs, 22) == [0]
    assert k.verify_batch(bytes([raw[0] & 0x7f]) + raw[1:], inputs, 22) == [2]
