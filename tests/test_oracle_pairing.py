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
