"""CPU check of the PRODUCT's warp-cooperative code — zero_chain_b200/csrc/pairing_lanes.cuh (the verifier's Fq12 spread over
six lanes: convolution products, the symmetric-pair squaring table, sparse line products, Granger-Scott squaring on lanes,
Frobenius, inversion, the merged three-pair Miller loop, the final exponentiation) and curve_coop.cuh (XYZZ doubling / addition
with the products of one stage on different lanes) — compiled with ZK_HOST_EMUL and a SIMT shim (one host thread per lane,
shuffles through a barrier; tests/host_emul/emul_lanes.cpp).  Checked against the oracle's independent pairing (oracle/pyref.py)
and group law; the PTX build of the same source is covered by the -m gpu tests (test_gpu_verify.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as pr
from tests.test_host_emul_pairing import Q, _p, arr_f12, f12_arr, fq_words, g1_arr, g2_arr, rand_f12, words_fq

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MUL, SQR, CYC, INV, CONJ, FROB1, FROB2, FROB3, FINAL, MULW2, EXPX = range(11)
BLS_X = 0xd201000000010000


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("emull")
    so, sop = str(d / "libemull.so"), str(d / "libemulp.so")
    inc = os.path.join(ROOT, "zero_chain_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-I", inc, "-o", so, os.path.join(HERE, "host_emul", "emul_lanes.cpp")])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", inc, "-o", sop, os.path.join(HERE, "host_emul", "emul_pairing.cpp")])
    lib = C.CDLL(so)
    lib.thread_version = C.CDLL(sop)           # pairing.cuh on one thread: the g2_prepare used for the Miller-loop inputs
    assert lib.emu_n_coeffs() == 68
    return lib


def _op(emu, op, a, b=None, lanes=6):
    o = np.zeros(144, np.uint32)
    A = f12_arr(a)
    B = f12_arr(b) if b is not None else None
    bad = emu.emu_lanes_f12(op, lanes, _p(A), _p(B) if B is not None else None, _p(o))
    assert bad == 0                            # every group of the warp computed the same slots
    return arr_f12(o)


def test_lane_products_against_the_oracle(emu):
    rng = pr.SplitMix64(606)
    for _ in range(4):
        a, b = rand_f12(rng), rand_f12(rng)
        assert _op(emu, MUL, a, b) == pr._f12_mul(a, b)
        assert _op(emu, SQR, a) == pr._f12_mul(a, a)                       # ZK_SQR_PAIRS: every unordered pair once, doubled / times xi
        assert pr._f12_mul(_op(emu, INV, a), a) == pr.F12_ONE
        assert _op(emu, CONJ, a) == pr._f12_pow(a, Q ** 6)
        for k, op in ((1, FROB1), (2, FROB2), (3, FROB3)):
            assert _op(emu, op, a) == pr._f12_pow(a, Q ** k)
        v = [0] * 12; v[2] = 1                                             # w^2 in the polynomial basis
        assert _op(emu, MULW2, a) == pr._f12_mul(a, v)
    # sparse and edge operands: 1, 0, a single slot, the largest coordinates
    one = list(pr.F12_ONE)
    e5 = [0] * 12; e5[5] = 1
    top = [Q - 1] * 12
    for a in (one, [0] * 12, e5, top):
        for b in (one, e5, top):
            assert _op(emu, MUL, a, b) == pr._f12_mul(a, b)
        assert _op(emu, SQR, a) == pr._f12_mul(a, a)


def test_cyclotomic_square_and_exp_x_on_lanes(emu):
    rng = pr.SplitMix64(707)
    g = pr._f12_pow(rand_f12(rng), (Q ** 6 - 1) * (Q ** 2 + 1))          # a member of the cyclotomic subgroup
    assert _op(emu, CYC, g) == pr._f12_mul(g, g)
    a = rand_f12(rng)
    assert _op(emu, CYC, a) != pr._f12_mul(a, a)                           # the shortcut is only valid on the subgroup
    # f^x, x = -0xd201000000010000: |x|-th power then the conjugate (= inverse on the subgroup)
    assert _op(emu, EXPX, g) == pr._f12_pow(pr._f12_pow(g, BLS_X), Q ** 6)


def test_a_whole_warp_agrees_with_one_group(emu):
    """32 lanes: five groups of six plus the two shadow lanes (30, 31 mirror 24, 25) run the same program."""
    rng = pr.SplitMix64(808)
    a, b = rand_f12(rng), rand_f12(rng)
    assert _op(emu, MUL, a, b, lanes=32) == pr._f12_mul(a, b)
    assert _op(emu, SQR, a, lanes=32) == pr._f12_mul(a, a)
    assert pr._f12_mul(_op(emu, INV, a, lanes=32), a) == pr.F12_ONE
    # all_lanes: AND over the six lanes of a group
    assert emu.emu_lanes_all(32, -1) == 32 and emu.emu_lanes_all(6, -1) == 6
    for off in range(6):
        assert emu.emu_lanes_all(32, off) == 0 and emu.emu_lanes_all(6, off) == 0


def _prepared(emu, q):
    c = np.zeros(68 * 72, np.uint32)
    emu.thread_version.emu_g2_prepare(_p(g2_arr(q)), _p(c))
    return c


def test_merged_miller_loop_and_final_exponentiation(emu):
    """miller_loop3 = the product of the three pairs' Miller loops (one accumulator, one squaring per bit); a pair with a
    point at infinity contributes 1; final_exponentiation of it = the product of the three reference pairings."""
    ps = [pr.ec_mul(pr.FQ, pr.G1_GEN, k) for k in (0x1234567, 0x89abcd, 3)]
    qs = [pr.ec_mul(pr.FQ2, pr.G2_GEN, k) for k in (0x7654321, 5, 0xfedcba987)]
    coeffs = np.concatenate([_prepared(emu, q) for q in qs])
    pts = np.concatenate([g1_arr(p) for p in ps])
    want = [pr.miller_loop_prepared(p, pr.g2_prepare(q)) for p, q in zip(ps, qs)]
    o = np.zeros(144, np.uint32)
    for skip in (0, 0b010, 0b101, 0b111):
        emu.emu_lanes_miller3(_p(pts), _p(coeffs), skip, _p(o))
        f = list(pr.F12_ONE)
        for k in range(3):
            if not (skip >> k) & 1:
                f = pr._f12_mul(f, want[k])
        assert arr_f12(o) == f, skip
    emu.emu_lanes_miller3(_p(pts), _p(coeffs), 0, _p(o))
    e = _op(emu, FINAL, arr_f12(o))
    ref = list(pr.F12_ONE)
    for p, q in zip(ps, qs):
        ref = pr._f12_mul(ref, pr.pairing_reference(p, q))               # independent affine Miller loop + plain power
    assert e == ref
    # bilinearity through the lanes: e(aP, Q) e(-P, aQ) e(P, Q) = e(P, Q)
    a = 0x5eed
    P, Qg = pr.G1_GEN, pr.G2_GEN
    ps = [pr.ec_mul(pr.FQ, P, a), pr.ec_neg(pr.FQ, P), P]
    qs = [Qg, pr.ec_mul(pr.FQ2, Qg, a), Qg]
    coeffs = np.concatenate([_prepared(emu, q) for q in qs])
    pts = np.concatenate([g1_arr(p) for p in ps])
    emu.emu_lanes_miller3(_p(pts), _p(coeffs), 0, _p(o))
    assert _op(emu, FINAL, arr_f12(o)) == pr.pairing_reference(P, Qg)


@pytest.mark.parametrize("g", [1, 2])
def test_cooperative_point_operations(emu, g):
    """curve_coop.cuh dbl / add on 4 and on 32 lanes = the one-thread XYZZ formulas (and the oracle's group law), incl. the
    exceptional cases the serial tails of the MSM meet: P + P, P + (-P), an infinity operand on either side."""
    w = 12 if g == 1 else 24
    fixed = co.g1_fixed_base if g == 1 else co.g2_fixed_base
    add = co.g1_add if g == 1 else co.g2_add
    dbl = co.g1_double if g == 1 else co.g2_double
    fn = emu.emu_coop_g1 if g == 1 else emu.emu_coop_g2
    pts = np.ascontiguousarray(fixed(co.ints_to_limbs([1, 2, 7, pr.R - 3, 0x123456789abcdef], 4)))
    inf = np.zeros(w, np.uint64)
    o = np.zeros(w, np.uint64)
    for lanes in (4, 32):
        for a in pts[: (5 if lanes == 4 else 2)]:
            for b in pts[: (5 if lanes == 4 else 2)]:
                a2, b3 = dbl(a), add(dbl(b), b)
                for op, want in ((0, dbl(a2)), (1, add(a2, b3)), (2, dbl(a2)), (3, inf), (4, b3), (5, a2)):
                    assert fn(op, lanes, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(o)) == 0, (op, lanes)
                    assert np.array_equal(o, want), (op, lanes)


def test_warp_products_of_the_batched_affine_rounds(emu):
    """msm_warp_scan.cuh ba_warp_products on 32 lanes: `others` = the product of the other 31 thread totals (prefix x suffix of two
    Kogge-Stone scans), `all` = the warp total in every lane — what lets ONE inversion per round serve every thread."""
    rng = pr.SplitMix64(909)
    for case in range(3):
        t = [rng.below(Q, 7) or 1 for _ in range(32)]
        if case == 1:
            t[0] = 1; t[31] = Q - 1; t[7] = 1                              # ones at the ends (threads without a division)
        if case == 2:
            t = [1] * 32; t[13] = 5
        T = np.array([w for v in t for w in fq_words(v)], dtype=np.uint32)
        O, A = np.zeros_like(T), np.zeros_like(T)
        emu.emu_warp_products_fq(_p(T), _p(O), _p(A))
        total = 1
        for v in t:
            total = total * v % Q
        for l in range(32):
            assert words_fq(A[12 * l:12 * l + 12]) == total
            assert words_fq(O[12 * l:12 * l + 12]) == total * pow(t[l], -1, Q) % Q
    # Fq2 totals (the G2 rounds)
    t2 = [(rng.below(Q, 7), rng.below(Q, 7)) for _ in range(32)]
    T = np.array([w for v in t2 for c in v for w in fq_words(c)], dtype=np.uint32)
    O, A = np.zeros_like(T), np.zeros_like(T)
    emu.emu_warp_products_fq2(_p(T), _p(O), _p(A))
    F = pr.FQ2
    total = (1, 0)
    for v in t2:
        total = F.mul(total, v)
    for l in range(32):
        got_all = (words_fq(A[24 * l:24 * l + 12]), words_fq(A[24 * l + 12:24 * l + 24]))
        got_oth = (words_fq(O[24 * l:24 * l + 12]), words_fq(O[24 * l + 12:24 * l + 24]))
        assert got_all == total and F.mul(got_oth, t2[l]) == total
