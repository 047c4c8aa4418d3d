"""CPU check of the PRODUCT's pairing code (zero_chain_b200/csrc/pairing.cuh) compiled with ZK_HOST_EMUL against the
oracle's independent pairing (oracle/pyref.py: polynomial-basis Fq12, affine Miller loop, plain-power final
exponentiation) and against the reference's fixture conf_vk.dat (tests/golden/)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyref as pr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
Q = pr.Q


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emulp") / "libemulp.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "zero_chain_b200", "csrc"),
                           "-o", so, os.path.join(HERE, "host_emul", "emul_pairing.cpp")])
    lib = C.CDLL(so)
    assert lib.emu_sizeof_fq12() == 576 and lib.emu_sizeof_coeff() == 288
    return lib


def fq_words(x):            # canonical int -> Montgomery LE u32 limbs
    m = pr.fq_to_mont(x % Q)
    return [(m >> (32 * i)) & 0xFFFFFFFF for i in range(12)]


def words_fq(w):
    return pr.fq_from_mont(sum(int(v) << (32 * i) for i, v in enumerate(w)))


def poly_to_tower(a):
    """polynomial basis (pyref) -> the 12 Fq values in tower memory order (w^i, v^j, u^k)."""
    out = []
    for i in range(2):
        for j in range(3):
            t = 2 * j + i
            out += [(a[t] + a[t + 6]) % Q, a[t + 6] % Q]
    return out


def tower_to_poly(c):
    a = [0] * 12
    n = 0
    for i in range(2):
        for j in range(3):
            t = 2 * j + i
            c0, c1 = c[n], c[n + 1]
            n += 2
            a[t] = (c0 - c1) % Q
            a[t + 6] = c1 % Q
    return a


def f12_arr(a):
    return np.array([w for v in poly_to_tower(a) for w in fq_words(v)], dtype=np.uint32)


def arr_f12(o):
    return tower_to_poly([words_fq(o[12 * i:12 * i + 12]) for i in range(12)])


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def rand_f12(rng):
    return [rng.below(Q, 7) for _ in range(12)]


def test_tower_arithmetic(emu):
    rng = pr.SplitMix64(77)
    o = np.zeros(144, np.uint32)
    for _ in range(6):
        a, b = rand_f12(rng), rand_f12(rng)
        A, B = f12_arr(a), f12_arr(b)
        emu.emu_f12_mul(_p(A), _p(B), _p(o)); assert arr_f12(o) == pr._f12_mul(a, b)
        emu.emu_f12_sqr(_p(A), _p(o)); assert arr_f12(o) == pr._f12_mul(a, a)
        emu.emu_f12_inv(_p(A), _p(o)); assert pr._f12_mul(arr_f12(o), a) == pr.F12_ONE
    # Granger-Scott squaring: valid exactly on the cyclotomic subgroup, i.e. after the easy part f^((q^6-1)(q^2+1))
    g = pr._f12_pow(rand_f12(rng), (Q ** 6 - 1) * (Q ** 2 + 1))
    emu.emu_f12_cyc_sqr(_p(f12_arr(g)), _p(o)); assert arr_f12(o) == pr._f12_mul(g, g)
    a = rand_f12(rng)
    A = f12_arr(a)
    emu.emu_f12_cyc_sqr(_p(A), _p(o)); assert arr_f12(o) != pr._f12_mul(a, a)
    for k in (1, 2, 3):
        emu.emu_f12_frob(_p(A), k, _p(o))
        assert arr_f12(o) == pr._f12_pow(a, Q ** k)


def g2_arr(q):
    return np.array(fq_words(q[0][0]) + fq_words(q[0][1]) + fq_words(q[1][0]) + fq_words(q[1][1]), dtype=np.uint32)


def g1_arr(p):
    return np.array(fq_words(p[0]) + fq_words(p[1]), dtype=np.uint32)


def test_prepare_miller_final_exp(emu):
    p = pr.ec_mul(pr.FQ, pr.G1_GEN, 0x1234567)
    q = pr.ec_mul(pr.FQ2, pr.G2_GEN, 0x7654321)
    co = np.zeros(68 * 72, np.uint32)
    emu.emu_g2_prepare(_p(g2_arr(q)), _p(co))
    want = pr.g2_prepare(q)
    assert len(want) == 68
    got = [tuple((words_fq(co[72 * i + 24 * j:72 * i + 24 * j + 12]), words_fq(co[72 * i + 24 * j + 12:72 * i + 24 * j + 24])) for j in range(3))
           for i in range(68)]
    assert got == [tuple(c) for c in want]
    f = np.zeros(144, np.uint32)
    emu.emu_miller(_p(g1_arr(p)), _p(co), _p(f))
    assert arr_f12(f) == pr.miller_loop_prepared(p, want)
    e = np.zeros(144, np.uint32)
    assert emu.emu_final_exp(_p(f), _p(e)) == 1
    assert arr_f12(e) == pr.pairing_reference(p, q)          # independent affine Miller loop + plain power, cubed
    z = np.zeros(144, np.uint32)
    assert emu.emu_final_exp(_p(z), _p(e)) == 0              # Engine::final_exponentiation -> None on zero


def test_compressed_decode(emu):
    """decode_compressed (codec.cuh): square roots, sign selection, flag / range / subgroup rejections, against pyref."""
    rng = pr.SplitMix64(5)
    o1, o2 = np.zeros(24, np.uint32), np.zeros(48, np.uint32)
    for _ in range(12):             # enough points to hit both branches of the Fq2 square root (d square / non-square)
        k = rng.below(pr.R, 4) or 1
        for neg in (False, True):
            p = pr.ec_mul(pr.FQ, pr.G1_GEN, k)
            q = pr.ec_mul(pr.FQ2, pr.G2_GEN, k + 1)
            if neg:
                p, q = pr.ec_neg(pr.FQ, p), pr.ec_neg(pr.FQ2, q)
            b1 = np.frombuffer(pr.g1_compressed(p), np.uint8)
            assert emu.emu_decode_g1c(_p(b1), _p(o1)) == 0
            assert (words_fq(o1[:12]), words_fq(o1[12:])) == p
            b2 = np.frombuffer(pr.g2_compressed(q), np.uint8)
            assert emu.emu_decode_g2c(_p(b2), _p(o2)) == 0
            assert ((words_fq(o2[:12]), words_fq(o2[12:24])), (words_fq(o2[24:36]), words_fq(o2[36:]))) == q
    good = pr.g1_compressed(pr.ec_mul(pr.FQ, pr.G1_GEN, 9))

    def code(b, g2=False):
        a = np.frombuffer(bytes(b), np.uint8)
        return (emu.emu_decode_g2c if g2 else emu.emu_decode_g1c)(_p(a), _p(o2))
    assert code(bytes([good[0] & 0x7F]) + good[1:]) == 1                       # UnexpectedCompressionMode
    assert code(bytes([0xC0]) + bytes(47)) == 0                                 # infinity decodes (Proof::read rejects it later)
    assert code(bytes([0xE0]) + bytes(47)) == 2 and code(bytes([0xC0]) + bytes(46) + b"\x01") == 2
    assert code(bytes([0x9F]) + b"\xff" * 47) == 3                              # x >= q
    x = 1
    while pr.FQ.sqrt((x ** 3 + 4) % pr.Q) is not None:
        x += 1
    assert code(bytes([0x80]) + x.to_bytes(47, "big")) == 4                     # NotOnCurve
    x = 0
    while True:
        y = pr.FQ.sqrt((x ** 3 + 4) % pr.Q)
        if y is not None and pr.ec_mul(pr.FQ, (x, y), pr.R) is not pr.INF:
            break
        x += 1
    assert code(pr.g1_compressed((x, y))) == 5                                  # NotInSubgroup
    # G2: an x whose right-hand side is a non-square, and the c1 = 0 branch of the square root
    xx = (1, 0)
    while pr.FQ2.sqrt(pr.FQ2.add(pr.FQ2.mul(pr.FQ2.mul(xx, xx), xx), pr.FQ2.b)) is not None:
        xx = (xx[0] + 1, 0)
    assert code(bytes([0x80]) + bytes(47) + xx[0].to_bytes(48, "big"), g2=True) == 4


def test_subgroup_checks_by_endomorphism(emu):
    """codec.cuh in_subgroup (phi / psi endomorphism tests) must agree with multiplication by r on members, on random curve
    points, on points of the cofactor torsion ([r] Q) and on member + torsion sums."""
    rng = pr.SplitMix64(404)

    def rand_point(F):
        while True:
            x = rng.below(Q, 7) if F is pr.FQ else (rng.below(Q, 7), rng.below(Q, 7))
            rhs = (x ** 3 + 4) % Q if F is pr.FQ else F.add(F.mul(F.mul(x, x), x), F.b)
            y = F.sqrt(rhs)
            if y is not None:
                return (x, y)
    for F, gen, arr, fn in ((pr.FQ, pr.G1_GEN, g1_arr, emu.emu_g1_subgroup), (pr.FQ2, pr.G2_GEN, g2_arr, emu.emu_g2_subgroup)):
        for k in (1, 2, 3, pr.R - 1, rng.below(pr.R, 4), rng.below(pr.R, 4)):
            assert fn(_p(arr(pr.ec_mul(F, gen, k)))) == 3
        n_tors = 0
        for _ in range(6):
            p = rand_point(F)
            assert fn(_p(arr(p))) == 0                                   # a random point is outside the subgroup
            t = pr.ec_mul(F, p, pr.R)                                    # cofactor torsion
            if t is not pr.INF:
                n_tors += 1
                assert fn(_p(arr(t))) == 0
                assert fn(_p(arr(pr.ec_add(F, t, pr.ec_mul(F, gen, 5))))) == 0
        assert n_tors
    # small-order points: the G1 cofactor is divisible by 3 and 11 — points of exactly those orders
    h1 = 0x396c8c005555e1568c00aaab0000aaab
    for small, mult in ((3, 1), (11, 2)):                       # E(Fq) is not cyclic: its 11-part is (Z/11)^2, exponent 11
        assert h1 % small ** mult == 0 and h1 % small ** (mult + 1) != 0
        while True:
            t = pr.ec_mul(pr.FQ, rand_point(pr.FQ), pr.R * (h1 // small ** mult))
            if t is not pr.INF:
                break
        assert pr.ec_mul(pr.FQ, t, small) is pr.INF and emu.emu_g1_subgroup(_p(g1_arr(t))) == 0
