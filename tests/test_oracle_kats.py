"""Pins the oracle (oracle/pyref.py big-int restatement AND oracle/zk_oracle.c) against the
reference's own known-answer vectors (tests/golden/kats.json, extracted by make_golden.py from
core/pairing/src/bls12_381/{fq,fr,fq2,ec}.rs) and its 4x1000-point encoding vector files
(core/pairing/src/bls12_381/tests/*.dat, compared by SHA-256: entry i = i*G, tests/mod.rs:55-79)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as pr

K = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kats.json")))


def G(name):
    return [sum(int(v, 16) << (64 * i) for i, v in enumerate(g)) for g in K["tests"][name]["groups"]]


def const(name):
    return [sum(int(v, 16) << (64 * i) for i, v in enumerate(g)) for g in K["consts"][name]]


# ------------------------------------------------------------------ constants
def test_constants():
    assert const("fq.rs::MODULUS")[0] == pr.Q
    assert const("fr.rs::MODULUS")[0] == pr.R
    assert const("fq.rs::R")[0] == pr.FQ_MONT_R and const("fr.rs::R")[0] == pr.FR_MONT_R
    assert const("fq.rs::R2")[0] == pr.FQ_MONT_R ** 2 % pr.Q
    assert const("fr.rs::R2")[0] == pr.FR_MONT_R ** 2 % pr.R
    assert const("fr.rs::GENERATOR")[0] == pr.fr_to_mont(7)
    assert const("fr.rs::ROOT_OF_UNITY")[0] == pr.fr_to_mont(pr.FR_ROOT_OF_UNITY)
    assert const("fq.rs::B_COEFF")[0] == pr.fq_to_mont(4)
    assert const("fq.rs::NEGATIVE_ONE")[0] == pr.fq_to_mont(pr.Q - 1)
    g1 = co.limbs_to_ints(co.g1_generator().reshape(2, 6))
    assert g1 == [const("fq.rs::G1_GENERATOR_X")[0], const("fq.rs::G1_GENERATOR_Y")[0]]
    assert g1 == [pr.fq_to_mont(pr.G1_GEN[0]), pr.fq_to_mont(pr.G1_GEN[1])]
    g2 = co.limbs_to_ints(co.g2_generator().reshape(4, 6))
    assert g2 == [const("fq.rs::G2_GENERATOR_X_C0")[0], const("fq.rs::G2_GENERATOR_X_C1")[0],
                  const("fq.rs::G2_GENERATOR_Y_C0")[0], const("fq.rs::G2_GENERATOR_Y_C1")[0]]
    assert g2 == [pr.fq_to_mont(c) for c in (*pr.G2_GEN[0], *pr.G2_GEN[1])]
    assert pr.ec_on_curve(pr.FQ, pr.G1_GEN) and pr.ec_on_curve(pr.FQ2, pr.G2_GEN)
    assert pr.ec_mul(pr.FQ, pr.G1_GEN, pr.R) is pr.INF
    assert co.g1_check(co.g1_generator()) == 3 and co.g2_check(co.g2_generator()) == 3


# ------------------------------------------------------------------ field KATs (raw Montgomery limbs)
@pytest.mark.parametrize("f,mod,mul,add,sub,sqr,into", [
    ("fq", pr.Q, co.fq_mul, co.fq_add, co.fq_sub, co.fq_sqr, co.fq_into_repr),
    ("fr", pr.R, co.fr_mul, co.fr_add, co.fr_sub, co.fr_sqr, co.fr_into_repr)])
def test_field_kats(f, mod, mul, add, sub, sqr, into):
    bits = 384 if f == "fq" else 256
    rinv = pow(1 << bits, -1, mod)
    # mul: fq.rs:2564-2588 / fr.rs:1241-1259  (a*b == c on raw Montgomery residues)
    a, b, c = G("%s.rs::test_%s_mul_assign" % (f, f))
    assert a * b * rinv % mod == c          # pyref (big-int definition of Montgomery product)
    assert mul(a, b) == c                   # C oracle
    # squaring: fq.rs:2635-2657 / fr.rs:1306-1326  (raw a, result given via from_repr => canonical)
    a, c = G("%s.rs::test_%s_squaring" % (f, f))
    assert a * a * rinv % mod == (c << bits) % mod
    assert into(sqr(a)) == c
    # add: groups g0, g0, g0+1, g3, g2+g3, q-1, x, y, x+y  (fq.rs:2330-2420)
    g = G("%s.rs::test_%s_add_assign" % (f, f))
    assert add(g[0], 0) == g[1] and add(g[0], 1) == g[2] and add(g[2], g[3]) == g[4]
    assert (g[2] + g[3]) % mod == g[4]
    assert add(g[5], 1) == 0 and add(g[6], g[7]) == g[8] and add(g[8], 1) == 0
    # sub: g0-g1=g2, g3-g4=g5, g6-0=g7 (fq.rs:2457-2530)
    g = G("%s.rs::test_%s_sub_assign" % (f, f))
    assert sub(g[0], g[1]) == g[2] == (g[0] - g[1]) % mod
    assert sub(g[3], g[4]) == g[5] == (g[3] - g[4]) % mod
    assert sub(g[6], 0) == g[7] and sub(0, 0) == 0


def test_field_random_vs_bigint():
    rng = pr.SplitMix64(11)
    for mod, bits, mul, add, sub, inv, frm, into, neg in (
            (pr.Q, 384, co.fq_mul, co.fq_add, co.fq_sub, co.fq_inv, co.fq_from_repr, co.fq_into_repr, co.fq_neg),
            (pr.R, 256, co.fr_mul, co.fr_add, co.fr_sub, co.fr_inv, co.fr_from_repr, co.fr_into_repr, co.fr_neg)):
        rinv = pow(1 << bits, -1, mod)
        edge = [0, 1, mod - 1, mod - 2, (1 << bits) % mod, 2]
        vals = edge + [rng.below(mod, bits // 64 + 1) for _ in range(300)]
        for i in range(len(vals) - 1):
            a, b = vals[i], vals[i + 1]
            assert mul(a, b) == a * b * rinv % mod
            assert add(a, b) == (a + b) % mod and sub(a, b) == (a - b) % mod
            assert neg(a) == (-a) % mod
            assert frm(a) == (a << bits) % mod and into(a) == a * rinv % mod
            if a:
                # inverse in Montgomery domain: inv(aR) = a^-1 R
                assert inv(a) == pow(a * rinv, -1, mod) * (1 << bits) % mod
        assert inv(0) is None
        assert frm(mod) is None and frm(mod + 5) is None     # from_repr rejects >= modulus


def _fq2_pack(c0, c1):
    return c0 | (c1 << 384)


def _fq2_unpack(v):
    return v & ((1 << 384) - 1), v >> 384


def test_fq2_kats():
    M = pr.fq_to_mont
    # mul KAT fq2.rs:369-431 (all via from_repr => canonical integers)
    a0, a1, b0, b1, c0, c1 = G("fq2.rs::test_fq2_mul")
    assert pr.FQ2.mul((a0, a1), (b0, b1)) == (c0, c1)
    assert _fq2_unpack(co.fq2_mul(_fq2_pack(M(a0), M(a1)), _fq2_pack(M(b0), M(b1)))) == (M(c0), M(c1))
    # squaring KAT fq2.rs:295-367 (third case)
    a0, a1, c0, c1 = G("fq2.rs::test_fq2_squaring")
    assert pr.FQ2.mul((a0, a1), (a0, a1)) == (c0, c1)
    assert _fq2_unpack(co.fq2_sqr(_fq2_pack(M(a0), M(a1)))) == (M(c0), M(c1))
    assert _fq2_unpack(co.fq2_sqr(_fq2_pack(M(1), M(1)))) == (0, M(2))          # (1+u)^2 = 2u
    assert _fq2_unpack(co.fq2_sqr(_fq2_pack(0, M(1)))) == (M(pr.Q - 1), 0)      # u^2 = -1
    # inverse KAT fq2.rs:433-480
    a0, a1, c0, c1 = G("fq2.rs::test_fq2_inverse")
    assert pr.FQ2.inv((a0, a1)) == (c0, c1)
    assert _fq2_unpack(co.fq2_inv(_fq2_pack(M(a0), M(a1)))) == (M(c0), M(c1))
    assert co.fq2_inv(0) is None
    rng = pr.SplitMix64(5)
    for _ in range(100):
        a, b = (rng.fq(), rng.fq()), (rng.fq(), rng.fq())
        assert _fq2_unpack(co.fq2_mul(_fq2_pack(M(a[0]), M(a[1])), _fq2_pack(M(b[0]), M(b[1])))) == tuple(M(x) for x in pr.FQ2.mul(a, b))
        assert _fq2_unpack(co.fq2_sqr(_fq2_pack(M(a[0]), M(a[1])))) == tuple(M(x) for x in pr.FQ2.mul(a, a))


# ------------------------------------------------------------------ curve KATs
def _g1(x, y):
    return co.ints_to_limbs([pr.fq_to_mont(x), pr.fq_to_mont(y)], 6).reshape(12)


def _g1_ints(p):
    if not np.any(p):
        return pr.INF
    x, y = co.limbs_to_ints(np.asarray(p).reshape(2, 6))
    return (pr.fq_from_mont(x), pr.fq_from_mont(y))


def _g2(p):
    (x0, x1), (y0, y1) = p
    return co.ints_to_limbs([pr.fq_to_mont(v) for v in (x0, x1, y0, y1)], 6).reshape(24)


def _g2_ints(p):
    if not np.any(p):
        return pr.INF
    v = [pr.fq_from_mont(x) for x in co.limbs_to_ints(np.asarray(p).reshape(4, 6))]
    return ((v[0], v[1]), (v[2], v[3]))


def test_g1_kats():
    # ec.rs:1070-1135 addition, 1138-1185 doubling, 1188-1272 same-y (add and mixed add)
    x1, y1, x2, y2, x3, y3 = G("ec.rs::test_g1_addition_correctness")
    assert pr.ec_add(pr.FQ, (x1, y1), (x2, y2)) == (x3, y3)
    assert _g1_ints(co.g1_add(_g1(x1, y1), _g1(x2, y2))) == (x3, y3)
    assert _g1_ints(co.g1_add_mixed(_g1(x1, y1), _g1(x2, y2))) == (x3, y3)
    x1, y1, x3, y3 = G("ec.rs::test_g1_doubling_correctness")
    assert pr.ec_double(pr.FQ, (x1, y1)) == (x3, y3)
    assert _g1_ints(co.g1_double(_g1(x1, y1))) == (x3, y3)
    assert _g1_ints(co.g1_add(_g1(x1, y1), _g1(x1, y1))) == (x3, y3)         # add falls back to double (ec.rs:394-397)
    assert _g1_ints(co.g1_add_mixed(_g1(x1, y1), _g1(x1, y1))) == (x3, y3)   # ec.rs:473-476
    x1, y1, x2, y2, x3, y3 = G("ec.rs::test_g1_same_y")
    assert y1 == y2
    assert pr.ec_add(pr.FQ, (x1, y1), (x2, y2)) == (x3, y3)
    assert _g1_ints(co.g1_add(_g1(x1, y1), _g1(x2, y2))) == (x3, y3)
    assert _g1_ints(co.g1_add_mixed(_g1(x1, y1), _g1(x2, y2))) == (x3, y3)
    # P + (-P) = infinity; infinity handling (ec.rs:357-365, 447-456)
    neg = _g1(x1, pr.Q - y1)
    z = np.zeros(12, np.uint64)
    assert not np.any(co.g1_add(_g1(x1, y1), neg)) and not np.any(co.g1_add_mixed(_g1(x1, y1), neg))
    assert np.array_equal(co.g1_add(z, _g1(x1, y1)), _g1(x1, y1)) and np.array_equal(co.g1_add_mixed(_g1(x1, y1), z), _g1(x1, y1))


def test_g2_kats():
    g = G("ec.rs::test_g2_addition_correctness")   # ec.rs:1868-1994
    p, q, s = ((g[0], g[1]), (g[2], g[3])), ((g[4], g[5]), (g[6], g[7])), ((g[8], g[9]), (g[10], g[11]))
    assert pr.ec_add(pr.FQ2, p, q) == s
    assert _g2_ints(co.g2_add(_g2(p), _g2(q))) == s and _g2_ints(co.g2_add_mixed(_g2(p), _g2(q))) == s
    g = G("ec.rs::test_g2_doubling_correctness")   # ec.rs:1996-2084
    p, s = ((g[0], g[1]), (g[2], g[3])), ((g[4], g[5]), (g[6], g[7]))
    assert pr.ec_double(pr.FQ2, p) == s
    assert _g2_ints(co.g2_double(_g2(p))) == s and _g2_ints(co.g2_add(_g2(p), _g2(p))) == s


def test_scalar_mul_vs_bigint():
    rng = pr.SplitMix64(3)
    for k in [0, 1, 2, pr.R - 1, pr.R, rng.fr(), rng.fr()]:
        assert _g1_ints(co.g1_mul(co.g1_generator(), k)) == pr.ec_mul(pr.FQ, pr.G1_GEN, k)
    for k in [0, 1, 3, pr.R - 1, rng.fr()]:
        assert _g2_ints(co.g2_mul(co.g2_generator(), k)) == pr.ec_mul(pr.FQ2, pr.G2_GEN, k)


# ------------------------------------------------------------------ encoding vector files
def _vector_file(group: int, compressed: bool, use_c: bool) -> bytes:
    """entry i = i*G, i = 0..999 (core/pairing/src/bls12_381/tests/mod.rs:55-79)."""
    out = bytearray()
    if use_c:
        gen = co.g1_generator() if group == 1 else co.g2_generator()
        cur = np.zeros(12 if group == 1 else 24, np.uint64)
        add = co.g1_add if group == 1 else co.g2_add
        enc = co.g1_encode if group == 1 else co.g2_encode
        for _ in range(1000):
            out += enc(cur, compressed)
            cur = add(cur, gen)
    else:
        F, gen = (pr.FQ, pr.G1_GEN) if group == 1 else (pr.FQ2, pr.G2_GEN)
        enc = {(1, False): pr.g1_uncompressed, (1, True): pr.g1_compressed,
               (2, False): pr.g2_uncompressed, (2, True): pr.g2_compressed}[(group, compressed)]
        cur = pr.INF
        for _ in range(1000):
            out += enc(cur)
            cur = pr.ec_add(F, cur, gen)
    return bytes(out)


@pytest.mark.parametrize("group,compressed", [(1, False), (1, True), (2, False), (2, True)])
def test_encoding_vector_files(group, compressed):
    name = "core/pairing/src/bls12_381/tests/g%d_%scompressed_valid_test_vectors.dat" % (group, "" if compressed else "un")
    want = K["files"][name]
    for use_c in (True, False):
        got = _vector_file(group, compressed, use_c)
        assert len(got) == want["size"]
        assert hashlib.sha256(got).hexdigest() == want["sha256"], (name, "C oracle" if use_c else "pyref")


def test_decode_roundtrip_and_rejects():
    pts = co.g1_fixed_base(co.ints_to_limbs([0, 1, 2, 12345, pr.R - 1], 4))
    assert not np.any(pts[0])
    enc = b"".join(co.g1_encode(p, False) for p in pts)
    assert np.array_equal(co.g1_decode_many(enc, checked=True), pts)
    for i, p in enumerate(pts):
        assert pr.g1_from_uncompressed(enc[96 * i:96 * i + 96]) == _g1_ints(p)
        assert pr.g1_from_compressed(co.g1_encode(p, True)) == _g1_ints(p)
    bad = bytearray(enc[96:192]); bad[95] ^= 1
    with pytest.raises(ValueError):
        co.g1_decode_many(bytes(bad), checked=True)         # NotOnCurve
    bad = bytearray(enc[96:192]); bad[0] |= 0x80
    with pytest.raises(ValueError):
        co.g1_decode_many(bytes(bad))                       # UnexpectedCompressionMode
    q2 = co.g2_fixed_base(co.ints_to_limbs([0, 1, 7, pr.R - 3], 4))
    enc2 = b"".join(co.g2_encode(p, False) for p in q2)
    assert np.array_equal(co.g2_decode_many(enc2, checked=True), q2)
    for p in q2:
        assert pr.g2_from_compressed(co.g2_encode(p, True)) == _g2_ints(p)
    # a point on the curve but outside the r-torsion must fail the checked decode (ec.rs:675-685)
    x = 4
    while True:
        y = pr.FQ.sqrt((x ** 3 + 4) % pr.Q)
        if y is not None and pr.ec_mul(pr.FQ, (x, y), pr.R) is not pr.INF:
            break
        x += 1
    raw = pr.g1_uncompressed((x, y))
    co.g1_decode_many(raw, checked=False)
    with pytest.raises(ValueError):
        co.g1_decode_many(raw, checked=True)
