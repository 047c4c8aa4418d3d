"""TEST INFRASTRUCTURE — Python big-integer restatement of the BLS12-381 arithmetic,
encodings and Groth16 algebra used on the prover hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
It is the *independent* pin for the C oracle (oracle/zk_oracle.c): it shares no code with
either the C oracle or the CUDA kernels (affine formulas + Python ints instead of
Jacobian/Montgomery limbs), and is itself pinned by the reference's known-answer vectors
(tests/golden/kats.json, extracted from core/pairing/src/bls12_381/{fq,fr,fq2,ec}.rs) and
by the 4x1000-point encoding vector files (core/pairing/src/bls12_381/tests/*.dat).

Reference anchors (file:line under the reference tree):
  Fq modulus / R / R2 / INV        core/pairing/src/bls12_381/fq.rs:5-43
  Fr modulus / R / R2 / GENERATOR  core/pairing/src/bls12_381/fr.rs:4-55
  Fq2 = Fq[u]/(u^2+1)              core/pairing/src/bls12_381/fq2.rs:109-158
  curve y^2 = x^3 + 4 / 4(1+u)     core/pairing/src/bls12_381/ec.rs:885-887,1567-1572
  point codecs                     core/pairing/src/bls12_381/ec.rs:686-867,1343-1549
  Proof codec                      core/bellman-verifier/src/lib.rs:55-110
  Parameters grammar               SURVEY.md §3.3 (upstream bellman 0.1.0 groth16::Parameters::write)
"""
from __future__ import annotations

import struct

# ---------------------------------------------------------------------------------------
# constants (fq.rs:5-13, fr.rs:4-10)
Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
FQ_MONT_R = (1 << 384) % Q
FR_MONT_R = (1 << 256) % R
FR_S = 32                      # fr.rs:47
FR_GENERATOR = 7               # fr.rs:38-44
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R - 1) >> FR_S, R)   # fr.rs:50-55 (Montgomery there)

G1_GEN = (
    0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
)
G2_GEN = (
    (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
     0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
    (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
     0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be),
)


def limbs64(x: int, n: int) -> list[int]:
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs64(l) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


def fq_to_mont(x): return (x << 384) % Q
def fq_from_mont(x): return (x * pow(1 << 384, -1, Q)) % Q
def fr_to_mont(x): return (x << 256) % R
def fr_from_mont(x): return (x * pow(1 << 256, -1, R)) % R


# ---------------------------------------------------------------------------------------
# field "classes" as plain function tables so G1/G2 share curve code
class _Fq:
    zero = 0
    one = 1
    b = 4
    @staticmethod
    def add(a, b): return (a + b) % Q
    @staticmethod
    def sub(a, b): return (a - b) % Q
    @staticmethod
    def mul(a, b): return (a * b) % Q
    @staticmethod
    def neg(a): return (-a) % Q
    @staticmethod
    def inv(a): return pow(a, -1, Q)
    @staticmethod
    def is_zero(a): return a == 0
    @staticmethod
    def sqrt(a):
        # q = 3 mod 4 (fq.rs:1152-1175)
        s = pow(a, (Q + 1) // 4, Q)
        return s if (s * s) % Q == a else None
    @staticmethod
    def gt(a, b):   # canonical integer order (fq.rs:708-713)
        return a > b


class _Fq2:
    zero = (0, 0)
    one = (1, 0)
    b = (4, 4)
    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
    @staticmethod
    def neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
    @staticmethod
    def inv(a):
        n = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
        return ((a[0] * n) % Q, (-a[1] * n) % Q)
    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0
    @staticmethod
    def gt(a, b):   # c1 compared first, then c0 (fq2.rs:21-30)
        return (a[1], a[0]) > (b[1], b[0])
    @staticmethod
    def sqrt(a):
        # generic: a = (x+yu)^2 ; use norm trick
        if a == (0, 0):
            return (0, 0)
        a0, a1 = a
        if a1 == 0:
            s = _Fq.sqrt(a0)
            if s is not None:
                return (s, 0)
            s = _Fq.sqrt((-a0) % Q)
            return (0, s) if s is not None else None
        n = _Fq.sqrt((a0 * a0 + a1 * a1) % Q)
        if n is None:
            return None
        half = pow(2, -1, Q)
        for nn in (n, (-n) % Q):
            x2 = ((a0 + nn) * half) % Q
            x = _Fq.sqrt(x2)
            if x is not None and x != 0:
                y = (a1 * pow(2 * x, -1, Q)) % Q
                if _Fq2.mul((x, y), (x, y)) == a:
                    return (x, y)
        return None


FQ, FQ2 = _Fq, _Fq2
INF = None   # affine infinity


# ---------------------------------------------------------------------------------------
# affine group law (mathematical definition; the reference's Jacobian formulas
# ec.rs:296-526 compute the same group law)
def ec_add(F, p, q):
    if p is INF:
        return q
    if q is INF:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if y1 == y2:
            return ec_double(F, p)
        return INF
    lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def ec_double(F, p):
    if p is INF:
        return INF
    x1, y1 = p
    if F.is_zero(y1):
        return INF
    xx = F.mul(x1, x1)
    lam = F.mul(F.add(F.add(xx, xx), xx), F.inv(F.add(y1, y1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x1)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def ec_neg(F, p):
    return INF if p is INF else (p[0], F.neg(p[1]))


def ec_mul(F, p, k: int):
    if k < 0:
        return ec_mul(F, ec_neg(F, p), -k)
    acc = INF
    for bit in bin(k)[2:] if k else "":
        acc = ec_double(F, acc)
        if bit == "1":
            acc = ec_add(F, acc, p)
    return acc


def ec_on_curve(F, p):
    if p is INF:
        return True
    x, y = p
    return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), F.b)


def ec_msm(F, bases, scalars):
    acc = INF
    for b, s in zip(bases, scalars):
        if s:
            acc = ec_add(F, acc, ec_mul(F, b, s))
    return acc


# ---------------------------------------------------------------------------------------
# encodings (ec.rs:686-867 G1, 1343-1549 G2)
def _be48(x: int) -> bytes:
    return x.to_bytes(48, "big")


def g1_uncompressed(p) -> bytes:
    if p is INF:
        return bytes([0x40]) + bytes(95)
    return _be48(p[0]) + _be48(p[1])


def g1_compressed(p) -> bytes:
    if p is INF:
        return bytes([0xC0]) + bytes(47)
    out = bytearray(_be48(p[0]))
    if FQ.gt(p[1], FQ.neg(p[1])):
        out[0] |= 0x20
    out[0] |= 0x80
    return bytes(out)


def g2_uncompressed(p) -> bytes:
    if p is INF:
        return bytes([0x40]) + bytes(191)
    (x0, x1), (y0, y1) = p
    return _be48(x1) + _be48(x0) + _be48(y1) + _be48(y0)


def g2_compressed(p) -> bytes:
    if p is INF:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), y = p
    out = bytearray(_be48(x1) + _be48(x0))
    if FQ2.gt(y, FQ2.neg(y)):
        out[0] |= 0x20
    out[0] |= 0x80
    return bytes(out)


def g1_from_uncompressed(b: bytes):
    assert len(b) == 96
    if b[0] & 0x80:
        raise ValueError("UnexpectedCompressionMode")
    if b[0] & 0x40:
        if any(b[1:]) or (b[0] & 0x3F):
            raise ValueError("UnexpectedInformation")
        return INF
    if b[0] & 0x20:
        raise ValueError("UnexpectedInformation")
    x = int.from_bytes(b[:48], "big")
    y = int.from_bytes(b[48:], "big")
    if x >= Q or y >= Q:
        raise ValueError("CoordinateDecodingError")
    return (x, y)


def g2_from_uncompressed(b: bytes):
    assert len(b) == 192
    if b[0] & 0x80:
        raise ValueError("UnexpectedCompressionMode")
    if b[0] & 0x40:
        if any(b[1:]) or (b[0] & 0x3F):
            raise ValueError("UnexpectedInformation")
        return INF
    if b[0] & 0x20:
        raise ValueError("UnexpectedInformation")
    v = [int.from_bytes(b[48 * i:48 * i + 48], "big") for i in range(4)]
    if any(c >= Q for c in v):
        raise ValueError("CoordinateDecodingError")
    return ((v[1], v[0]), (v[3], v[2]))


def g1_from_compressed(b: bytes):
    assert len(b) == 48 and b[0] & 0x80
    if b[0] & 0x40:
        return INF
    greatest = bool(b[0] & 0x20)
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = FQ.sqrt((x * x * x + 4) % Q)
    if y is None:
        raise ValueError("NotOnCurve")
    ny = FQ.neg(y)
    return (x, y if (y < ny) ^ greatest else ny)   # ec.rs:102-123


def g2_from_compressed(b: bytes):
    assert len(b) == 96 and b[0] & 0x80
    if b[0] & 0x40:
        return INF
    greatest = bool(b[0] & 0x20)
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    x = (x0, x1)
    y = FQ2.sqrt(FQ2.add(FQ2.mul(FQ2.mul(x, x), x), FQ2.b))
    if y is None:
        raise ValueError("NotOnCurve")
    ny = FQ2.neg(y)
    return (x, y if FQ2.gt(ny, y) ^ greatest else ny)


def proof_bytes(a, b, c) -> bytes:
    """core/bellman-verifier/src/lib.rs:55-65: compressed a (G1) || b (G2) || c (G1)."""
    return g1_compressed(a) + g2_compressed(b) + g1_compressed(c)


# ---------------------------------------------------------------------------------------
# Parameters file (upstream bellman groth16::Parameters::{read,write})
def params_write(vk, h, l, a, b_g1, b_g2) -> bytes:
    """vk = dict(alpha_g1,beta_g1,beta_g2,gamma_g2,delta_g1,delta_g2,ic=[...])."""
    out = bytearray()
    out += g1_uncompressed(vk["alpha_g1"]) + g1_uncompressed(vk["beta_g1"])
    out += g2_uncompressed(vk["beta_g2"]) + g2_uncompressed(vk["gamma_g2"])
    out += g1_uncompressed(vk["delta_g1"]) + g2_uncompressed(vk["delta_g2"])
    out += struct.pack(">I", len(vk["ic"]))
    for p in vk["ic"]:
        out += g1_uncompressed(p)
    for vec, enc in ((h, g1_uncompressed), (l, g1_uncompressed), (a, g1_uncompressed),
                     (b_g1, g1_uncompressed), (b_g2, g2_uncompressed)):
        out += struct.pack(">I", len(vec))
        for p in vec:
            out += enc(p)
    return bytes(out)


def params_layout(buf: bytes) -> dict:
    """Offsets/lengths only (no point decoding): returns dict name -> (offset, count)."""
    off = 96 + 96 + 192 + 192 + 96 + 192
    lay = {}
    (n,) = struct.unpack_from(">I", buf, off); off += 4
    lay["ic"] = (off, n); off += 96 * n
    for name, sz in (("h", 96), ("l", 96), ("a", 96), ("b_g1", 96), ("b_g2", 192)):
        (n,) = struct.unpack_from(">I", buf, off); off += 4
        lay[name] = (off, n); off += sz * n
    lay["end"] = (off, 0)
    return lay


# ---------------------------------------------------------------------------------------
# Fr NTT (mathematical definition of upstream bellman EvaluationDomain::fft: out[k] = sum_j a[j] w^(jk))
def omega(log_n: int) -> int:
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_S - log_n), R)


def ntt(a, log_n, w=None):
    n = 1 << log_n
    assert len(a) == n
    w = omega(log_n) if w is None else w
    a = list(a)
    # bit reversal
    for i in range(n):
        j = int(format(i, "0%db" % log_n)[::-1], 2) if log_n else 0
        if i < j:
            a[i], a[j] = a[j], a[i]
    m = 1
    for _ in range(log_n):
        wm = pow(w, n // (2 * m), R)
        for k in range(0, n, 2 * m):
            t = 1
            for j in range(m):
                u = a[k + j]
                v = a[k + j + m] * t % R
                a[k + j] = (u + v) % R
                a[k + j + m] = (u - v) % R
                t = t * wm % R
        m *= 2
    return a


def intt(a, log_n):
    n = 1 << log_n
    ninv = pow(n, -1, R)
    return [x * ninv % R for x in ntt(a, log_n, pow(omega(log_n), -1, R))]


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def h_coeffs(a_ev, b_ev, c_ev, log_n):
    """Quotient polynomial exactly as upstream bellman create_proof:
    ifft; coset_fft; a*b-c; divide_by_z_on_coset; icoset_fft; drop last coeff."""
    n = 1 << log_n
    pad = lambda v: list(v) + [0] * (n - len(v))
    g = FR_GENERATOR
    def coset_fft(ev):
        co = intt(pad(ev), log_n)
        co = [c * pow(g, i, R) % R for i, c in enumerate(co)]
        return ntt(co, log_n)
    A, B, C = coset_fft(a_ev), coset_fft(b_ev), coset_fft(c_ev)
    zinv = pow((pow(g, n, R) - 1) % R, -1, R)
    Hc = [((x * y - z) % R) * zinv % R for x, y, z in zip(A, B, C)]
    co = intt(Hc, log_n)
    ginv = pow(g, -1, R)
    co = [c * pow(ginv, i, R) % R for i, c in enumerate(co)]
    return co[: n - 1]


# ---------------------------------------------------------------------------------------
# deterministic RNG shared by python tests (SplitMix64)
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def below(self, m: int, words: int) -> int:
        x = 0
        for i in range(words):
            x |= self.next() << (64 * i)
        return x % m

    def fr(self) -> int:
        return self.below(R, 5)

    def fq(self) -> int:
        return self.below(Q, 7)


# ---------------------------------------------------------------------------------------
# Pairing (verifier-side oracle).  Test infrastructure only, written for clarity:
# Fq12 is represented as Fq[w]/(w^12 - 2 w^6 + 2) — the same field as the reference's tower
# Fq2[v]/(v^3 - (u+1)), Fq6[w]/(w^2 - v) (fq6.rs, fq12.rs) because u = w^6 - 1 satisfies u^2 = -1.
# The Miller loop uses the textbook affine formulas on E(Fq12) after untwisting Q; the final
# exponentiation is the plain power (q^12 - 1)/r.  Pinned by the reference's fixture: conf_vk.dat[0:576]
# = Fq12::write(e(alpha_g1, beta_g2)) (core/bellman-verifier/src/lib.rs:174-196, verifier.rs:15-30).
BLS_X = 0xd201000000010000          # |x|; the BLS parameter is -x (mod.rs:23-25)


def _f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for i in range(22, 11, -1):       # w^12 = 2 w^6 - 2
        c = t[i]
        if c:
            t[i - 6] += 2 * c
            t[i - 12] -= 2 * c
    return [x % Q for x in t[:12]]


def _f12_add(a, b): return [(x + y) % Q for x, y in zip(a, b)]
def _f12_sub(a, b): return [(x - y) % Q for x, y in zip(a, b)]
def _f12_scalar(a, k): return [x * k % Q for x in a]
F12_ONE = [1] + [0] * 11
F12_ZERO = [0] * 12


def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _f12_inv(a):
    """Extended Euclid on polynomials over Fq modulo w^12 - 2 w^6 + 2."""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [2, 0, 0, 0, 0, 0, (-2) % Q, 0, 0, 0, 0, 0, 1]
    while _poly_deg(low):
        dl, dh = _poly_deg(low), _poly_deg(high)
        r = [0] * 13
        tmp = list(high)
        inv_lead = pow(low[dl], -1, Q)
        for i in range(dh - dl, -1, -1):
            r[i] = tmp[dl + i] * inv_lead % Q
            for c in range(dl + 1):
                tmp[c + i] = (tmp[c + i] - r[i] * low[c]) % Q
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q
                new[i + j] = (new[i + j] - low[i] * r[j]) % Q
        lm, low, hm, high = nm, new, lm, low
    c = pow(low[0], -1, Q)
    return [x * c % Q for x in lm[:12]]


def _f12_pow(a, e):
    r, b = F12_ONE, a
    while e:
        if e & 1:
            r = _f12_mul(r, b)
        b = _f12_mul(b, b)
        e >>= 1
    return r


def _fq2_to_f12(c):              # c0 + c1 u, u = w^6 - 1
    r = [0] * 12
    r[0] = (c[0] - c[1]) % Q
    r[6] = c[1] % Q
    return r


def _untwist(q):
    """E'(Fq2) -> E(Fq12): (x, y) -> (x / w^2, y / w^3)  (w^6 = 1 + u)."""
    x, y = _fq2_to_f12(q[0]), _fq2_to_f12(q[1])
    w2 = [0, 0, 1] + [0] * 9
    w3 = [0, 0, 0, 1] + [0] * 8
    return (_f12_mul(x, _f12_inv(w2)), _f12_mul(y, _f12_inv(w3)))


def miller_loop(p, q):
    """f_{|x|,Q}(P) conjugated for the negative BLS parameter; p in G1 (affine ints), q in G2 (affine Fq2 pairs)."""
    if p is INF or q is INF:
        return F12_ONE
    xp = [p[0]] + [0] * 11
    yp = [p[1]] + [0] * 11
    Qx, Qy = _untwist(q)
    Tx, Ty = Qx, Qy
    f = F12_ONE

    def line(lam, x1, y1):       # l(P) = (y_P - y_1) - lam (x_P - x_1)
        return _f12_sub(_f12_sub(yp, y1), _f12_mul(lam, _f12_sub(xp, x1)))
    for bit in bin(BLS_X)[3:]:
        lam = _f12_mul(_f12_scalar(_f12_mul(Tx, Tx), 3), _f12_inv(_f12_scalar(Ty, 2)))
        f = _f12_mul(_f12_mul(f, f), line(lam, Tx, Ty))
        nx = _f12_sub(_f12_mul(lam, lam), _f12_scalar(Tx, 2))
        Ty = _f12_sub(_f12_mul(lam, _f12_sub(Tx, nx)), Ty)
        Tx = nx
        if bit == "1":
            lam = _f12_mul(_f12_sub(Qy, Ty), _f12_inv(_f12_sub(Qx, Tx)))
            f = _f12_mul(f, line(lam, Tx, Ty))
            nx = _f12_sub(_f12_sub(_f12_mul(lam, lam), Tx), Qx)
            Ty = _f12_sub(_f12_mul(lam, _f12_sub(Tx, nx)), Ty)
            Tx = nx
    # x < 0: f_{x,Q} = 1 / f_{|x|,Q} up to factors killed by the final exponentiation; conjugation (w -> -w... the
    # q^6-Frobenius) equals inversion on the cyclotomic subgroup, so invert here and let the exponentiation finish it.
    return _f12_inv(f)


def final_exponentiation(f):
    return _f12_pow(f, (Q ** 12 - 1) // R)


def pairing(p, q):
    """Reduced pairing with the plain exponent (q^12 - 1)/r."""
    return final_exponentiation(miller_loop(p, q))


def pairing_reference(p, q):
    """The value the reference's Engine::pairing returns: its final_exponentiation (mod.rs:104-160) uses the
    x-addition chain of eprint 2016/130, which yields the CUBE of the plain reduced pairing.  Established against the
    fixture conf_vk.dat[0:576] (tests/test_oracle_pairing.py); irrelevant for verification, where only == 1 matters."""
    return _f12_pow(pairing(p, q), 3)


def f12_to_tower_bytes(a) -> bytes:
    """Fq12::write order (fq12.rs:29-45, fq6.rs:31-48, fq2.rs:40-44): for w^i (i=0,1), v^j (j=0..2), u^k (k=0,1): 48-byte BE."""
    out = bytearray()
    for i in range(2):
        for j in range(3):
            t = 2 * j + i
            c1 = a[t + 6] % Q
            c0 = (a[t] + a[t + 6]) % Q
            out += c0.to_bytes(48, "big") + c1.to_bytes(48, "big")
    return bytes(out)


def groth16_verify(vk, proof_abc, public_inputs) -> bool:
    """verify_proof (core/bellman-verifier/src/verifier.rs:32-63): e(A,B) = e(alpha,beta) e(sum x_i ic_i, gamma) e(C,delta).
    vk: dict(alpha_g1, beta_g2, gamma_g2, delta_g2, ic=[...]) affine; public_inputs exclude the leading ONE."""
    a, b, c = proof_abc
    acc = vk["ic"][0]
    for x, pt in zip(public_inputs, vk["ic"][1:]):
        acc = ec_add(FQ, acc, ec_mul(FQ, pt, x))
    m = miller_loop(a, b)
    m = _f12_mul(m, miller_loop(ec_neg(FQ, vk["alpha_g1"]), vk["beta_g2"]))
    m = _f12_mul(m, miller_loop(ec_neg(FQ, acc), vk["gamma_g2"]))
    m = _f12_mul(m, miller_loop(ec_neg(FQ, c), vk["delta_g2"]))
    return final_exponentiation(m) == F12_ONE


# ---------------------------------------------------------------------------------------
# PreparedVerifyingKey (core/bellman-verifier/src/lib.rs:110-245) and the strict Proof::read (lib.rs:67-108).
# The prepared form of a G2 point is the list of line coefficients its Miller loop consumes
# (G2Prepared::from_affine, core/pairing/src/bls12_381/mod.rs:163-359).  Restated from the curve equations: with
# T = (X, Y, Z) Jacobian on E'(Fq2),
#   doubling  (EFD dbl-2009-l):  X3 = 9X^4 - 8XY^2, Z3 = 2YZ, Y3 = 3X^2 (4XY^2 - X3) - 8Y^4
#             line = (2 Z3 Z^2, -6 X^2 Z^2, 6 X^3 - 4 Y^2)
#   addition of the affine base point (x2, y2) (EFD madd-2007-bl):  H = x2 Z^2 - X, r = 2 (y2 Z^3 - Y), Z3 = 2ZH,
#             X3 = r^2 - 4H^3 - 8XH^2, Y3 = r (4XH^2 - X3) - 8YH^3;  line = (2 Z3, -2 r, 2 (r x2 - y2 Z3))
# Pinned bit-for-bit by the shipped conf_vk.dat / anony_vk.dat (tests/test_oracle_pairing.py): the file's two coefficient
# tables are prepare(-gamma_g2) and prepare(-delta_g2) of the VerifyingKey inside conf_pk.dat / anony_pk.dat.
def g2_prepare(q):
    F = FQ2
    if q is INF:
        return []
    def k(a, n): return ((a[0] * n) % Q, (a[1] * n) % Q)
    sq = lambda a: F.mul(a, a)
    X, Y, Z = q[0], q[1], F.one
    x2, y2 = q
    out = []

    def dbl():
        nonlocal X, Y, Z
        xx, yy, zz = sq(X), sq(Y), sq(Z)
        xyy4 = k(F.mul(X, yy), 4)
        e = k(xx, 3)
        X3 = F.sub(sq(e), k(xyy4, 2))
        Z3 = k(F.mul(Y, Z), 2)
        Y3 = F.sub(F.mul(e, F.sub(xyy4, X3)), k(sq(yy), 8))
        line = (k(F.mul(Z3, zz), 2), F.neg(k(F.mul(e, zz), 2)), F.sub(k(F.mul(e, X), 2), k(yy, 4)))
        X, Y, Z = X3, Y3, Z3
        out.append(line)

    def add():
        nonlocal X, Y, Z
        zz = sq(Z)
        h = F.sub(F.mul(x2, zz), X)
        r = k(F.sub(F.mul(y2, F.mul(Z, zz)), Y), 2)
        hh = sq(h)
        Z3 = k(F.mul(Z, h), 2)
        v4 = k(F.mul(X, hh), 4)
        h34 = k(F.mul(h, hh), 4)
        X3 = F.sub(F.sub(sq(r), h34), k(v4, 2))
        Y3 = F.sub(F.mul(r, F.sub(v4, X3)), k(F.mul(Y, h34), 2))
        line = (k(Z3, 2), F.neg(k(r, 2)), k(F.sub(F.mul(r, x2), F.mul(y2, Z3)), 2))
        X, Y, Z = X3, Y3, Z3
        out.append(line)

    bits = bin(BLS_X)[3:]             # below the leading one; the last bit's doubling closes the list (mod.rs:338-352)
    for b in bits[:-1]:
        dbl()
        if b == "1":
            add()
    dbl()
    assert bits[-1] == "0"
    return out


def _fq2_bytes(a) -> bytes:            # Fq2::write: c0 then c1 (fq2.rs:40-44)
    return a[0].to_bytes(48, "big") + a[1].to_bytes(48, "big")


def g2_prepared_write(coeffs, infinity=False) -> bytes:
    """G2Prepared::write (core/pairing/src/bls12_381/ec.rs:1631-1650)."""
    out = bytearray(len(coeffs).to_bytes(4, "big"))
    for c in coeffs:
        out += _fq2_bytes(c[0]) + _fq2_bytes(c[1]) + _fq2_bytes(c[2])
    out += b"\x01" if infinity else b"\x00"
    return bytes(out)


def pvk_write(vk) -> bytes:
    """prepare_verifying_key (verifier.rs:15-30) followed by PreparedVerifyingKey::write (lib.rs:183-202)."""
    out = bytearray(f12_to_tower_bytes(pairing_reference(vk["alpha_g1"], vk["beta_g2"])))
    for g in (vk["gamma_g2"], vk["delta_g2"]):
        out += g2_prepared_write(g2_prepare(ec_neg(FQ2, g)))
    out += len(vk["ic"]).to_bytes(4, "big")
    for p in vk["ic"]:
        out += g1_uncompressed(p)
    return bytes(out)


def vk_read(buf: bytes) -> dict:
    """VerifyingKey::read — the head of Parameters::write (alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2, ic)."""
    o = 0
    def g1():
        nonlocal o
        p = g1_from_uncompressed(buf[o:o + 96]); o += 96; return p
    def g2():
        nonlocal o
        p = g2_from_uncompressed(buf[o:o + 192]); o += 192; return p
    vk = dict(alpha_g1=g1(), beta_g1=g1(), beta_g2=g2(), gamma_g2=g2(), delta_g1=g1(), delta_g2=g2())
    n = int.from_bytes(buf[o:o + 4], "big"); o += 4
    vk["ic"] = [g1() for _ in range(n)]
    vk["_size"] = o
    return vk


def vk_write(vk) -> bytes:
    out = g1_uncompressed(vk["alpha_g1"]) + g1_uncompressed(vk["beta_g1"]) + g2_uncompressed(vk["beta_g2"])
    out += g2_uncompressed(vk["gamma_g2"]) + g1_uncompressed(vk["delta_g1"]) + g2_uncompressed(vk["delta_g2"])
    out += len(vk["ic"]).to_bytes(4, "big")
    return out + b"".join(g1_uncompressed(p) for p in vk["ic"])


def _compressed_strict(b: bytes, g2: bool):
    """Compressed::into_affine (ec.rs:796-838 + subgroup check ec.rs:775-794): raises ValueError like GroupDecodingError."""
    if not b[0] & 0x80:
        raise ValueError("UnexpectedCompressionMode")
    if b[0] & 0x40:
        if (b[0] & 0x3F) or any(b[1:]):
            raise ValueError("UnexpectedInformation")
        return INF
    F = FQ2 if g2 else FQ
    if g2:
        x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
        x0 = int.from_bytes(b[48:], "big")
        if x0 >= Q or x1 >= Q:
            raise ValueError("CoordinateDecodingError")
    else:
        if int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big") >= Q:
            raise ValueError("CoordinateDecodingError")
    p = g2_from_compressed(b) if g2 else g1_from_compressed(b)
    if ec_mul(F, p, R) is not INF:
        raise ValueError("NotInSubgroup")
    return p


def proof_read(b: bytes):
    """Proof::read (lib.rs:67-108): returns (a, b, c); raises ValueError('InvalidData') / ValueError('PointInfinity')."""
    assert len(b) == 192
    pts = []
    for chunk, g2 in ((b[:48], False), (b[48:144], True), (b[144:], False)):
        try:
            p = _compressed_strict(chunk, g2)
        except ValueError:
            raise ValueError("InvalidData")
        if p is INF:
            raise ValueError("PointInfinity")
        pts.append(p)
    return tuple(pts)


def miller_loop_prepared(p, coeffs):
    """Engine::miller_loop for one (G1Affine, G2Prepared) pair (mod.rs:40-102): ell() places coeffs.2 at 1, coeffs.1 * x_P at
    v and coeffs.0 * y_P at v w (mul_by_014), i.e. at w^0, w^2 and w^3 of the polynomial basis used here."""
    if p is INF or not coeffs:
        return F12_ONE
    def ell(c):
        r = [0] * 12
        for pos, val in ((0, c[2]), (2, (c[1][0] * p[0] % Q, c[1][1] * p[0] % Q)), (3, (c[0][0] * p[1] % Q, c[0][1] * p[1] % Q))):
            r[pos] = (r[pos] + val[0] - val[1]) % Q
            r[pos + 6] = (r[pos + 6] + val[1]) % Q
        return r
    it = iter(coeffs)
    f = F12_ONE
    bits = bin(BLS_X)[3:]
    for b in bits[:-1]:
        f = _f12_mul(f, ell(next(it)))
        if b == "1":
            f = _f12_mul(f, ell(next(it)))
        f = _f12_mul(f, f)
    f = _f12_mul(f, ell(next(it)))
    # conjugation = the q^6 Frobenius: w -> -w
    return [x if i % 2 == 0 else (-x) % Q for i, x in enumerate(f)]


def verify_prepared(pvk_alpha_beta, neg_gamma_coeffs, neg_delta_coeffs, ic, proof_abc, public_inputs):
    """verify_proof (verifier.rs:32-63) on prepared data; returns bool, raises ValueError('MalformedVerifyingKey')."""
    if len(public_inputs) + 1 != len(ic):
        raise ValueError("MalformedVerifyingKey")
    a, b, c = proof_abc
    acc = ic[0]
    for x, pt in zip(public_inputs, ic[1:]):
        acc = ec_add(FQ, acc, ec_mul(FQ, pt, x))
    m = miller_loop_prepared(a, g2_prepare(b))
    m = _f12_mul(m, miller_loop_prepared(acc, neg_gamma_coeffs))
    m = _f12_mul(m, miller_loop_prepared(c, neg_delta_coeffs))
    return _f12_pow(final_exponentiation(m), 3) == pvk_alpha_beta
