"""TEST INFRASTRUCTURE — ctypes binding of oracle/libzkoracle.so (the C restatement of the
reference's CPU path, oracle/zk_oracle.c).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

Flat layouts (numpy uint64, little-endian limbs):
  Fq 6 limbs Montgomery; Fr 4 limbs (Montgomery or canonical as each function states);
  G1 affine limb form (n,12) = x|y, infinity = zeros; G2 affine limb form (n,24) = x.c0|x.c1|y.c0|y.c1.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzkoracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("zk_oracle.c", "field_tmpl.inc", "curve_tmpl.inc", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.zko_fq_mul_bench.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def ints_to_limbs(vals, n_limbs):
    out = np.zeros((len(vals), n_limbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(n_limbs):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64)
    arr = arr.reshape(-1, arr.shape[-1])
    return [sum(int(v) << (64 * j) for j, v in enumerate(row)) for row in arr]


# ---- field ops on single elements given as python ints holding RAW limbs (Montgomery residues) ----
def _bin(name, nl):
    def f(a: int, b: int) -> int:
        A, B, O = ints_to_limbs([a], nl), ints_to_limbs([b], nl), np.zeros((1, nl), np.uint64)
        getattr(lib(), name)(_p(A), _p(B), _p(O))
        return limbs_to_ints(O)[0]
    return f


def _un(name, nl, ret=False):
    def f(a: int):
        A, O = ints_to_limbs([a], nl), np.zeros((1, nl), np.uint64)
        r = getattr(lib(), name)(_p(A), _p(O))
        if ret and r != 0:
            return None
        return limbs_to_ints(O)[0]
    return f


fq_mul, fq_add, fq_sub = _bin("zko_fq_mul", 6), _bin("zko_fq_add", 6), _bin("zko_fq_sub", 6)
fq_sqr, fq_neg, fq_into_repr = _un("zko_fq_sqr", 6), _un("zko_fq_neg", 6), _un("zko_fq_into_repr", 6)
fq_inv, fq_from_repr = _un("zko_fq_inv", 6, True), _un("zko_fq_from_repr", 6, True)
fr_mul, fr_add, fr_sub = _bin("zko_fr_mul", 4), _bin("zko_fr_add", 4), _bin("zko_fr_sub", 4)
fr_sqr, fr_neg, fr_into_repr = _un("zko_fr_sqr", 4), _un("zko_fr_neg", 4), _un("zko_fr_into_repr", 4)
fr_inv, fr_from_repr = _un("zko_fr_inv", 4, True), _un("zko_fr_from_repr", 4, True)
fq2_mul = _bin("zko_fq2_mul", 12)
fq2_sqr = _un("zko_fq2_sqr", 12)
fq2_inv = _un("zko_fq2_inv", 12, True)


# ---- points (numpy limb form) ----------------------------------------------------------------
def g1_generator():
    o = np.zeros(12, np.uint64); lib().zko_g1_generator(_p(o)); return o


def g2_generator():
    o = np.zeros(24, np.uint64); lib().zko_g2_generator(_p(o)); return o


def _pt_bin(name, w):
    def f(a, b):
        a, b, o = _u64(a), _u64(b), np.zeros(w, np.uint64)
        getattr(lib(), name)(_p(a), _p(b), _p(o)); return o
    return f


def _pt_un(name, w):
    def f(a):
        a, o = _u64(a), np.zeros(w, np.uint64)
        getattr(lib(), name)(_p(a), _p(o)); return o
    return f


g1_add, g1_add_mixed, g1_double = _pt_bin("zko_g1_add", 12), _pt_bin("zko_g1_add_mixed", 12), _pt_un("zko_g1_double", 12)
g2_add, g2_add_mixed, g2_double = _pt_bin("zko_g2_add", 24), _pt_bin("zko_g2_add_mixed", 24), _pt_un("zko_g2_double", 24)


def g1_mul(a, k: int):
    a, kk, o = _u64(a), ints_to_limbs([k], 4), np.zeros(12, np.uint64)
    lib().zko_g1_mul(_p(a), _p(kk), _p(o)); return o


def g2_mul(a, k: int):
    a, kk, o = _u64(a), ints_to_limbs([k], 4), np.zeros(24, np.uint64)
    lib().zko_g2_mul(_p(a), _p(kk), _p(o)); return o


def g1_check(a) -> int:
    return lib().zko_g1_check(_p(_u64(a)))


def g2_check(a) -> int:
    return lib().zko_g2_check(_p(_u64(a)))


def g1_encode(a, compressed: bool) -> bytes:
    o = np.zeros(48 if compressed else 96, np.uint8)
    lib().zko_g1_encode(_p(_u64(a)), int(compressed), _p(o)); return o.tobytes()


def g2_encode(a, compressed: bool) -> bytes:
    o = np.zeros(96 if compressed else 192, np.uint8)
    lib().zko_g2_encode(_p(_u64(a)), int(compressed), _p(o)); return o.tobytes()


def g1_decode_many(buf: bytes, checked=False):
    n = len(buf) // 96
    b = np.frombuffer(buf, np.uint8).copy(); o = np.zeros((n, 12), np.uint64)
    e = lib().zko_g1_decode_many(_p(b), C.c_size_t(n), int(checked), _p(o))
    if e:
        raise ValueError("GroupDecodingError %d" % e)
    return o


def g2_decode_many(buf: bytes, checked=False):
    n = len(buf) // 192
    b = np.frombuffer(buf, np.uint8).copy(); o = np.zeros((n, 24), np.uint64)
    e = lib().zko_g2_decode_many(_p(b), C.c_size_t(n), int(checked), _p(o))
    if e:
        raise ValueError("GroupDecodingError %d" % e)
    return o


def g1_fixed_base(scalars, base=None, enc=False):
    """scalars: (n,4) canonical -> (n,12) limb-form points s_i*base, or bytes (n*96) if enc."""
    s = _u64(scalars, (-1, 4)); n = s.shape[0]
    base = g1_generator() if base is None else _u64(base)
    o = np.zeros(n * 96, np.uint8)
    lib().zko_g1_fixed_base_many(_p(base), _p(s), C.c_size_t(n), int(enc), _p(o))
    return o.tobytes() if enc else o.view(np.uint64).reshape(n, 12)


def g2_fixed_base(scalars, base=None, enc=False):
    s = _u64(scalars, (-1, 4)); n = s.shape[0]
    base = g2_generator() if base is None else _u64(base)
    o = np.zeros(n * 192, np.uint8)
    lib().zko_g2_fixed_base_many(_p(base), _p(s), C.c_size_t(n), int(enc), _p(o))
    return o.tobytes() if enc else o.view(np.uint64).reshape(n, 24)


def g1_msm(bases, scalars, density=None):
    b, s = _u64(bases, (-1, 12)), _u64(scalars, (-1, 4))
    d = None if density is None else np.ascontiguousarray(density, np.uint8)
    o = np.zeros(12, np.uint64)
    e = lib().zko_g1_msm(_p(b), _p(s), C.c_size_t(s.shape[0]), _p(d) if d is not None else None, _p(o))
    if e:
        raise ValueError("multiexp error %d" % e)
    return o


def g2_msm(bases, scalars, density=None):
    b, s = _u64(bases, (-1, 24)), _u64(scalars, (-1, 4))
    d = None if density is None else np.ascontiguousarray(density, np.uint8)
    o = np.zeros(24, np.uint64)
    e = lib().zko_g2_msm(_p(b), _p(s), C.c_size_t(s.shape[0]), _p(d) if d is not None else None, _p(o))
    if e:
        raise ValueError("multiexp error %d" % e)
    return o


# ---- Fr vectors --------------------------------------------------------------------------------
def fr_to_mont(a):
    a = _u64(a, (-1, 4)); o = np.empty_like(a)
    lib().zko_fr_from_repr_many(_p(a), _p(o), C.c_size_t(a.shape[0])); return o


def fr_from_mont(a):
    a = _u64(a, (-1, 4)); o = np.empty_like(a)
    lib().zko_fr_into_repr_many(_p(a), _p(o), C.c_size_t(a.shape[0])); return o


def fr_mul_many(a, b):
    a, b = _u64(a, (-1, 4)), _u64(b, (-1, 4)); o = np.empty_like(a)
    lib().zko_fr_mul_many(_p(a), _p(b), _p(o), C.c_size_t(a.shape[0])); return o


NTT_FFT, NTT_IFFT, NTT_COSET_FFT, NTT_ICOSET_FFT = 0, 1, 2, 3


def fr_ntt(data_mont, log_n: int, mode: int):
    """Montgomery-form Fr (n,4), natural order in/out; returns a new array."""
    a = _u64(data_mont, (-1, 4)).copy()
    assert a.shape[0] == 1 << log_n
    e = lib().zko_fr_ntt(_p(a), C.c_uint(log_n), int(mode))
    if e:
        raise ValueError("ntt error %d" % e)
    return a


def h_coeffs(a, b, c):
    a, b, c = _u64(a, (-1, 4)), _u64(b, (-1, 4)), _u64(c, (-1, 4))
    n = a.shape[0]; m = 1
    while m < n:
        m *= 2
    o = np.zeros((m - 1, 4), np.uint64)
    e = lib().zko_h_coeffs(_p(a), _p(b), _p(c), C.c_size_t(n), _p(o))
    if e:
        raise ValueError("h error %d" % e)
    return o


# ---- Parameters + prover -------------------------------------------------------------------------
class Params:
    def __init__(self, buf: bytes, checked: bool = False):
        self._h = C.c_void_p()
        b = np.frombuffer(buf, np.uint8)
        e = lib().zko_params_read(_p(b), C.c_size_t(len(buf)), int(checked), C.byref(self._h))
        if e:
            raise ValueError("Parameters::read error %d" % e)
        cnt = np.zeros(5, np.uint64)
        lib().zko_params_counts(self._h, _p(cnt))
        self.n_ic, self.n_h, self.n_l, self.n_a, self.n_b = (int(x) for x in cnt)

    def export(self, which: str):
        idx = {"ic": 0, "h": 1, "l": 2, "a": 3, "b_g1": 4, "b_g2": 5}[which]
        n = [self.n_ic, self.n_h, self.n_l, self.n_a, self.n_b, self.n_b][idx]
        o = np.zeros((n, 24 if idx == 5 else 12), np.uint64)
        lib().zko_params_export(self._h, idx, _p(o)); return o

    def prove(self, a, b, c, inputs, aux, a_aux_density, b_input_density, b_aux_density, r: int, s: int) -> bytes:
        a, b, c = _u64(a, (-1, 4)), _u64(b, (-1, 4)), _u64(c, (-1, 4))
        inputs, aux = _u64(inputs, (-1, 4)), _u64(aux, (-1, 4))
        d1, d2, d3 = (np.ascontiguousarray(x, np.uint8) for x in (a_aux_density, b_input_density, b_aux_density))
        rr, ss = ints_to_limbs([r], 4), ints_to_limbs([s], 4)
        out = np.zeros(192, np.uint8)
        e = lib().zko_groth16_prove(self._h, _p(a), _p(b), _p(c), C.c_size_t(a.shape[0]),
                                    _p(inputs), C.c_size_t(inputs.shape[0]), _p(aux), C.c_size_t(aux.shape[0]),
                                    _p(d1), _p(d2), _p(d3), _p(rr), _p(ss), _p(out))
        if e:
            raise ValueError("create_proof error %d" % e)
        return out.tobytes()

    def __del__(self):
        try:
            if self._h:
                lib().zko_params_free(self._h); self._h = None
        except Exception:
            pass


def num_threads() -> int:
    return lib().zko_num_threads()


def set_num_threads(n: int):
    lib().zko_set_num_threads(int(n))


# ---- verifier side (oracle/pairing_oracle.inc) ----------------------------------------------------------------
class PreparedVerifyingKey:
    """PreparedVerifyingKey<Bls12> in the C oracle: `prepare` = prepare_verifying_key(vk) (verifier.rs:15-30) from the
    VerifyingKey encoding, `read` = PreparedVerifyingKey::read (lib.rs:204-245)."""

    def __init__(self, h):
        self._h = h

    @staticmethod
    def prepare(vk_bytes: bytes) -> "PreparedVerifyingKey":
        h = C.c_void_p()
        b = np.frombuffer(vk_bytes, np.uint8)
        r = lib().zko_pvk_prepare(_p(b), C.c_size_t(len(vk_bytes)), C.byref(h))
        if r:
            raise ValueError("zko_pvk_prepare: %d" % r)
        return PreparedVerifyingKey(h)

    @staticmethod
    def read(buf: bytes) -> "PreparedVerifyingKey":
        h = C.c_void_p()
        b = np.frombuffer(buf, np.uint8)
        r = lib().zko_pvk_read(_p(b), C.c_size_t(len(buf)), C.byref(h))
        if r:
            raise ValueError("zko_pvk_read: %d" % r)
        return PreparedVerifyingKey(h)

    def write(self) -> bytes:
        lib().zko_pvk_size.restype = C.c_size_t
        o = np.zeros(lib().zko_pvk_size(self._h), np.uint8)
        lib().zko_pvk_write(self._h, _p(o))
        return o.tobytes()

    def verify_batch(self, proofs: bytes, inputs_limbs, n_inputs: int):
        """verdicts per proof: 1 Ok(true), 0 Ok(false), 2 InvalidData, 3 PointInfinity; ValueError on MalformedVerifyingKey."""
        n = len(proofs) // 192
        pb = np.frombuffer(proofs, np.uint8) if n else np.zeros(1, np.uint8)
        inp = _u64(inputs_limbs).reshape(-1) if n * n_inputs else np.zeros(4, np.uint64)
        out = np.zeros(max(n, 1), np.uint8)
        r = lib().zko_verify_batch(self._h, C.c_size_t(n), _p(pb), _p(inp), C.c_size_t(n_inputs), _p(out))
        if r:
            raise ValueError("MalformedVerifyingKey" if r == -9 else "zko_verify_batch: %d" % r)
        return [int(v) for v in out[:n]]

    def __del__(self):
        try:
            lib().zko_pvk_free(self._h)
        except Exception:
            pass


def pairing(g1_uncompressed: bytes, g2_uncompressed: bytes) -> bytes:
    o = np.zeros(576, np.uint8)
    r = lib().zko_pairing(_p(np.frombuffer(g1_uncompressed, np.uint8)), _p(np.frombuffer(g2_uncompressed, np.uint8)), _p(o))
    if r:
        raise ValueError("zko_pairing: %d" % r)
    return o.tobytes()
