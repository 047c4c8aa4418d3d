"""CPU checks of the drop-in boundary: the C-ABI shared library loads, exports every symbol that
include/zkb200.h declares, the ctypes binding covers exactly those symbols, and — with no GPU —
compute entry points fail loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "zkb200.h")


def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", src)))


def test_header_library_binding_agree():
    from zero_chain_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    assert sorted(_lib.SIGNATURES) == names                     # binding covers the header exactly
    L = _lib.lib()                                              # loads and binds (AttributeError on a missing symbol)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH]).decode()
    exported = set(re.findall(r" T (zk_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert b"sm_100a" in L.zk_version()


def test_library_is_sm100a_and_has_no_oracle_dependency():
    from zero_chain_b200 import _lib
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.SO_PATH]).decode()
    assert "sm_100a" in out and "sm_90" not in out
    needed = subprocess.check_output(["readelf", "-d", _lib.SO_PATH]).decode()
    assert "zkoracle" not in needed                             # product never links the oracle


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zero_chain_b200 import groth16 as zk
    with pytest.raises(zk.ZkError) as e:
        zk.Context(0)
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)


def test_hot_kernel_sass_uses_fused_wide_multiplies():
    """The Montgomery product must compile to IMAD / IMAD.WIDE.U32.X carry chains (no local-memory spills)."""
    from zero_chain_b200 import _lib
    listing = subprocess.check_output(["cuobjdump", "-lelf", _lib.SO_PATH]).decode()
    assert "sm_100a" in listing
    names = subprocess.check_output("cuobjdump -sass %s | grep 'Function :'" % _lib.SO_PATH, shell=True).decode()
    # the two kernels that carry the G1 bucket additions: the XYZZ pass and the batched-affine backward pass (first round)
    for key in ("k_accumulateI2FpI8FqParamsELi3", "k_ba_backwardI2FpI8FqParamsELb1ELi4"):
        fn = [l.split(":")[1].strip() for l in names.splitlines() if key in l]
        assert len(fn) == 1, (key, names)
        sass = subprocess.check_output(["cuobjdump", "-sass", "-fun", fn[0], _lib.SO_PATH], stderr=subprocess.STDOUT).decode()
        assert sass.count("IMAD.WIDE.U32") > 500 and "STL" not in sass and "LDL" not in sass, key


def test_lane_verifier_kernels_keep_fq12_out_of_local_memory():
    """The lane-parallel Miller loop must not spill its Fq12 state (round 1's thread-per-proof kernels carried 1 301 / 4 285 LDL/STL
    instructions and 18 GB of local-memory write-back per 32 k proofs)."""
    from zero_chain_b200 import _lib
    names = subprocess.check_output("cuobjdump -sass %s | grep 'Function :'" % _lib.SO_PATH, shell=True).decode()
    for key, limit in (("k_miller_lanes", 100), ("k_verify_final_lanes", 1000)):
        fn = [l.split(":")[1].strip() for l in names.splitlines() if key in l]
        assert len(fn) == 1, (key, names)
        sass = subprocess.check_output(["cuobjdump", "-sass", "-fun", fn[0], _lib.SO_PATH], stderr=subprocess.STDOUT).decode()
        body = [l for l in sass.splitlines() if "/*" in l and ("LDL" in l or "STL" in l)]
        n = len(set(l.split("*/")[0] for l in body))          # one line per instruction address (the listing repeats encodings)
        assert n < limit, (key, n)
