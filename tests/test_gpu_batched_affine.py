"""GPU parity tests of the batched-affine bucket rounds (csrc/msm_batchaff.cuh) against the oracle's multiexp restatement.
By default only MSMs with >= 2^22 entries take that path; here the threshold is set to 0 (zk_ctx_set_opt — a tuning option,
results must not depend on it) and the number of rounds is forced, so that small, ragged and degenerate inputs go through every
branch: odd leftovers, empty buckets, P + P (doubling), P + (-P) (infinity as an intermediate result and as an operand of the
next round), heavy buckets, zero scalars, batches, tables and per-call bases, G1 and G2."""
import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as pr
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy

pytestmark = pytest.mark.gpu


@pytest.fixture()
def actx():
    c = zk.Context(0)
    c.set_opt(zk.Context.OPT_AFFINE_MIN_ENTRIES, 0)
    yield c
    c.close()


def _enc(group, p):
    return (co.g1_encode if group == 1 else co.g2_encode)(p, False)


@pytest.mark.parametrize("levels", [1, 2, 3, 6])
@pytest.mark.parametrize("n,c,tables", [(1, 5, True), (37, 4, True), (3000, 7, True), (3000, 7, False), (20000, 10, True), (50000, 16, True)])
def test_g1_rounds_match_oracle(actx, levels, n, c, tables):
    actx.set_opt(zk.Context.OPT_AFFINE_LEVELS, levels)
    bases = co.g1_fixed_base(sy.random_fr_limbs(n, 11 * n + c))
    scal = sy.random_fr_limbs(n, 11 * n + c + 1)
    edge = [0, 1, pr.R - 1, 2, (1 << 255) % pr.R, pr.R - (1 << 128)]
    scal[: min(n, len(edge))] = co.ints_to_limbs(edge[: min(n, len(edge))], 4)
    if n > 500:
        scal[100:300, :] = 0; scal[100:300, 0] = 1                   # a heavy bucket
        scal[300:340, :] = 0                                          # zero scalars
    b = zk.Bases(actx, 1, bases, window_bits=c, precompute=tables)
    assert zk.multiexp(b, scal) == _enc(1, co.g1_msm(bases, scal))
    b.free()


@pytest.mark.parametrize("levels", [1, 3, 5])
def test_g1_degenerate_pairs(actx, levels):
    """Equal points in one bucket (doubling), opposite points (infinity), and both mixed with ordinary additions."""
    actx.set_opt(zk.Context.OPT_AFFINE_LEVELS, levels)
    g = co.g1_fixed_base(co.ints_to_limbs([1, 2, 3, 5], 4))
    # 64 copies of G with the same scalar: every pair of every round is a doubling
    same = np.repeat(g[:1], 64, axis=0)
    b = zk.Bases(actx, 1, same, window_bits=5, precompute=True)
    assert zk.multiexp(b, co.ints_to_limbs([3] * 64, 4)) == pr.g1_uncompressed(pr.ec_mul(pr.FQ, pr.G1_GEN, 192))
    # s and r - s alternate: P + (-P) = O in round 1, O + O afterwards
    assert zk.multiexp(b, co.ints_to_limbs([3, pr.R - 3] * 32, 4)) == pr.g1_uncompressed(pr.INF)
    # O as ONE operand of a later round: three cancelling pairs and one survivor per bucket
    assert zk.multiexp(b, co.ints_to_limbs(([3, pr.R - 3] * 3 + [3, 0]) * 8, 4)) == pr.g1_uncompressed(pr.ec_mul(pr.FQ, pr.G1_GEN, 24))
    b.free()
    # a mix: repeated and distinct bases, random scalars, against the oracle
    n = 4000
    idx = np.random.RandomState(3).randint(0, 4, size=n)
    bases = g[idx]
    scal = sy.random_fr_limbs(n, 99)
    scal[::7] = scal[0]                                                # equal (base, scalar) pairs land in the same buckets
    for c, tables in ((6, True), (6, False), (9, True)):
        b = zk.Bases(actx, 1, bases, window_bits=c, precompute=tables)
        assert zk.multiexp(b, scal) == _enc(1, co.g1_msm(bases, scal))
        b.free()


@pytest.mark.parametrize("levels", [1, 2, 4])
@pytest.mark.parametrize("n,c,tables", [(500, 6, True), (3000, 8, False), (12402, 0, True)])
def test_g2_rounds_match_oracle(actx, levels, n, c, tables):
    actx.set_opt(zk.Context.OPT_AFFINE_LEVELS, levels)
    bases = co.g2_fixed_base(sy.random_fr_limbs(n, 177 + n))
    scal = sy.random_fr_limbs(n, 178 + n)
    scal[:3] = co.ints_to_limbs([0, 1, pr.R - 1], 4)
    scal[10:60, :] = 0; scal[10:60, 0] = 2
    b = zk.Bases(actx, 2, bases, window_bits=c, precompute=tables)
    assert zk.multiexp(b, scal) == _enc(2, co.g2_msm(bases, scal))
    b.free()
    if n == 500:
        same = np.repeat(bases[:1], 32, axis=0)
        b = zk.Bases(actx, 2, same, window_bits=5, precompute=True)
        want = _enc(2, co.g2_mul(bases[0], 32 * 3 % pr.R))
        assert zk.multiexp(b, co.ints_to_limbs([3] * 32, 4)) == want
        assert zk.multiexp(b, co.ints_to_limbs([5, pr.R - 5] * 16, 4)) == pr.g2_uncompressed(pr.INF)
        b.free()


def test_batch_and_default_heuristic(actx):
    """A batch of scalar vectors against one table (what the prover does) with the automatic number of rounds."""
    import torch
    actx.set_opt(zk.Context.OPT_AFFINE_LEVELS, -1)
    n, batch = 6000, 5
    bases = co.g1_fixed_base(sy.random_fr_limbs(n, 9))
    b = zk.Bases(actx, 1, bases, window_bits=8)
    scal = sy.random_fr_limbs(n * batch, 10).reshape(batch, n, 4)
    scal[1, :, :] = 0; scal[1, :, 0] = np.arange(n) & 1                 # a 0/1 vector: two huge buckets
    scal[2, :, :] = 0                                                   # an all-zero vector: result O
    d = torch.from_numpy(scal.view(np.int64)).cuda()
    torch.cuda.synchronize()
    got = zk.multiexp_device(b, d.data_ptr(), n, batch)
    for k in range(batch):
        assert got[96 * k:96 * k + 96] == _enc(1, co.g1_msm(bases, scal[k])), k
    b.free()


def test_result_does_not_depend_on_the_option():
    n = 30000
    bases = co.g1_fixed_base(sy.random_fr_limbs(n, 21))
    scal = sy.random_fr_limbs(n, 22)
    outs = []
    for min_entries, levels in ((1 << 40, -1), (0, -1), (0, 2), (0, 8)):
        c = zk.Context(0)
        c.set_opt(zk.Context.OPT_AFFINE_MIN_ENTRIES, min_entries)
        c.set_opt(zk.Context.OPT_AFFINE_LEVELS, levels)
        b = zk.Bases(c, 1, bases, window_bits=12)
        outs.append(zk.multiexp(b, scal))
        b.free(); c.close()
    assert outs[0] == outs[1] == outs[2] == outs[3] == _enc(1, co.g1_msm(bases, scal))
