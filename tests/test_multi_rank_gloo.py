"""world_size-2 test (gloo, CPU) of the multi-GPU host logic used by bench.py: bases sharded by contiguous
index range, one partial result per rank, all-gather of the fixed-size partials, fold on every rank.
The per-rank MSM is stood in for by the oracle here (no GPU in this container); on the B200 box the same
plumbing runs with zk_msm_partial_device / zk_points_fold over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_range(n, world, rank):
    """rank k of G owns [k n/G, (k+1) n/G) — SURVEY.md §8(e)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    from oracle import coracle as co
    from zero_chain_b200 import synthetic as sy
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bases = co.g1_fixed_base(sy.random_fr_limbs(n, 11))            # every rank derives the same global vectors
    scal = sy.random_fr_limbs(n, 12)
    lo, hi = shard_range(n, world, rank)
    part = co.g1_msm(bases[lo:hi], scal[lo:hi])                    # this rank's shard only
    mine = torch.from_numpy(part.view(np.int64).copy())
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    acc = np.zeros(12, np.uint64)
    for g in gathered:                                             # fold (every rank computes the same sum)
        acc = co.g1_add(acc, g.numpy().view(np.uint64))
    full = co.g1_msm(bases, scal)
    ok = np.array_equal(acc, full)
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(out_path, "w").write("ok" if int(t.item()) == 1 else "mismatch")
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    for n in (1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))


def test_world2_allgather_fold(tmp_path):
    out = str(tmp_path / "res.txt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 3000, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def _closed_form_worker(rank, world, port, n, out_path):
    """bench.py's N > 1 parity check, on CPU: every rank holds a shard of base scalars b_i (bases b_i G) and of scalars s_i, computes its
    part of sum s_i b_i mod r with bench.dot_mod_r, the parts travel through bench.gather_ints, and the closed-form point
    (sum mod r) G must equal the oracle's MSM over all bases — the comparison bench.py makes with the NCCL-folded GPU result."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import coracle as co
    from zero_chain_b200 import synthetic as sy
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b_all = [sy.random_fr_limbs(n, 7 + r) for r in range(world)]              # bench.py: base scalars seeded 7 + rank
    s_all = [bench.make_scalars(n, r, 3) for r in range(world)]
    parts = bench.gather_ints(dist, torch, world, bench.dot_mod_r(s_all[rank], b_all[rank]))
    want_scalar = sum(sum(x * y for x, y in zip(co.limbs_to_ints(s_all[r]), co.limbs_to_ints(b_all[r]))) for r in range(world)) % bench.R_MODULUS
    ok = sum(parts) % bench.R_MODULUS == want_scalar
    if rank == 0:
        bases = np.concatenate([co.g1_fixed_base(b) for b in b_all])
        full = co.g1_encode(co.g1_msm(bases, np.concatenate(s_all)), False)
        ok = ok and bench.closed_form_g1(sum(parts)) == full
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(out_path, "w").write("ok" if int(t.item()) == 1 else "mismatch")
    dist.destroy_process_group()


def test_world2_closed_form_checker(tmp_path):
    out = str(tmp_path / "res2.txt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_closed_form_worker, args=(2, port, 2000, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
