#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native Groth16 prover hot path.

Metric (BASELINE.json): G1 Pippenger MSM Mop/s at 2^20 bases (configs[1]), whole-job aggregate over
N GPUs, with proofs/sec for the confidential_transfer-shaped circuit reported in "secondary".

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N ...            # reference arm: the CPU restatement of the
                                                           # reference's bellman/pairing path (oracle/) on
                                                           # the box's host cores; rank 0 only

A "step" is one complete MSM of 2^20 terms per GPU: scalars -> digits -> counting sort -> bucket
accumulation -> bucket reduction -> canonical affine result (96 bytes, bit-identical to the oracle).
N > 1 is weak scaling: every rank owns a 2^20-base shard of an N*2^20-term MSM (bases partitioned by
index range); the 192-byte partial results are exchanged with one NCCL all-gather and
folded on every rank.  `value` has scalars resident in HBM; `e2e` goes through the C-ABI call with
scalars in pinned HOST memory (host->device copy and the 96-byte device->host result inside the timed
region).  Only the cpu_baseline leg and --impl reference touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 20
N_SETS = 8                     # distinct scalar vectors cycled through: 8 x 32 MiB = 256 MiB > 126 MB L2
KERNELS_PER_MSM = 24           # 20-bit windows at 2^20: digits, tile_hist, col_scan, 3 x scan_block, scan_add, scatter, fine_sort,
                               # pick_task_len, len_hist, len_scan, len_place, accumulate, combine_serial, combine_warp,
                               # rowcol_stage1, 2 x seg_sums, bit_sums, sum_points, finish_bits, join_rowcol, encode_xyzz
                               # (counted from the ncu launch list profiles/r01_launches_msm_2p20.csv; N > 1 adds the fold kernel)
ALGO_MODMUL_PER_TERM = 176     # 11 (mixed add) x ceil(255/16) windows — the FIXED convention of SURVEY.md §8(d) / BASELINE.md §3,
                               # independent of the window size the library actually uses
ALGO_BYTES_PER_TERM = 128      # 96 B base + 32 B scalar


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def make_scalars(n, rank, k):
    from zero_chain_b200 import synthetic as sy
    return sy.random_fr_limbs(n, 1000 + 97 * rank + k)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from zero_chain_b200 import groth16 as zk
    from zero_chain_b200 import synthetic as sy

    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to STDOUT when the first communicator is created; stdout must carry ONE JSON
        # line, so fd 1 points at stderr while the process group and its communicator come up.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True          # the 192-byte all-gather belongs to the latency-bound tail of an MSM
            dist.init_process_group("nccl", pg_options=opts, device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n = 1 << args.log_n
    ctx = zk.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    # ---- setup (untimed): this rank's shard of the bases, generated on the device, + window tables ----
    t0 = time.time()
    base_scalars = sy.random_fr_limbs(n, 7 + rank)
    bases_limbs = zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, base_scalars)      # uniform random subgroup points
    bases = zk.Bases(ctx, 1, bases_limbs, window_bits=args.window_bits, precompute=True)
    setup_s = time.time() - t0
    h_sets = [make_scalars(n, rank, k) for k in range(N_SETS)]
    d_sets = [torch.from_numpy(h.view(np.int64)).cuda() for h in h_sets]
    pinned = [torch.from_numpy(h.view(np.int64)).pin_memory() for h in h_sets]
    psz = zk.partial_size(1)
    d_part = torch.zeros(psz, dtype=torch.uint8, device="cuda")
    d_all = torch.zeros(psz * world, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def step_device(k):
        d = d_sets[k % N_SETS]
        if world == 1:
            return zk.multiexp_device(bases, d.data_ptr(), n)
        zk.multiexp_partial_device(bases, d.data_ptr(), n, d_part.data_ptr())
        with torch.cuda.stream(stream):                 # NCCL all-gather enqueued on the library's stream: no host sync needed
            dist.all_gather_into_tensor(d_all, d_part)
        return zk.points_fold(ctx, 1, d_all.data_ptr(), world)

    d_stage = torch.empty_like(d_sets[0])

    def step_e2e(k):
        h = pinned[k % N_SETS]
        if world == 1:
            return zk.multiexp(bases, h.numpy().view(np.uint64).reshape(-1, 4))    # C-ABI call with a HOST buffer
        with torch.cuda.stream(stream):
            d_stage.copy_(h, non_blocking=True)
        zk.multiexp_partial_device(bases, d_stage.data_ptr(), n, d_part.data_ptr())
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(d_all, d_part)
        return zk.points_fold(ctx, 1, d_all.data_ptr(), world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for k in range(warmup):
            fn(k)
        barrier()
        if profile:
            ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        last = None
        for k in range(steps):
            last = fn(warmup + k)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        prof = ctx.profile_read() if profile else None
        if profile:
            ctx.profile(False)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), last, prof

    # world == 1: successive MSMs are independent jobs, and the reference's multiexp returns a future — the timed loop keeps two
    # of them in flight on two contexts (zk_msm_begin / zk_msm_end), so the latency-bound tail of one MSM and the upload of the
    # next scalars overlap the accumulation of the other.  The blocking single-call numbers are reported beside it.
    ctx2 = zk.Context(local)

    def timed_pipelined(begin, steps, warmup, profile=False):
        ctxs = [ctx, ctx2]
        res = {}

        def run(k0, k1):
            inflight = [None, None]
            for k in range(k0, k1):
                c = k % 2
                if inflight[c] is not None:
                    res[inflight[c]] = zk.multiexp_end(ctxs[c], bases)
                begin(ctxs[c], k)
                inflight[c] = k
            for k in sorted(x for x in inflight if x is not None):
                res[k] = zk.multiexp_end(ctxs[k % 2], bases)
        run(0, warmup)
        barrier()
        if profile:
            ctx.profile(True); ctx2.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(warmup, warmup + steps)
        e1.record(stream)                      # every MSM has been collected on the host, so this is after all of the work
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        prof = None
        if profile:
            a, b_ = ctx.profile_read(), ctx2.profile_read()
            prof = (a[0] + b_[0], a[1] + b_[1])
            ctx.profile(False); ctx2.profile(False)
        return ms, res[warmup + steps - 1], prof

    if world == 1:
        begin_device = lambda c, k: zk.multiexp_device_begin(c, bases, d_sets[k % N_SETS].data_ptr(), n)
        begin_e2e = lambda c, k: zk.multiexp_begin(c, bases, pinned[k % N_SETS].numpy().view(np.uint64).reshape(-1, 4))
    else:
        # per context: its partial, the gathered partials, a staging buffer for the e2e arm, and torch views of its two streams
        lanes = {}
        for c in (ctx, ctx2):
            lanes[id(c)] = dict(part=torch.zeros(psz, dtype=torch.uint8, device="cuda"), all=torch.zeros(psz * world, dtype=torch.uint8, device="cuda"),
                                stage=torch.empty_like(d_sets[0]), main=torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local)),
                                tail=torch.cuda.ExternalStream(zk.tail_stream(c), device=torch.device("cuda", local)))

        def begin_multi(c, d_ptr):
            L = lanes[id(c)]
            zk.multiexp_partial_device_begin(c, bases, d_ptr, n, L["part"].data_ptr())
            with torch.cuda.stream(L["tail"]):          # the all-gather is ordered after the partial on the context's tail stream
                dist.all_gather_into_tensor(L["all"], L["part"])
            zk.points_fold_begin(c, 1, L["all"].data_ptr(), world)

        begin_device = lambda c, k: begin_multi(c, d_sets[k % N_SETS].data_ptr())

        def begin_e2e(c, k):
            L = lanes[id(c)]
            with torch.cuda.stream(L["main"]):
                L["stage"].copy_(pinned[k % N_SETS], non_blocking=True)
            begin_multi(c, L["stage"].data_ptr())

    # modmul roofline calibrated live on this GPU (register-resident independent Fq products)
    modmul_peak, _ = zk.bench_modmul(ctx, zk.FIELD_FQ, 148 * 4, 256, 3000)

    sampler = ClockSampler(local) if rank == 0 else None
    W = max(4, args.warmup)            # the last timed step (W + steps - 1) picks the scalar set the CPU check uses
    ms_dev, res_dev, prof = timed_pipelined(begin_device, args.steps, W, profile=True)
    clocks = sampler.stop() if sampler else None
    ms_e2e, res_e2e, _ = timed_pipelined(begin_e2e, args.steps, W)
    ms_b, res_b, _ = timed(step_device, args.steps, W)
    ms_be, res_be, _ = timed(step_e2e, args.steps, W)
    if not (res_b == res_dev and res_be == res_e2e):
        raise SystemExit("PARITY FAILURE: pipelined and blocking MSM results differ")
    blocking = {"device_ms_per_step": ms_b / args.steps, "device_mops": n * world * args.steps / (ms_b * 1e-3) / 1e6,
                "e2e_ms_per_step": ms_be / args.steps, "e2e_mops": n * world * args.steps / (ms_be * 1e-3) / 1e6,
                "api": "zk_msm_device / zk_msm (N > 1: zk_msm_partial_device + all-gather + zk_points_fold), one call at a time"}

    total_terms = n * world
    value = total_terms * args.steps / (ms_dev * 1e-3) / 1e6
    e2e_value = total_terms * args.steps / (ms_e2e * 1e-3) / 1e6
    hbm_peak, hbm_how = peaks()
    acc_ms, acc_launches = prof
    acc_avg_s = acc_ms * 1e-3 / max(1, acc_launches)
    algo_modmul = ALGO_MODMUL_PER_TERM * n            # per launch: one launch processes one rank's n terms
    roofline = {
        "kernel": "zkmsm::k_accumulate<Fq> (bucket accumulation)",
        "bound": "int32-modmul",                       # IMAD issue rate, not HBM, not tensor
        "achieved": algo_modmul / acc_avg_s, "peak": modmul_peak, "unit": "Fq-modmul/s",
        "frac": algo_modmul / acc_avg_s / modmul_peak,
        "peak_how": "zk_bench_modmul: register-resident independent Fq Montgomery products, measured in this run",
        "avg_launch_ms": acc_avg_s * 1e3, "launches": acc_launches, "share_of_step": acc_ms / ms_dev,
        # the convention above counts 11 products x 16 windows per term; the kernel actually executes 10 products (XYZZ mixed
        # addition) x W windows per term, so with W = 13 (20-bit windows) `frac` can exceed 1 — `executed_frac` is the pipe efficiency
        "note": "frac uses SURVEY 8(d)'s fixed algorithmic count (176 products per term = 11 x 16 windows); the kernel executes "
                "10 x %d per term (XYZZ mixed addition, %d-bit windows from precomputed tables), so frac can exceed 1 — executed_frac is "
                "the pipe efficiency of the kernel as run" % (255 // bases.window_bits + 1, bases.window_bits),
        "executed_modmul_per_launch": 10 * n * (255 // bases.window_bits + 1),
        "executed_frac": 10 * n * (255 // bases.window_bits + 1) / acc_avg_s / modmul_peak,
        "whole_msm_frac": ALGO_MODMUL_PER_TERM * total_terms * args.steps / (ms_dev * 1e-3) / (modmul_peak * world),
        "hbm": {"bound": "hbm", "achieved": ALGO_BYTES_PER_TERM * n / acc_avg_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": ALGO_BYTES_PER_TERM * n / acc_avg_s / 1e9 / hbm_peak, "peak_how": hbm_how},
        "traffic": None,
    }
    try:
        roofline["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_accumulate_dram_bytes_per_launch"]
    except Exception:
        pass

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import coracle as co
        t = time.time()
        want = co.g1_msm(bases_limbs, h_sets[(W + args.steps - 1) % N_SETS])
        dt = time.time() - t
        ok = co.g1_encode(want, False) == res_dev == res_e2e
        cpu_baseline = {"value": n / dt / 1e6, "unit": "Mop/s", "cores": co.num_threads(), "kind": "port",
                        "sample": "one full 2^%d-term MSM (same bases and scalars as the last timed GPU step), oracle/zk_oracle.c "
                                  "bellman-style Pippenger, %.2f s" % (args.log_n, dt),
                        "matches_gpu_result": bool(ok)}
        if not ok:
            raise SystemExit("PARITY FAILURE: GPU MSM result differs from the oracle")

    # batched proving at N > 1 is "replicas only" (SURVEY.md §8e): every rank proves its own 256-proof batch with a resident
    # CRS, no data-path collective; aggregate = all proofs / slowest rank
    replicas = None
    if world > 1 and args.secondary:
        try:
            g = prove_metrics(ctx, zk, sy, args, batch=256, steps=2, cpu=False)
            t = torch.tensor([g["ms_per_batch"], g["from_witness"]["ms_per_batch"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            replicas = {"metric": g["metric"], "scaling": "weak (replicas, no collective)", "batch_per_gpu": 256,
                        "e2e_proofs_per_sec": world * 256 / (float(t[0]) * 1e-3), "from_witness_proofs_per_sec": world * 256 / (float(t[1]) * 1e-3),
                        "ms_per_batch_max_over_ranks": float(t[0])}
        except Exception as e:
            replicas = {"error": repr(e)}
    if rank == 0:
        line = {
            "metric": "g1_msm_mops_2^%d" % args.log_n, "value": value, "unit": "Mop/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32-limb Montgomery (Fq 12x32, Fr 8x32), integer", "data": "synthetic",
            "config": {"workload": "G1 Pippenger MSM, 2^%d uniform-random subgroup bases per GPU (bases sharded by index range, "
                                   "partial sums all-gathered over NCCL), uniform Fr scalars" % args.log_n,
                       "window_bits": bases.window_bits, "precomputed_window_tables": True,
                       "l2_policy": "inputs larger than L2: %d distinct 32 MiB scalar vectors cycled, %.2f GiB window tables gathered randomly" % (N_SETS, (255 // bases.window_bits + 1) * n * 96 / 2**30),
                       "setup_s_untimed": round(setup_s, 2)},
            "e2e": {"value": e2e_value, "unit": "Mop/s", "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": 96 * world,
                    "ms_per_step": ms_e2e / args.steps,
                    "api": "zk_msm_begin / zk_msm_end (C ABI futures, scalars in pinned host memory, two in flight)" if world == 1 else
                           "H2D copy of the scalars + zk_msm_partial_device_begin + NCCL all-gather + zk_points_fold_begin / zk_msm_end, two in flight"},
            "gpu_launches": (KERNELS_PER_MSM + (1 if world > 1 else 0)) * args.steps * world,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if blocking is not None:
            line["config"]["pipelining"] = "two MSMs in flight on two contexts (futures), tail kernels on a high-priority stream"
            line["blocking_call"] = blocking
        if args.secondary and world == 1:
            try:
                line["secondary"] = secondary_metrics(ctx, zk, sy, args)
            except Exception as e:      # the headline line must still print
                line["secondary"] = {"error": repr(e)}
        if replicas is not None:
            line["secondary"] = {"groth16_replicas": replicas}
        print(json.dumps(line), flush=True)
    # ordered teardown: tensors that were used on the library's stream must be released before the stream is
    # destroyed with the context (their allocator blocks record events on it when freed)
    if world > 1:
        lanes.clear()
    del d_sets, pinned, d_stage, d_part, d_all
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    bases.free()
    ctx2.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def secondary_metrics(ctx, zk, sy, args):
    """NTT 2^22 and batched proving of the confidential_transfer-shaped circuit (short runs)."""
    import torch
    import ctypes as C
    from zero_chain_b200 import _lib
    out = {}
    stream = torch.cuda.ExternalStream(ctx.stream)
    # Fr NTT 2^22, device resident
    logn = 22
    d = torch.from_numpy(sy.random_fr_limbs(1 << logn, 5).view(np.int64)).cuda()
    torch.cuda.synchronize()
    L = _lib.lib()
    for _ in range(3):
        L.zk_ntt_fr_device(ctx._h, C.c_void_p(d.data_ptr()), logn, 0)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    reps = 10
    for _ in range(reps):
        L.zk_ntt_fr_device(ctx._h, C.c_void_p(d.data_ptr()), logn, 0)
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fr_peak, _ = zk.bench_modmul(ctx, zk.FIELD_FR, 148 * 4, 256, 3000)
    out["ntt_fr_2^22"] = {"ms": ms, "melem_per_s": (1 << logn) / ms / 1e3, "algo_modmul_frac": (1 << (logn - 1)) * logn / (ms * 1e-3) / fr_peak,
                          "fr_modmul_peak": fr_peak}
    out["groth16"] = prove_metrics(ctx, zk, sy, args)
    return out


def prove_metrics(ctx, zk, sy, args, batch=256, steps=3, cpu=True):
    """proofs/sec for the confidential_transfer-shaped synthetic circuit: batch of 256 witnesses in
    pinned host memory -> zk_groth16_prove_batch (one C-ABI call per step, H2D of every witness and D2H of the proofs
    inside the timed region); CPU: the oracle's create_proof on the same CRS and witness, all host threads."""
    import torch
    r1cs = sy.make_r1cs(seed=1, **sy.CONF_SHAPE)
    dens = sy.densities(r1cs)
    g1 = lambda s: zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, s)
    g2 = lambda s: zk.scalar_mul_many(ctx, 2, zk.G2_GENERATOR, s)
    crs = sy.make_toy_crs(r1cs, g1, g2, seed=2)                      # toy CRS (known trapdoor), exact Parameters::write bytes
    t = time.time()
    params = zk.Parameters.read(ctx, crs.params_bytes, checked=True)
    load_s = time.time() - t
    n_w = 8                                                          # distinct witnesses, cycled with distinct (r, s)
    ws = []
    for k in range(n_w):
        z = sy.make_witness(r1cs, 100 + k)
        a, b, c = sy.evaluate(r1cs, z)
        ws.append([sy.ints_to_limbs(v) for v in (a, b, c, z[:r1cs.n_inputs], z[r1cs.n_inputs:])])
    def pinned(j):
        arr = np.stack([ws[k % n_w][j] for k in range(batch)])
        return torch.from_numpy(arr.view(np.int64)).pin_memory()
    bufs = [pinned(j) for j in range(5)]
    views = [t_.numpy().view(np.uint64) for t_ in bufs]
    rng = sy.SplitMix64(4242)
    rs = sy.ints_to_limbs([rng.fr() for _ in range(batch)]); ss = sy.ints_to_limbs([rng.fr() for _ in range(batch)])
    h2d = sum(v.nbytes for v in views) + rs.nbytes + ss.nbytes
    zk.create_proof_batch_raw(params, batch, *views, *dens, rs, ss)          # warm-up (allocations, NTT tables)
    zk.create_proof_batch_raw(params, batch, *views, *dens, rs, ss)
    t = time.perf_counter()
    for _ in range(steps):
        proofs = zk.create_proof_batch_raw(params, batch, *views, *dens, rs, ss)
    dt = (time.perf_counter() - t) / steps
    # two batches in flight: a second context (own streams and workspace, same resident CRS) driven by a second host thread, so
    # the upload of one batch and the latency-bound tails of its MSMs overlap the other's kernels (ctypes releases the GIL)
    two = None
    try:
        import threading
        ctx_b = zk.Context(ctx.device)
        params_b = zk.Parameters(ctx_b, params._h, [params.n_ic, params.n_h, params.n_l, params.n_a, params.n_b_g1, params.n_b_g2])
        outs = [None, None]

        def worker(i, prm, reps):
            for _ in range(reps):
                outs[i] = zk.create_proof_batch_raw(prm, batch, *views, *dens, rs, ss)
        worker(1, params_b, 1)                                          # warm-up of the second context (allocations, lanes)
        th = [threading.Thread(target=worker, args=(i, prm, steps)) for i, prm in enumerate((params, params_b))]
        t = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt2 = (time.perf_counter() - t) / (2 * steps)
        if outs[0] != proofs or outs[1] != proofs:
            raise SystemExit("PARITY FAILURE: proofs made with two batches in flight differ")
        two = {"e2e_proofs_per_sec": batch / dt2, "ms_per_batch": dt2 * 1e3, "batches_in_flight": 2,
               "how": "two contexts on one GPU, one host thread each, blocking zk_groth16_prove_batch calls"}
        params_b._h = None                                              # the handle belongs to `params`
        ctx_b.close()
    except SystemExit:
        raise
    except Exception as e:
        two = {"error": repr(e)}
    # same batch straight from the assignments: the fixed constraint system is resident on the device and the GPU
    # evaluates <A_j,z>, <B_j,z>, <C_j,z> itself (zk_groth16_prove_witness_batch)
    cs = zk.ConstraintSystem(ctx, r1cs.n_inputs, r1cs.n_aux, r1cs.A, r1cs.B, r1cs.C)
    zk.create_proof_from_witness_batch(cs, params, batch, views[3], views[4], rs, ss)
    t = time.perf_counter()
    for _ in range(steps):
        proofs_w = zk.create_proof_from_witness_batch(cs, params, batch, views[3], views[4], rs, ss)
    dt_w = (time.perf_counter() - t) / steps
    if proofs_w != proofs:
        raise SystemExit("PARITY FAILURE: witness-path proofs differ from the evaluation-path proofs")
    h2d_w = views[3].nbytes + views[4].nbytes + rs.nbytes + ss.nbytes
    cs.free()
    lat = 1e9
    for _ in range(4):                   # first call loads the single-domain kernels lazily; report the steady-state latency
        t = time.perf_counter()
        single = zk.create_proof_batch_raw(params, 1, *[v[:1] for v in views], *dens, rs[:1], ss[:1])
        lat = min(lat, time.perf_counter() - t)
    assert single == proofs[:192]
    # CPU port of the reference path on the same CRS / witness
    cpu_block = None
    if cpu:
        from oracle import coracle as co
        op = co.Params(crs.params_bytes, checked=False)
        r0 = sum(int(x) << (64 * i) for i, x in enumerate(rs[0])); s0 = sum(int(x) << (64 * i) for i, x in enumerate(ss[0]))
        w0 = [v[0] for v in views]
        t = time.perf_counter(); want = op.prove(*w0, *dens, r0, s0); cpu_dt = time.perf_counter() - t
        t = time.perf_counter(); op.prove(*w0, *dens, r0, s0); cpu_dt = min(cpu_dt, time.perf_counter() - t)
        if want != proofs[:192]:
            raise SystemExit("PARITY FAILURE: GPU proof bytes differ from the oracle")
        cpu_block = {"value": 1.0 / cpu_dt, "unit": "proofs/s", "cores": co.num_threads(), "kind": "port",
                     "sample": "oracle create_proof, best of 2, same CRS/witness", "matches_gpu_proof_bytes": True}
    params.free()
    try:
        verify_block = verify_metrics(ctx, zk, crs.params_bytes, proofs, np.ascontiguousarray(views[3][:, 1:, :]), cpu)
    except SystemExit:
        raise
    except Exception as e:
        verify_block = {"error": repr(e)}
    return {"metric": "proofs_per_sec (confidential_transfer shape: 19974 constraints, 23 inputs, domain 2^15; synthetic R1CS, toy CRS)",
            "e2e_proofs_per_sec": batch / dt, "batch": batch, "steps": steps, "ms_per_batch": dt * 1e3, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": 192 * batch, "single_proof_latency_ms": lat * 1e3, "params_load_checked_s": load_s,
            "from_witness": {"e2e_proofs_per_sec": batch / dt_w, "ms_per_batch": dt_w * 1e3, "h2d_bytes_per_step": int(h2d_w),
                             "api": "zk_groth16_prove_witness_batch (constraint system resident, GPU evaluates the R1CS rows)"},
            "two_batches_in_flight": two, "cpu_baseline": cpu_block, "verify": verify_block,
            "timing": "host wall clock around synchronous C-ABI calls (each call ends with a stream synchronise)"}


def verify_metrics(ctx, zk, vk_bytes, proofs, inputs, cpu=True, n_v=8192):
    """verifications/sec: the proofs the prover just made, replicated to a block-import sized batch, through
    zk_groth16_verify_batch (host buffers: H2D of proofs + public inputs, D2H of the verdicts inside the timed region) and
    zk_groth16_verify_batch_device (resident inputs, CUDA events on the library's stream); CPU: the oracle's verify_proof."""
    import torch
    import ctypes as C
    from zero_chain_b200 import _lib
    L = _lib.lib()
    t = time.perf_counter()
    pvk = zk.PreparedVerifyingKey.prepare(ctx, vk_bytes)
    prep_s = time.perf_counter() - t
    batch = len(proofs) // 192
    n_in = inputs.shape[1]
    reps = (n_v + batch - 1) // batch
    pb = np.tile(np.frombuffer(proofs, np.uint8), reps)[:192 * n_v].copy()
    inp = np.tile(inputs.reshape(batch, -1), (reps, 1))[:n_v].copy()
    hp, hi = torch.from_numpy(pb).pin_memory(), torch.from_numpy(inp.view(np.int64)).pin_memory()
    out = np.zeros(n_v, np.uint8)
    call = lambda: zk._ck(L.zk_groth16_verify_batch(ctx._h, pvk._h, n_v, C.c_void_p(hp.data_ptr()), C.c_void_p(hi.data_ptr()), n_in,
                                                    out.ctypes.data_as(C.c_void_p)))
    call()
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); call(); best = min(best, time.perf_counter() - t)
    if not (out == 1).all():
        raise SystemExit("PARITY FAILURE: the GPU verifier rejected a proof made by the GPU prover")
    dp, di = hp.cuda(), hi.cuda()
    dv = torch.zeros(n_v, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream)
    torch.cuda.synchronize()
    dev = lambda: zk.verify_proofs_device(pvk, n_v, dp.data_ptr(), di.data_ptr(), n_in, dv.data_ptr())
    dev(); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        dev()
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1) / 3
    assert bool((dv == 1).all())
    # a tampered copy: one public input bumped in every 5th proof -> exactly those are rejected
    bad = inp.copy(); bad[::5, 0] ^= 1
    hb = torch.from_numpy(bad.view(np.int64))
    zk._ck(L.zk_groth16_verify_batch(ctx._h, pvk._h, n_v, C.c_void_p(hp.data_ptr()), C.c_void_p(hb.data_ptr()), n_in, out.ctypes.data_as(C.c_void_p)))
    want = np.ones(n_v, np.uint8); want[::5] = 0
    if not (out == want).all():
        raise SystemExit("PARITY FAILURE: the GPU verifier accepted a proof with a wrong public input")
    cpu_block = None
    if cpu:
        from oracle import coracle as co
        k = co.PreparedVerifyingKey.prepare(vk_bytes)
        if k.write() != pvk.write():
            raise SystemExit("PARITY FAILURE: prepared verifying key differs from the oracle's")
        n_c = 8 * co.num_threads()
        t = time.perf_counter(); v = k.verify_batch(pb[:192 * n_c].tobytes(), bad[:n_c], n_in); cpu_dt = time.perf_counter() - t
        if v != [int(x) for x in want[:n_c]]:
            raise SystemExit("PARITY FAILURE: oracle verdicts differ from the GPU's")
        cpu_block = {"value": n_c / cpu_dt, "unit": "verifications/s", "cores": co.num_threads(), "kind": "port",
                     "sample": "%d proofs (every 5th with a wrong input), oracle/pairing_oracle.inc verify_proof, %.2f s" % (n_c, cpu_dt),
                     "matches_gpu_verdicts": True}
    pvk.free()
    return {"metric": "verifications_per_sec (Proof::read + verify_proof, %d public inputs, one prepared key)" % n_in, "batch": n_v,
            "e2e_verifications_per_sec": n_v / best, "e2e_ms_per_batch": best * 1e3, "h2d_bytes_per_step": int(pb.nbytes + inp.nbytes),
            "d2h_bytes_per_step": n_v, "device_verifications_per_sec": n_v / (ms_dev * 1e-3), "device_ms_per_batch": ms_dev,
            "prepare_verifying_key_s": prep_s, "cpu_baseline": cpu_block}


def run_reference(args):
    """Reference arm: the CPU restatement of the reference's bellman/pairing path (oracle/zk_oracle.c;
    the reference itself is Rust and cannot be built here: no cargo/rustc, bellman un-vendored)."""
    world, rank = env_int("WORLD_SIZE", 1), env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import coracle as co
    from zero_chain_b200 import synthetic as sy
    co.build()
    n_full = 1 << args.log_n
    n = min(n_full, 1 << 18)              # bounded sample of the workload: 2^18 of the 2^20 terms per step
    bases = co.g1_fixed_base(sy.random_fr_limbs(n, 7))
    sets = [make_scalars(n, 0, k) for k in range(4)]
    for k in range(args.warmup):
        co.g1_msm(bases, sets[k % 4])
    t0 = time.time()
    for k in range(args.steps):
        co.g1_msm(bases, sets[(args.warmup + k) % 4])
    dt = time.time() - t0
    value = n * args.steps / dt / 1e6
    cores = co.num_threads()
    line = {"impl": "reference", "metric": "g1_msm_mops_2^%d" % args.log_n, "value": value, "unit": "Mop/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64-limb Montgomery, integer", "data": "synthetic",
            "config": {"workload": "G1 Pippenger MSM (bellman multiexp restatement, c = ceil(ln n), one thread per window), "
                                   "bounded sample: 2^18 of the 2^%d terms per step" % args.log_n},
            "cpu_baseline": {"value": value, "unit": "Mop/s", "cores": cores, "kind": "port",
                             "sample": "2^18-term MSM per step, %d steps, all %d host threads" % (args.steps, cores)},
            "e2e": {"value": value, "unit": "Mop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", dest="log_n", type=int, default=LOG_N)
    ap.add_argument("--window-bits", dest="window_bits", type=int, default=0, help="0 = library default (20 bits from 2^20 terms, else <= 16)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
