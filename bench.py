#!/usr/bin/env python3
"""bench.py -- the hot path's headline benchmark (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): synthetic
1000 taxa x 100 000 site patterns, GTR + Weibull(4), a batch of 64 ELBO parameter draws per step;
one step = log-likelihood + full gradient of all 64 draws.  Per-GPU work is fixed (weak scaling):
with N ranks the alignment has N x 100 000 patterns, sharded by pattern, and every step ends with one
NCCL all-reduce of the [64, 1+grad] result block.

`value`  = pattern x category likelihood+gradient evaluations per second with the parameters
           already resident in HBM (kernels only), whole job.
`e2e`    = the same through the C ABI `phylo_b200_eval_batch` with HOST buffers: parameter packing,
           H2D, kernels, D2H inside the timed region.
`parity`  = the timed run's own draw-0 row against the CPU oracle on ALL patterns of this rank's shard
           (north-star tolerances 1e-10 / 1e-8); with N > 1 also DS1 (committed fixture) split N ways through
           ShardedLikelihood over NCCL against the oracle, and the all-reduce of the timed block against the sum
           of the ranks' local rows.  A failed check makes the process exit non-zero.
`config4` = BASELINE.json configs[3]: 10 000 taxa x 1 000 000 patterns, GTR+W4, 2 draws, the alignment
           generated on the GPUs and sharded N ways (STRONG scaling: total work fixed), one NCCL all-reduce.
`--impl reference` times the CPU oracle (oracle/phylo_oracle.c, a port of the reference algorithm --
the reference's own C++ needs Eigen/Stan Math, which are not installed) with every host thread
on a bounded pattern slice of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S_TAXA, L_PATTERNS, N_CAT, N_DRAWS = 1000, 100_000, 4, 64
METRIC = "site-pattern x category L+grad evals/sec (fp64)"
UNIT = "pattern*category evals/s"
WORKLOAD = "synthetic 1000 taxa x 100k patterns, GTR+W4, batch of 64 ELBO parameter draws"


def algorithmic_bytes(S, L, C):
    """B_vg = 32 L C (5S-9) + 2 S L + 8 L per value+gradient evaluation."""
    return 32.0 * L * C * (5 * S - 9) + 2.0 * S * L + 8.0 * L


def design_bytes(S, L, C):
    """Minimum DRAM traffic of the depth-first sweep actually implemented: one write and one read of
    every internal partial (32 B) and of its rescale byte, tip codes once per sweep, weights once.  (The message-
    statistic sweep stores the S-2 messages of the non-root internal nodes instead of S-1 partials: one row of 2S-3
    less, 0.05 % at S = 1000 -- the same formula is kept.)"""
    return (32.0 + 1.0) * L * C * (2 * S - 3) + 2.0 * S * L + 8.0 * L


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(n_shards=1, shard=0):
    from phylostan_b200 import synth
    prob = synth.make_problem(S_TAXA, L_PATTERNS, N_CAT, seed=synth.SEED_DATA + shard)
    if shard:  # every shard shares the tree and base branch lengths of shard 0, with its own columns
        base = synth.make_problem(S_TAXA, 32, N_CAT, seed=synth.SEED_DATA)
        rng = np.random.default_rng(synth.SEED_DATA + shard)
        tm, w = synth.simulate_alignment(base.peel, base.blens, L_PATTERNS, N_CAT, rng)
        prob = synth.SynthProblem(S_TAXA, L_PATTERNS, N_CAT, base.peel, tm, w, base.blens)
    draws = synth.make_draws(prob, N_DRAWS)
    return prob, draws


# ----------------------------------------------------------------------------- reference arm

REF_SAMPLE_PATTERNS = 32768   # the same bounded sample at every N


def host_threads():
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to its children; the CPU arm is
    meant to use the whole box, so the count comes from the affinity mask, not from the environment."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def shared_config(world):
    """The workload description both arms print (identical keys and values for the same N)."""
    return {"workload": WORKLOAD, "taxa": S_TAXA, "patterns_per_gpu": L_PATTERNS, "categories": N_CAT,
            "draws_per_step": N_DRAWS,
            "parallelism": f"pattern-sharded x{world}, one NCCL all-reduce per step" if world > 1 else "single GPU",
            "l2": "per-step working set (2.4 GB partial scratch + 130 MB matrices) exceeds the 126 MB L2"}


def run_reference(args):
    """CPU oracle ("port") with all host threads; one step = value+gradient of ONE draw on a fixed
    32 768-pattern slice, scaled to the metric's unit."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import oracle as O
    cores = host_threads()
    prob, draws = build_problem()
    sample_L = min(L_PATTERNS, REF_SAMPLE_PATTERNS)
    sl = slice(0, sample_L)
    bl, rates, freqs, rs, ps = draws

    def step(i):
        d = i % N_DRAWS
        O.loglik_grad(prob.peel, prob.tipmask[:, sl], prob.weights[sl], O.GTR, bl[d], rates[d], freqs[d], rs[d],
                      ps[d], dp_eigen=True, nthreads=cores)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = args.steps * sample_L * N_CAT / dt
    sample = f"{sample_L} of {L_PATTERNS} patterns x 1 of {N_DRAWS} draws per step, value+gradient, {cores} OpenMP threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(world),
            "tree_evals_per_s": value / (L_PATTERNS * N_CAT),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm

RTOL_LOGL, TOL_GRAD = 1e-10, 1e-8   # north-star parity tolerances (fp64)


def parity_block(row, want_flat, label):
    """|dlogL| / |logL| and max_k |dg_k| / max(1, |g_k|) of one packed GPU row against the oracle's."""
    row, want = np.asarray(row, dtype=np.float64), np.asarray(want_flat, dtype=np.float64)
    e_l = abs(row[0] - want[0]) / abs(want[0])
    e_g = float(np.max(np.abs(row[1:] - want[1:]) / np.maximum(1.0, np.abs(want[1:]))))
    return {"what": label, "logL_rel": float(e_l), "grad_max_err": e_g, "tol_logL_rel": RTOL_LOGL, "tol_grad": TOL_GRAD,
            "ok": bool(np.isfinite(e_l) and np.isfinite(e_g) and e_l <= RTOL_LOGL and e_g <= TOL_GRAD)}


def cpu_baseline_and_parity(prob, draws, local_row0, want_cpu):
    """Oracle on this box's host cores: draw 0 on ALL patterns of this rank's shard with every thread (the
    parity reference of the timed run AND the all-threads CPU baseline), and a 1-thread sample."""
    from oracle import oracle as O
    bl, rates, freqs, rs, ps = draws
    cores = host_threads()
    O.loglik_grad(prob.peel, prob.tipmask[:, :64], prob.weights[:64], O.GTR, bl[0], rates[0], freqs[0], rs[0],
                  ps[0], dp_eigen=True, nthreads=cores)  # warm
    t0 = time.perf_counter()
    want = O.loglik_grad(prob.peel, prob.tipmask, prob.weights, O.GTR, bl[0], rates[0], freqs[0], rs[0], ps[0],
                         dp_eigen=True, nthreads=cores)
    dt = time.perf_counter() - t0
    par = parity_block(local_row0, want.flat(), f"timed run's draw 0 vs CPU oracle on all {prob.L} patterns of rank 0's shard")
    cpu = None
    if want_cpu:
        n1 = 8192
        t1 = time.perf_counter()
        O.loglik_grad(prob.peel, prob.tipmask[:, :n1], prob.weights[:n1], O.GTR, bl[0], rates[0], freqs[0], rs[0],
                      ps[0], dp_eigen=True, nthreads=1)
        d1 = time.perf_counter() - t1
        v, v1 = prob.L * N_CAT / dt, n1 * N_CAT / d1
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{prob.L} of {L_PATTERNS} patterns x 1 draw, value+gradient, {dt:.1f} s; "
                         f"1 thread on {n1} patterns: {v1:.4g} {UNIT}",
               "value_1thread": v1, "tree_evals_per_s": v / (L_PATTERNS * N_CAT)}
    return par, cpu


def sharded_parity(world, rank, local, stream):
    """DS1 (27 taxa x 934 patterns, GTR+W4, unrooted; committed fixture tests/golden/DS1.npz) split `world` ways
    through ShardedLikelihood over NCCL, against the CPU oracle on the whole alignment."""
    from phylostan_b200 import likelihood as lk, sharded, encode
    z = np.load(os.path.join(ROOT, "tests", "golden", "DS1.npz"))
    peel, tm, w = z["peel"], z["tipmask"], z["weights"]
    S = tm.shape[0]
    rng = np.random.default_rng(20261020)
    B = 3
    bl = rng.exponential(0.05, size=(B, 2 * S - 3)) + 1e-4
    rates, freqs = rng.dirichlet(np.ones(6) * 3, size=B), rng.dirichlet(np.ones(4) * 5, size=B)
    rs = np.stack([encode.weibull_rates(x, N_CAT) for x in rng.uniform(0.3, 1.5, size=B)])
    ps = rng.dirichlet(np.ones(N_CAT) * 4, size=B)
    lo, hi = sharded.shard_bounds(tm.shape[1], world, rank)
    with lk.TreeLikelihood(peel, tm[:, lo:hi], w[lo:hi], model="GTR", categories=N_CAT, rooted=False, device=local) as lik:
        lik.set_stream(stream.cuda_stream)
        rows = sharded.ShardedLikelihood(lik).packed(bl, rates, freqs, rs, ps)
    if rank != 0:
        return None
    from oracle import oracle as O
    worst = {"logL_rel": 0.0, "grad_max_err": 0.0}
    for d in range(B):
        want = O.loglik_grad(peel, tm, w, O.GTR, bl[d], rates[d], freqs[d], rs[d], ps[d], rooted=False, dp_eigen=True)
        pb = parity_block(rows[d], want.flat(), "")
        worst = {k: max(worst[k], pb[k]) for k in worst}
    worst.update(what=f"DS1 split {world} ways over NCCL (ShardedLikelihood) vs CPU oracle, {B} draws",
                 ok=bool(worst["logL_rel"] <= RTOL_LOGL and worst["grad_max_err"] <= TOL_GRAD))
    return worst


# BASELINE.json configs[3]
C4_TAXA, C4_PATTERNS, C4_DRAWS, C4_SEED = 10_000, 1_000_000, 2, 20261021


def run_config4(args, world, rank, local, stream):
    """10 000 taxa x 1 000 000 patterns, GTR+W4, 2 draws, sharded `world` ways by pattern: STRONG scaling (the
    alignment is fixed, each rank holds 1 000 000 / world patterns and the full tree), one NCCL all-reduce of
    the [2, 1+grad] block per step.  The alignment is simulated ON the GPU (torch) and handed to the library as
    device pointers (phylo_b200_create_device); parity is checked against the CPU oracle on a 192-pattern
    slice of rank 0's shard, run through the same kernel variant (same tiling, same stack slots)."""
    import torch
    import torch.distributed as dist
    from phylostan_b200 import likelihood as lk, sharded, synth
    dev = torch.device("cuda", local)
    rng = np.random.default_rng(C4_SEED)
    peel = synth.coalescent_peel_fast(C4_TAXA, rng)
    blens = np.clip(rng.exponential(0.02, size=2 * C4_TAXA - 2), 1e-4, 0.5)
    lo, hi = sharded.shard_bounds(C4_PATTERNS, world, rank)
    Lr = hi - lo
    t0 = time.perf_counter()
    tm_d, w_d = synth.simulate_alignment_device(peel, blens, Lr, N_CAT, dev, C4_SEED + 1 + rank)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    probe = synth.SynthProblem(C4_TAXA, Lr, N_CAT, peel, None, None, blens)
    bl, rates, freqs, rs, ps = synth.make_draws(probe, C4_DRAWS)
    t0 = time.perf_counter()
    lik = lk.TreeLikelihood(peel, model="GTR", categories=N_CAT, device=local, device_tips=(tm_d.data_ptr(), Lr, w_d.data_ptr()))
    t_create = time.perf_counter() - t0
    n_slice = 192
    tm_s, w_s = tm_d[:, :n_slice].cpu().numpy(), w_d[:n_slice].cpu().numpy()
    del tm_d, w_d
    torch.cuda.empty_cache()
    lik.set_stream(stream.cuda_stream)
    B = C4_DRAWS
    lik.upload(bl, rates, freqs, rs, ps)
    lik.run(B, True)
    out_t = sharded.device_out_tensor(lik, B)
    info = lik.info()

    def step():
        lik.run(B, True)
        if world > 1:
            dist.all_reduce(out_t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        step()
    barrier()
    steps = max(2, min(args.steps, 4))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    lik.set_timing(True)
    lik.run(B, True)
    kt = lik.get_timing()
    lik.set_timing(False)
    ar_ms = 0.0
    if world > 1:   # the collective alone (160 KB payload)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.all_reduce(out_t)
        barrier()
        a0.record(stream)
        for _ in range(20):
            dist.all_reduce(out_t)
        a1.record(stream)
        barrier()
        ar_ms = a0.elapsed_time(a1) / 20
    t = torch.tensor([dev_ms, kt["sweep_ms"], ar_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, sweep_ms, ar_ms = (float(x) for x in t.cpu())
    lik.close()
    # parity: the same tree / draws on a small slice, through the same kernel variant, against the oracle
    par = None
    if rank == 0:
        from oracle import oracle as O
        mask = tm_s
        with lk.TreeLikelihood(peel, mask, w_s, model="GTR", categories=N_CAT, device=local) as small:
            small.set_tiling(info["patterns_per_thread"], 1)
            small.set_stack_slots(info["stack_slots"])
            got = small.value_grad(bl[0], rates[0], freqs[0], rs[0], ps[0])
            sinfo = small.info()
        want = O.loglik_grad(peel, mask, w_s, O.GTR, bl[0], rates[0], freqs[0], rs[0], ps[0], dp_eigen=True,
                             nthreads=host_threads())
        row = np.concatenate([[got.log_P], got.grad])
        par = parity_block(row, want.flat(), f"{n_slice}-pattern slice of rank 0's shard, same kernel variant "
                                             f"(K={sinfo['patterns_per_thread']}, {sinfo['stack_slots']} of "
                                             f"{sinfo['stack_depth']} stack slots on chip) vs CPU oracle")
    if rank != 0:
        return None
    ms_step = dev_ms / steps
    evals = B / (ms_step * 1e-3)
    peak, _ = measured_peak_gbs()
    des = design_bytes(C4_TAXA, Lr, N_CAT) * B     # per sweep launch on one GPU
    return {"workload": "synthetic 10k taxa x 1M patterns, GTR+W4, site patterns sharded across N GPUs, one NCCL "
                        "all-reduce of log-lik + gradient per step",
            "scaling": "strong", "taxa": C4_TAXA, "patterns_total": C4_PATTERNS, "patterns_per_gpu": Lr,
            "draws_per_step": B, "steps": steps, "ms_per_step": ms_step, "tree_evals_per_s": evals,
            "value": evals * C4_PATTERNS * N_CAT, "unit": UNIT,
            "sweep_ms": sweep_ms, "allreduce_ms": ar_ms, "allreduce_bytes": int(B * lik.nout * 8),
            "kernel": ("sweep_kernel<double,K,GRAD,TIPS,128,DEEP" if info["stack_slots"] < info["stack_depth"]
                       else "sweep_kernel<double,K,GRAD,TIPS,128") + (",MSG>" if info["message_statistic"] else ">"),
            "tiling": {k: info[k] for k in ("stack_depth", "stack_slots", "patterns_per_thread", "threads_per_cta",
                                            "grid", "smem_bytes", "tiles", "scratch_bytes", "message_statistic",
                                            "cherry_tables", "post_order_tables")},
            "design_bytes_per_launch": des, "dram_frac_of_measured_peak": des / (sweep_ms * 1e-3) / 1e9 / peak,
            "survey_Bvg_frac_all_gpus": algorithmic_bytes(C4_TAXA, C4_PATTERNS, N_CAT) * evals / (world * peak * 1e9),
            "setup_s": {"simulate_on_gpu": t_gen, "create_device": t_create},
            "data": "synthetic, simulated on the GPU with torch (GTR + Weibull(4) down a Kingman coalescent tree)",
            "parity": par}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from phylostan_b200 import likelihood as lk, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (libphylo_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    prob, draws = build_problem(world, rank)
    bl, rates, freqs, rs, ps = draws
    lik = lk.TreeLikelihood(prob.peel, prob.tipmask, prob.weights, model="GTR", categories=N_CAT, device=local)
    stream = torch.cuda.Stream(device=local)   # a real (non-default) stream shared by the library,
    torch.cuda.set_stream(stream)              # the timing events and NCCL
    lik.set_stream(stream.cuda_stream)
    if args.k or args.pb:
        lik.set_tiling(args.k, args.pb)
    ref64 = None
    if args.fp32:  # error of the fp32 mode against the fp64 path on the same draws, then switch
        lik.upload(bl, rates, freqs, rs, ps)
        lik.run(N_DRAWS, True)
        ref64 = lik.download(N_DRAWS)
        lik.set_precision(32)
    B = N_DRAWS

    lik.upload(bl, rates, freqs, rs, ps)
    lik.run(B, True)
    out_t = sharded.device_out_tensor(lik, B)   # device result block [B, nout] for the NCCL all-reduce

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        lik.run(B, True)
        if world > 1:
            dist.all_reduce(out_t)

    def e2e_step():
        if world == 1:  # the call a user (or the Stan shim) makes: phylo_b200_eval_batch on host arrays
            vg = lik.value_grad(bl, rates, freqs, rs, ps)
            return np.concatenate([vg.log_P[:, None], vg.grad_blens, vg.grad_subst, vg.grad_freqs, vg.grad_rs,
                                   vg.grad_ps], axis=1)
        lik.upload(bl, rates, freqs, rs, ps)       # host packing + H2D
        lik.run(B, True)
        dist.all_reduce(out_t)
        return lik.download(B)                     # D2H + sync

    # ---- device-resident timing (value) with per-kernel CUDA events
    for _ in range(args.warmup):
        resident_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lik.set_timing(True)
    kern = {"pmat_ms": [], "sweep_ms": [], "contract_ms": [], "total_ms": []}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        resident_step()
        if args.kernel_times:
            for k, v in lik.get_timing().items():
                kern[k].append(v)
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    if not args.kernel_times:  # one extra, separately timed step for the per-kernel split
        resident_step()
        for k, v in lik.get_timing().items():
            kern[k].append(v)
    lik.set_timing(False)
    res = lik.download(B)          # the timed configuration's result block (all-reduced when N > 1)
    # the same step without the collective: this rank's own rows, for the parity checks below
    lik.run(B, True)
    local_rows = lik.download(B)

    # ---- end-to-end timing through the host-buffer API
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None

    times = torch.tensor([dev_ms / 1e3, wall, e2e_wall], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, wall_s, e2e_s = (float(x) for x in times.cpu())
    info = lik.info()

    # ---- parity of what was just timed
    parity = {}
    if world > 1:   # the all-reduced block must be the sum of the ranks' local rows (draw 0 and the last draw)
        mine = torch.from_numpy(local_rows[[0, B - 1]].copy()).to(f"cuda:{local}")
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        total = torch.stack(gathered).sum(0).cpu().numpy()
        pb0 = parity_block(res[0], total[0], "")
        pb1 = parity_block(res[B - 1], total[1], "")
        parity["allreduce"] = {"what": "NCCL all-reduced block vs the sum of the ranks' local rows (draws 0 and 63)",
                               "logL_rel": max(pb0["logL_rel"], pb1["logL_rel"]),
                               "grad_max_err": max(pb0["grad_max_err"], pb1["grad_max_err"]),
                               "ok": pb0["ok"] and pb1["ok"]}
        parity["sharded"] = sharded_parity(world, rank, local, stream)
    cpu = None
    if rank == 0:
        e2e_vs_resident = parity_block(res_e2e[0], res[0], "")
        parity["timed_run"], cpu = cpu_baseline_and_parity(prob, draws, local_rows[0], world == 1 and not args.no_cpu)
        parity["e2e_vs_resident"] = {"what": "host-buffer API result vs device-resident result, draw 0",
                                     "logL_rel": e2e_vs_resident["logL_rel"],
                                     "grad_max_err": e2e_vs_resident["grad_max_err"], "ok": e2e_vs_resident["ok"]}
    # ---- secondary numbers of the same handle, reported separately (not the headline): the value-only sweep and the
    #      optional fp32-with-scaling mode with its error against the fp64 rows just timed
    extras = None
    if world == 1 and not args.fp32 and not args.no_extras:
        def timed(grad, n=3):
            lik.run(B, grad)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                lik.run(B, grad)
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        lik.upload(bl, rates, freqs, rs, ps)
        ms_value = timed(False)
        # the same value + gradient step WITHOUT the cherry tables (every internal node's message stored by the post-order
        # and read back by the pre-order: a third more scratch traffic); checked against the rows of the timed run
        lik.set_cherry_tables(False)
        ms_cherry = timed(True)
        rch = lik.download(B)
        cherry_used = lik.info()["cherry_tables"]
        lik.set_cherry_tables(True)
        lik.set_precision(32)
        ms32 = timed(True)
        r32 = lik.download(B)
        lik.set_precision(64)
        g64, g32 = local_rows[:, 1:], r32[:, 1:]
        extras = {
            "without_cherry_tables": {"ms_per_step": ms_cherry, "tree_evals_per_s": B / ms_cherry * 1e3,
                                      "value": B * L_PATTERNS * N_CAT / ms_cherry * 1e3, "unit": UNIT, "used": int(cherry_used),
                                      "logL_max_rel_vs_default": float(np.max(np.abs(rch[:, 0] - local_rows[:, 0]) / np.abs(local_rows[:, 0]))),
                                      "grad_max_err_vs_default": float(np.max(np.abs(rch[:, 1:] - local_rows[:, 1:])
                                                                              / np.maximum(1.0, np.abs(local_rows[:, 1:])))),
                                      "note": "phylo_b200_set_cherry_tables(0): whole step"},
            "value_only": {"ms_per_step": ms_value, "tree_evals_per_s": B / ms_value * 1e3,
                           "value": B * L_PATTERNS * N_CAT / ms_value * 1e3, "unit": UNIT.replace("evals", "value-only evals")},
            "fp32_with_scaling": {"ms_per_step": ms32, "tree_evals_per_s": B / ms32 * 1e3, "value": B * L_PATTERNS * N_CAT / ms32 * 1e3,
                                  "unit": UNIT, "dtype": "f32 partials / f64 sums",
                                  "error_vs_fp64": {
                                      "logL_max_rel": float(np.max(np.abs(r32[:, 0] - local_rows[:, 0]) / np.abs(local_rows[:, 0]))),
                                      "grad_max_abs_over_max_abs": float(np.max(np.abs(g32 - g64)) / np.max(np.abs(g64))),
                                      "grad_median_rel": float(np.median(np.abs(g32 - g64) / np.maximum(1e-300, np.abs(g64))))},
                                  "note": "optional mode, outside the 1e-10 / 1e-8 parity guarantee"}}
    lik.close()
    del out_t

    config4 = None
    if not args.no_config4 and not args.fp32:
        config4 = run_config4(args, world, rank, local, stream)

    ok = True
    if rank == 0:
        Lg = L_PATTERNS * world
        units_per_step = B * Lg * N_CAT
        value = args.steps * units_per_step / dev_s
        e2e_value = args.steps * units_per_step / e2e_s
        sweep_ms = statistics.mean(kern["sweep_ms"])
        peak, peak_src = measured_peak_gbs()
        alg = algorithmic_bytes(S_TAXA, L_PATTERNS, N_CAT) * B          # per sweep launch on one GPU
        des = design_bytes(S_TAXA, L_PATTERNS, N_CAT) * B
        achieved = des / (sweep_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "sweep_traffic.json")
        if os.path.exists(tp):
            # a constant of one ncu capture: only reported when THIS run launched the same kernel shape on the same
            # problem shape (otherwise null -- a stale capture must not pass for a measurement)
            tj = json.load(open(tp))
            cap = tj.get("capture", {})
            same = (cap.get("taxa") == S_TAXA and cap.get("patterns") == L_PATTERNS and cap.get("categories") == N_CAT
                    and cap.get("precision") == (32 if args.fp32 else 64)
                    and all(cap.get(k) == info[k] for k in ("patterns_per_thread", "threads_per_cta", "stack_slots", "smem_bytes",
                                                            "message_statistic", "sweep_variant", "cherry_tables",
                                                            "post_order_tables")))
            traffic = tj.get("dram_bytes_per_evaluation") if same else None
            traffic = traffic * B if traffic else None   # one launch sweeps B draws
        if not args.fp32:
            parity["ok"] = all(v["ok"] for v in parity.values() if isinstance(v, dict))
            if config4 is not None and config4.get("parity") is not None:
                parity["ok"] = parity["ok"] and config4["parity"]["ok"]
            ok = parity["ok"]
        cfg = shared_config(world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": cfg,
            "tiling": {k: info[k] for k in ("stack_depth", "stack_slots", "patterns_per_thread", "threads_per_cta", "grid",
                                            "smem_bytes", "tiles", "message_statistic", "sweep_variant", "cherry_tables",
                                            "post_order_tables")},
            "tree_evals_per_s": value / (Lg * N_CAT),
            "node_updates_per_s": value * (S_TAXA - 1),
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "kernels_ms": {k: statistics.mean(v) for k, v in kern.items()},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "h2d_bytes_per_step": int(B * _param_stride(lik) * 8),
                    "d2h_bytes_per_step": int(B * lik.nout * 8)},
            "gpu_launches": int(args.steps * info["kernel_launches"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": "sweep_kernel<double,4,GRAD,TIPS,128" + (",MSG>" if info["message_statistic"] else ">"),
                         "launch_ms": sweep_ms,
                         "algorithmic_bytes_per_launch": des, "peak_source": peak_src,
                         "survey_model_bytes_per_launch": alg,
                         "frac_vs_survey_model": alg / (sweep_ms * 1e-3) / 1e9 / peak,
                         "frac_of_measured_traffic": (traffic / (sweep_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the capture named in "
                                           "profiles/sweep_traffic.json (a constant of that capture, not of this run; null when this run's "
                                           "kernel shape differs from the captured one)",
                         "note": "algorithmic bytes = the depth-first sweep's own DRAM traffic, 33 L C (2S-3) + 2SL + 8L per "
                                 "evaluation (one write + one read of every internal partial / message and its rescale "
                                 "byte, tip codes once per sweep) -- the basis of this number since round 1 (26.6 GB per "
                                 "evaluation on config 3); with cherry tables the kernel moves a third less than that "
                                 "(traffic, ncu; frac_of_measured_traffic is the DRAM fraction actually sustained); "
                                 "frac_vs_survey_model uses SURVEY 8(d)'s level-synchronous B_vg, which this design does "
                                 "not move, and may exceed 1"},
            "cpu_baseline": cpu, "clocks": clocks,
            "parity": parity,
            "config4": config4,
            "extras": extras,
            "checksum_logL_draw0": float(res[0, 0]),
        }
        if ref64 is not None and world == 1:
            g64, g32 = ref64[:, 1:], res[:, 1:]
            line["fp32_error_vs_fp64"] = {
                "logL_max_rel": float(np.max(np.abs(res[:, 0] - ref64[:, 0]) / np.abs(ref64[:, 0]))),
                "grad_max_abs_over_max_abs": float(np.max(np.abs(g32 - g64)) / np.max(np.abs(g64))),
                "grad_median_rel": float(np.median(np.abs(g32 - g64) / np.maximum(1e-300, np.abs(g64))))}
            line["roofline"]["note"] += "; fp32 mode moves half the bytes per entry, byte counts above are the fp64 ones"
        print(json.dumps(line), flush=True)
    if world > 1:
        flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
        dist.broadcast(flag, 0)
        ok = int(flag.item()) == 0
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: PARITY FAILED (see the parity block of the JSON line)")


def _param_stride(lik):
    # doubles per draw in the packed parameter block (see csrc/phylo_b200.cu create_common)
    nn, C = 2 * lik.S - 1, lik.C
    ntheta = 0 if lik.model == 0 else lik.nsubst + 4
    o = nn + 2 * C + 4 + 4 + 48 + 16 * ntheta
    return (o + 1) & ~1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k", type=int, default=0, help="patterns per thread (tuning)")
    ap.add_argument("--pb", type=int, default=0, help="pattern blocks per CTA (tuning)")
    ap.add_argument("--kernel-times", action="store_true", help="read per-kernel events every step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the 1-thread cpu_baseline sample (the parity oracle run stays)")
    ap.add_argument("--no-config4", action="store_true", help="skip the 10k x 1M strong-scaling block")
    ap.add_argument("--no-extras", action="store_true", help="skip the value-only / fp32-mode block (N = 1 only)")
    ap.add_argument("--fp32", action="store_true",
                    help="optional fp32-with-scaling mode (reported separately; NOT the headline, dtype f32)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
