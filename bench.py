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
    every internal partial (32 B) and of its rescale byte, tip codes once per sweep, weights once."""
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

def run_reference(args):
    """CPU oracle ("port") with all host threads; one step = value+gradient of ONE draw on a
    pattern slice sized for a few seconds of work, scaled to the metric's unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    cores = O.num_threads()
    prob, draws = build_problem()
    sample_L = min(L_PATTERNS, 2048 * cores)
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
            "config": {"workload": WORKLOAD, "taxa": S_TAXA, "patterns": L_PATTERNS, "categories": N_CAT,
                       "draws": N_DRAWS, "model": None},
            "tree_evals_per_s": value / (L_PATTERNS * N_CAT),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line["config"].pop("model")
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm

def cpu_baseline_sample(prob, draws):
    """Oracle on this box's host cores, ~10-20 s: 1 thread and all threads, one draw, pattern slice."""
    from oracle import oracle as O
    bl, rates, freqs, rs, ps = draws
    cores = O.num_threads()
    out = {}
    for label, nt, nL in (("1 thread", 1, 16384), ("all threads", cores, min(L_PATTERNS, 8192 * cores))):
        sl = slice(0, nL)
        O.loglik_grad(prob.peel, prob.tipmask[:, :64], prob.weights[:64], O.GTR, bl[0], rates[0], freqs[0], rs[0],
                      ps[0], dp_eigen=True, nthreads=nt)  # warm
        t0 = time.perf_counter()
        O.loglik_grad(prob.peel, prob.tipmask[:, sl], prob.weights[sl], O.GTR, bl[0], rates[0], freqs[0], rs[0], ps[0],
                      dp_eigen=True, nthreads=nt)
        dt = time.perf_counter() - t0
        out[label] = (nL * N_CAT / dt, nt, nL, dt)
    v, nt, nL, dt = out["all threads"]
    v1 = out["1 thread"][0]
    return {"value": v, "unit": UNIT, "cores": nt, "kind": "port",
            "sample": f"{nL} of {L_PATTERNS} patterns x 1 draw, value+gradient, {dt:.1f} s; "
                      f"1 thread on {out['1 thread'][2]} patterns: {v1:.4g} {UNIT}",
            "value_1thread": v1, "tree_evals_per_s": v / (L_PATTERNS * N_CAT)}


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
    res = lik.download(B)

    # ---- end-to-end timing through the host-buffer API
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None

    times = torch.tensor([dev_ms / 1e3, wall, e2e_wall], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, wall_s, e2e_s = (float(x) for x in times.cpu())
    info = lik.info()

    if rank == 0:
        Lg = L_PATTERNS * world
        units_per_step = B * Lg * N_CAT
        value = args.steps * units_per_step / dev_s
        e2e_value = args.steps * units_per_step / e2e_s
        sweep_ms = statistics.mean(kern["sweep_ms"])
        peak, peak_src = measured_peak_gbs()
        alg = algorithmic_bytes(S_TAXA, L_PATTERNS, N_CAT) * B          # per sweep launch on one GPU
        des = design_bytes(S_TAXA, L_PATTERNS, N_CAT) * B
        achieved = alg / (sweep_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "sweep_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_evaluation")
            traffic = traffic * B if traffic else None   # one launch sweeps B draws
        cpu = cpu_baseline_sample(prob, draws) if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "taxa": S_TAXA, "patterns_per_gpu": L_PATTERNS, "categories": N_CAT,
                       "draws_per_step": B, "parallelism": f"pattern-sharded x{world}, one NCCL all-reduce per step"
                       if world > 1 else "single GPU",
                       "l2": "per-step working set (2.4 GB partial scratch + 130 MB matrices) exceeds the 126 MB L2",
                       "tiling": {k: info[k] for k in ("stack_depth", "patterns_per_thread", "threads_per_cta", "grid",
                                                       "smem_bytes", "tiles")}},
            "tree_evals_per_s": value / (Lg * N_CAT),
            "node_updates_per_s": value * (S_TAXA - 1),
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "kernels_ms": {k: statistics.mean(v) for k, v in kern.items()},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "h2d_bytes_per_step": int(B * _param_stride(lik) * 8),
                    "d2h_bytes_per_step": int(B * lik.nout * 8)},
            "gpu_launches": int(args.steps * info["kernel_launches"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "sweep_kernel<K,GRAD>", "launch_ms": sweep_ms,
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                         "design_bytes_per_launch": des, "design_achieved": des / (sweep_ms * 1e-3) / 1e9,
                         "design_frac": des / (sweep_ms * 1e-3) / 1e9 / peak,
                         "note": "achieved uses SURVEY 8(d)'s level-synchronous byte count B_vg; the depth-first "
                                 "sweep keeps child reads and all q traffic in shared memory, so its own minimum "
                                 "DRAM traffic is design_bytes (~0.41 B_vg) and frac may exceed 1"},
            "cpu_baseline": cpu, "clocks": clocks,
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
    lik.close()
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
