"""Batched ADVI over pattern shards: every rank (one per GPU) holds the full tree and its slice of the patterns
and runs the same driver; each likelihood call ends in one NCCL all-reduce of the [B, nout] device block.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/advi_sharded.py [taxa] [patterns_per_gpu] [grad_samples] [iterations]
Weak scaling: the alignment has N x patterns_per_gpu patterns."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from phylostan_b200 import advi, encode as E, likelihood as lk, sharded, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
Lg = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
gs = int(sys.argv[3]) if len(sys.argv) > 3 else 64
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
prob = synth.make_problem(S, Lg * world, 4, structured=False)      # same seed on every rank: same alignment
lo, hi = sharded.shard_bounds(Lg * world, world, rank)
stream = torch.cuda.Stream(device=local_rank)
torch.cuda.set_stream(stream)
lik = lk.TreeLikelihood(E.unrooted_swap(prob.peel), prob.tipmask[:, lo:hi], prob.weights[lo:hi], model="GTR",
                        categories=4, rooted=False, device=local_rank)
lik.set_stream(stream.cuda_stream)
m = advi.UnrootedModel(sharded.ShardedLikelihood(lik), "GTR")
z0 = np.zeros(m.dim)
z0[m.slices["blens"]] = np.log(0.02)
t0 = time.perf_counter()
fit = advi.advi(m, iter=iters, grad_samples=gs, elbo_samples=gs, eval_elbo=max(iters // 2, 1), eta=0.1, seed=1, init=z0,
                output_samples=0)
dt = time.perf_counter() - t0
if world > 1:
    mus = [torch.zeros(m.dim, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(mus, torch.as_tensor(fit.mu, device="cuda"))
    same = all(bool(torch.equal(mus[0], x)) for x in mus)
else:
    same = True
if rank == 0:
    print(f"{world} GPU(s), {S} taxa x {Lg * world} patterns, dim {m.dim}: {fit.likelihood_calls} library calls "
          f"({fit.likelihood_draws} draws) in {dt:.1f} s = {dt / fit.likelihood_calls * 1e3:.0f} ms per call; ELBO "
          f"{fit.elbo_trace[0][1]:.1f} -> {fit.elbo_trace[-1][1]:.1f}; identical variational mean on every rank: {same}")
lik.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
