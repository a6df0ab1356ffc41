"""N > 1 host logic on CPU: world_size-2 gloo run of the pattern-sharded evaluation with the oracle
standing in for each rank's GPU, checked against the unsharded oracle."""
import os
import socket

import numpy as np
import pytest

from phylostan_b200 import sharded


def test_shard_bounds_cover_exactly():
    for L in (1, 7, 238, 100_000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [sharded.shard_bounds(L, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == L
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from conftest import load_dataset
    from oracle import oracle as O
    from phylostan_b200 import encode as E
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = load_dataset("DS1")
    S, L = d["tipmask"].shape
    lo, hi = sharded.shard_bounds(L, world, rank)
    rng = np.random.default_rng(3)
    B = 3
    bl = rng.exponential(0.05, (B, 2 * S - 3)) + 1e-4
    rates, freqs = rng.dirichlet(np.ones(6), B), rng.dirichlet(np.ones(4) * 5, B)
    rs, ps = np.stack([E.weibull_rates(0.4 + 0.2 * i, 4) for i in range(B)]), rng.dirichlet(np.ones(4) * 3, B)

    def local(bl, rates, freqs, rs, ps, sl=slice(lo, hi)):
        rows = []
        for i in range(len(bl)):
            r = O.loglik_grad(d["peel"], d["tipmask"][:, sl], d["weights"][sl], O.GTR, bl[i], rates[i], freqs[i],
                              rs[i], ps[i], rooted=False, nthreads=1)
            rows.append(r.flat())
        return np.stack(rows)

    got = sharded.ShardedLikelihood(local).packed(bl, rates, freqs, rs, ps)
    want = local(bl, rates, freqs, rs, ps, slice(0, L))
    q.put((rank, float(np.max(np.abs(got - want) / np.maximum(1, np.abs(want)))), got.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_unsharded():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, shape in res:
        assert err < 1e-10, (rank, err)
        assert shape == (3, 1 + 51 + 6 + 4 + 4 + 4)


def _advi_worker(rank, world, port, q):
    """Every rank runs the same ADVI driver on its pattern shard; the all-reduce makes them agree."""
    import torch.distributed as dist
    from conftest import load_dataset
    from oracle import oracle as O
    from phylostan_b200 import advi
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = load_dataset("DS1")
    S, L = d["tipmask"].shape
    L = 120                                                    # a slice keeps the oracle quick
    lo, hi = sharded.shard_bounds(L, world, rank)

    def local(bl, rates, freqs, rs, ps, sl=slice(lo, hi)):
        return np.stack([O.loglik_grad(d["peel"], d["tipmask"][:, sl], d["weights"][sl], O.GTR, bl[i], rates[i], freqs[i],
                                       rs[i], ps[i], rooted=False, nthreads=1).flat() for i in range(len(bl))])

    lik = sharded.ShardedLikelihood(local, layout=(2 * S - 3, 6, 4))
    m = advi.UnrootedModel(lik, "GTR")
    z0 = np.full(m.dim, -2.5)
    fit = advi.advi(m, iter=12, grad_samples=2, elbo_samples=4, eval_elbo=6, eta=0.1, seed=5, init=z0, output_samples=0)
    # the same run on the unsharded alignment, in this process
    whole = sharded.ShardedLikelihood.__new__(sharded.ShardedLikelihood)
    whole.local, whole.world, whole._is_gpu = (lambda *a: local(*a, sl=slice(0, L))), 1, False
    whole.bcount, whole.nsubst, whole.C = 2 * S - 3, 6, 4
    ref = advi.advi(advi.UnrootedModel(whole, "GTR"), iter=12, grad_samples=2, elbo_samples=4, eval_elbo=6, eta=0.1, seed=5,
                    init=z0, output_samples=0)
    q.put((rank, fit.mu, float(np.max(np.abs(fit.mu - ref.mu))), fit.elbo_trace[-1][1], ref.elbo_trace[-1][1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_advi_over_pattern_shards():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_advi_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1])                  # both ranks hold the same variational mean
    for rank, mu, err, e_sharded, e_whole in res:
        assert err < 1e-8 and abs(e_sharded - e_whole) < 1e-6 * abs(e_whole), (rank, err, e_sharded, e_whole)


def _gpu_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from conftest import load_dataset
    from phylostan_b200 import encode as E
    from phylostan_b200 import likelihood as lk
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = load_dataset("fluA")
    S, L = d["tipmask"].shape
    lo, hi = sharded.shard_bounds(L, world, rank)
    rng = np.random.default_rng(3)
    B = 4
    bl = rng.exponential(0.05, (B, 2 * S - 2)) + 1e-4
    rates, freqs = rng.dirichlet(np.ones(6), B), rng.dirichlet(np.ones(4) * 5, B)
    rs, ps = np.stack([E.weibull_rates(0.4 + 0.2 * i, 4) for i in range(B)]), rng.dirichlet(np.ones(4) * 3, B)
    # no set_stream here: ShardedLikelihood itself must put the kernels and the all-reduce on one stream
    lik = lk.TreeLikelihood(d["peel"], d["tipmask"][:, lo:hi], d["weights"][lo:hi], model="GTR", categories=4, device=rank)
    sh = sharded.ShardedLikelihood(lik)
    got = sh.packed(bl, rates, freqs, rs, ps)
    for _ in range(20):  # a race between the sweep and the collective would show up as a changing result
        again = sh.packed(bl, rates, freqs, rs, ps)
        assert np.array_equal(again[:, 0], got[:, 0]) or np.allclose(again, got, rtol=1e-12, atol=1e-12)
    q.put((rank, got))
    dist.barrier()
    lik.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_nccl_allreduce_matches_single_gpu(datasets):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from phylostan_b200 import encode as E
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = datasets["fluA"]
    S = d["tipmask"].shape[0]
    rng = np.random.default_rng(3)
    B = 4
    bl = rng.exponential(0.05, (B, 2 * S - 2)) + 1e-4
    rates, freqs = rng.dirichlet(np.ones(6), B), rng.dirichlet(np.ones(4) * 5, B)
    rs, ps = np.stack([E.weibull_rates(0.4 + 0.2 * i, 4) for i in range(B)]), rng.dirichlet(np.ones(4) * 3, B)
    for i in range(B):
        want = O.loglik_grad(d["peel"], d["tipmask"], d["weights"], O.GTR, bl[i], rates[i], freqs[i], rs[i], ps[i]).flat()
        for rank in (0, 1):
            err = np.abs(res[rank][i] - want) / np.maximum(1, np.abs(want))
            assert abs(res[rank][i][0] - want[0]) <= 1e-10 * abs(want[0]) and err.max() <= 1e-8
