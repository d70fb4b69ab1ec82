"""world_size-2 gloo test (CPU) of the batch-sharding host logic: the mean of the per-rank losses
equals the full-batch loss, per-image gradients match after accounting for the shard size, and
timings reduce as max-over-ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import torch_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monodepth2_jl_b200 import dist as D
    torch.set_num_threads(1)
    N, C, H, W = 4, 1, 24, 40
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=5, dtype=torch.float64)
    K, invK = O.make_K(W, H, dtype=torch.float64)
    xs, ds, rs, ts = D.shard_batch((x, disps, rv, tv), rank, world)
    ds = [d.clone().requires_grad_(True) for d in ds]
    loss = O.view_synthesis_loss(xs, ds, rs, ts, K, invK)
    loss.backward()
    gl = D.global_loss(loss, xs.shape[0])
    ms = D.max_over_ranks(10.0 + rank)
    shared = [torch.full((3,), float(rank + 1), dtype=torch.float64)]
    D.allreduce_mean_(shared)
    # (numpy arrays travel by value; torch tensors would be shared through the worker's resource sharer, which is
    # gone if the worker exits before the parent unpickles them)
    q.put((rank, gl.item(), ms, shared[0].tolist(), [d.grad.numpy().copy() for d in ds], D.shard_bounds(N, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    from monodepth2_jl_b200 import dist as D
    for n in (1, 7, 8, 12):
        for w in (1, 2, 3, 8):
            b = [D.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharding_matches_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, C, H, W = 4, 1, 24, 40
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=5, dtype=torch.float64)
    K, invK = O.make_K(W, H, dtype=torch.float64)
    dfull = [d.clone().requires_grad_(True) for d in disps]
    full = O.view_synthesis_loss(x, dfull, rv, tv, K, invK)
    full.backward()
    for rank, gl, ms, shared, grads, (lo, hi) in res:
        assert abs(gl - full.item()) < 1e-12          # mean of means == global mean (equal shards)
        assert ms == 11.0                              # max over ranks
        assert shared == [1.5, 1.5, 1.5]               # all-reduce average of a replicated parameter gradient
        for g, gf in zip(grads, dfull):
            # a rank's per-image gradient is for the mean over ITS images: 1/world of it is the global one
            assert torch.allclose(torch.from_numpy(g) / world, gf.grad[lo:hi], atol=1e-14)


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from monodepth2_jl_b200 import dist as D
    from monodepth2_jl_b200.train_step import GradientBuckets
    out = {}
    for overlap in (True, False, None):
        torch.manual_seed(0)                                    # replicated parameters
        net = torch.nn.Sequential(torch.nn.Linear(6, 33), torch.nn.Tanh(), torch.nn.Linear(33, 17), torch.nn.Tanh(), torch.nn.Linear(17, 1))
        unused = torch.nn.Parameter(torch.ones(5))              # a parameter that never gets a gradient
        net.register_parameter("unused", unused)
        gb = GradientBuckets(net, bucket_bytes=256, overlap=overlap)
        assert len(gb.buckets) >= 3
        for step in range(2):                                   # second step: counters re-armed, gradient zeroed
            torch.manual_seed(100 + rank + 10 * step)
            xin = torch.randn(4, 6)
            gb.arm()
            net(xin).pow(2).mean().backward()
            scale = gb.finish()
        out[str(overlap)] = (gb.grad.clone().numpy(), scale, gb.calls)
    # weighted mean for unequal shards: rank r holds r + 1 images whose mean gradient is (r + 1)
    g = [torch.full((2,), float(rank + 1), dtype=torch.float64)]
    D.allreduce_mean_(g, n_local=rank + 1)
    q.put((rank, out, g[0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_buckets_sum_over_ranks():
    """the one collective of the training step (train_step.GradientBuckets): overlapped (hook-driven, asynchronous),
    blocking, and none; buckets with a parameter that received no gradient are still reduced"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import numpy as np
    local = [res[r][1]["None"][0] for r in range(world)]                 # un-reduced local gradients
    for r in range(world):
        out = res[r][1]
        assert out["None"][1] == 1.0 and out["None"][2] == 0
        for mode in ("True", "False"):
            g, scale, calls = out[mode]
            assert scale == 0.5 and calls > 0
            assert np.allclose(g, local[0] + local[1], rtol=1e-6, atol=1e-9)
        assert np.array_equal(out["True"][0], out["False"][0])
        assert np.allclose(res[r][2], [(1 * 1 + 2 * 2) / 3.0] * 2, rtol=1e-14)  # n_local-weighted mean
    assert not np.allclose(local[0], local[1])
