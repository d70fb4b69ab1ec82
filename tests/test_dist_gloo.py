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
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
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
