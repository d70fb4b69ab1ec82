"""Row F1 of the path: the training step around the loss (src/Monodepth.jl:156-176) --

    x = device(x); [auto_loss = automasking_loss(...)]; grads = gradient(theta) do train_loss(model, x, ...) end;
    Flux.Optimise.update!(ADAM(1e-4), theta, grads)

-- as a data-parallel step, one process per GPU: every rank runs the model and the fused loss kernels on its shard
of the batch, the PARAMETER gradients (nothing else) are summed over the ranks with NCCL in buckets that start while
the backward pass is still producing the earlier layers' gradients, and one fused ADAM launch (md2_adam_step) applies
the mean.  The ResNet-18 encoder / depth decoder / pose decoder are NOT part of this library (`north_star`: they stay
on the host framework's layers): `StandInModel` builds them from the host framework's own conv layers (torch.nn ->
cuDNN) with the reference's architecture (src/model.jl, src/depth_decoder.jl, src/pose_decoder.jl) and random
weights, so that the step has the real shapes, parameter count (the all-reduce payload) and call pattern.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .ops import SSIM, Backproject, Project
from .training import Adam, AsyncViz, Params, Pose, TrainCache, train_loss

_F32 = torch.float32


# ---------------------------------------------------------------------------------------------------------------
# stand-in for the reference's Model (host-framework layers; architecture of src/model.jl)
# ---------------------------------------------------------------------------------------------------------------
class _BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False); self.b1 = nn.BatchNorm2d(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False); self.b2 = nn.BatchNorm2d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.b1(self.c1(x)))
        y = self.b2(self.c2(y))
        return F.relu(y + (x if self.down is None else self.down(x)))


class _ResNet18Stages(nn.Module):
    """ResidualNetwork(18; classes=nothing) evaluated with Val(:stages): the five feature maps (64, 64, 128, 256, 512
    channels at 1/2 ... 1/32 resolution)"""
    stages = (64, 64, 128, 256, 512)

    def __init__(self, in_channels=3):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU())
        self.pool = nn.MaxPool2d(3, 2, 1)
        chans, layers, cin = (64, 128, 256, 512), [], 64
        for i, c in enumerate(chans):
            layers.append(nn.Sequential(_BasicBlock(cin, c, 1 if i == 0 else 2), _BasicBlock(c, c, 1)))
            cin = c
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        f = [self.stem(x)]
        y = self.pool(f[0])
        for layer in self.layers:
            y = layer(y)
            f.append(y)
        return f


class _DecoderBlock(nn.Module):        # conv 3x3 over pad_reflect(x, 1)   (src/depth_decoder.jl:1-5)
    def __init__(self, cin, cout, act):
        super().__init__()
        self.conv, self.act = nn.Conv2d(cin, cout, 3, padding=1, padding_mode="reflect"), act

    def forward(self, x):
        return self.act(self.conv(x))


class _BranchBlock(nn.Module):         # src/depth_decoder.jl:7-19
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.c1, self.c2 = _DecoderBlock(cin, cout, F.elu), _DecoderBlock(cout + cskip, cout, F.elu)

    def forward(self, x, skip):
        y = F.interpolate(self.c1(x), scale_factor=2, mode="bilinear", align_corners=True)
        return self.c2(y if skip is None else torch.cat([y, skip], dim=1))


class _DepthDecoder(nn.Module):        # src/depth_decoder.jl:21-66, scale_levels 2:5 -> disparities at 1/8, 1/4, 1/2, 1
    def __init__(self, encoder_channels, scale_levels=(2, 3, 4, 5)):
        super().__init__()
        dec = [256, 128, 64, 32, 16]
        enc = list(encoder_channels)[::-1]
        cin = [enc[0]] + dec[:-1]
        cskip = enc[1:] + [0]
        self.branches, self.heads, b0 = nn.ModuleList(), nn.ModuleList(), 0
        for lv in scale_levels:
            self.branches.append(nn.ModuleList([_BranchBlock(cin[b], cskip[b], dec[b]) for b in range(b0, lv)]))
            self.heads.append(_DecoderBlock(dec[lv - 1], 1, torch.sigmoid))
            b0 = lv

    def forward(self, feats):
        x, skips, out, b = feats[-1], feats[-2::-1], [], 0
        for branch, head in zip(self.branches, self.heads):
            for blk in branch:
                x = blk(x, skips[b] if b < len(skips) else None)
                b += 1
            out.append(head(x))
        return out


class _PoseDecoder(nn.Module):         # src/pose_decoder.jl:7-33
    def __init__(self, cenc):
        super().__init__()
        self.squeezer = nn.Conv2d(cenc, 256, 1)
        self.pose = nn.Sequential(nn.Conv2d(512, 256, 3, padding=1), nn.ReLU(), nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(), nn.Conv2d(256, 6, 1))

    def forward(self, fa, fb):
        sq = torch.cat([F.relu(self.squeezer(fa)), F.relu(self.squeezer(fb))], dim=1)
        p = 1e-2 * self.pose(sq).mean(dim=(2, 3))
        return Pose(p[:, :3].contiguous(), p[:, 3:].contiguous())


class StandInModel(nn.Module):
    """`model(x, source_ids, target_id) -> (disparities, poses)` like the reference's Model (src/model.jl:8-20):
    x (N,L,C,H,W); all L frames go through the encoder at once; the depth decoder sees the target frame's features,
    the pose decoder the (earlier, later) frame pair of every source."""

    def __init__(self, in_channels=3):
        super().__init__()
        self.encoder = _ResNet18Stages(in_channels)
        self.depth_decoder = _DepthDecoder(self.encoder.stages)
        self.pose_decoder = _PoseDecoder(self.encoder.stages[-1])

    def forward(self, x, source_ids, target_id):
        N, L, C, H, W = x.shape
        if H % 32 or W % 32:
            raise ValueError("the ResNet-18 encoder / U-Net decoder pair needs image sides that are multiples of 32")
        feats = [f.reshape(N, L, *f.shape[1:]) for f in self.encoder(x.reshape(N * L, C, H, W))]
        disparities = self.depth_decoder([f[:, target_id] for f in feats])
        last = feats[-1]
        poses = [self.pose_decoder(last[:, i], last[:, target_id]) if i < target_id else self.pose_decoder(last[:, target_id], last[:, i])
                 for i in source_ids]
        return disparities, poses


# ---------------------------------------------------------------------------------------------------------------
# flat parameter / gradient storage, bucketed gradient all-reduce overlapped with the backward pass
# ---------------------------------------------------------------------------------------------------------------
class FlatParameters:
    """All parameters of a module as views of ONE contiguous buffer, their gradients as views of another, in reverse
    registration order (the order in which the backward pass finishes them), so that a bucket of gradients is one
    contiguous slice: one NCCL call per bucket and one fused ADAM launch over the whole model."""

    def __init__(self, module, bucket_bytes=8 << 20):
        params = [p for p in module.parameters() if p.requires_grad][::-1]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev = params[0].device
        pad = lambda n: (n + 3) & ~3                                         # 16-byte aligned views (float4 ADAM path)
        offs, total = [], 0
        for p in params:
            offs.append(total); total += pad(p.numel())
        self.data = torch.zeros(total, device=dev, dtype=_F32)
        self.grad = torch.zeros(total, device=dev, dtype=_F32)
        self.params = params
        with torch.no_grad():
            for p, o in zip(params, offs):
                self.data[o:o + p.numel()].copy_(p.reshape(-1))
                p.data = self.data[o:o + p.numel()].view_as(p)
                p.grad = self.grad[o:o + p.numel()].view_as(p)
        # buckets: consecutive parameters up to bucket_bytes
        self.buckets, self.bucket_of, lo, cur = [], {}, 0, 0
        for k, (p, o) in enumerate(zip(params, offs)):
            self.bucket_of[p] = len(self.buckets)
            cur += 1
            end = o + pad(p.numel())
            if (end - lo) * 4 >= bucket_bytes or k == len(params) - 1:
                self.buckets.append((lo, end, cur)); lo, cur = end, 0
        self.total = total

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:                                                # (a backward that replaced .grad is re-bound)
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or p.grad.data_ptr() >= self.grad.data_ptr() + 4 * self.total:
                raise RuntimeError("a parameter gradient left the flat buffer (do not call zero_grad(set_to_none=True))")


class GradientBuckets(FlatParameters):
    """The one collective of a training step: SUM all-reduce of the parameter gradients over the ranks, one call per
    bucket, started from autograd's post-accumulate hooks while the backward pass is still running.  Pure
    torch.distributed (NCCL on GPUs; the host-side logic is tested with gloo on CPU, tests/test_dist_gloo.py).

    overlap: True  -- a bucket is reduced asynchronously as soon as the backward pass has produced all of it
             False -- one blocking pass over the buckets after backward (the exposed-communication baseline)
             None  -- no all-reduce at all (what a single process does; used to measure the exposed time)"""

    def __init__(self, module, bucket_bytes=8 << 20, overlap=True):
        super().__init__(module, bucket_bytes)
        self.overlap = overlap
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self._pending, self._handles, self._armed = [0] * len(self.buckets), [], False
        self.calls = 0                       # all-reduce calls issued so far
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._on_grad)

    def _reduce(self, b, async_op):
        lo, hi, _ = self.buckets[b]
        self.calls += 1
        return dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, async_op=async_op)

    def _on_grad(self, p):
        if self.world == 1 or not self.overlap or not self._armed:
            return
        b = self.bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._handles.append(self._reduce(b, True))

    def arm(self):
        """before backward: zero the flat gradient, reset the per-bucket counters"""
        self.zero_grad()
        self._pending = [n for (_, _, n) in self.buckets]
        self._handles = []
        self._armed = True

    def finish(self):
        """after backward: every bucket holds the SUM over the ranks when this returns (on the current stream).
        Returns the factor that turns the sum into the mean (the optimiser folds it into its update)."""
        self._armed = False
        if self.world == 1 or self.overlap is None:
            return 1.0
        if self.overlap:
            for b, n in enumerate(self._pending):      # parameters that received no gradient this step
                if n:
                    self._handles.append(self._reduce(b, True))
            for h in self._handles:
                h.wait()
        else:
            for b in range(len(self.buckets)):
                self._reduce(b, False)
        return 1.0 / self.world


class DataParallelTrainer:
    """One training step of the reference's loop (src/Monodepth.jl:145-176) per `step(x)`; with torch.distributed
    initialised it is the batch-sharded data-parallel step: x is this rank's shard (equal shards on all ranks), the
    parameter gradients are all-reduced in buckets overlapped with backward (GradientBuckets) and the mean is folded
    into the fused ADAM launch (`grad_scale`)."""

    def __init__(self, model, cache: TrainCache, params: Params, lr=1e-4, bucket_bytes=8 << 20, overlap=True):
        self.model, self.cache, self.hp = model, cache, params
        self.flat = GradientBuckets(model, bucket_bytes, overlap)
        self.opt = Adam([self.flat.data], lr=lr)
        self.steps = 0
        self.viz = AsyncViz(self.flat.data.device)

    def step(self, x, do_visualization=False):
        """x (n_local,L,C,H,W) on this rank's device.  Returns (loss [device scalar], ticket): ticket is None or an
        AsyncViz ticket whose .get() gives train_loss's (vis_disparity, vis_warped, vis_loss) host copies -- the copy
        runs on a side stream and is normally collected one step later, when the log is written."""
        c, hp = self.cache, self.hp
        self.flat.arm()
        # (auto_loss = None with parameters.automasking: the fused call forms the automask map itself, src/Monodepth.jl:159-164)
        loss, vd, vw, vl = train_loss(self.model, x, None, c, hp, do_visualization, viz=self.viz if do_visualization else None)
        loss.backward()
        scale = self.flat.finish()
        self.opt.step([self.flat.grad], grad_scale=scale)
        self.steps += 1
        return loss.detach(), (vd if do_visualization else None)

    # -- checkpoint / resume (the reference dumps the model every 500 steps, src/Monodepth.jl:189-192) ------
    def state_dict(self):
        return dict(model=self.model.state_dict(), optimizer=self.opt.state_dict(), steps=self.steps)

    def load_state_dict(self, sd):
        with torch.no_grad():
            self.model.load_state_dict(sd["model"])       # (copies into the flat views)
        self.opt.load_state_dict(sd["optimizer"])
        self.steps = int(sd["steps"])

    def save_checkpoint(self, path):
        torch.save(self.state_dict(), path)

    def load_checkpoint(self, path):
        self.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))


def make_training_setup(W, H, device, *, channels=3, batch_size=8, automasking=False, lr=1e-4, overlap=True, seed=0,
                        bucket_bytes=8 << 20, disparity_smoothness=1e-3):
    """what the reference's train() builds before its loop (src/Monodepth.jl:100-131), with the stand-in model:
    returns (trainer, model, cache, parameters).  Every rank seeds the model identically (replicated parameters)."""
    from .synthetic import make_K
    torch.manual_seed(seed)
    dev = torch.device(device)
    model = StandInModel(channels).to(dev).train()
    K, invK = make_K(W, H)
    hp = Params(target_size=(W, H), batch_size=batch_size, disparity_smoothness=disparity_smoothness, automasking=automasking)
    scales = [1.0 / 2.0 ** (5 - level) for level in (2, 3, 4, 5)]                       # src/Monodepth.jl:106-107
    cache = TrainCache(SSIM(), Backproject(W, H), Project(W, H), K.to(dev), invK.to(dev), 1, (0, 2), scales)
    return DataParallelTrainer(model, cache, hp, lr=lr, bucket_bytes=bucket_bytes, overlap=overlap), model, cache, hp
