"""Sharded optimizer state for data-parallel fine-tuning (training row: the reference runs DeepSpeed
ZeRO-2, scripts/zero2.json; this is the optimizer-state half of it, ZeRO-1).

Every rank keeps fp32 master weights, m and v for ONE contiguous slice of every trainable tensor (1/world of the 12
bytes per parameter: 81.6 GB -> 10.2 GB per rank for the 6.8 B-parameter full fine-tune on 8 GPUs). A step is:
gradients averaged over the ranks (dist.average_gradients, replicated), fused AdamW (b200_adamw_step) on the local
slice writing the bf16 working weights in place, then one all-gather per tensor of the updated bf16 slices (in place
when the tensor divides evenly, through a padded staging buffer otherwise). Replicas stay bit-identical: every element
is updated by exactly one rank with the arithmetic of the unsharded step.
"""
import torch
import torch.distributed as dist

from .. import _lib as L


def slice_range(n, rank, world):
    """Contiguous slice [lo, hi) of a flat tensor of n elements owned by `rank`; slices are ceil(n / world) long."""
    s = -(-n // world)
    lo = min(rank * s, n)
    return lo, min(lo + s, n), s


class ShardedAdamW:
    def __init__(self, params, source, names, group, betas=(0.9, 0.999), eps=1e-8, adamw=None):
        """params: {name: full bf16 working tensor (contiguous; updated in place)}; source: {name: tensor} the fp32
        master slices are cut from; adamw: the update kernel (default: libb200mmor's b200_adamw_step through
        _lib.adamw_step; the CPU gloo test injects a torch restatement to exercise the partition / gather logic)."""
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.params, self.names = params, list(names)
        self.betas, self.eps = betas, eps
        self.adamw = adamw if adamw is not None else L.adamw_step
        self.range, self.master, self.m, self.v = {}, {}, {}, {}
        for k in self.names:
            p = params[k]
            if not p.is_contiguous():
                raise ValueError(f"{k}: working weights must be contiguous")
            lo, hi, _ = slice_range(p.numel(), self.rank, self.world)
            self.range[k] = (lo, hi)
            self.master[k] = source[k].detach().reshape(-1)[lo:hi].to(p.device, torch.float32).clone()
            self.m[k] = torch.zeros_like(self.master[k])
            self.v[k] = torch.zeros_like(self.master[k])

    def state_bytes(self):
        return sum(3 * 4 * t.numel() for t in self.master.values())

    def step(self, grads, step, lr_of, wd_of, clip_coef=None):
        """grads: {name: fp32 gradient of the full tensor, identical on every rank}. Names without a gradient are
        skipped (on every rank alike: see dist.align_optional_gradients)."""
        todo = [k for k in self.names if k in grads]
        for k in todo:
            lo, hi = self.range[k]
            if hi > lo:
                g = grads[k].contiguous().view(-1)[lo:hi]
                self.adamw(self.master[k], self.params[k].view(-1)[lo:hi], g, self.m[k], self.v[k], lr_of(k),
                           self.betas[0], self.betas[1], self.eps, wd_of(k), step, clip_coef=clip_coef)
        stage_in = stage_out = None
        for k in todo:
            flat = self.params[k].view(-1)
            n = flat.numel()
            lo, hi, s = slice_range(n, self.rank, self.world)
            if s * self.world == n:
                dist.all_gather_into_tensor(flat, flat[lo:hi], group=self.group)
                continue
            if stage_in is None or stage_in.numel() < s:
                stage_in = torch.zeros(s, dtype=flat.dtype, device=flat.device)
                stage_out = torch.empty(s * self.world, dtype=flat.dtype, device=flat.device)
            inp, out = stage_in[:s], stage_out[:s * self.world]
            inp.zero_()
            inp[:hi - lo] = flat[lo:hi]
            dist.all_gather_into_tensor(out, inp, group=self.group)
            flat.copy_(out[:n])
