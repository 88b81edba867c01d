"""Sharded optimizer state and sharded gradients for data-parallel fine-tuning (training row: the
reference runs DeepSpeed ZeRO-2, scripts/zero2.json). `ShardedAdamW.step` is the optimizer-state half (ZeRO-1:
gradients replicated by an all-reduce); `ShardedAdamW.step_from_local_grads` is the ZeRO-2 step: every gradient is
reduce-scattered (bf16 payload by default, as SURVEY 8e sizes it: 13.5 GB instead of 27 GB for the full fine-tune), so
a rank only ever holds the averaged gradient of the slice it updates, the clipping norm is assembled from the slices'
squared norms with one scalar all-reduce, and the update + bf16 all-gather are those of the ZeRO-1 step.

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


def reduce_scatter_mean(flat, rank, world, group, comm_dtype=torch.bfloat16):
    """flat: this rank's 1-D gradient of a whole tensor. Returns the fp32 mean over the ranks of the slice
    slice_range(n, rank, world) -- the only part this rank's optimizer shard needs. The payload crosses the wire in
    `comm_dtype` (bf16: half the bytes of the fp32 all-reduce; every element is reduced exactly once, for its owner, so
    replicas cannot diverge). NCCL: reduce_scatter_tensor on a buffer padded to world equal slices; gloo (CPU tests) has
    no reduce-scatter, so the padded buffer is all-reduced and cut, which yields the same sums."""
    n = flat.numel()
    lo, hi, s = slice_range(n, rank, world)
    buf = torch.zeros(s * world, dtype=comm_dtype, device=flat.device)
    buf[:n] = flat.reshape(-1)
    if flat.is_cuda:
        out = torch.empty(s, dtype=comm_dtype, device=flat.device)
        dist.reduce_scatter_tensor(out, buf, op=dist.ReduceOp.SUM, group=group)
    else:
        wire = buf.float() if comm_dtype == torch.bfloat16 else buf       # gloo: no bf16 arithmetic
        dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=group)
        out = wire[rank * s:(rank + 1) * s].to(comm_dtype)                # (exactly NCCL's bf16 sum for two ranks)
    return (out[:hi - lo].float() / world).contiguous()


class ShardedAdamW:
    def __init__(self, params, source, names, group, betas=(0.9, 0.999), eps=1e-8, adamw=None, sq_norm=None):
        """params: {name: full bf16 working tensor (contiguous; updated in place)}; source: {name: tensor} the fp32
        master slices are cut from; adamw: the update kernel (default: libb200mmor's b200_adamw_step through
        _lib.adamw_step; the CPU gloo test injects a torch restatement to exercise the partition / gather logic)."""
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.params, self.names = params, list(names)
        self.betas, self.eps = betas, eps
        self.adamw = adamw if adamw is not None else L.adamw_step
        # sq_norm(t, out2) adds sum(t^2) to out2[0] (default: b200_grad_sq_norm; the gloo test injects torch)
        self.sq_norm = sq_norm if sq_norm is not None else (
            lambda t, out2: L.grad_sq_norm(t, out2=out2, accumulate=True, max_norm=0.0))
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

    def step_from_local_grads(self, grads, step, lr_of, wd_of, max_norm=0.0, comm_dtype=torch.bfloat16):
        """ZeRO-2 step. grads: {name: THIS rank's gradient of the full tensor (any float dtype)}; the same names on
        every rank (dist.align_optional_gradients). Entries are popped as they are reduced, so the full-size
        gradients are released tensor by tensor. Returns out2 = [global squared gradient norm, clip coefficient]
        (torch.nn.utils.clip_grad_norm_ semantics, like b200_grad_sq_norm) -- identical on every rank."""
        todo = [k for k in self.names if k in grads]
        dev = self.params[self.names[0]].device
        slices = {}
        for k in todo:
            slices[k] = reduce_scatter_mean(grads.pop(k), self.rank, self.world, self.group, comm_dtype)
        out2 = torch.zeros(2, dtype=torch.float32, device=dev)
        for k in todo:
            if slices[k].numel():
                self.sq_norm(slices[k], out2)
        dist.all_reduce(out2[:1], op=dist.ReduceOp.SUM, group=self.group)       # slices partition every tensor
        out2[1] = torch.clamp(max_norm / (out2[0].sqrt() + 1e-6), max=1.0) if max_norm > 0 else 1.0
        self.step(slices, step, lr_of, wd_of, clip_coef=out2[1:], sliced=True)
        return out2

    def step(self, grads, step, lr_of, wd_of, clip_coef=None, sliced=False):
        """grads: {name: fp32 gradient of the full tensor, identical on every rank} or, with sliced=True,
        {name: this rank's slice of it} (step_from_local_grads). Names without a gradient are skipped (on every rank
        alike: see dist.align_optional_gradients)."""
        todo = [k for k in self.names if k in grads]
        for k in todo:
            lo, hi = self.range[k]
            if hi > lo:
                g = grads[k] if sliced else grads[k].contiguous().view(-1)[lo:hi]
                self.adamw(self.master[k], self.params[k].view(-1)[lo:hi], g, self.m[k], self.v[k], lr_of(k),
                           self.betas[0], self.betas[1], self.eps, wd_of(k), step, clip_coef=clip_coef)
        self.gather_params(todo)

    def refresh_from_masters(self):
        """bf16 working weights <- this rank's fp32 master slices, then the all-gather of a step (resume: after the
        slices were loaded from a checkpoint, train/checkpoint.py)."""
        for k in self.names:
            lo, hi = self.range[k]
            if hi > lo:
                self.params[k].view(-1)[lo:hi].copy_(self.master[k].to(self.params[k].dtype))
        self.gather_params(self.names)

    def gather_params(self, todo):
        """All-gather of the updated bf16 slices (in place when the tensor divides evenly, staged otherwise)."""
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
