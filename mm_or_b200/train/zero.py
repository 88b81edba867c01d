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
from .trace import phase


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
        every rank (dist.align_optional_gradients). Entries are popped as they are packed, so the full-size
        gradients are released tensor by tensor. The whole gradient set crosses the wire in ONE reduce-scatter and the
        updated slices come back in ONE all-gather (flat (world, S) buffers; round 1 issued two collectives and five
        staging kernels per tensor -- 400 tensors -- and spent 500 ms of an 815 ms step there on 2 GPUs). Returns out2 = [global squared gradient norm, clip coefficient]
        (torch.nn.utils.clip_grad_norm_ semantics, like b200_grad_sq_norm) -- identical on every rank."""
        todo = [k for k in self.names if k in grads]
        dev = self.params[self.names[0]].device
        lay = self._flat_layout(tuple(todo), comm_dtype, dev)
        W, S = self.world, lay["S"]
        send = lay["send"].view(W, S)
        # pack: row r of the send buffer = rank r's slice of every tensor, side by side (one strided copy per tensor,
        # which is also the fp32 -> wire-dtype conversion); the zero padding of ragged tensors is never overwritten
        with phase("zero2.pack"):
            for k in todo:
                self._scatter_rows(send, grads.pop(k).reshape(-1), *lay["at"][k])
        # ONE reduce-scatter for the whole gradient set: this rank receives the sums of exactly the slices it updates
        recv = lay["recv"]
        with phase("zero2.reduce_scatter"):
            if dev.type == "cuda":
                dist.reduce_scatter_tensor(recv, lay["send"], op=dist.ReduceOp.SUM, group=self.group)
            else:   # gloo (CPU tests) has no reduce-scatter: all-reduce the buffer and cut -- the same sums
                wire = lay["send"].float() if comm_dtype == torch.bfloat16 else lay["send"]   # gloo: no bf16 arithmetic
                dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group)
                recv.copy_(wire.view(W, S)[self.rank].to(comm_dtype))   # (exactly NCCL's bf16 sum for two ranks)
        # The mean over the ranks is never materialised: recv holds the SUMS in the wire dtype; the squared norm of the
        # mean is sum(recv^2) / W^2 and AdamW multiplies every gradient by clip / W (one factor, applied in fp32 inside
        # the kernel). For W a power of two this is bit-identical to dividing first.
        with phase("zero2.mean_norm"):
            out2 = torch.zeros(2, dtype=torch.float32, device=dev)
            self.sq_norm(recv, out2)                                     # the padding is zero: it adds nothing
            dist.all_reduce(out2[:1], op=dist.ReduceOp.SUM, group=self.group)   # slices partition every tensor
            out2[0] /= float(W * W)
            out2[1] = torch.clamp(max_norm / (out2[0].sqrt() + 1e-6), max=1.0) if max_norm > 0 else 1.0
            scale = (out2[1:] / float(W)).contiguous()                   # clip coefficient x 1 / W
        # AdamW on this rank's slices; the updated bf16 slices are written straight into the all-gather send buffer
        ag_send = lay["ag_send"]
        with phase("zero2.adamw"):
            for k in todo:
                off, s, n = lay["at"][k]
                lo, hi = self.range[k]
                if hi > lo:
                    self.adamw(self.master[k], ag_send[off:off + hi - lo], recv[off:off + hi - lo], self.m[k], self.v[k],
                               lr_of(k), self.betas[0], self.betas[1], self.eps, wd_of(k), step, clip_coef=scale)
        # ONE all-gather of the updated slices (the reduce-scatter's send buffer is free again: it receives them)
        gathered = lay["send"] if lay["send"].dtype == ag_send.dtype else lay["ag_recv"]
        with phase("zero2.all_gather"):
            if dev.type == "cuda":
                dist.all_gather_into_tensor(gathered, ag_send, group=self.group)
            else:
                raw = ag_send.view(torch.uint8)
                parts = [torch.empty_like(raw) for _ in range(W)]
                dist.all_gather(parts, raw, group=self.group)
                for r, part in enumerate(parts):
                    gathered.view(W, S)[r] = part.view(ag_send.dtype)
        g2 = gathered.view(W, S)
        with phase("zero2.unpack"):
            for k in todo:
                self._gather_rows(self.params[k].view(-1), g2, *lay["at"][k])
        # (the gathered buffer doubles as the next step's send buffer: its padding is still zero, because every rank's
        # ag_send is zero wherever AdamW does not write)
        return out2

    # ---- flat layout of a gradient set -----------------------------------------------------------------------------
    def _flat_layout(self, todo, comm_dtype, dev):
        """Column offsets of every tensor's slice inside a (world, S) buffer, S = sum of the slice lengths; buffers are
        allocated once per (set of tensors, wire dtype) and reused by every step."""
        key = (todo, comm_dtype)
        cache = self.__dict__.setdefault("_layouts", {})
        if key in cache:
            return cache[key]
        cache.clear()                       # one gradient set at a time: do not hoard 13 GB buffers of stale sets
        at, off = {}, 0
        pdt = self.params[todo[0]].dtype
        for k in todo:
            n = self.params[k].numel()
            s = -(-n // self.world)
            at[k] = (off, s, n)
            off += s
        S = (off + 7) // 8 * 8              # 16-byte aligned rows
        lay = dict(at=at, S=S, send=torch.zeros(self.world * S, dtype=comm_dtype, device=dev),
                   recv=torch.empty(S, dtype=comm_dtype, device=dev), ag_send=torch.zeros(S, dtype=pdt, device=dev))
        if comm_dtype != pdt:
            lay["ag_recv"] = torch.empty(self.world * S, dtype=pdt, device=dev)
        cache[key] = lay
        return lay

    def _scatter_rows(self, buf2d, flat, off, s, n):
        """buf2d[r, off:off + s] <- flat[r * s:(r + 1) * s] for every rank r (the last rows of a ragged tensor are short
        or empty; what they leave untouched is zero)."""
        full = n // s
        if full:
            buf2d[:full, off:off + s].copy_(flat[:full * s].view(full, s))
        rem = n - full * s
        if rem:
            buf2d[full, off:off + rem].copy_(flat[full * s:])

    def _gather_rows(self, flat, buf2d, off, s, n):
        full = n // s
        if full:
            flat[:full * s].view(full, s).copy_(buf2d[:full, off:off + s])
        rem = n - full * s
        if rem:
            flat[full * s:].copy_(buf2d[full, off:off + rem])

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
