"""Multi-GPU plumbing of the MM2SG inference path: one process per GPU, weights replicated, samples
sharded across ranks, one exchange step -- the all-gather of projected visual tokens before LLM fusion.

The reference has no multi-GPU inference path (its eval runs one process on one GPU, scene_graph_prediction/main.py:
57-60; only training is data parallel through HF Trainer / DeepSpeed). BASELINE.json's north_star defines the
partitioning: every rank encodes the views of its own samples (CLIP ViT + pooler + projector), the projected tokens
(B_local, T_vis, 4096) are all-gathered over NVLink, and every rank then decodes a slice of the gathered batch.

Two exchange implementations:
  * `all_gather_tokens`  -- torch.distributed all_gather_into_tensor (NCCL on GPUs; gloo in the CPU tests)
  * `PeerGather`         -- symmetric-memory buffer whose peer pointers are handed to the projector GEMM so that its
                            epilogue stores every output tile straight into all ranks' copies (b200_projector_pack
                            with `peers`): the transfer overlaps the GEMM tile by tile and no separate collective runs.
Host logic here is device agnostic so that world_size-2 gloo tests on CPU cover it.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block partition of n units over `world` ranks (first n % world ranks get one extra)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def decode_owner(rank, world, shift=0):
    """Which rank's encoded samples `rank` decodes. shift = 0: its own (the gathered copy is then only read locally);
    shift != 0 rotates the assignment, which is how the tests prove that the gathered tokens are the peers' tokens."""
    return (rank + shift) % world


def all_gather_tokens(local, group=None):
    """local (B_local, T, D), same shape on every rank -> (world, B_local, T, D) ordered by rank."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    else:  # gloo has no all_gather_into_tensor for every dtype (bf16): go through a list of views
        src = local.contiguous()
        if src.dtype == torch.bfloat16:
            raw = src.view(torch.uint8)
            parts = [torch.empty_like(raw) for _ in range(world)]
            dist.all_gather(parts, raw, group=group)
            for r, p in enumerate(parts):
                out[r] = p.view(torch.bfloat16)
        else:
            parts = [torch.empty_like(src) for _ in range(world)]
            dist.all_gather(parts, src, group=group)
            for r, p in enumerate(parts):
                out[r] = p
    return out


def all_gather_objects_equal(value, group=None):
    """True iff `value` (small python object) is identical on every rank; used to validate batch geometry."""
    world = dist.get_world_size(group)
    got = [None] * world
    dist.all_gather_object(got, value, group=group)
    return all(g == got[0] for g in got), got


def average_gradients(grads, names, group=None):
    """Sum every named gradient over the ranks (all_reduce, NCCL on GPUs / gloo in the CPU tests) and divide by the
    world size. Contiguous tensors are reduced in place (no extra copy of a 27 GB gradient set); a strided view is
    replaced by a reduced contiguous copy. Returns the dict restricted to `names`."""
    world = dist.get_world_size(group)
    out = {}
    for k in names:
        t = grads[k] if grads[k].is_contiguous() else grads[k].contiguous()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t /= world
        grads[k] = out[k] = t
    return out


def host_side_group(group):
    """A gloo group over the ranks of `group` for tiny host-side agreements (a few flags per step). Exchanging them
    through NCCL would put the all-reduce behind the whole backward on the stream and make the host wait for it --
    i.e. serialise the Python launch loop of the optimizer step with the GPU (measured: ~100 ms per fine-tune step)."""
    import os
    if dist.get_backend(group) == "gloo":
        return group
    if os.environ.get("MASTER_ADDR", "") in ("127.0.0.1", "localhost", "::1"):
        # one node: gloo must not try to resolve the container's hostname (it may not resolve): use the loopback device
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    return dist.new_group(ranks=dist.get_process_group_ranks(group), backend="gloo")


def align_optional_gradients(grads, optional, group=None, host_group=None):
    """Gradients of modality-specific parameters (audio projection, seg-mask CNN, embed_tokens) exist only on ranks
    whose local batch carried the modality. After this call every rank holds the same keys, so the per-key collectives
    of average_gradients line up: a gradient present on ANY rank exists everywhere (zeros where the local batch did not
    produce it -- what DDP does for parameters unused on a rank), one present nowhere stays absent everywhere (the
    optimizer skips it, like .grad = None). `optional`: {name: shape}. host_group: a gloo group over the same ranks
    (host_side_group) -- the flags then travel host to host and the call does not wait for the GPU."""
    names = sorted(optional)
    if not names:
        return grads
    dev = next(iter(grads.values())).device if grads else torch.device("cpu")
    if host_group is not None:      # host tensors over gloo: no device synchronisation (see host_side_group)
        flags = torch.tensor([1 if k in grads else 0 for k in names], dtype=torch.int32)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=host_group)
    else:
        flags = torch.tensor([1 if k in grads else 0 for k in names], dtype=torch.int32, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    for k, f in zip(names, flags.tolist()):
        if f and k not in grads:
            grads[k] = torch.zeros(tuple(optional[k]), dtype=torch.float32, device=dev)
    return grads


class PeerGather:
    """Symmetric bf16 buffer on every rank, viewed as (world, B_local, T, D), with every peer's base pointer, for the
    fused projector-GEMM + all-gather epilogue. Needs CUDA, NVLink peer access and torch symmetric memory.
    The allocation (a collective rendezvous) happens once per capacity: a call with another batch geometry that fits
    the capacity only re-views the same memory (`set_shape`), so successive workloads of different shapes -- the bench
    runs configs[1], then configs[2] (580 tokens x 64 samples), then configs[3] -- never allocate a second symmetric
    buffer next to a live one (which hung on 2 and 8 GPUs when it was tried, round 2)."""

    def __init__(self, b_local, t_vis, hidden, group=None, capacity=0):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.capacity = max(int(capacity), self.world * b_local * t_vis * hidden)
        self.flat = symm.empty((self.capacity,), dtype=torch.bfloat16,
                               device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm.rendezvous(self.flat, self.group)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        if len(self.peer_ptrs) != self.world:
            raise RuntimeError("symmetric memory rendezvous returned %d peers for world %d"
                               % (len(self.peer_ptrs), self.world))
        self.set_shape(b_local, t_vis, hidden)

    def fits(self, b_local, t_vis, hidden):
        return self.world * b_local * t_vis * hidden <= self.capacity

    def set_shape(self, b_local, t_vis, hidden):
        """Re-view the buffer for another batch geometry (same on every rank); no allocation, no collective."""
        if not self.fits(b_local, t_vis, hidden):
            raise ValueError("PeerGather: geometry exceeds the allocated capacity")
        self.shape = (self.world, b_local, t_vis, hidden)
        self.buf = self.flat[:self.world * b_local * t_vis * hidden].view(self.shape)

    def slot_offset_bytes(self):
        """Byte offset of this rank's (B_local, T, D) slot inside every peer's buffer."""
        _, b, t, d = self.shape
        return self.rank * b * t * d * 2

    def barrier(self):
        """All peers' stores into this rank's buffer have landed (device-side signal pads) -- call after the GEMM."""
        self.hdl.barrier()
