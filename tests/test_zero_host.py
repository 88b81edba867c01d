"""world_size-2 gloo test (CPU) of the sharded optimizer's host logic (mm_or_b200/train/zero.py): slice ownership,
the in-place and the staged all-gather of updated bf16 slices, skipped parameters. The update arithmetic itself is the
CUDA kernel b200_adamw_step (parity vs torch.optim.AdamW: tests/test_gpu_train_ops.py); here a torch restatement of
that kernel is injected so that the partition / gather logic can run without a GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mm_or_b200.train.zero import ShardedAdamW, slice_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torch_adamw(master, param, grad, m, v, lr, beta1, beta2, eps, weight_decay, step, clip_coef=None):
    """The arithmetic of adamw_kernel (csrc/train.cu): decoupled decay, bias-corrected step, bf16 working copy."""
    g = grad.float() * (float(clip_coef) if clip_coef is not None else 1.0)
    master.mul_(1.0 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    master.sub_((lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps))
    param.copy_(master.to(param.dtype))


def test_slice_range_covers_every_element_once():
    for n in (1, 7, 16, 37, 1024):
        for world in (1, 2, 3, 8):
            got = [slice_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            assert all(hi - lo <= s for lo, hi, s in got)


def _params():
    g = torch.Generator().manual_seed(5)
    shapes = {"even": (8, 16), "ragged": (37,), "tiny": (1,), "skipped": (6,)}
    return {k: torch.randn(s, generator=g) for k, s in shapes.items()}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        src = _params()
        names = sorted(src)
        params = {k: v.to(torch.bfloat16).contiguous() for k, v in src.items()}
        opt = ShardedAdamW(params, src, names, dist.group.WORLD, adamw=torch_adamw)
        assert opt.state_bytes() < 3 * 4 * sum(v.numel() for v in src.values())          # only a slice is held
        # unsharded reference with the same arithmetic
        ref_p = {k: v.to(torch.bfloat16) for k, v in src.items()}
        ref = {k: dict(master=v.clone().reshape(-1), m=torch.zeros(v.numel()), v=torch.zeros(v.numel()))
               for k, v in src.items()}
        g = torch.Generator().manual_seed(77)                                              # same gradients on every rank
        for step in (1, 2, 3):
            grads = {k: torch.randn(src[k].shape, generator=g) for k in names if k != "skipped"}
            clip = torch.tensor([0.5])
            lr_of = lambda k: 1e-2 if k == "even" else 3e-3
            wd_of = lambda k: 0.1 if k == "ragged" else 0.0
            opt.step(grads, step, lr_of, wd_of, clip_coef=clip)
            for k in grads:
                r = ref[k]
                torch_adamw(r["master"], ref_p[k].view(-1), grads[k].reshape(-1), r["m"], r["v"], lr_of(k), 0.9, 0.999,
                            1e-8, wd_of(k), step, clip_coef=clip)
            for k in names:
                assert torch.equal(params[k], ref_p[k]), (k, step)                        # every slice arrived everywhere
                lo, hi, _ = slice_range(src[k].numel(), rank, world)
                if k != "skipped":
                    assert torch.equal(opt.master[k], ref[k]["master"][lo:hi]), (k, step)
        assert torch.equal(params["skipped"], src["skipped"].to(torch.bfloat16))
        # resume: slices saved per rank, loaded into a fresh optimizer whose working weights are stale
        import tempfile
        from mm_or_b200.train import checkpoint as C
        tmp = os.path.join(tempfile.gettempdir(), f"b200_zero_resume_{port}_rank{rank}.pt")
        C.save_optimizer(tmp, opt.master, opt.m, opt.v, 3, 1e-2, 3e-3)
        stale = {k: torch.zeros_like(v) for k, v in params.items()}
        opt2 = ShardedAdamW(stale, src, names, dist.group.WORLD, adamw=torch_adamw)
        assert C.load_optimizer(tmp, opt2.master, opt2.m, opt2.v) == (3, 1e-2, 3e-3)
        opt2.refresh_from_masters()
        os.remove(tmp)
        for k in names:
            assert torch.equal(stale[k], params[k]) or k == "skipped", k     # "skipped" never moved: master == source
        assert torch.equal(stale["skipped"], src["skipped"].to(torch.bfloat16))
        q.put((rank, "ok"))
    except Exception as e:  # surfaced by the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_adamw_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def torch_sq_norm(t, out2):
    out2[0] += (t.float() ** 2).sum()


def _zero2_worker(rank, world, port, q):
    """ZeRO-2 step (reduce-scattered gradients): every rank contributes a DIFFERENT gradient; the result must equal the
    unsharded step on the mean gradient (rounded as the wire format rounds it), clipped by the GLOBAL norm."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for comm in (torch.float32, torch.bfloat16):
            src = _params()
            names = sorted(src)
            params = {k: v.to(torch.bfloat16).contiguous() for k, v in src.items()}
            opt = ShardedAdamW(params, src, names, dist.group.WORLD, adamw=torch_adamw, sq_norm=torch_sq_norm)
            ref_p = {k: v.to(torch.bfloat16) for k, v in src.items()}
            ref = {k: dict(master=v.clone().reshape(-1), m=torch.zeros(v.numel()), v=torch.zeros(v.numel()))
                   for k, v in src.items()}
            lr_of = lambda k: 1e-2 if k == "even" else 3e-3
            wd_of = lambda k: 0.1 if k == "ragged" else 0.0
            max_norm = 0.7
            for step in (1, 2):
                per_rank = []
                for r in range(world):          # every rank can rebuild all contributions: the expectation is local
                    g = torch.Generator().manual_seed(1000 * step + r)
                    per_rank.append({k: torch.randn(src[k].shape, generator=g) for k in names if k != "skipped"})
                mine = {k: v.clone() for k, v in per_rank[rank].items()}
                out2 = opt.step_from_local_grads(mine, step, lr_of, wd_of, max_norm=max_norm, comm_dtype=comm)
                assert not mine                                            # full-size gradients were released
                mean = {}
                for k in per_rank[0]:
                    tot = sum(p[k].to(comm).float() for p in per_rank).to(comm).float()
                    mean[k] = tot / world
                sq = sum((v ** 2).sum() for v in mean.values())
                coef = torch.clamp(max_norm / (sq.sqrt() + 1e-6), max=1.0)
                assert torch.allclose(out2[0], sq, rtol=1e-5) and torch.allclose(out2[1], coef, rtol=1e-5)
                for k in mean:
                    r_ = ref[k]
                    torch_adamw(r_["master"], ref_p[k].view(-1), mean[k].reshape(-1), r_["m"], r_["v"], lr_of(k), 0.9,
                                0.999, 1e-8, wd_of(k), step, clip_coef=out2[1])
                for k in names:
                    assert torch.equal(params[k], ref_p[k]), (k, step, comm)
                    lo, hi, _ = slice_range(src[k].numel(), rank, world)
                    if k != "skipped":
                        assert torch.equal(opt.master[k], ref[k]["master"][lo:hi]), (k, step, comm)
        q.put((rank, "ok"))
    except Exception as e:  # surfaced by the parent
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-800:]))
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_zero2_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_zero2_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def _zero2_world3_worker(rank, world, port, q):
    """The flat (world, S) layout of the ZeRO-2 step with a world size that divides nothing: ragged last rows, a rank
    whose slice of a tiny tensor is EMPTY, the alignment padding of a row. fp32 wire (the mean over 3 ranks is not a
    power-of-two scaling, so sum * (clip / 3) and (sum / 3) * clip differ in the last bit: tolerance 1e-6 relative on the
    masters, bf16 working weights equal up to one ulp)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(9)
        shapes = {"a": (7,), "b": (2,), "c": (5, 13), "d": (1,), "e": (64,)}
        src = {k: torch.randn(s, generator=g) for k, s in shapes.items()}
        names = sorted(src)
        params = {k: v.to(torch.bfloat16).contiguous() for k, v in src.items()}
        opt = ShardedAdamW(params, src, names, dist.group.WORLD, adamw=torch_adamw, sq_norm=torch_sq_norm)
        ref = {k: dict(master=v.clone().reshape(-1), m=torch.zeros(v.numel()), v=torch.zeros(v.numel()))
               for k, v in src.items()}
        ref_p = {k: v.to(torch.bfloat16) for k, v in src.items()}
        lr_of, wd_of = (lambda k: 5e-3), (lambda k: 0.05 if k == "c" else 0.0)
        for step in (1, 2, 3):
            per_rank = []
            for r in range(world):
                gr = torch.Generator().manual_seed(100 * step + r)
                per_rank.append({k: torch.randn(src[k].shape, generator=gr) for k in names})
            mine = {k: v.clone() for k, v in per_rank[rank].items()}
            out2 = opt.step_from_local_grads(mine, step, lr_of, wd_of, max_norm=0.3, comm_dtype=torch.float32)
            assert not mine
            mean = {k: sum(p[k] for p in per_rank) / world for k in names}
            sq = sum((v ** 2).sum() for v in mean.values())
            coef = torch.clamp(0.3 / (sq.sqrt() + 1e-6), max=1.0)
            assert torch.allclose(out2[0], sq, rtol=1e-5) and torch.allclose(out2[1], coef, rtol=1e-5)
            for k in names:
                r_ = ref[k]
                torch_adamw(r_["master"], ref_p[k].view(-1), mean[k].reshape(-1), r_["m"], r_["v"], lr_of(k), 0.9, 0.999,
                            1e-8, wd_of(k), step, clip_coef=coef)
                lo, hi, _ = slice_range(src[k].numel(), rank, world)
                assert opt.master[k].numel() == hi - lo
                assert torch.allclose(opt.master[k], r_["master"][lo:hi], rtol=1e-5, atol=1e-7), (k, step)
                assert torch.allclose(params[k].float(), ref_p[k].float(), rtol=1e-2, atol=1e-3), (k, step)
            # every rank holds the same working weights bit for bit (each element was updated by exactly one rank)
            flat = torch.cat([params[k].float().reshape(-1) for k in names])
            gathered = [torch.zeros_like(flat) for _ in range(world)]
            dist.all_gather(gathered, flat)
            assert all(torch.equal(gathered[0], x) for x in gathered)
        lo, hi, _ = slice_range(2, 2, 3)
        assert (lo, hi) == (2, 2) or rank != 2                                # rank 2 owns nothing of tensor "b"
        q.put((rank, "ok"))
    except Exception as e:  # surfaced by the parent
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-800:]))
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_zero2_world3_ragged_gloo():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_zero2_world3_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok"), (2, "ok")], results
