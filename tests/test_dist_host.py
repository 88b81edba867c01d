"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in mm_or_b200/dist.py: sample partition, decode-slice
ownership and the visual-token all-gather layout that generate() relies on under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mm_or_b200 import dist as D
from mm_or_b200.model.pack import plan_pack


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            got = [D.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)


def test_decode_owner_is_a_permutation():
    for world in (1, 2, 8):
        for shift in (0, 1, 3):
            assert sorted(D.decode_owner(r, world, shift) for r in range(world)) == list(range(world))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, T, Dh = 3, 5, 16
        g = torch.Generator().manual_seed(1234)
        everyone = torch.randn(world, B, T, Dh, generator=g)            # identical on all ranks
        for dtype in (torch.float32, torch.bfloat16):
            local = everyone[rank].to(dtype)
            gathered = D.all_gather_tokens(local)
            assert gathered.shape == (world, B, T, Dh) and gathered.dtype == dtype
            assert torch.equal(gathered, everyone.to(dtype)), "gathered tokens are not ordered by rank"
        same, got = D.all_gather_objects_equal((B, T, Dh))
        assert same and len(got) == world
        same, _ = D.all_gather_objects_equal(rank)
        assert not same
        # the slice a rank decodes with shift 1 is its neighbour's; pack rows must index that slice
        owner = D.decode_owner(rank, world, 1)
        vis = D.all_gather_tokens(everyone[rank])[owner]
        ids = np.full((B, 9), 7, dtype=np.int64)
        ids[:, 2] = -200
        plan = plan_pack(ids, None, None, T, "left", None)
        src = plan.src.reshape(B, plan.L).astype(np.int64)
        vis_ids = np.where(src <= -2, np.arange(B)[:, None] * T + (-2 - src), -2)
        rows = vis.reshape(B * T, Dh)
        packed = torch.zeros(B, plan.L, Dh)
        sel = torch.from_numpy(vis_ids >= 0)
        packed[sel] = rows[torch.from_numpy(vis_ids[vis_ids >= 0])]
        assert torch.equal(packed[:, 2:2 + T], everyone[owner])
        # data-parallel fine-tuning: gradients are averaged over ranks, identically on every rank
        grads = {"a": torch.full((4, 3), float(rank + 1)), "b": torch.arange(6.0).view(2, 3).t() * (rank + 1)}
        avg = D.average_gradients(grads, ["a", "b"])
        assert torch.allclose(avg["a"], torch.full((4, 3), (1 + world) / 2.0))
        assert torch.allclose(avg["b"], torch.arange(6.0).view(2, 3).t() * (1 + world) / 2.0)
        # modality-specific gradients: present on rank 0 only -> zeros appear on rank 1; present nowhere -> stay absent
        g2 = {"w": torch.ones(2, 2)}
        if rank == 0:
            g2["audio"] = torch.full((3,), 4.0)
        D.align_optional_gradients(g2, {"audio": (3,), "seg": (2, 5)})
        assert sorted(g2) == ["audio", "w"] and g2["audio"].shape == (3,)
        avg2 = D.average_gradients(g2, sorted(g2))
        assert torch.allclose(avg2["audio"], torch.full((3,), 4.0 / world))
        q.put((rank, "ok"))
    except Exception as e:  # surfaced by the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_all_gather_layout_world2_gloo():
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
