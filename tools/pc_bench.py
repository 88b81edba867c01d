"""Times the point-cloud branch (PointTransformerV3 -> project_pc token) on cuda:0 for synthetic clouds of a realistic
size (the reference's clouds are fused depth maps of the room at 1 cm voxels). Prints one JSON line.
  python tools/pc_bench.py [--points 60000] [--clouds 2] [--iters 5]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def surface_cloud(n, seed, extent=400):
    """n distinct voxels on a wavy surface inside an extent^2 footprint (points of a depth-fused room are surfaces)."""
    g = torch.Generator().manual_seed(seed)
    xy = torch.randperm(extent * extent, generator=g)[:n]
    x, y = (xy // extent).float(), (xy % extent).float()
    z = (20 + 10 * torch.sin(x / 23.0) + 8 * torch.cos(y / 17.0)).round()
    v = torch.stack([x, y, z], 1)
    xyz = (v + 0.5) * 0.01
    return torch.cat([xyz, torch.rand(n, 3, generator=g)], 1)


def unique_voxels(clouds, grid_size=0.01):
    """Drops the points that share a voxel of the batch-wide 1 cm grid with an earlier point of the same cloud (the
    grid is computed with the reference's fp32 formula, pointtransformerv3.py:96-98; duplicates are undefined for
    submanifold convolution)."""
    for _ in range(8):
        lo = torch.cat(clouds)[:, :3].min(0)[0]
        changed = False
        for i, c in enumerate(clouds):
            gcoord = torch.div(c[:, :3] - lo, torch.tensor(grid_size), rounding_mode="trunc").long()
            key = (gcoord[:, 0] * 65536 + gcoord[:, 1]) * 65536 + gcoord[:, 2]
            order = torch.argsort(key, stable=True)
            first = torch.ones(len(key), dtype=torch.bool)
            first[order[1:]] = key[order[1:]] != key[order[:-1]]
            if not bool(first.all()):
                clouds[i] = c[first]
                changed = True
        if not changed:
            return clouds
    raise RuntimeError("could not make the voxels unique")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=60000)
    ap.add_argument("--clouds", type=int, default=2)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    from mm_or_b200 import _lib as L
    from mm_or_b200.model import point_transformer as PT
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(0)
    sd = {}
    model = PT.PointTransformerV3()
    for name, shape, kind in PT.weight_specs():
        if kind == "v":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "g":
            t = 1 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "w":
            fan = 1
            for s in shape[1:]:
                fan *= s
            t = torch.randn(shape, generator=g) / fan ** 0.5
        else:
            t = 0.05 * torch.randn(shape, generator=g)
        sd[name] = t
    model.load_weights(sd, "", "cuda:0")
    clouds = unique_voxels([surface_cloud(a.points, s) for s in range(a.clouds)])
    torch.manual_seed(0)
    model(clouds)
    torch.cuda.synchronize()
    L.prof_enable(True)
    n0 = L.launch_count()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        out = model(clouds)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.iters
    fam = L.prof_collect()["pointcloud"]
    L.prof_enable(False)
    print(json.dumps({"what": "PointTransformerV3 cls_mode + project_pc, fp32", "clouds": a.clouds,
                      "points_per_cloud": [len(c) for c in clouds], "wall_ms": round(wall * 1e3, 2),
                      "kernel_ms": round(fam["ms"] / a.iters, 2), "launches": (L.launch_count() - n0) // a.iters,
                      "gflop": round(fam["flops"] / a.iters / 1e9, 1),
                      "tflops_fp32": round(fam["flops"] / max(fam["ms"], 1e-9) / 1e9, 2),
                      "finite": bool(torch.isfinite(out.float()).all())}))


if __name__ == "__main__":
    main()
