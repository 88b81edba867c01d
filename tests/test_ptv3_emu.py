"""CPU: the point-cloud KERNELS (mm_or_b200/csrc/ptv3.cu, compiled for the host against the kernel emulator of
tests/emu/) and their host orchestration (mm_or_b200/model/point_transformer.py) against the oracle and the fixtures
recorded from the reference. Integer outputs bit-exact; fp32 features within summation-order tolerance."""
import os

import pytest
import torch

import emu_lib
import golden_cases as gc
from mm_or_b200.model import point_transformer as PT
from oracle import ptv3_oracle as P

TOL = 2e-4


@pytest.fixture(scope="module")
def ops():
    return emu_lib.ops()


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def test_grid_coords_and_codes_bit_exact(ops):
    clouds = P.dedupe_clouds([P.synth_cloud(900, seed=3), P.synth_cloud(400, seed=4, box=(20, 20, 3))])
    pts = torch.cat(clouds)
    batch = torch.cat([torch.full((len(c),), i, dtype=torch.int32) for i, c in enumerate(clouds)])
    grid, mx = ops.grid_coords(pts, 0.01)
    ref = P.grid_coords(pts[:, :3], 0.01)
    assert torch.equal(grid, ref) and int(mx) == int(ref.max())
    depth = int(mx).bit_length()
    for k, name in enumerate(P.ORDERS):
        assert torch.equal(ops.encode(grid, batch, len(pts), depth, k), P.encode(ref, batch.long(), depth, name)), name


@pytest.mark.parametrize("order", range(4))
def test_codes_match_reference_fixture(ops, order):
    fx = torch.load(os.path.join(gc.GOLDEN_DIR, "ptv3_codes.pt"))
    for depth, c in fx.items():
        got = ops.encode(c["grid"].contiguous(), c["batch"].to(torch.int32), len(c["grid"]), depth, order)
        assert torch.equal(got, c[P.ORDERS[order]]), (depth, order)


def test_argsort_neighbors_pool_plan(ops):
    cloud = P.dedupe_clouds([P.synth_cloud(1500, seed=6)])[0]
    n = len(cloud)
    batch = torch.zeros(n, dtype=torch.int32)
    grid, mx = ops.grid_coords(cloud, 0.01)
    depth = int(mx).bit_length()
    code = ops.encode(grid, batch, n, depth, 0)
    zc, order = ops.argsort(code, n, 3 * depth + 1)
    assert torch.equal(order.long(), torch.argsort(code, stable=True)) and torch.equal(zc, code[order.long()])
    g = ops.gather_rows(grid, order, n)
    assert torch.equal(g, grid[order.long()])
    dup = torch.zeros(1, dtype=torch.int32)
    for k in (3, 5):
        nbr = ops.neighbors(zc, g, batch, n, depth, k, dup)
        assert torch.equal(nbr.long(), P.neighbor_table(g, batch.long(), k))
    assert int(dup) == 0
    seg, n_out, grid_o, batch_o = ops.pool_plan(zc, g, batch, n, 1)
    parent, counts = torch.unique(zc >> 3, return_counts=True)
    m = int(n_out)
    assert m == len(parent)
    assert torch.equal(seg[:m + 1].long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))
    assert torch.equal(grid_o[:m], g[seg[:m].long()] >> 1)
    # the pooled level is born sorted: its z codes are the parents' codes
    assert torch.equal(ops.encode(grid_o[:m].contiguous(), batch_o[:m].contiguous(), m, depth - 1, 0), parent)
    # duplicate voxel -> flag
    g2 = g.clone()
    g2[5] = g2[4]
    zc2 = zc.clone()
    zc2[5] = zc2[4]
    ops.neighbors(zc2, g2, batch, n, depth, 3, dup)
    assert int(dup) == 1


def test_gather_gemm_epilogues(ops):
    g = torch.Generator().manual_seed(0)
    M, K, N, taps = 150, 24, 70, 5
    a = torch.randn(M, K, generator=g)
    w = torch.randn(taps * K, N, generator=g)
    idx = torch.randint(-1, M, (M, taps), generator=g).to(torch.int32)
    bias, scale, shift = torch.randn(N, generator=g), torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    acc = torch.zeros(M, N)
    for t in range(taps):
        ok = idx[:, t] >= 0
        acc[ok] += a[idx[ok, t].long()] @ w[t * K:(t + 1) * K]
    ref = torch.nn.functional.gelu((acc + bias) * scale + shift) + res
    got = ops.gemm(a, w, M, K, N, idx=idx, taps=taps, bias=bias, bn=(scale, shift), act=2, residual=res)
    assert rel(got, ref) < 1e-5
    plain = ops.gemm(a, w[:K].contiguous(), M, K, N)
    assert rel(plain, a @ w[:K]) < 1e-5
    out = torch.zeros(M, 3, N, dtype=torch.bfloat16)
    ops.gemm(a, w[:K].contiguous(), M, K, N, out=out[:, 1], ldc=3 * N, out_bf16=True)
    assert torch.equal(out[:, 1], (a @ w[:K]).to(torch.bfloat16)) or rel(out[:, 1], a @ w[:K]) < 4e-3
    assert float(out[:, 0].abs().sum()) == 0 and float(out[:, 2].abs().sum()) == 0


def test_layernorm_segment_max_cloud_mean(ops):
    g = torch.Generator().manual_seed(1)
    for C in (32, 48, 512):
        x, res = torch.randn(37, C, generator=g), torch.randn(37, C, generator=g)
        gam, bet = torch.randn(C, generator=g), torch.randn(C, generator=g)
        ref = torch.nn.functional.layer_norm(x, (C,), gam, bet, 1e-5)
        assert rel(ops.layernorm(x, 37, C, gam, bet, 1e-5), ref) < 1e-5
        assert rel(ops.layernorm(x, 37, C, gam, bet, 1e-5, residual=res), ref + res) < 1e-5
    x = torch.randn(40, 64, generator=g)
    seg = torch.tensor([0, 1, 4, 4 + 20, 40], dtype=torch.int32)
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
    ref = torch.stack([x[seg[i]:seg[i + 1]].max(0).values for i in range(4)])
    assert torch.equal(ops.segment_max(x, seg, 4, 64, (sc, sh), 0), ref * sc + sh)
    out = torch.zeros(5, 64)
    ops.cloud_mean(x, seg, 4, 64, torch.tensor([4, 0, 2, 1], dtype=torch.int32), out)
    for b, r in enumerate([4, 0, 2, 1]):
        assert rel(out[r], x[seg[b]:seg[b + 1]].mean(0)) < 1e-6
    assert float(out[3].abs().sum()) == 0


def test_patch_attention_matches_flash_semantics(ops):
    g = torch.Generator().manual_seed(2)
    K, C, H = 64, 32, 2                         # small patch size: the kernel takes patches, not K
    counts = [150, 40, 64]
    n = sum(counts)
    qkv = torch.randn(n, 3 * C, generator=g) * 2
    off = 0
    order = []
    for c in counts:                            # a serialized order: a permutation inside each cloud
        order.append(off + torch.randperm(c, generator=g))
        off += c
    order = torch.cat(order).to(torch.int32)
    pat = PT.patch_descriptors(counts, K)
    got = ops.patch_attention(qkv, order, torch.tensor(pat, dtype=torch.int32), len(pat), max(p[1] for p in pat), n, C, H)
    pad, unpad, cu = P.patch_plan(counts, K)
    inverse = torch.empty(n, dtype=torch.long)
    inverse[order.long()] = torch.arange(n)
    q = qkv[order.long()[pad]]
    ref = P.varlen_attention_fp16(q.half().reshape(-1, 3, H, C // H), cu, H, (C // H) ** -0.5).float()[unpad[inverse]]
    assert rel(got, ref) < 1e-3                 # both rounded to fp16 at the end: differences are 1-ulp fp16 flips
    assert (got - ref).abs().max().item() < 4e-3


def _model(ops):
    sd = P.synth_weights()
    return PT.PointTransformerV3().load_weights(sd, P.PT, "cpu", ops=ops), sd


def test_encode_pc_matches_oracle_and_reference_fixture(ops):
    """The whole branch on the golden case: (B, 1024) bf16 tokens vs the reference's recorded fp32 pc_feats."""
    fx = torch.load(os.path.join(gc.GOLDEN_DIR, "ptv3_encode.pt"))
    clouds = P.dedupe_clouds([P.synth_cloud(2500, seed=1), None, P.synth_cloud(700, seed=2, box=(30, 30, 3))])
    model, sd = _model(ops)
    torch.manual_seed(fx["shuffle_seed"])
    out = model(clouds)
    assert out.dtype == torch.bfloat16 and out.shape == (3, 1024)
    ref = fx["pc_feats"]
    assert rel(out, ref) < 4e-3                                  # one bf16 rounding of the result
    assert torch.equal(out[1], sd[P.PT + "project_pc.bias"].to(torch.bfloat16))
    # fp32 features of the last level, before pooling, against the oracle (canonical order = z order here)
    torch.manual_seed(fx["shuffle_seed"])
    pts = torch.cat([c for c in clouds if c is not None])
    batch = torch.cat([torch.full((len(c),), j, dtype=torch.int32)
                       for j, c in enumerate(c for c in clouds if c is not None)])
    spts, nbr5, levels = model.plan(pts, batch, 2)
    feat = model.features(spts, nbr5, levels)
    key = ((levels[-1].batch.long() * 4096 + levels[-1].grid[:, 0]) * 4096 + levels[-1].grid[:, 1]) * 4096 + \
        levels[-1].grid[:, 2]
    o = torch.argsort(key)
    assert torch.equal(key[o], fx["enc4"]["key"])
    assert rel(feat[o], fx["enc4"]["feat"]) < TOL
    assert [lv.n for lv in levels] == [fx[f"enc{s}"]["feat"].shape[0] for s in range(5)]


def test_duplicate_voxels_raise(ops):
    model, _ = _model(ops)
    c = P.synth_cloud(200, seed=8, box=(10, 10, 3))
    c = torch.cat([c, c[:1] + 1e-5])
    with pytest.raises(ValueError):
        model([c])
    with pytest.raises(ValueError):
        model([torch.zeros(10, 5)])
