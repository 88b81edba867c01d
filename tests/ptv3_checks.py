"""Shared checks of the point-cloud operators (csrc/ptv3.cu) and their host orchestration against the oracle and the
fixtures recorded from the reference. Run twice: on the CPU through the kernel emulator (tests/test_ptv3_emu.py) and on
the GPU through libb200mmor.so (tests/test_gpu_zz_pointcloud.py). Inputs are created on the host and moved to
`ops.device`; results are compared on the host. Integer outputs bit-exact; fp32 features within summation-order
tolerance (the tolerance is written next to each comparison)."""
import os

import pytest
import torch

import golden_cases as gc
from mm_or_b200.model import point_transformer as PT
from oracle import ptv3_oracle as P

TOL_F32 = 2e-4      # fp32 pipeline vs fp32 oracle / reference: summation order and erf / exp implementations only


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def eq(a, b):
    return torch.equal(a.cpu(), b.cpu())


def golden_clouds():
    return P.dedupe_clouds([P.synth_cloud(2500, seed=1), None, P.synth_cloud(700, seed=2, box=(30, 30, 3))])


def check_grid_coords_and_codes_bit_exact(ops):
    d = lambda t: t.to(ops.device)
    clouds = P.dedupe_clouds([P.synth_cloud(900, seed=3), P.synth_cloud(400, seed=4, box=(20, 20, 3))])
    pts = torch.cat(clouds)
    batch = torch.cat([torch.full((len(c),), i, dtype=torch.int32) for i, c in enumerate(clouds)])
    grid, mx = ops.grid_coords(d(pts), 0.01)
    ref = P.grid_coords(pts[:, :3], 0.01)
    assert eq(grid, ref) and int(mx) == int(ref.max())
    depth = int(mx).bit_length()
    for k, name in enumerate(P.ORDERS):
        assert eq(ops.encode(grid, d(batch), len(pts), depth, k), P.encode(ref, batch.long(), depth, name)), name


def check_codes_match_reference_fixture(ops, order):
    """int64 codes recorded from the reference's serialization/{z_order,hilbert}.py: bit-exact."""
    d = lambda t: t.to(ops.device)
    fx = torch.load(os.path.join(gc.GOLDEN_DIR, "ptv3_codes.pt"))
    for depth, c in fx.items():
        got = ops.encode(d(c["grid"].contiguous()), d(c["batch"].to(torch.int32)), len(c["grid"]), depth, order)
        assert eq(got, c[P.ORDERS[order]]), (depth, order)


def check_argsort_neighbors_pool_plan(ops):
    d = lambda t: t.to(ops.device)
    cloud = P.dedupe_clouds([P.synth_cloud(1500, seed=6)])[0]
    n = len(cloud)
    batch = d(torch.zeros(n, dtype=torch.int32))
    grid, mx = ops.grid_coords(d(cloud), 0.01)
    depth = int(mx).bit_length()
    code = ops.encode(grid, batch, n, depth, 0)
    zc, order = ops.argsort(code, n, 3 * depth + 1)
    assert eq(order.long(), torch.argsort(code.cpu(), stable=True)) and eq(zc, code.cpu()[order.cpu().long()])
    g = ops.gather_rows(grid, order, n)
    assert eq(g, grid.cpu()[order.cpu().long()])
    dup = d(torch.zeros(1, dtype=torch.int32))
    for k in (3, 5):
        nbr = ops.neighbors(zc, g, batch, n, depth, k, dup)
        assert eq(nbr.long(), P.neighbor_table(g.cpu(), batch.cpu().long(), k))
    assert int(dup) == 0
    seg, n_out, grid_o, batch_o = ops.pool_plan(zc, g, batch, n, 1)
    parent, counts = torch.unique(zc.cpu() >> 3, return_counts=True)
    m = int(n_out)
    assert m == len(parent)
    assert eq(seg[:m + 1].long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))
    assert eq(grid_o[:m], g.cpu()[seg.cpu()[:m].long()] >> 1)
    # the pooled level is born sorted: its z codes are the parents' codes
    assert eq(ops.encode(grid_o[:m].contiguous(), batch_o[:m].contiguous(), m, depth - 1, 0), parent)
    # duplicate voxel -> flag
    g2, zc2 = g.clone(), zc.clone()
    g2[5] = g2[4]
    zc2[5] = zc2[4]
    ops.neighbors(zc2, g2, batch, n, depth, 3, dup)
    assert int(dup) == 1


def check_gather_gemm_epilogues(ops):
    d = lambda t: t.to(ops.device)
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
    got = ops.gemm(d(a), d(w), M, K, N, idx=d(idx), taps=taps, bias=d(bias), bn=(d(scale), d(shift)), act=2,
                   residual=d(res))
    assert rel(got, ref) < 1e-5
    w1 = d(w[:K].contiguous())
    assert rel(ops.gemm(d(a), w1, M, K, N), a @ w[:K]) < 1e-5
    out = d(torch.zeros(M, 3, N, dtype=torch.bfloat16))
    ops.gemm(d(a), w1, M, K, N, out=out[:, 1], ldc=3 * N, out_bf16=True)
    assert rel(out[:, 1], a @ w[:K]) < 4e-3                               # one bf16 rounding
    assert float(out[:, 0].float().abs().sum()) == 0 and float(out[:, 2].float().abs().sum()) == 0
    # a shape that is not a multiple of any tile and K = 6 with 125 taps (the stem)
    M, K, N, taps = 77, 6, 32, 125
    a, w = torch.randn(M, K, generator=g), torch.randn(taps * K, N, generator=g) * 0.1
    idx = torch.randint(-1, M, (M, taps), generator=g).to(torch.int32)
    acc = torch.zeros(M, N)
    for t in range(taps):
        ok = idx[:, t] >= 0
        acc[ok] += a[idx[ok, t].long()] @ w[t * K:(t + 1) * K]
    assert rel(ops.gemm(d(a), d(w), M, K, N, idx=d(idx), taps=taps), acc) < 1e-5


def check_gather_gemm_tiled_shapes(ops):
    """The fast path (K, N multiples of 4): every tile width, partial tiles in M and N, absent taps, every epilogue."""
    d = lambda t: t.to(ops.device)
    g = torch.Generator().manual_seed(3)
    # small M: 64-row tiles, taps split over CTAs (27 -> 9 x 3, 3 -> 3 x 1, 2 -> 2 x 1); M = 38000: 128-row tiles
    for M, K, N, taps in [(300, 32, 32, 27), (129, 64, 64, 27), (257, 32, 96, 1), (140, 128, 128, 1), (70, 256, 256, 3),
                          (33, 512, 1024, 1), (200, 36, 44, 2), (5, 8, 4, 1), (38000, 16, 32, 1), (19000, 8, 64, 2),
                          (9600, 8, 128, 1), (38000, 8, 64, 1), (38000, 8, 128, 1), (5120, 8, 32, 27)]:   # last: 4 ragged splits (7, 7, 7, 6)
        a = torch.randn(M, K, generator=g)
        w = torch.randn(taps * K, N, generator=g) / (taps * K) ** 0.5
        idx = torch.randint(-1, M, (M, taps), generator=g).to(torch.int32) if taps > 1 else None
        bias, scale, shift = torch.randn(N, generator=g), torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
        res = torch.randn(M, N, generator=g)
        acc = torch.zeros(M, N)
        for t in range(taps):
            if idx is None:
                acc += a @ w
            else:
                ok = idx[:, t] >= 0
                acc[ok] += a[idx[ok, t].long()] @ w[t * K:(t + 1) * K]
        ref = torch.nn.functional.gelu((acc + bias) * scale + shift) + res
        got = ops.gemm(d(a), d(w), M, K, N, idx=None if idx is None else d(idx), taps=taps, bias=d(bias),
                       bn=(d(scale), d(shift)), act=2, residual=d(res))
        assert rel(got, ref) < 1e-5, (M, K, N, taps)
        assert rel(ops.gemm(d(a), d(w), M, K, N, idx=None if idx is None else d(idx), taps=taps), acc) < 1e-5


def check_layernorm_segment_max_cloud_mean(ops):
    d = lambda t: t.to(ops.device)
    g = torch.Generator().manual_seed(1)
    for C in (32, 48, 512):
        x, res = torch.randn(37, C, generator=g), torch.randn(37, C, generator=g)
        gam, bet = torch.randn(C, generator=g), torch.randn(C, generator=g)
        ref = torch.nn.functional.layer_norm(x, (C,), gam, bet, 1e-5)
        assert rel(ops.layernorm(d(x), 37, C, d(gam), d(bet), 1e-5), ref) < 1e-5
        assert rel(ops.layernorm(d(x), 37, C, d(gam), d(bet), 1e-5, residual=d(res)), ref + res) < 1e-5
    x = torch.randn(40, 64, generator=g)
    seg = torch.tensor([0, 1, 4, 4 + 20, 40], dtype=torch.int32)
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
    ref = torch.stack([x[seg[i]:seg[i + 1]].max(0).values for i in range(4)])
    assert rel(ops.segment_max(d(x), d(seg), 4, 64, (d(sc), d(sh)), 0), ref * sc + sh) < 1e-6
    out = d(torch.zeros(5, 64))
    ops.cloud_mean(d(x), d(seg), 4, 64, d(torch.tensor([4, 0, 2, 1], dtype=torch.int32)), out)
    for b, r in enumerate([4, 0, 2, 1]):
        assert rel(out[r], x[seg[b]:seg[b + 1]].mean(0)) < 1e-6
    assert float(out[3].abs().sum()) == 0


def check_patch_attention_matches_flash_semantics(ops):
    d = lambda t: t.to(ops.device)
    g = torch.Generator().manual_seed(2)
    # small patch sizes: the kernel takes patch descriptors, not K. (K, cloud sizes): one short patch, full patches, a
    # topped-up last patch; 300 exercises the second query of a thread and partial key tiles
    for K, counts, C, H in [(64, [150, 40, 64], 32, 2), (300, [700, 150, 300], 64, 4)]:
        n = sum(counts)
        qkv = torch.randn(n, 3 * C, generator=g) * 2
        off, order = 0, []
        for c in counts:                            # a serialized order: a permutation inside each cloud
            order.append(off + torch.randperm(c, generator=g))
            off += c
        order = torch.cat(order).to(torch.int32)
        pat = PT.patch_descriptors(counts, K)
        got = ops.patch_attention(d(qkv), d(order), d(torch.tensor(pat, dtype=torch.int32)), len(pat),
                                  max(p[1] for p in pat), n, C, H).cpu()
        pad, unpad, cu = P.patch_plan(counts, K)
        inverse = torch.empty(n, dtype=torch.long)
        inverse[order.long()] = torch.arange(n)
        q = qkv[order.long()[pad]]
        ref = P.varlen_attention_fp16(q.half().reshape(-1, 3, H, C // H), cu, H, (C // H) ** -0.5).float()
        ref = ref[unpad[inverse]]
        assert rel(got, ref) < 1e-3                 # both rounded to fp16 at the end: differences are 1-ulp fp16 flips
        assert (got - ref).abs().max().item() < 4e-3


def make_model(ops):
    sd = P.synth_weights()
    return PT.PointTransformerV3().load_weights(sd, P.PT, ops.device, ops=ops), sd


def check_encode_pc_matches_oracle_and_reference_fixture(ops):
    """The whole branch on the golden case: (B, 1024) bf16 tokens vs the reference's recorded fp32 pc_feats."""
    fx = torch.load(os.path.join(gc.GOLDEN_DIR, "ptv3_encode.pt"))
    clouds = golden_clouds()
    model, sd = make_model(ops)
    torch.manual_seed(fx["shuffle_seed"])
    out = model(clouds)
    assert out.dtype == torch.bfloat16 and out.shape == (3, 1024)
    assert rel(out, fx["pc_feats"]) < 4e-3                                   # one bf16 rounding of the result
    assert eq(out[1], sd[P.PT + "project_pc.bias"].to(torch.bfloat16))
    # fp32 features of the last level, before pooling, against the reference's (canonical order = z order here)
    torch.manual_seed(fx["shuffle_seed"])
    real = [c for c in clouds if c is not None]
    pts = torch.cat(real).to(ops.device)
    batch = torch.cat([torch.full((len(c),), j, dtype=torch.int32) for j, c in enumerate(real)]).to(ops.device)
    spts, nbr5, levels = model.plan(pts, batch, len(real))
    feat = model.features(spts, nbr5, levels).cpu()
    last = levels[-1]
    grid, b = last.grid.cpu(), last.batch.cpu()
    key = ((b.long() * 4096 + grid[:, 0]) * 4096 + grid[:, 1]) * 4096 + grid[:, 2]
    o = torch.argsort(key)
    assert eq(key[o], fx["enc4"]["key"])
    assert rel(feat[o], fx["enc4"]["feat"]) < TOL_F32
    assert [lv.n for lv in levels] == [fx[f"enc{s}"]["feat"].shape[0] for s in range(5)]


def check_edge_cases_match_oracle(ops):
    """All clouds missing; a single point; 1024 and 1025 points (exactly one patch / one full + a one-point patch that
    attends to the last 1024 points); ragged batch order (cloud, None, None, cloud)."""
    model, sd = make_model(ops)
    bias = sd[P.PT + "project_pc.bias"].to(torch.bfloat16)
    out = model([None, None])
    assert eq(out, torch.stack([bias, bias]))
    cases = [[P.synth_cloud(1, seed=30, box=(4, 4, 2))],
             P.dedupe_clouds([P.synth_cloud(1024, seed=31, box=(30, 30, 2))]),
             P.dedupe_clouds([P.synth_cloud(1031, seed=32, box=(30, 30, 2)), None, None,
                              P.synth_cloud(60, seed=33, box=(8, 8, 2))])]
    while len(cases[2][0]) > 1025:                       # trim to exactly 1025 points after de-duplication
        cases[2][0] = cases[2][0][:1025]
        cases[2] = P.dedupe_clouds(cases[2])
    for clouds in cases:
        torch.manual_seed(5)
        got = model(clouds)
        torch.manual_seed(5)
        ref = P.encode_pc(sd, clouds)
        assert rel(got, ref) < 4e-3, [None if c is None else len(c) for c in clouds]      # one bf16 rounding


def check_entry_points_reject_bad_arguments(ops):
    """Every b200_pc_* entry point returns -2 and leaves a message (no launch) for arguments it cannot serve."""
    d = lambda t: t.to(ops.device)
    lib, p, st = ops.lib, ops.ptr, ops.stream
    f32, i32, i64 = (lambda *s: d(torch.zeros(*s))), (lambda *s: d(torch.zeros(*s, dtype=torch.int32))), \
        (lambda *s: d(torch.zeros(*s, dtype=torch.int64)))
    x, g, b, c = f32(8, 6), i32(8, 3), i32(8), i64(8)
    bad = [
        lib.b200_pc_grid_coords(p(x), 6, 0, 0.01, p(f32(3)), p(g), p(i32(1)), st()),            # n = 0
        lib.b200_pc_grid_coords(p(x), 6, 8, 0.0, p(f32(3)), p(g), p(i32(1)), st()),             # grid size 0
        lib.b200_pc_encode(p(g), p(b), 8, 17, 0, p(c), st()),                                   # depth > 16
        lib.b200_pc_encode(p(g), p(b), 8, 4, 4, p(c), st()),                                    # unknown order
        lib.b200_pc_argsort(p(c), 8, 12, p(c), p(b), p(f32(1)), 4, st()),                       # workspace too small
        lib.b200_pc_neighbors(p(c), p(g), p(b), 8, 4, 4, p(i32(8, 64)), p(i32(1)), st()),       # kernel size 4
        lib.b200_pc_pool_plan(p(c), p(g), p(b), 8, 5, p(i32(9)), p(i32(1)), p(g), p(b), st()),  # pooling depth 5
        lib.b200_pc_gemm_f32(p(x), 6, None, 2, p(f32(12, 4)), None, None, None, 0, None, 0, p(f32(8, 4)), 4, 0, 8, 4, 6,
                             None, 0, st()),                                                    # taps without indices
        lib.b200_pc_gemm_f32(p(x), 6, None, 1, p(f32(6, 4)), None, None, None, 1, None, 0, p(f32(8, 4)), 4, 0, 8, 4, 6,
                             None, 0, st()),                                                    # unsupported activation
        lib.b200_pc_layernorm_f32(p(x), 4, p(f32(6)), p(f32(6)), 1e-5, None, 0, p(f32(8, 6)), 6, 8, 6, st()),  # ld < C
        lib.b200_pc_patch_attention(p(f32(8, 96)), 96, p(b), p(i32(1, 4)), 1, 8, 32, 4, 0.25, p(f32(8, 32)), 32,
                                    st()),                                                      # head_dim 8
        lib.b200_pc_segment_max(p(x), 6, p(i32(3)), 2, 6, p(f32(6)), None, 0, p(f32(2, 6)), 6, st()),  # scale w/o shift
        lib.b200_pc_cloud_mean(p(x), 6, p(i32(2)), 0, 6, p(i32(1)), p(f32(1, 6)), 6, st()),     # no clouds
    ]
    assert all(rc == -2 for rc in bad), bad
    assert b"bad argument" in lib.b200_last_error() or b"b200_pc_" in lib.b200_last_error()


check_entry_points_reject_bad_arguments.no_launch = True      # its point is that nothing is launched


def check_bad_inputs_raise(ops):
    model, _ = make_model(ops)
    c = P.synth_cloud(200, seed=8, box=(10, 10, 3))
    with pytest.raises(ValueError):
        model([torch.cat([c, c[:1] + 1e-5])])          # two points in one voxel
    with pytest.raises(ValueError):
        model([torch.zeros(10, 5)])


ALL = [check_grid_coords_and_codes_bit_exact, check_argsort_neighbors_pool_plan, check_gather_gemm_epilogues,
       check_gather_gemm_tiled_shapes,
       check_layernorm_segment_max_cloud_mean, check_patch_attention_matches_flash_semantics,
       check_encode_pc_matches_oracle_and_reference_fixture, check_edge_cases_match_oracle, check_entry_points_reject_bad_arguments,
       check_bad_inputs_raise]
