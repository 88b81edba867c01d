"""CPU: the point-cloud oracle (oracle/ptv3_oracle.py) against fixtures recorded from the reference's own code
(tests/golden/make_ptv3_golden.py): serialization codes bit-exact, PointTransformerV3 features within fp32 tolerance."""
import os

import pytest
import torch

import golden_cases as gc
from oracle import ptv3_oracle as P

CODES = os.path.join(gc.GOLDEN_DIR, "ptv3_codes.pt")
ENC = os.path.join(gc.GOLDEN_DIR, "ptv3_encode.pt")
TOL_F32 = 2e-4      # fp32 pipeline vs fp32 reference: summation order only (the fp16 attention output is rounded alike)


def ptv3_case():
    clouds = [P.synth_cloud(2500, seed=1), None, P.synth_cloud(700, seed=2, box=(30, 30, 3))]
    return P.dedupe_clouds(clouds)


def canonical(t):
    key = ((t["batch"].long() * 4096 + t["grid"][:, 0]) * 4096 + t["grid"][:, 1]) * 4096 + t["grid"][:, 2]
    o = torch.argsort(key)
    return key[o], t["feat"][o]


@pytest.mark.parametrize("order", P.ORDERS)
def test_codes_bit_exact(order):
    fx = torch.load(CODES)
    for depth, c in fx.items():
        mine = P.encode(c["grid"], c["batch"], depth, order)
        assert mine.dtype == torch.int64 and torch.equal(mine, c[order]), (order, depth)


def test_hilbert_is_hierarchical():
    """code >> 3 of a point equals the code of its parent voxel one level up: what SerializedPooling relies on
    (pointtransformerv3.py:655) and what lets the B200 path derive the next level's codes by shifting."""
    g = torch.Generator().manual_seed(0)
    grid = torch.randint(0, 1 << 9, (500, 3), generator=g, dtype=torch.int32)
    b = torch.zeros(500, dtype=torch.long)
    for o in P.ORDERS:
        assert torch.equal(P.encode(grid, b, 9, o) >> 3, P.encode(grid >> 1, b, 8, o))


def test_patch_plan_matches_reference_shapes():
    pad, unpad, cu = P.patch_plan([2500, 700, 1024, 3], 1024)
    assert cu.tolist() == [0, 1024, 2048, 3072, 3772, 4796, 4799]
    assert len(pad) == 4799 and len(unpad) == 2500 + 700 + 1024 + 3
    # the topped-up patch of cloud 0 holds exactly its last 1024 points
    assert sorted(pad[2048:3072].tolist()) == list(range(2500 - 1024, 2500))
    assert torch.equal(pad[unpad], torch.arange(len(unpad)))


def test_encode_pc_matches_reference():
    fx = torch.load(ENC)
    clouds = ptv3_case()
    assert [None if c is None else len(c) for c in clouds] == fx["n_points"]
    sd = P.synth_weights()
    trace = []
    torch.manual_seed(fx["shuffle_seed"])
    out = P.encode_pc(sd, clouds, trace=trace)
    for t in trace:
        key, feat = canonical(t)
        ref = fx[t["stage"]]
        assert torch.equal(key, ref["key"]), t["stage"]
        tol = TOL_F32 if ref["feat"].dtype == torch.float32 else 2e-3      # fp16-compressed slices
        err = ((feat - ref["feat"].float()).norm() / ref["feat"].float().norm()).item()
        assert err < tol, (t["stage"], err)
    err = ((out - fx["pc_feats"]).norm() / fx["pc_feats"].norm()).item()
    assert err < TOL_F32, err
    # the missing cloud's row is project_pc(0) = the bias
    assert torch.allclose(out[1], sd[P.PT + "project_pc.bias"])


def test_explicit_perms_equal_global_rng():
    clouds = [P.synth_cloud(300, seed=4, box=(12, 12, 3))]
    clouds = P.dedupe_clouds(clouds)
    sd = P.synth_weights()
    torch.manual_seed(9)
    perms = [torch.randperm(4) for _ in range(5)]
    torch.manual_seed(9)
    a = P.encode_pc(sd, clouds)
    b = P.encode_pc(sd, clouds, perms=perms)
    assert torch.equal(a, b)


def test_full_model_with_point_clouds_matches_reference_fixture():
    """pc + audio + seg-masks through the whole oracle (token order pooled, pc, audio, seg x 3; prefill logits; greedy
    continuation) against tests/golden/pc_left.pt, recorded from the reference's own LlavaLlamaForCausalLM.forward
    (tests/golden/make_pc_golden.py)."""
    from helpers import oracle_cfg
    from oracle import mm2sg_oracle as O
    fx = torch.load(os.path.join(gc.GOLDEN_DIR, "pc_left.pt"))
    cfg = gc.small_config()
    sd = gc.small_weights(cfg)
    sd.update(P.synth_weights())
    sd = gc.bf16_round(sd)
    case = gc.make_case(cfg, "extras_left")
    pcs = P.dedupe_clouds([P.synth_cloud(1300, seed=21), None, P.synth_cloud(300, seed=22, box=(16, 16, 3))])
    assert [None if c is None else len(c) for c in pcs] == fx["n_points"]
    torch.manual_seed(fx["shuffle_seed"])
    with torch.no_grad():
        out = O.multimodal_prefill(sd, oracle_cfg(cfg), case["input_ids"], case["attention_mask"], case["images"],
                                   audio=case["audio"], segmasks=case["segmasks"], pc=pcs, padding_side="left")
        toks, lg = O.greedy_decode(sd, oracle_cfg(cfg), out["logits"][:, -1], out["kv"], out["mask"],
                                   fx["greedy_ids"].shape[1], stop_on_eos=False)
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
    assert out["visual"].shape[1] == 581
    assert rel(out["visual"][:, 570:], fx["visual_tail"]) < TOL_F32
    assert rel(out["logits"][:, -1], fx["logits_last"]) < TOL_F32
    assert torch.equal(toks, fx["greedy_ids"])
    assert rel(lg, fx["greedy_logits"]) < 1e-3                      # fp16-compressed fixture
