"""CPU: the oracle restatement reproduces the golden outputs recorded from the reference's own model code
(tests/golden/make_golden.py). This is what pins the oracle."""
import os

import pytest
import torch

import golden_cases as gc
from helpers import oracle_cfg, rel_err
from oracle import mm2sg_oracle as O


@pytest.fixture(scope="module")
def setup():
    torch.set_grad_enabled(False)
    cfg = gc.small_config()
    return cfg, oracle_cfg(cfg), gc.bf16_round(gc.small_weights(cfg))


@pytest.mark.parametrize("name", list(gc.CASES))
def test_oracle_matches_reference_golden(setup, name):
    cfg, ocfg, sd = setup
    g = torch.load(os.path.join(gc.GOLDEN_DIR, name + ".pt"))
    case = gc.make_case(cfg, name)
    out = O.multimodal_prefill(sd, ocfg, case["input_ids"], case["attention_mask"], case["images"], case.get("labels"),
                               case.get("audio"), case.get("segmasks"), padding_side=case["side"])
    assert list(out["logits"].shape) == g["logits_shape"].tolist()
    feats = O.clip_tower_forward(sd, torch.cat(case["images"], 0), ocfg.vit)
    assert rel_err(feats[:, ::48, ::8], g["vit_slice"]) < 2e-3            # golden slices are stored in fp16
    assert rel_err(out["visual"][:, ::24, ::4], g["visual_slice"]) < 2e-3
    assert rel_err(out["visual"][:, 570:], g["visual_tail"]) < 2e-3
    mask = out["mask"]
    if case["side"] == "left":
        assert rel_err(out["logits"][:, -1], g["logits_last"]) < 1e-5
    rows = out["logits"][:, ::37]
    m = mask[:, ::37]
    assert rel_err(rows[m], g["logits_rows"].float()[m]) < 2e-3
    if "modified_labels" in g:
        assert torch.equal(out["modified_labels"], g["modified_labels"])
        vw = torch.linspace(0.2, 1.0, cfg.vocab_size)
        assert abs(O.weighted_ce(out["logits"], out["modified_labels"], vw).item() - g["weighted_loss"].item()) < 1e-4
        sl = out["logits"][..., :-1, :].reshape(-1, cfg.vocab_size)
        tl = out["modified_labels"][..., 1:].reshape(-1)
        assert abs(torch.nn.functional.cross_entropy(sl, tl).item() - g["hf_loss"].item()) < 1e-4
    if "greedy_ids" in g:
        steps = g["greedy_ids"].shape[1]
        toks, lg = O.greedy_decode(sd, ocfg, out["logits"][:, -1], out["kv"], out["mask"], steps, stop_on_eos=False)
        assert rel_err(lg, g["greedy_logits"]) < 2e-3
        assert torch.equal(toks, g["greedy_ids"])


def test_oracle_matches_reference_chain_fixture(setup):
    """The exact-token fixture (reference greedy ids under the chain weight set): all ids distinct, margins large."""
    cfg, ocfg, _ = setup
    sd = gc.bf16_round(gc.small_weights(cfg, chain=True))
    g = torch.load(os.path.join(gc.GOLDEN_DIR, "infer_left_chain.pt"))
    case = gc.make_case(cfg, "infer_left")
    out = O.multimodal_prefill(sd, ocfg, case["input_ids"], case["attention_mask"], case["images"], padding_side="left")
    toks, lg = O.greedy_decode(sd, ocfg, out["logits"][:, -1], out["kv"], out["mask"], gc.CHAIN_STEPS, stop_on_eos=False)
    assert torch.equal(toks, g["greedy_ids"]) and rel_err(lg, g["greedy_logits"]) < 2e-3
    assert all(len(set(r)) >= 8 for r in g["greedy_ids"].tolist())
    assert g["min_margin"].item() > 3.0
    # teacher forcing with the same ids is the same computation
    toks2, lg2 = O.greedy_decode(sd, ocfg, out["logits"][:, -1], out["kv"], out["mask"], gc.CHAIN_STEPS,
                                 stop_on_eos=False, forced_tokens=g["greedy_ids"])
    assert torch.equal(toks2, toks) and torch.equal(lg2, lg)


def test_token_weights_formula():
    # train/train.py:1316-1322
    import math
    w = O.token_weights({5: 100.0, 7: 10.0}, 16)
    assert abs(w[5].item() - 1 / (math.log(100.0) + 1)) < 1e-7
    assert abs(w[0].item() - w[5].item() / 100) < 1e-9


def test_training_gradients_match_reference_backward():
    """Loss and parameter gradients of the oracle under torch autograd against tests/golden/train_extras_right.pt,
    recorded from the reference model's own backward (tests/golden/make_train_golden.py): decoder, projector, pooler,
    trainable CLIP layer, audio projection, seg-mask CNN and embed_tokens."""
    import sys
    sys.path.insert(0, os.path.join(gc.GOLDEN_DIR))
    import make_train_golden as MT
    fx = torch.load(os.path.join(gc.GOLDEN_DIR, "train_extras_right.pt"))
    cfg = gc.small_config()
    sd = gc.bf16_round(gc.small_weights(cfg))
    case = gc.make_case(cfg, "train_extras_right")
    w = MT.vocab_weight(cfg)
    with torch.enable_grad():
        params = {k: sd[k].clone().requires_grad_(True) for k in MT.PROBE}
        out = O.multimodal_prefill({**sd, **params}, oracle_cfg(cfg), case["input_ids"], case["attention_mask"],
                                   case["images"], labels=case["labels"], audio=case["audio"],
                                   segmasks=case["segmasks"], padding_side="right")
        loss = O.weighted_ce(out["logits"], out["modified_labels"], w)
        loss.backward()
    assert abs(loss.item() - fx["loss"].item()) < 1e-4
    for k in MT.PROBE:
        got, ref = MT.compress(k, params[k].grad.detach()), fx["grads"][k]
        assert got.shape == ref.shape, k
        assert ((got - ref).norm() / (ref.norm() + 1e-20)).item() < 5e-4, k
