"""Parity at BASELINE.json's FULL size (configs[1]: CLIP ViT-L/14-336 with all 23 consumed layers, the 2-layer BERT
pooler, mlp2x_gelu projector, Llama-7B with 32 layers and the 32000-entry vocabulary, 6 views, bf16, random-init weights
as in bench.py) through properties that do not need the CPU oracle at that size (the oracle needs 40 s per inference
there; it is compared at reduced depth in the other files):

  * run-to-run determinism: the same call twice gives bit-identical ids and logits (no atomics, fixed reduction orders);
  * the CUDA-graph decode loop equals the eager one bit for bit;
  * batch-composition independence: a sample decoded alone (other padding, other batch size -> other tile shapes,
    split counts and cluster sizes) gives the same logits up to bf16 rounding of different summation splits, and the
    same greedy ids wherever the top-2 margin exceeds that difference;
  * the prompt (with its -200 placeholder) is echoed back in front of the new tokens, left padding included.

Needs ~45 GB of HBM (and, for the oracle comparisons at the end of the file, ~30 GB of host memory and ~2 min of CPU)."""
import pytest
import torch

from mm_or_b200.synth import make_state_dict, synth_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model7b():
    torch.set_grad_enabled(False)
    from mm_or_b200.config import LlavaConfig
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a GPU with room for the 7B model")
    cfg = LlavaConfig(num_hidden_layers=32, tokenizer_padding_side="left", mv_type="learned")
    sd = make_state_dict(cfg, seed=0, device="cuda", dtype=torch.bfloat16)
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device="cuda")
    _HOST_WEIGHTS["sd"] = {k: v.cpu() for k, v in sd.items()}         # the same bf16 values, for the CPU oracle
    del sd
    torch.cuda.empty_cache()
    yield cfg, model
    del model
    _HOST_WEIGHTS.clear()
    torch.cuda.empty_cache()


_HOST_WEIGHTS = {}


class _Widening(dict):
    """bf16 weights held once on the host; the fp32 oracle sees fp32 tensors (widened per access, so the 7B parameter
    set costs 13.5 GB of host memory, not 27)."""

    def __getitem__(self, k):
        return dict.__getitem__(self, k).float()


def oracle_weights():
    import psutil
    sd = _HOST_WEIGHTS["sd"]
    if "wide" not in _HOST_WEIGHTS:
        if psutil.virtual_memory().available > 80e9:
            _HOST_WEIGHTS["wide"] = {k: v.float() for k, v in sd.items()}      # widen once: decode steps stream it
        else:
            _HOST_WEIGHTS["wide"] = _Widening(sd)
    return _HOST_WEIGHTS["wide"]


def _oracle_compare(cfg, model, b, steps, tol):
    """GPU generate (eager loop, per-step logits) vs the fp32 CPU oracle, teacher-forced with the GPU's own ids so that
    both sides see identical inputs at every step (random-init 7B logits have near-ties). Returns the per-step errors."""
    import os
    from oracle import mm2sg_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ids, images = b["input_ids"], b["images"]
    Lin = ids.shape[1]
    out, lg = model.generate(ids, images=images, do_sample=False, use_cache=True, max_new_tokens=steps,
                             stop_on_eos=False, return_logits=True)
    gen = out[:, Lin:].cpu()
    lg = lg.float().cpu()
    sd, ocfg = oracle_weights(), O.cfg_from_llava(cfg)
    ref = O.multimodal_prefill(sd, ocfg, ids, b["attention_mask"], [im.float() for im in images], padding_side="left",
                               last_only=True)
    toks, ref_lg = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], steps, stop_on_eos=False,
                                   forced_tokens=gen)
    assert lg.shape == ref_lg.shape == (ids.shape[0], steps, cfg.vocab_size)
    rel = lambda a: [((a[:, s] - ref_lg[:, s]).norm() / ref_lg[:, s].norm()).item() for s in range(steps)]
    per_step = rel(lg)
    # The band: the SAME oracle evaluated in bf16 (the dtype the reference itself runs in, model/builder.py:43,153,166)
    # against its fp32 self, on the same inputs and forced ids. Through 23 + 2 + 32 transformer blocks at full width a
    # bf16 evaluation of the reference's algorithm is itself ~3 % away from fp32 (printed below), so the stated tolerance
    # at this size is: every step within max(tol, 1.25 x that step's bf16 band) -- no further from the fp32 truth than the
    # reference's own arithmetic is, with a quarter of slack for the different summation orders of tiled kernels.
    sdb = _HOST_WEIGHTS["sd"]
    refb = O.multimodal_prefill(sdb, ocfg, ids, b["attention_mask"], [im.to(torch.bfloat16) for im in images],
                                padding_side="left", last_only=True)
    _, refb_lg = O.greedy_decode(sdb, ocfg, refb["logits"][:, -1], refb["kv"], refb["mask"], steps, stop_on_eos=False,
                                 forced_tokens=gen)
    band = rel(refb_lg.float())
    print("per-step logit rel err vs fp32 oracle: B200", [round(e, 4) for e in per_step],
          "| bf16 oracle (band)", [round(e, 4) for e in band])
    for s in range(steps):
        assert per_step[s] < max(tol, 1.25 * band[s]), (s, per_step[s], band[s])
    assert max(per_step) < 5e-2
    err = (lg - ref_lg).abs().max().item()
    top2 = ref_lg.topk(2, -1).values
    safe = (top2[..., 0] - top2[..., 1]) > 2 * err
    assert torch.equal(gen[safe], toks[safe])                 # same greedy id wherever the margin allows a verdict
    return ref, per_step, float(safe.float().mean())


def test_full_size_determinism_graph_and_batch_independence(model7b):
    cfg, model = model7b
    B, V, steps = 3, 6, 6
    b = synth_batch(cfg, B, V, 64, seed=91, jitter=9, image_pos=40, dtype=torch.bfloat16)
    ids, images = b["input_ids"], b["images"]
    Lin = ids.shape[1]
    kw = dict(do_sample=False, use_cache=True, max_new_tokens=steps, stop_on_eos=False)
    out1, lg1 = model.generate(ids, images=images, return_logits=True, **kw)              # eager decode loop
    out2, lg2 = model.generate(ids, images=images, return_logits=True, **kw)
    assert torch.equal(out1, out2) and torch.equal(lg1, lg2)                               # deterministic
    assert torch.isfinite(lg1).all()
    out_graph = model.generate(ids, images=images, **kw)                                    # CUDA-graph decode loop
    assert torch.equal(out_graph, out1)
    assert out1.shape == (B, Lin + steps) and torch.equal(out1[:, :Lin].cpu(), ids)         # prompt echoed, pads kept
    assert int((out1[:, Lin:] < 0).sum()) == 0 and int(out1[:, Lin:].max()) < cfg.vocab_size
    # every sample alone: its own (shorter) left padding, batch 1 => other tiles / splits / cluster sizes
    for r in range(B):
        n = int(b["attention_mask"][r].sum())
        solo_ids = ids[r:r + 1, Lin - n:]
        solo, lg = model.generate(solo_ids, images=[images[r]], return_logits=True, **kw)
        ref = lg1[r:r + 1].float()
        err = (lg.float() - ref).abs().max().item()
        scale = ref.abs().max().item()
        # two bf16 evaluations of a 32-layer decoder with different summation splits: each is within TOL_E2E of the
        # exact result (test_config1_full_size_against_oracle), so their difference is bounded by twice that; measured
        # on B200: 2.6 % of the largest logit on one element, 1.x % relative Frobenius
        assert ((lg.float() - ref).norm() / ref.norm()).item() < 3e-2, r
        assert err < 5e-2 * scale, (r, err, scale)
        top2 = ref.topk(2, -1).values
        safe = ((top2[..., 0] - top2[..., 1]) > 2 * err)[0].cpu()
        assert torch.equal(solo[0, n:].cpu()[safe], out1[r, Lin:].cpu()[safe])


def test_config1_full_size_against_oracle(model7b):
    """BASELINE configs[1] AT ITS REAL SIZE against the CPU oracle in fp32: 23-layer ViT-L over 2 x 6 views, pooler,
    projector, pack to L ~ 831 with +-16 tokens of jitter (left padding), 32-layer 7B prefill, then 8 teacher-forced
    decode steps -- the tile shapes, cluster split-K sizes and depths that bench.py times. Tolerance: TOL_STAGE on the
    projected visual tokens, TOL_E2E (3e-2, relative Frobenius) on the last-position logits of the prefill and of
    every decode step -- widened, where the bf16 evaluation of the oracle itself (the reference's dtype) is further than
    that from fp32, to 1.25 x that band (see _oracle_compare); greedy ids equal wherever the oracle's top-2 margin exceeds
    twice the logit error."""
    from helpers import TOL_E2E, TOL_STAGE, rel_err
    cfg, model = model7b
    b = synth_batch(cfg, 2, 6, 256, seed=93, jitter=16, image_pos=40, dtype=torch.bfloat16)
    ref, per_step, safe_frac = _oracle_compare(cfg, model, b, 9, TOL_E2E)
    assert 815 <= ref["mask"].shape[1] <= 831 and ref["visual"].shape[1] == 576
    pooled = model.encode_images_pooled(torch.cat(b["images"], 0).cuda(), [6, 6], None, None, None)
    visual = model.get_model().mm_projector(pooled)
    assert rel_err(visual, ref["visual"]) < TOL_STAGE
    print("configs[1] full size: per-step logit rel err", [round(e, 4) for e in per_step], "decidable ids", safe_frac)


def test_config0_full_size_against_oracle(model7b):
    """BASELINE configs[0] at its real size: one 336x336 frame (list form, one view), 256 prompt tokens, greedy 32 new
    tokens, batch 1 -- every step's logits against the fp32 CPU oracle (teacher-forced with the GPU's ids)."""
    from helpers import TOL_E2E
    cfg, model = model7b
    b = synth_batch(cfg, 1, 1, 256, seed=94, jitter=0, image_pos=40, dtype=torch.bfloat16)
    ref, per_step, safe_frac = _oracle_compare(cfg, model, b, 32, TOL_E2E)
    assert ref["mask"].shape[1] == 831
    print("configs[0] full size: max per-step logit rel err", round(max(per_step), 4), "decidable ids", safe_frac)
