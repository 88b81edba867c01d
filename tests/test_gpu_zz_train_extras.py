"""GPU: seg-mask CNN fine-tune kernels and the embed_tokens gradient (csrc/train_extras.cu) through the C ABI against
torch autograd over the oracle (same checks as the CPU emulator run)."""
import pytest
import torch

import train_extras_checks as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend():
    from mm_or_b200 import _lib as L
    torch.cuda.set_device(0)
    return dict(cdll=L.lib(), device=torch.device("cuda:0"), stream_fn=L.stream_ptr, ptr_fn=L.ptr)


@pytest.mark.parametrize("check", C.ALL, ids=lambda f: f.__name__[6:])
def test_train_extras(backend, check):
    from mm_or_b200 import _lib as L
    n0 = L.launch_count()
    check(backend)
    torch.cuda.synchronize()
    assert L.launch_count() > n0


def test_fine_tune_step_with_extra_modalities():
    """Fine-tune step with audio + seg-mask (+ frozen point-cloud) tokens and a trainable embed_tokens: loss and the
    gradients of project_audio, the seg-mask CNN and embed_tokens vs torch autograd over the oracle; parameters of a
    modality that is absent from the batch are skipped by the optimizer step."""
    import golden_cases as gc
    from helpers import oracle_cfg, rel_err
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train import llama as T
    from mm_or_b200.train.step import FineTuner
    from oracle import mm2sg_oracle as O
    from oracle import ptv3_oracle as P
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.small_weights(cfg)
    sd.update(P.synth_weights())
    sd = gc.bf16_round(sd)
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    case = gc.make_case(cfg, "train_extras_right")
    pcs = P.dedupe_clouds([None, P.synth_cloud(400, seed=41, box=(16, 16, 3)), None])
    w = torch.rand(cfg.vocab_size, generator=torch.Generator().manual_seed(9)) + 0.05
    ft = FineTuner(model, sd, lr=1e-3, max_grad_norm=0.1, first_trainable_clip_layer=1, vocab_weight=w,
                   train_embed_tokens=True)
    assert "model.embed_tokens.weight" in ft.names and "model.image_pooler.project_audio.weight" in ft.names
    assert not any("point_transformer" in k for k in ft.names)
    torch.manual_seed(3)
    loss, wsum, grads = ft.forward_backward(case["input_ids"], case["labels"], case["attention_mask"], case["images"],
                                            pc=pcs, audio=case["audio"], segmasks=case["segmasks"])
    named = T.unfuse_grads(grads, cfg)
    names = [k for k in ft.names if "project_audio" in k or "segmasks_encoder" in k or k == "model.embed_tokens.weight"
             or k.startswith("model.mm_projector.")]
    with torch.enable_grad():
        params = {k: sd[k].clone().float().requires_grad_(True) for k in names}
        torch.manual_seed(3)
        ref = O.multimodal_prefill({**sd, **params}, oracle_cfg(cfg), case["input_ids"], case["attention_mask"],
                                   case["images"], labels=case["labels"], audio=case["audio"],
                                   segmasks=case["segmasks"], pc=pcs, padding_side="right")
        ref_loss = O.weighted_ce(ref["logits"], ref["modified_labels"], w)
        ref_loss.backward()
    assert abs(float(loss) - float(ref_loss.detach())) < 3e-2 * abs(float(ref_loss.detach()))
    bad = [(k, rel_err(named[k], params[k].grad)) for k in names
           if float(params[k].grad.norm()) > 1e-7 and rel_err(named[k], params[k].grad) > 0.12]
    assert not bad, bad
    # a step without audio / seg-masks leaves their parameters untouched and still updates the rest
    before = {k: ft.master[k].clone() for k in ft.names if "project_audio" in k or "segmasks_encoder" in k}
    ft.train_step(case["input_ids"], case["labels"], case["attention_mask"], case["images"])
    assert all(torch.equal(before[k], ft.master[k]) for k in before)
    ft.train_step(case["input_ids"], case["labels"], case["attention_mask"], case["images"], audio=case["audio"],
                  segmasks=case["segmasks"])
    assert any(not torch.equal(before[k], ft.master[k]) for k in before)


def test_fine_tune_gradients_against_reference_backward_fixture():
    """Loss and gradients of the fine-tune step against tests/golden/train_extras_right.pt, recorded from the
    reference model's own backward (tests/golden/make_train_golden.py). Written after the round's GPU budget was
    spent: same code path as test_fine_tune_step_with_extra_modalities (which ran green), not yet run itself."""
    import os
    import sys
    import golden_cases as gc
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train import llama as T
    from mm_or_b200.train.step import FineTuner
    sys.path.insert(0, gc.GOLDEN_DIR)
    import make_train_golden as MT
    fx = torch.load(os.path.join(gc.GOLDEN_DIR, "train_extras_right.pt"))
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    case = gc.make_case(cfg, "train_extras_right")
    ft = FineTuner(model, sd, lr=1e-3, max_grad_norm=0.1, first_trainable_clip_layer=1, vocab_weight=MT.vocab_weight(cfg),
                   train_embed_tokens=True)
    loss, wsum, grads = ft.forward_backward(case["input_ids"], case["labels"], case["attention_mask"], case["images"],
                                            audio=case["audio"], segmasks=case["segmasks"])
    named = T.unfuse_grads(grads, cfg)
    assert abs(float(loss) - float(fx["loss"])) < 3e-2 * abs(float(fx["loss"]))
    bad = []
    for k in MT.PROBE:
        if k not in named:                       # frozen in this configuration (none of the probes should be)
            bad.append((k, "missing"))
            continue
        got, ref = MT.compress(k, named[k].float().cpu()), fx["grads"][k]
        err = ((got - ref).norm() / (ref.norm() + 1e-20)).item()
        if err > 0.12:                            # bf16 pipeline vs fp32 reference, as in the oracle comparison
            bad.append((k, err))
    assert not bad, bad

