"""Gradient accumulation of the fine-tune step (reference recipe: gradient_accumulation_steps 4, README.md:143).
Written without GPU access; its own file, late in the suite, so that a surprise here cannot hide other results."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_gradient_accumulation_matches_one_batch():
    """gradient_accumulation_steps (reference recipe: 4, README.md:143): the same micro-batch fed twice with the loss
    gradient scaled by 1/2 must give the single-batch gradients -- 0.5 g + 0.5 g is exact in binary floating point, so
    the comparison is tight -- and a micro-batch WITHOUT audio / seg-masks after one with them must leave those
    gradients as the first micro-batch produced them (scaled). Not yet run on hardware (written without GPU access)."""
    import golden_cases as gc
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train.step import FineTuner
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    case = gc.make_case(cfg, "train_extras_right")
    args = (case["input_ids"], case["labels"], case["attention_mask"], case["images"])
    extras = dict(audio=case["audio"], segmasks=case["segmasks"])
    ft = FineTuner(model, sd, lr=1e-3, max_grad_norm=0.1, first_trainable_clip_layer=1, train_embed_tokens=True)
    loss1, _, g1 = ft.forward_backward(*args, **extras)
    g1 = {k: v.clone() for k, v in g1.items()}
    _, _, g2 = ft.forward_backward(*args, grad_scale=0.5, **extras)
    loss2, _, g2 = ft.forward_backward(*args, grads=g2, accumulate=True, grad_scale=0.5, **extras)
    assert float(loss1) == float(loss2) and sorted(g1) == sorted(g2)
    for k in g1:
        assert torch.allclose(g2[k].float(), g1[k].float(), rtol=1e-4, atol=1e-9), k
    # second micro-batch without the extra modalities: their gradients stay those of the first one
    _, _, g3 = ft.forward_backward(*args, grad_scale=0.5, **extras)
    audio_half = g3["model.image_pooler.project_audio.weight"].clone()
    _, _, g3 = ft.forward_backward(*args, grads=g3, accumulate=True, grad_scale=0.5)
    assert torch.equal(g3["model.image_pooler.project_audio.weight"], audio_half)
    assert torch.allclose(audio_half.float(), 0.5 * g1["model.image_pooler.project_audio.weight"].float(), rtol=1e-4,
                          atol=1e-9)
    # and the public entry point steps once over both micro-batches
    mb = dict(input_ids=args[0], labels=args[1], attention_mask=args[2], images=args[3], **extras)
    before = {k: ft.master[k].clone() for k in ft.names}
    loss, out2 = ft.train_step_accumulated([mb, mb])
    assert ft.step_count == 1 and abs(float(loss) - float(loss1)) < 1e-5 * abs(float(loss1))
    assert any(not torch.equal(before[k], ft.master[k]) for k in before)


@pytest.mark.parametrize("with_lora", [False, True])
def test_activation_recomputation_is_bit_identical(with_lora):
    """recompute_activations (the reference's gradient checkpointing, train.py:1148): the backward re-runs every decoder
    layer's forward from its saved input with the same kernels and the same dropout seeds, so loss and every gradient
    equal the stored-activation step bit for bit -- full fine-tuning and the LoRA recipe with adapter dropout."""
    import golden_cases as gc
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train.lora import LoraState
    from mm_or_b200.train.step import FineTuner
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.bf16_round(gc.small_weights(cfg))
    case = gc.make_case(cfg, "train_right")
    args = (case["input_ids"], case["labels"], case["attention_mask"], case["images"])
    out = []
    for recompute in (False, True):
        model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
        lora = LoraState(cfg, r=8, alpha=16, device="cuda", seed=5, init_b="random", dropout=0.1) if with_lora else None
        ft = FineTuner(model, sd, lr=1e-3, max_grad_norm=0.1, first_trainable_clip_layer=1, lora=lora,
                       recompute_activations=recompute)
        loss, _, g = ft.forward_backward(*args)
        out.append((float(loss), {k: v.clone() for k, v in g.items()}))
    (l0, g0), (l1, g1) = out
    assert l0 == l1 and sorted(g0) == sorted(g1) and len(g0) > 10
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k
