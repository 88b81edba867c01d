"""Records the REFERENCE model's training-side outputs: loss and parameter gradients of the reference's own
LlavaLlamaForCausalLM (imported unmodified through oracle/ref_shim.py) under torch autograd, for a right-padded batch
with audio and seg-masks and the class-weighted CE of LLaVATrainer.compute_loss (train/llava_trainer.py:143-167,
restated here on the reference's logits / modified_labels because the Trainer object itself needs a dataset).
-> tests/golden/train_extras_right.pt. Pins the oracle's backward (what the GPU fine-tune step is compared with) to the
reference's backward, including the modality encoders of the image pooler (trainable per train/train.py:1257-1261).

Run in the build container only:   python tests/golden/make_train_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch

import golden_cases as gc
import make_golden as MG
from oracle import mm2sg_oracle as O
from oracle.ref_shim import build_reference_model

# a spread of parameter kinds: decoder, projector, pooler BERT, CLIP (trainable layer), audio, seg-mask CNN, embeddings
PROBE = ["model.layers.1.self_attn.q_proj.weight", "model.layers.0.mlp.down_proj.weight", "model.norm.weight",
         "lm_head.weight", "model.embed_tokens.weight", "model.mm_projector.0.weight", "model.mm_projector.2.bias",
         "model.image_pooler.bert.encoder.layer.1.attention.self.value.weight",
         "model.image_pooler.bert.embeddings.position_embeddings.weight",
         "model.vision_tower.vision_tower.vision_model.encoder.layers.1.mlp.fc1.weight",
         "model.image_pooler.project_audio.weight", "model.image_pooler.project_audio.bias",
         "model.image_pooler.segmasks_encoder.embedding.weight", "model.image_pooler.segmasks_encoder.conv1.weight",
         "model.image_pooler.segmasks_encoder.conv3.bias", "model.image_pooler.segmasks_encoder.conv5.weight"]


def vocab_weight(cfg):
    return torch.linspace(0.2, 1.0, cfg.vocab_size)


def weighted_loss(logits, modified_labels, w):
    sl = logits[..., :-1, :].reshape(-1, logits.shape[-1]).float()
    tl = modified_labels[..., 1:].reshape(-1)
    return torch.nn.functional.cross_entropy(sl, tl, weight=w)          # llava_trainer.py:156-167


def compress(name, g):
    """Keep fixtures small: full tensor for small parameters, a strided sample of about 8 k elements for matrices."""
    if g.numel() <= 8192:
        return g.clone()
    g2 = g.reshape(g.shape[0], -1)
    step = max(2, int((g2.numel() / 8192) ** 0.5) + 1)
    return g2[::step, ::step + 1].clone()


def main():
    cfg = gc.small_config()
    ocfg = MG.oracle_cfg(cfg)
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = build_reference_model(cfg, sd)
    model.config.tokenizer_padding_side = "right"
    model.config.mv_type = "learned"
    for p in model.parameters():
        p.requires_grad_(True)
    case = gc.make_case(cfg, "train_extras_right")
    w = vocab_weight(cfg)
    with torch.enable_grad():
        ref = model(input_ids=case["input_ids"], attention_mask=case["attention_mask"], labels=case["labels"],
                    images=case["images"], audio=case["audio"], segmasks=case["segmasks"])
        loss = weighted_loss(ref.logits, ref["modified_labels"], w)
        loss.backward()
    named = dict(model.named_parameters())
    ref_grads = {k: named[k].grad.detach().float() for k in PROBE}
    # the oracle's autograd on the same batch
    with torch.enable_grad():
        params = {k: sd[k].clone().requires_grad_(True) for k in PROBE}
        orc = O.multimodal_prefill({**sd, **params}, ocfg, case["input_ids"], case["attention_mask"], case["images"],
                                   labels=case["labels"], audio=case["audio"], segmasks=case["segmasks"],
                                   padding_side="right")
        oloss = O.weighted_ce(orc["logits"], orc["modified_labels"], w)
        oloss.backward()
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-20)).item()
    errs = {k: rel(params[k].grad, ref_grads[k]) for k in PROBE}
    print("loss reference %.6f oracle %.6f" % (loss.item(), oloss.item()))
    for k, e in errs.items():
        print("  %-90s %.2e  |g| %.3e" % (k, e, ref_grads[k].norm().item()))
    assert abs(loss.item() - oloss.item()) < 1e-4 and max(errs.values()) < 5e-4, errs
    torch.save({"loss": loss.detach().float(), "grads": {k: compress(k, g) for k, g in ref_grads.items()},
                "grad_norms": {k: g.norm() for k, g in ref_grads.items()}},
               os.path.join(gc.GOLDEN_DIR, "train_extras_right.pt"))


if __name__ == "__main__":
    main()
