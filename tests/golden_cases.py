"""Shared definition of the small parity configuration and golden cases (used by tests/golden/make_golden.py, which
records the reference's outputs, and by the CPU / GPU parity tests, which regenerate the identical weights and
inputs from their seeds)."""
import os

import torch

from mm_or_b200.config import LlavaConfig
from mm_or_b200.synth import make_state_dict, synth_batch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WEIGHT_SEED = 7


def small_config(**kw):
    """Full-width CLIP ViT-L geometry but 3 layers (2 consumed: select_layer = -2), the reference's fixed-size
    pooler, and a 2-layer 4-head Llama (hidden 512, head_dim 128) with a 512-entry vocabulary."""
    d = dict(hidden_size=512, intermediate_size=1408, num_hidden_layers=2, num_attention_heads=4, vocab_size=512,
             max_position_embeddings=2048, mm_vision_config=dict(num_hidden_layers=3), pad_token_id=0)
    d.update(kw)
    return LlavaConfig(**d)


# the exact-token weight set (mm_or_b200.synth.make_state_dict: one-hot-per-context component in lm_head): the oracle's
# top-2 margin is > 4 logits at every step on the small configuration, against a bf16 logit error of a few 1e-2
CHAIN = dict(chain=0.5, chain_embed_scale=8.0)
CHAIN_STEPS = 24


def small_weights(cfg, peaked=0.0, dtype=torch.float32, chain=False):
    kw = CHAIN if chain else {}
    return make_state_dict(cfg, seed=WEIGHT_SEED, device="cpu", dtype=dtype, peaked_lm_head=peaked, **kw)


def bf16_round(sd):
    """Weights as the GPU build sees them (bf16), widened back to fp32 for the oracle."""
    return {k: v.to(torch.bfloat16).float() for k, v in sd.items()}


CASES = {
    # name: synth_batch kwargs + padding side + whether labels are attached
    "infer_left": dict(batch=2, views=[3, 2], text_len=24, jitter=5, image_pos=5, audio=False, segmasks=False,
                       side="left", labels=False, seed=1),
    "extras_left": dict(batch=3, views=[2, 1, 2], text_len=20, jitter=4, image_pos=3, audio=True, segmasks=True,
                        side="left", labels=False, seed=2),
    "train_right": dict(batch=2, views=[2, 3], text_len=28, jitter=6, image_pos=4, audio=False, segmasks=False,
                        side="right", labels=True, seed=3),
}


# cases without a recorded fixture (compared against the oracle only)
EXTRA_CASES = {
    "train_extras_right": dict(batch=3, views=[2, 1, 2], text_len=22, jitter=5, image_pos=4, audio=True, segmasks=True,
                               side="right", labels=True, seed=4),
}


def make_case(cfg, name, dtype=torch.float32):
    c = CASES[name] if name in CASES else EXTRA_CASES[name]
    b = synth_batch(cfg, c["batch"], max(c["views"]), c["text_len"], seed=c["seed"], jitter=c["jitter"],
                    image_pos=c["image_pos"], audio=c["audio"], segmasks=c["segmasks"], dtype=dtype)
    b["images"] = [im[:v].contiguous() for im, v in zip(b["images"], c["views"])]
    ids = b["input_ids"]
    if c["side"] == "right":   # move the left padding produced by synth_batch to the right
        out = torch.zeros_like(ids)
        for r in range(ids.shape[0]):
            row = ids[r][ids[r] != 0]
            out[r, :len(row)] = row
        ids = out
        b["input_ids"] = ids
        b["attention_mask"] = ids.ne(0)
    if c["labels"]:
        g = torch.Generator().manual_seed(c["seed"] + 99)
        lab = ids.clone()
        lab[:, :ids.shape[1] // 2] = -100                      # prompt part is ignored, answer part is supervised
        lab[ids == 0] = -100
        lab[ids == -200] = -100
        b["labels"] = lab
    b["side"] = c["side"]
    return b
