"""Records what the REFERENCE's own prepare_inputs_labels_for_multimodal (model/llava_arch.py:188-353, imported
unmodified through oracle/ref_shim.py) does with ragged, interior-padded, truncated, left- and right-padded batches:
-> tests/golden/pack_cases.pt (integers only; the bar is bit-exact).

The vision side is replaced by marker features on the reference INSTANCE (encode_images_pooled returns vectors that
spell (sample, token index)), and embed_tokens is loaded with a table that spells the token id, so that the returned
inputs_embeds can be decoded back into "which source row landed where" without running a ViT.

Run in the build container only:   python tests/golden/make_pack_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import golden_cases as gc
from oracle.ref_shim import build_reference_model

IGNORE, IMAGE = -100, -200
PAD_ROW, VISUAL_BASE = -1, -2            # mm_or_b200/model/pack.py encoding of a decoded row


def random_batch(rng, B, Lt, with_labels, interior_pad, text_only_row, vocab):
    ids = torch.zeros(B, Lt, dtype=torch.long)
    for b in range(B):
        n = int(rng.integers(3, Lt + 1))
        row = torch.from_numpy(rng.integers(3, vocab, n))
        if not (text_only_row and b == 0):
            row[int(rng.integers(0, n))] = IMAGE
        if interior_pad and n > 4:
            row[2] = 0
        ids[b, Lt - n:] = row
    labels = None
    if with_labels:
        labels = ids.clone()
        labels[ids <= 0] = IGNORE
    return ids, ids.ne(0), labels


def cases(vocab):
    rng = np.random.default_rng(11)
    out = []
    for side in ("left", "right"):
        for with_labels in (False, True):
            for max_len in (None, 40):
                for trial in range(4):
                    ids, mask, labels = random_batch(rng, 4, 20, with_labels, trial % 2 == 1, trial == 3, vocab)
                    out.append(dict(side=side, max_len=max_len, ids=ids, mask=mask, labels=labels,
                                    t_vis=int(rng.integers(1, 33))))
    return out


def main():
    torch.set_grad_enabled(False)
    cfg = gc.small_config()
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = build_reference_model(cfg, sd)
    model.config.mv_type = "learned"
    D, V = cfg.hidden_size, cfg.vocab_size
    table = torch.zeros(V, D)
    table[:, 0] = torch.arange(V, dtype=torch.float32)
    table[:, 1] = 1.0                                          # flag: text row
    model.get_model().embed_tokens.weight.data.copy_(table)
    records = []
    for c in cases(V):
        B, t_vis = c["ids"].shape[0], c["t_vis"]
        feats = torch.zeros(B, t_vis, D)
        feats[:, :, 0] = torch.arange(t_vis, dtype=torch.float32)[None, :]
        feats[:, :, 1] = 2.0                                   # flag: visual row
        feats[:, :, 2] = torch.arange(B, dtype=torch.float32)[:, None]
        model.encode_images_pooled = lambda *a, _f=feats: _f   # instance attribute shadows the method
        model.config.tokenizer_padding_side = c["side"]
        model.config.tokenizer_model_max_length = c["max_len"]
        pos_in = torch.arange(c["ids"].shape[1])[None].expand(B, -1).clone()
        images = [torch.zeros(1, 3, 2, 2) for _ in range(B)]
        _, pos, am, _, emb, lab = model.prepare_inputs_labels_for_multimodal(
            c["ids"], pos_in, c["mask"], None, c["labels"], images, None, None, None, None)
        flag, val, samp = emb[..., 1].round().long(), emb[..., 0].round().long(), emb[..., 2].round().long()
        src = torch.where(flag == 1, val, torch.where(flag == 2, VISUAL_BASE - val, torch.full_like(val, PAD_ROW)))
        assert bool(((flag != 2) | (samp == torch.arange(B)[:, None])).all())       # a row only holds its own visuals
        records.append(dict(side=c["side"], max_len=c["max_len"], t_vis=t_vis, ids=c["ids"], mask=c["mask"],
                            labels=c["labels"], src=src.to(torch.int32), out_labels=lab, out_mask=am.bool(),
                            out_pos=pos))
    del model.encode_images_pooled
    torch.save(records, os.path.join(gc.GOLDEN_DIR, "pack_cases.pt"))
    print(len(records), "cases recorded")


if __name__ == "__main__":
    main()
