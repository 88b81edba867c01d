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

IGNORE, IMAGE, DESCRIPTOR = -100, -200, 18610
PAD_ROW, VISUAL_BASE, DESC_BASE = -1, -2, -(1 << 20)   # mm_or_b200/model/pack.py encoding of a decoded row


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


def descriptor_cases(vocab, D):
    """Batches whose prompts carry VIS_DESCRIPTOR placeholders (llava_arch.py:243,253,278-294) together with
    vis_descriptor_embs: fewer descriptors than placeholders (dummy zero row), more than placeholders (ignored), 2-D
    descriptors (several rows), a text-only row, truncation, both padding sides, the bare-list form for a batch of 1."""
    rng = np.random.default_rng(23)
    out = []
    for side in ("left", "right"):
        for with_labels in (False, True):
            for max_len in (None, 30):
                for trial in range(3):
                    B = 1 if trial == 2 else 3
                    Lt = 24
                    ids = torch.zeros(B, Lt, dtype=torch.long)
                    embs = []
                    for b in range(B):
                        n = int(rng.integers(8, Lt + 1))
                        row = torch.from_numpy(rng.integers(3, vocab, n))
                        text_only = trial == 1 and b == 0
                        n_desc = int(rng.integers(0, 4))
                        if not text_only:
                            spots = rng.choice(n, size=1 + n_desc, replace=False)
                            row[int(spots[0])] = IMAGE
                            for sp in spots[1:]:
                                row[int(sp)] = DESCRIPTOR
                        ids[b, Lt - n:] = row
                        n_emb = max(0, n_desc + int(rng.integers(-1, 2)))
                        per, r0 = [], 0
                        for j in range(n_emb):
                            k = 1 if rng.integers(0, 2) == 0 else int(rng.integers(2, 4))
                            e = torch.zeros(k, D)
                            e[:, 0] = torch.arange(r0, r0 + k, dtype=torch.float32)
                            e[:, 1] = 3.0                           # flag: descriptor row
                            e[:, 2] = float(b)
                            r0 += k
                            per.append(e[0] if k == 1 and rng.integers(0, 2) == 0 else e)   # 1-D and (1, D) forms
                        embs.append(per)
                    labels = None
                    if with_labels:
                        labels = ids.clone()
                        labels[(ids <= 0) | (ids == DESCRIPTOR)] = IGNORE
                    out.append(dict(side=side, max_len=max_len, ids=ids, mask=ids.ne(0), labels=labels,
                                    t_vis=int(rng.integers(1, 9)),
                                    embs=embs[0] if B == 1 and len(embs[0]) > 0 else embs))   # bare list: :279-280
    return out


def multi_image_cases(vocab):
    """Rows with SEVERAL <image> placeholders: the reference walks `image_features` with one running index over the
    batch (llava_arch.py:239,245,264-265), so `images` holds more view groups than there are rows. Also a text-only
    row (it skips a block) and a batch that asks for more blocks than there are (the reference's IndexError)."""
    rng = np.random.default_rng(37)
    out = []
    for side in ("left", "right"):
        for with_labels in (False, True):
            for max_len in (None, 36):
                for trial in range(3):
                    B, Lt = 3, 18
                    ids = torch.zeros(B, Lt, dtype=torch.long)
                    need = 0
                    for b in range(B):
                        n = int(rng.integers(6, Lt + 1))
                        row = torch.from_numpy(rng.integers(3, vocab, n))
                        n_img = 0 if (trial == 1 and b == 1) else int(rng.integers(1, 4))
                        for sp in rng.choice(n, size=n_img, replace=False):
                            row[int(sp)] = IMAGE
                        need += max(n_img, 1)
                        ids[b, Lt - n:] = row
                    labels = None
                    if with_labels:
                        labels = ids.clone()
                        labels[ids <= 0] = IGNORE
                    out.append(dict(side=side, max_len=max_len, ids=ids, mask=ids.ne(0), labels=labels,
                                    t_vis=int(rng.integers(1, 7)), n_blocks=need - (1 if trial == 2 else 0)))
    return out


def record(model, c, D):
    B, t_vis = c["ids"].shape[0], c["t_vis"]
    NB = c.get("n_blocks", B)                              # entries of `images` = blocks of image features
    feats = torch.zeros(NB, t_vis, D)
    feats[:, :, 0] = torch.arange(t_vis, dtype=torch.float32)[None, :]
    feats[:, :, 1] = 2.0                                   # flag: visual row
    feats[:, :, 2] = torch.arange(NB, dtype=torch.float32)[:, None]
    model.encode_images_pooled = lambda *a, _f=feats: _f   # instance attribute shadows the method
    model.config.tokenizer_padding_side = c["side"]
    model.config.tokenizer_model_max_length = c["max_len"]
    pos_in = torch.arange(c["ids"].shape[1])[None].expand(B, -1).clone()
    images = [torch.zeros(1, 3, 2, 2) for _ in range(NB)]
    try:
        _, pos, am, _, emb, lab = model.prepare_inputs_labels_for_multimodal(
            c["ids"], pos_in, c["mask"], None, c["labels"], images, c.get("embs"), None, None, None)
    except IndexError as e:                                 # more placeholders than feature blocks
        return dict(side=c["side"], max_len=c["max_len"], t_vis=t_vis, ids=c["ids"], mask=c["mask"],
                    labels=c["labels"], n_blocks=NB, raises="IndexError", message=str(e))
    flag, val, samp = emb[..., 1].round().long(), emb[..., 0].round().long(), emb[..., 2].round().long()
    src = torch.where(flag == 1, val, torch.where(flag == 2, VISUAL_BASE - val,
                      torch.where(flag == 3, DESC_BASE - val, torch.full_like(val, PAD_ROW))))
    if "n_blocks" not in c:
        assert bool(((flag < 2) | (samp == torch.arange(B)[:, None])).all())   # a row only holds its own visuals
    rec = dict(side=c["side"], max_len=c["max_len"], t_vis=t_vis, ids=c["ids"], mask=c["mask"],
               labels=c["labels"], src=src.to(torch.int32), out_labels=lab, out_mask=am.bool(), out_pos=pos)
    if "n_blocks" in c:
        rec["n_blocks"] = NB
        rec["blocks"] = torch.where(flag == 2, samp, torch.full_like(samp, -1)).to(torch.int32)
    if "embs" in c:
        bare = type(c["embs"][0]) is not list
        embs = [c["embs"]] if bare else c["embs"]
        rec["desc_rows"] = [[1 if e.ndim == 1 else int(e.shape[0]) for e in per] for per in embs]
        rec["bare_list"] = bare
    return rec


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
    records = [record(model, c, D) for c in cases(V)]
    multi_records = [record(model, c, D) for c in multi_image_cases(V)]
    del model.encode_images_pooled
    # descriptor cases need hidden 4096: the reference's dummy descriptor is a hard-coded zeros(4096) (llava_arch.py:286)
    import contextlib
    import io
    cfg4 = gc.small_config(hidden_size=4096, intermediate_size=64, num_hidden_layers=1, num_attention_heads=32)
    model = build_reference_model(cfg4, gc.bf16_round(gc.small_weights(cfg4)))
    model.config.mv_type = "learned"
    D4 = cfg4.hidden_size
    table = torch.zeros(V, D4)
    table[:, 0] = torch.arange(V, dtype=torch.float32)
    table[:, 1] = 1.0
    model.get_model().embed_tokens.weight.data.copy_(table)
    with contextlib.redirect_stdout(io.StringIO()):            # the reference prints when it substitutes a dummy row
        desc_records = [record(model, c, D4) for c in descriptor_cases(V, D4)]
    del model.encode_images_pooled
    torch.save(records, os.path.join(gc.GOLDEN_DIR, "pack_cases.pt"))
    torch.save(desc_records, os.path.join(gc.GOLDEN_DIR, "pack_desc_cases.pt"))
    torch.save(multi_records, os.path.join(gc.GOLDEN_DIR, "pack_multi_image_cases.pt"))
    print(len(records), "+", len(desc_records), "+", len(multi_records), "cases recorded;",
          sum("raises" in r for r in multi_records), "of the multi-image cases raise IndexError in the reference")


if __name__ == "__main__":
    main()
