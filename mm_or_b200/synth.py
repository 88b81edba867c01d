"""Seeded synthetic weights and inputs for MM2SG (there is no network for checkpoints or datasets).

`make_state_dict` produces tensors under the reference's state_dict names, so the same
dictionary can be loaded into the reference model (tests/golden/make_golden.py), the CPU oracle and this package.
Generation order and the per-tensor seeds are fixed: tensor `name` is drawn from a generator seeded with
hash(seed, name), which makes every tensor independent of which other tensors are requested.
"""
import zlib

import torch

from .config import LlavaConfig

VIT = "model.vision_tower.vision_tower.vision_model."
POOL = "model.image_pooler."

POOLER_GEOMETRY = dict(hidden=1024, heads=8, layers=2, ffn=4096, max_pos=576 * 7, keep=576, eps=1e-12)
SEG_CHANNELS = [8, 64, 128, 256, 512, 1024]


def _gen(seed, name, device):
    g = torch.Generator(device=device)
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(shape, std, seed, name, device, dtype, mean=0.0):
    g = _gen(seed, name, device)
    t = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std + mean
    return t.to(dtype)


def weight_specs(config: LlavaConfig, include_lm=True, include_vision=True):
    """[(name, shape, kind)] with kind in {'w' (linear/conv/embedding), 'g' (norm gain), 'b' (bias)}"""
    D, F, V = config.hidden_size, config.intermediate_size, config.vocab_size
    specs = []
    if include_lm:
        specs.append(("model.embed_tokens.weight", (V, D), "w"))
        for i in range(config.num_hidden_layers):
            p = f"model.layers.{i}."
            for n in ("q", "k", "v", "o"):
                specs.append((p + f"self_attn.{n}_proj.weight", (D, D), "w"))
            specs.append((p + "mlp.gate_proj.weight", (F, D), "w"))
            specs.append((p + "mlp.up_proj.weight", (F, D), "w"))
            specs.append((p + "mlp.down_proj.weight", (D, F), "w"))
            specs.append((p + "input_layernorm.weight", (D,), "g"))
            specs.append((p + "post_attention_layernorm.weight", (D,), "g"))
        specs.append(("model.norm.weight", (D,), "g"))
        specs.append(("lm_head.weight", (V, D), "w"))
    if include_vision:
        vc = config.vision_config()
        d, f, P = vc["hidden_size"], vc["intermediate_size"], vc["patch_size"]
        T = (vc["image_size"] // P) ** 2 + 1
        specs += [(VIT + "embeddings.class_embedding", (d,), "w"),
                  (VIT + "embeddings.patch_embedding.weight", (d, 3, P, P), "w"),
                  (VIT + "embeddings.position_embedding.weight", (T, d), "w"),
                  (VIT + "pre_layrnorm.weight", (d,), "g"), (VIT + "pre_layrnorm.bias", (d,), "b")]
        for i in range(vc["num_hidden_layers"]):
            p = VIT + f"encoder.layers.{i}."
            for n in ("layer_norm1", "layer_norm2"):
                specs += [(p + n + ".weight", (d,), "g"), (p + n + ".bias", (d,), "b")]
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                specs += [(p + f"self_attn.{n}.weight", (d, d), "w"), (p + f"self_attn.{n}.bias", (d,), "b")]
            specs += [(p + "mlp.fc1.weight", (f, d), "w"), (p + "mlp.fc1.bias", (f,), "b"),
                      (p + "mlp.fc2.weight", (d, f), "w"), (p + "mlp.fc2.bias", (d,), "b")]
        specs += [(VIT + "post_layernorm.weight", (d,), "g"), (VIT + "post_layernorm.bias", (d,), "b")]
        g = POOLER_GEOMETRY
        h, hf = g["hidden"], g["ffn"]
        b = POOL + "bert."
        specs += [(b + "embeddings.word_embeddings.weight", (1, h), "w"),
                  (b + "embeddings.position_embeddings.weight", (g["max_pos"], h), "w"),
                  (b + "embeddings.token_type_embeddings.weight", (2, h), "w"),
                  (b + "embeddings.LayerNorm.weight", (h,), "g"), (b + "embeddings.LayerNorm.bias", (h,), "b")]
        for i in range(g["layers"]):
            p = b + f"encoder.layer.{i}."
            for n in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
                specs += [(p + n + ".weight", (h, h), "w"), (p + n + ".bias", (h,), "b")]
            specs += [(p + "attention.output.LayerNorm.weight", (h,), "g"), (p + "attention.output.LayerNorm.bias", (h,), "b"),
                      (p + "intermediate.dense.weight", (hf, h), "w"), (p + "intermediate.dense.bias", (hf,), "b"),
                      (p + "output.dense.weight", (h, hf), "w"), (p + "output.dense.bias", (h,), "b"),
                      (p + "output.LayerNorm.weight", (h,), "g"), (p + "output.LayerNorm.bias", (h,), "b")]
        specs += [(b + "pooler.dense.weight", (h, h), "w"), (b + "pooler.dense.bias", (h,), "b")]
        specs += [(POOL + "project_audio.weight", (h, 512), "w"), (POOL + "project_audio.bias", (h,), "b")]
        specs += [(POOL + "segmasks_encoder.embedding.weight", (30, 8), "e")]
        for i in range(5):
            specs += [(POOL + f"segmasks_encoder.conv{i + 1}.weight", (SEG_CHANNELS[i + 1], SEG_CHANNELS[i], 3, 3), "c"),
                      (POOL + f"segmasks_encoder.conv{i + 1}.bias", (SEG_CHANNELS[i + 1],), "b")]
        specs += [("model.mm_projector.0.weight", (D, config.mm_hidden_size), "w"), ("model.mm_projector.0.bias", (D,), "b"),
                  ("model.mm_projector.2.weight", (D, D), "w"), ("model.mm_projector.2.bias", (D,), "b")]
    return specs


def make_tensor(name, shape, kind, seed, device, dtype):
    if kind == "w":
        return _normal(shape, 0.02, seed, name, device, dtype)
    if kind == "e":
        return _normal(shape, 1.0, seed, name, device, dtype)
    if kind == "c":  # conv: keep activations O(1) through 5 ReLU layers
        fan_in = shape[1] * shape[2] * shape[3]
        return _normal(shape, (2.0 / fan_in) ** 0.5, seed, name, device, dtype)
    if kind == "g":
        return _normal(shape, 0.1, seed, name, device, dtype, mean=1.0)
    return _normal(shape, 0.02, seed, name, device, dtype)


def chain_successor(vocab):
    """The fixed successor map of the `chain` weight set: t -> (t + stride) mod vocab, stride coprime with vocab (one
    cycle through the whole vocabulary, so a greedy continuation never repeats a token before `vocab` steps)."""
    import math
    stride = vocab // 3 + 1
    while math.gcd(stride, vocab) != 1:
        stride += 1
    return (torch.arange(vocab) + stride) % vocab


def make_state_dict(config: LlavaConfig, seed=0, device="cpu", dtype=torch.float32, include_lm=True,
                    include_vision=True, peaked_lm_head=0.0, chain=0.0, chain_embed_scale=16.0):
    """peaked_lm_head > 0 scales lm_head (kept for the older tests; it scales the rounding error with the margin).
    chain > 0 builds the exact-token weight set: `embed_tokens` is scaled by `chain_embed_scale` so that the current
    token's embedding dominates the residual stream, and row succ(t) of lm_head gets `chain` x unit(embed[t]) added
    (one one-hot-like component per context token). The greedy continuation is then t, succ(t), succ(succ(t)), ... --
    all distinct -- with a top-2 margin of ~chain * sqrt(hidden) logits, far above the bf16 error of a logit, while
    every decoder layer still contributes to the logits that the tests compare."""
    sd = {}
    for name, shape, kind in weight_specs(config, include_lm, include_vision):
        t = make_tensor(name, shape, kind, seed, device, dtype)
        if name == "lm_head.weight" and peaked_lm_head > 0:
            t = (t.float() * peaked_lm_head).to(dtype)
        sd[name] = t
    if chain > 0 and include_lm:
        e = sd["model.embed_tokens.weight"].float() * chain_embed_scale
        succ = chain_successor(config.vocab_size).to(e.device)
        w = sd["lm_head.weight"].float()
        w[succ] += chain * e / e.norm(dim=1, keepdim=True)
        sd["model.embed_tokens.weight"] = e.to(dtype)
        sd["lm_head.weight"] = w.to(dtype)
    return sd


def synth_batch(config: LlavaConfig, batch, views, text_len, seed=0, jitter=0, image_pos=40, audio=False,
                segmasks=False, dtype=torch.float32):
    """Synthetic inputs in the shape ModelWrapper.forward feeds generate() (scene_graph_prediction_model.py:117-231):
    images: list of (V, 3, S, S); input_ids (B, Ltext) with one IMAGE_TOKEN_INDEX, left-padded with pad id 0;
    ids avoid 0 (pad), 1/2 (bos/eos) and VIS_DESCRIPTOR_TOKEN_INDEX (SURVEY.md 7 'traps')."""
    from .constants import IMAGE_TOKEN_INDEX, VIS_DESCRIPTOR_TOKEN_INDEX
    g = torch.Generator().manual_seed(seed + 12345)
    vc = config.vision_config()
    S = vc["image_size"]
    images = [torch.randn(views, 3, S, S, generator=g).to(torch.bfloat16).to(dtype) for _ in range(batch)]
    lens = [text_len - (int(torch.randint(0, jitter + 1, (1,), generator=g)) if jitter else 0) for _ in range(batch)]
    L = max(lens)
    ids = torch.zeros(batch, L, dtype=torch.long)
    for b, n in enumerate(lens):
        row = torch.randint(3, config.vocab_size, (n,), generator=g)
        row[row == VIS_DESCRIPTOR_TOKEN_INDEX] = 3
        row[min(image_pos, n - 1)] = IMAGE_TOKEN_INDEX
        ids[b, L - n:] = row  # left padding (tokenizer_padding_side = 'left', scene_graph_prediction_model.py:54)
    out = {"images": images, "input_ids": ids, "attention_mask": ids.ne(0)}
    if audio:
        # bf16-representable: the reference rounds audio embeddings to bf16 on entry (builder.py:156)
        out["audio"] = [torch.randn(512, generator=g).to(torch.bfloat16).to(dtype) if b % 3 != 2 else None
                        for b in range(batch)]
    if segmasks:
        out["segmasks"] = [[torch.randint(0, 30, (32, 32), generator=g, dtype=torch.uint8) for _ in range(3 - b % 2)]
                           if b % 4 != 3 else None for b in range(batch)]
    return out
