"""CPU oracle for the MM2SG hot path -- TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32 by default) restatement of the arithmetic the reference executes on the path named by
BASELINE.json: CLIP ViT tower -> BERT image pooler (+ audio / seg-mask tokens) -> mlp2x_gelu projector ->
multimodal token pack -> Llama decoder prefill / greedy decode / weighted CE. Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product (mm_or_b200/) never does.

Where the arithmetic lives in un-vendored third-party code (transformers==4.31.0, SGG/requirements.txt:4) the
published algorithm is restated and the reference call site is cited. Paths are relative to
/root/reference/scene_graph_generation/LLaVA/llava/.

Pinning: tests/golden/make_golden.py imports the reference's own LlavaLlamaForCausalLM through a runtime shim
(oracle/ref_shim.py; works only where /root/reference exists), loads the same synthetic weights, and records the
reference outputs in tests/golden/*.pt; tests/test_oracle_golden.py checks this restatement against those files.
Parity is therefore pinned to the reference's Python code executed on transformers 5.5.0 building blocks (the
reference ships no tests or golden vectors of its own).

All functions take `sd`, a dict of tensors keyed by the reference's state_dict names.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100          # constants.py:7
IMAGE_TOKEN_INDEX = -200     # constants.py:8
VIS_DESCRIPTOR_TOKEN_INDEX = 18610  # constants.py:10

VIT = "model.vision_tower.vision_tower.vision_model."
POOL = "model.image_pooler."


@dataclass
class VitCfg:
    hidden: int = 1024
    heads: int = 16
    layers: int = 24
    ffn: int = 4096
    image: int = 336
    patch: int = 14
    select_layer: int = -2   # mm_vision_select_layer
    eps: float = 1e-5


@dataclass
class PoolerCfg:             # multimodal_projector/builder.py:68-80
    hidden: int = 1024
    heads: int = 8
    layers: int = 2
    ffn: int = 4096
    max_pos: int = 576 * 7
    keep: int = 576          # builder.py:175
    eps: float = 1e-12


@dataclass
class LlmCfg:
    hidden: int = 4096
    heads: int = 32
    layers: int = 32
    ffn: int = 11008
    vocab: int = 32000
    eps: float = 1e-5
    rope_theta: float = 10000.0
    max_pos: int = 4096


@dataclass
class Mm2sgCfg:
    vit: VitCfg = field(default_factory=VitCfg)
    pooler: PoolerCfg = field(default_factory=PoolerCfg)
    llm: LlmCfg = field(default_factory=LlmCfg)


def cfg_from_llava(cfg) -> Mm2sgCfg:
    """Oracle geometry from a LlavaConfig-like attribute bag (mm_or_b200.config.LlavaConfig)."""
    vc = cfg.vision_config()
    return Mm2sgCfg(
        vit=VitCfg(hidden=vc["hidden_size"], heads=vc["num_attention_heads"], layers=vc["num_hidden_layers"],
                   ffn=vc["intermediate_size"], image=vc["image_size"], patch=vc["patch_size"],
                   select_layer=cfg.mm_vision_select_layer),
        pooler=PoolerCfg(),
        llm=LlmCfg(hidden=cfg.hidden_size, heads=cfg.num_attention_heads, layers=cfg.num_hidden_layers,
                   ffn=cfg.intermediate_size, vocab=cfg.vocab_size, eps=cfg.rms_norm_eps,
                   rope_theta=cfg.rope_theta, max_pos=cfg.max_position_embeddings))


def _lin(x, sd, name, bias=True):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"] if bias else None)


# ----------------------------------------------------------------------------------------------------------------
# (1) CLIP ViT tower -- CLIPVisionTower.forward + feature_select (multimodal_encoder/clip_encoder.py:29-51) calling
#     HF CLIPVisionModel(output_hidden_states=True); algorithm of HF 4.31 CLIPVisionTransformer restated.
# ----------------------------------------------------------------------------------------------------------------
def clip_vit_hidden_states(sd, pixels: torch.Tensor, cfg: VitCfg, n_layers: Optional[int] = None) -> List[torch.Tensor]:
    p = VIT
    n_layers = cfg.layers if n_layers is None else n_layers
    x = F.conv2d(pixels, sd[p + "embeddings.patch_embedding.weight"], None, stride=cfg.patch)  # no bias
    x = x.flatten(2).transpose(1, 2)                                                           # (N, P, D)
    cls = sd[p + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][None]
    x = F.layer_norm(x, (cfg.hidden,), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], cfg.eps)
    hs = [x]
    hd = cfg.hidden // cfg.heads
    for l in range(n_layers):
        q = p + f"encoder.layers.{l}."
        a = F.layer_norm(x, (cfg.hidden,), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], cfg.eps)
        N, T, _ = a.shape
        qh = (_lin(a, sd, q + "self_attn.q_proj") * hd ** -0.5).view(N, T, cfg.heads, hd).transpose(1, 2)
        kh = _lin(a, sd, q + "self_attn.k_proj").view(N, T, cfg.heads, hd).transpose(1, 2)
        vh = _lin(a, sd, q + "self_attn.v_proj").view(N, T, cfg.heads, hd).transpose(1, 2)
        pr = torch.softmax(qh @ kh.transpose(-1, -2), dim=-1)            # no mask, dropout 0
        ctx = (pr @ vh).transpose(1, 2).reshape(N, T, cfg.hidden)
        x = x + _lin(ctx, sd, q + "self_attn.out_proj")
        m = F.layer_norm(x, (cfg.hidden,), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], cfg.eps)
        h = _lin(m, sd, q + "mlp.fc1")
        h = h * torch.sigmoid(1.702 * h)                                  # quick_gelu
        x = x + _lin(h, sd, q + "mlp.fc2")
        hs.append(x)
    return hs


def clip_tower_forward(sd, pixels: torch.Tensor, cfg: VitCfg) -> torch.Tensor:
    """hidden_states[select_layer][:, 1:] (clip_encoder.py:29-33); layers past the selected one are dead work."""
    need = cfg.layers + 1 + cfg.select_layer if cfg.select_layer < 0 else cfg.select_layer
    hs = clip_vit_hidden_states(sd, pixels, cfg, n_layers=need)
    return hs[need][:, 1:]


# ----------------------------------------------------------------------------------------------------------------
# (2) pad_embeddings (llava_arch.py:143-170) + ImageEmbeddingPooler (multimodal_projector/builder.py:169-190)
# ----------------------------------------------------------------------------------------------------------------
def pad_embeddings(embs: List[torch.Tensor]):
    B = len(embs)
    img_len, D = embs[0].shape[1], embs[0].shape[2]
    vmax = max(e.shape[0] for e in embs)
    out = torch.zeros(B, vmax, img_len, D, dtype=embs[0].dtype)
    mask = torch.zeros(B, vmax * img_len, dtype=torch.bool)
    for i, e in enumerate(embs):
        out[i, :e.shape[0]] = e
        mask[i, :e.shape[0] * img_len] = True
    return out.flatten(1, 2), mask


def bert_pooler_forward(sd, emb: torch.Tensor, mask: torch.Tensor, cfg: PoolerCfg) -> torch.Tensor:
    """HF BertModel(inputs_embeds=emb, attention_mask=mask)['last_hidden_state'][:, :keep], eval mode (dropout 0)."""
    p = POOL + "bert."
    B, S, D = emb.shape
    e = emb + sd[p + "embeddings.token_type_embeddings.weight"][0] + sd[p + "embeddings.position_embeddings.weight"][:S]
    e = F.layer_norm(e, (D,), sd[p + "embeddings.LayerNorm.weight"], sd[p + "embeddings.LayerNorm.bias"], cfg.eps)
    add = torch.zeros(B, 1, 1, S, dtype=e.dtype)
    add.masked_fill_(~mask[:, None, None, :], torch.finfo(e.dtype).min)
    hd = D // cfg.heads
    for l in range(cfg.layers):
        q = p + f"encoder.layer.{l}."
        qh = _lin(e, sd, q + "attention.self.query").view(B, S, cfg.heads, hd).transpose(1, 2)
        kh = _lin(e, sd, q + "attention.self.key").view(B, S, cfg.heads, hd).transpose(1, 2)
        vh = _lin(e, sd, q + "attention.self.value").view(B, S, cfg.heads, hd).transpose(1, 2)
        pr = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(hd) + add, dim=-1)
        ctx = (pr @ vh).transpose(1, 2).reshape(B, S, D)
        e = F.layer_norm(e + _lin(ctx, sd, q + "attention.output.dense"), (D,),
                         sd[q + "attention.output.LayerNorm.weight"], sd[q + "attention.output.LayerNorm.bias"], cfg.eps)
        h = F.gelu(_lin(e, sd, q + "intermediate.dense"))                 # exact (erf) GELU
        e = F.layer_norm(e + _lin(h, sd, q + "output.dense"), (D,),
                         sd[q + "output.LayerNorm.weight"], sd[q + "output.LayerNorm.bias"], cfg.eps)
    return e[:, :cfg.keep]


def encode_audio(sd, audios, length: int, dtype) -> torch.Tensor:
    """_encode_audio (builder.py:150-159): zeros for missing rows, then Linear(512, 1024) (bias applies to zeros)."""
    feats = torch.zeros(length, 512, dtype=dtype)
    if audios is not None:
        for i, a in enumerate(audios):
            if a is not None:
                feats[i] = a.to(dtype)
    return _lin(feats, sd, POOL + "project_audio")


def segmask_features(sd, maps: torch.Tensor) -> torch.Tensor:
    """SegmentationMapFeatureExtractor.forward (segmentation_map_feature_extractor.py:53-75). maps (n, 32, 32)."""
    p = POOL + "segmasks_encoder."
    x = F.embedding(maps.long(), sd[p + "embedding.weight"]).permute(0, 3, 1, 2)
    for i in range(1, 6):
        x = F.relu(F.conv2d(x, sd[p + f"conv{i}.weight"], sd[p + f"conv{i}.bias"], stride=2, padding=1))
    return x.squeeze(-1).squeeze(-1)


def encode_segmasks(sd, segmasks, dtype) -> torch.Tensor:
    """_encode_segmasks (builder.py:161-167): (B, 3, 1024), zero rows for missing maps."""
    out = torch.zeros(len(segmasks), 3, 1024, dtype=dtype)
    for i, sm in enumerate(segmasks):
        if sm is not None:
            out[i, :len(sm)] = segmask_features(sd, torch.stack(list(sm))).to(dtype)
    return out


def pooler_with_extras(sd, emb, mask, cfg: PoolerCfg, audio=None, segmasks=None, pc=None) -> torch.Tensor:
    """ImageEmbeddingPooler.forward (builder.py:169-190). Token order: pooled[0:keep], pc, audio, seg0, seg1, seg2.
    The point-cloud token is computed in fp32 with autocast off and cast to the pooler dtype (builder.py:176-177)."""
    out = bert_pooler_forward(sd, emb, mask, cfg)
    extra = []
    if pc is not None:
        from . import ptv3_oracle
        extra.append(ptv3_oracle.encode_pc({k: v.float() for k, v in sd.items() if k.startswith(ptv3_oracle.PT)},
                                           pc).to(out.dtype).unsqueeze(1))
    if audio is not None:
        extra.append(encode_audio(sd, audio, len(out), out.dtype).unsqueeze(1))
    if segmasks is not None:
        extra.extend(encode_segmasks(sd, segmasks, out.dtype).transpose(0, 1).unsqueeze(2))
    return torch.cat([out] + extra, dim=1) if extra else out


def mm_projector(sd, x: torch.Tensor) -> torch.Tensor:
    """mlp2x_gelu (builder.py:46-53) applied at llava_arch.py:182."""
    return _lin(F.gelu(_lin(x, sd, "model.mm_projector.0")), sd, "model.mm_projector.2")


def encode_images_pooled(sd, images: List[torch.Tensor], cfg: Mm2sgCfg, audio=None, segmasks=None,
                         pc=None) -> torch.Tensor:
    """llava_arch.py:172-183 for the `type(images) is list or ndim == 5` branch (:203-207)."""
    concat = torch.cat([im for im in images], dim=0)
    feats = clip_tower_forward(sd, concat, cfg.vit)
    split = torch.split(feats, [im.shape[0] for im in images], dim=0)
    emb, mask = pad_embeddings(list(split))
    pooled = pooler_with_extras(sd, emb, mask, cfg.pooler, audio, segmasks, pc)
    return mm_projector(sd, pooled)


# ----------------------------------------------------------------------------------------------------------------
# (3) multimodal pack -- prepare_inputs_labels_for_multimodal (llava_arch.py:188-353), with / without vis_descriptor_embs
# ----------------------------------------------------------------------------------------------------------------
DESC_BASE = -(1 << 20)


def pack_plan(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], labels: Optional[torch.Tensor],
              t_vis: int, padding_side: str = "right", max_len: Optional[int] = None, desc_rows=None,
              n_blocks: Optional[int] = None, return_blocks: bool = False):
    """Pure index arithmetic of the pack: returns (src, labels, mask, position_ids) with src (B, L) int64 where
    src >= 0 is a token id to embed, -1 a zero pad row, -2 - j the j-th visual token, DESC_BASE - r the r-th row of
    the sample's concatenated vis_descriptor_embs. Mirrors :235-338 including the quirk that text following a
    VIS_DESCRIPTOR token is dropped when no descriptor embeddings are given (:253-294) and truncation to
    tokenizer_model_max_length (:302-306). desc_rows: per sample the row counts of its descriptor tensors (:278-294:
    descriptor j replaces the j-th VIS_DESCRIPTOR placeholder, one zero row when the sample has too few).
    n_blocks / return_blocks: the reference indexes image_features with ONE running counter over the batch
    (cur_image_idx, :239,245,264-265): every <image> placeholder takes the next block, a text-only row skips one, an
    index past the last block raises IndexError. return_blocks adds a fifth result, blk (B, L): the block a visual
    position came from (-1 elsewhere)."""
    B = input_ids.shape[0]
    n_blocks = B if n_blocks is None else n_blocks
    cur_image_idx = 0
    rblocks = []
    am = torch.ones_like(input_ids, dtype=torch.bool) if attention_mask is None else attention_mask.bool()
    lab = torch.full_like(input_ids, IGNORE_INDEX) if labels is None else labels
    rows, rlabels = [], []
    for b in range(B):
        ids = input_ids[b][am[b]]
        lb = lab[b][am[b]]
        n_img = int((ids == IMAGE_TOKEN_INDEX).sum())
        if n_img == 0:
            if cur_image_idx >= n_blocks:
                raise IndexError(f"index {cur_image_idx} is out of bounds for dimension 0 with size {n_blocks}")
            cur_image_idx += 1                                                   # :250
            rows.append(ids.clone())
            rlabels.append(lb.clone())
            rblocks.append(torch.full((len(ids),), -1, dtype=torch.long))
            continue
        cut = [-1] + torch.where((ids == IMAGE_TOKEN_INDEX) | (ids == VIS_DESCRIPTOR_TOKEN_INDEX))[0].tolist() + [len(ids)]
        chunks = [ids[cut[i] + 1:cut[i + 1]] for i in range(len(cut) - 1)]
        lchunks = [lb[cut[i] + 1:cut[i + 1]] for i in range(len(cut) - 1)]
        r, rl, rb = [], [], []
        for i in range(n_img + 1):
            r.append(chunks[i])
            rl.append(lchunks[i])
            rb.append(torch.full((len(chunks[i]),), -1, dtype=torch.long))
            if i < n_img:
                if cur_image_idx >= n_blocks:
                    raise IndexError(f"index {cur_image_idx} is out of bounds for dimension 0 with size {n_blocks}")
                r.append(-2 - torch.arange(t_vis))
                rl.append(torch.full((t_vis,), IGNORE_INDEX, dtype=lb.dtype))
                rb.append(torch.full((t_vis,), cur_image_idx, dtype=torch.long))
                cur_image_idx += 1                                               # :265
        if desc_rows is not None:
            n_desc = int((ids == VIS_DESCRIPTOR_TOKEN_INDEX).sum())
            done = 0
            for j in range(n_desc):
                if j < len(desc_rows[b]):
                    k = int(desc_rows[b][j])
                    d = DESC_BASE - torch.arange(done, done + k)
                    done += k
                else:
                    d = torch.full((1,), -1, dtype=torch.long)          # dummy zeros(4096) row (:284-286)
                r.append(d)
                rl.append(torch.full((len(d),), IGNORE_INDEX, dtype=lb.dtype))
                rb.append(torch.full((len(d),), -1, dtype=torch.long))
                r.append(chunks[n_img + j + 1])
                rl.append(lchunks[n_img + j + 1])
                rb.append(torch.full((len(chunks[n_img + j + 1]),), -1, dtype=torch.long))
        rows.append(torch.cat(r))
        rlabels.append(torch.cat(rl))
        rblocks.append(torch.cat(rb))
    if max_len is not None:
        rows = [r[:max_len] for r in rows]
        rlabels = [r[:max_len] for r in rlabels]
        rblocks = [r[:max_len] for r in rblocks]
    L = max(len(r) for r in rows)
    src = torch.full((B, L), -1, dtype=torch.long)
    out_l = torch.full((B, L), IGNORE_INDEX, dtype=lab.dtype)
    mask = torch.zeros(B, L, dtype=torch.bool)
    pos = torch.zeros(B, L, dtype=torch.long)
    blk = torch.full((B, L), -1, dtype=torch.long)
    for b, (r, rl, rb) in enumerate(zip(rows, rlabels, rblocks)):
        n = len(r)
        if n == 0:
            continue
        sl = slice(L - n, L) if padding_side == "left" else slice(0, n)
        src[b, sl] = r
        out_l[b, sl] = rl
        mask[b, sl] = True
        pos[b, sl] = torch.arange(n)
        blk[b, sl] = rb
    if return_blocks:
        return src, out_l, mask, pos, blk
    return src, out_l, mask, pos


def pack_embeds(sd, src: torch.Tensor, visual: torch.Tensor, desc=None, blk=None) -> torch.Tensor:
    """Materialise inputs_embeds (B, L, D) from a pack plan and the projected visual tokens (blocks, T_vis, D); desc:
    per sample the (R_b, D) concatenation of its descriptor tensors (rows addressed by DESC_BASE - r); blk: the block
    of every visual position (pack_plan(return_blocks=True)), default block b for row b."""
    table = sd["model.embed_tokens.weight"]
    B, L = src.shape
    out = torch.zeros(B, L, table.shape[1], dtype=visual.dtype)
    for b in range(B):
        txt = src[b] >= 0
        out[b, txt] = table[src[b, txt]].to(visual.dtype)
        vis = (src[b] <= -2) & (src[b] > DESC_BASE)
        out[b, vis] = visual[(blk[b, vis] if blk is not None else b), (-2 - src[b, vis])]
        dsc = src[b] <= DESC_BASE
        if bool(dsc.any()):
            out[b, dsc] = desc[b][DESC_BASE - src[b, dsc]].to(visual.dtype)
    return out


# ----------------------------------------------------------------------------------------------------------------
# (4) Llama decoder -- HF LlamaForCausalLM.forward reached from language_model/llava_llama.py:93 (4.31 eager attention)
# ----------------------------------------------------------------------------------------------------------------
def rope_tables(cfg: LlmCfg, dtype=torch.float32):
    hd = cfg.hidden // cfg.heads
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.outer(torch.arange(cfg.max_pos, dtype=torch.float32), inv)      # (max_pos, hd/2)
    return fr.cos().to(dtype), fr.sin().to(dtype)


def _rmsnorm(x, w, eps):
    xf = x.float()
    return w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)


def _rope(x, cos, sin):  # x (B, H, T, hd); cos/sin (B, T, hd/2)
    hd = x.shape[-1]
    c = torch.cat([cos, cos], -1)[:, None]
    s = torch.cat([sin, sin], -1)[:, None]
    rot = torch.cat([-x[..., hd // 2:], x[..., :hd // 2]], -1)
    return x * c + rot * s


def llama_forward(sd, x: torch.Tensor, mask: torch.Tensor, pos: torch.Tensor, cfg: LlmCfg, past=None,
                  last_only: bool = False, return_hidden: bool = False):
    """x (B, T, D) embeddings for the new positions; mask (B, past + T) bool key mask; pos (B, T) position ids.
    Returns (logits, kv) with kv a list of (k, v) per layer shaped (B, H, past + T, hd)."""
    B, T, D = x.shape
    hd = D // cfg.heads
    cos_t, sin_t = rope_tables(cfg)
    cos, sin = cos_t[pos].to(x.dtype), sin_t[pos].to(x.dtype)
    P = 0 if past is None else past[0][0].shape[2]
    S = P + T
    qi = torch.arange(P, S)[:, None]
    kj = torch.arange(S)[None, :]
    vis = (kj <= qi)[None, None] & mask[:, None, None, :S]                     # causal & key padding
    add = torch.zeros(B, 1, T, S, dtype=torch.float32).masked_fill_(~vis, torch.finfo(torch.float32).min)
    kv = []
    for l in range(cfg.layers):
        p = f"model.layers.{l}."
        a = _rmsnorm(x, sd[p + "input_layernorm.weight"], cfg.eps)
        q = _lin(a, sd, p + "self_attn.q_proj", False).view(B, T, cfg.heads, hd).transpose(1, 2)
        k = _lin(a, sd, p + "self_attn.k_proj", False).view(B, T, cfg.heads, hd).transpose(1, 2)
        v = _lin(a, sd, p + "self_attn.v_proj", False).view(B, T, cfg.heads, hd).transpose(1, 2)
        q, k = _rope(q, cos, sin), _rope(k, cos, sin)
        if past is not None:
            k = torch.cat([past[l][0], k], dim=2)
            v = torch.cat([past[l][1], v], dim=2)
        kv.append((k, v))
        sc = (q @ k.transpose(-1, -2)).float() / math.sqrt(hd) + add
        pr = torch.softmax(sc, dim=-1, dtype=torch.float32).to(x.dtype)
        ctx = (pr @ v).transpose(1, 2).reshape(B, T, D)
        x = x + _lin(ctx, sd, p + "self_attn.o_proj", False)
        b = _rmsnorm(x, sd[p + "post_attention_layernorm.weight"], cfg.eps)
        x = x + _lin(F.silu(_lin(b, sd, p + "mlp.gate_proj", False)) * _lin(b, sd, p + "mlp.up_proj", False),
                     sd, p + "mlp.down_proj", False)
    if return_hidden:
        return x, kv
    h = _rmsnorm(x[:, -1:] if last_only else x, sd["model.norm.weight"], cfg.eps)
    return F.linear(h, sd["lm_head.weight"]).float(), kv


# ----------------------------------------------------------------------------------------------------------------
# (5) end-to-end: LlavaLlamaForCausalLM.forward (llava_llama.py:54-106) and the greedy loop the reference drives via
#     model.generate(do_sample=False) (scene_graph_prediction_model.py:221-231; decode-step inputs llava_arch.py:192-201)
# ----------------------------------------------------------------------------------------------------------------
def multimodal_prefill(sd, cfg: Mm2sgCfg, input_ids, attention_mask, images, labels=None, audio=None, segmasks=None,
                       padding_side="right", max_len=None, last_only=False, pc=None, vis_descriptor_embs=None):
    """vis_descriptor_embs: as the reference takes it (llava_arch.py:278-290) -- one list of (D,) / (k, D) tensors per
    sample, or the bare list for a batch of one."""
    visual = encode_images_pooled(sd, images, cfg, audio, segmasks, pc)
    desc_rows = desc = None
    if vis_descriptor_embs is not None:
        embs = vis_descriptor_embs if type(vis_descriptor_embs[0]) is list else [vis_descriptor_embs]
        desc_rows = [[1 if e.ndim == 1 else e.shape[0] for e in per] for per in embs]
        D = visual.shape[-1]
        desc = [torch.cat([e.reshape(-1, D) for e in per]) if per else torch.zeros(0, D) for per in embs]
    src, mlabels, mask, pos, blk = pack_plan(input_ids, attention_mask, labels, visual.shape[1], padding_side, max_len,
                                             desc_rows=desc_rows, n_blocks=visual.shape[0], return_blocks=True)
    emb = pack_embeds(sd, src, visual, desc, blk)
    logits, kv = llama_forward(sd, emb, mask, pos, cfg.llm, last_only=last_only)
    return {"logits": logits, "kv": kv, "mask": mask, "pos": pos, "modified_labels": mlabels, "inputs_embeds": emb,
            "visual": visual}


def greedy_decode(sd, cfg: Mm2sgCfg, logits_last, kv, mask, max_new_tokens, eos_id=2, pad_id=0, stop_on_eos=True,
                  logits_dtype=None, forced_tokens=None):
    """HF greedy_search semantics: argmax of the last-position logits (optionally rounded to `logits_dtype` first, as
    the bf16 reference does), finished rows emit pad, stop when every row has emitted EOS or at max_new_tokens.
    forced_tokens (B, n): teacher forcing for the tests -- the token FED to step s + 1 is forced_tokens[:, s] instead
    of the argmax (the argmax is still what is returned), so two implementations can be compared step by step on
    identical inputs even where their argmax differs by a rounding-level tie."""
    B = logits_last.shape[0]
    unfinished = torch.ones(B, dtype=torch.long)
    out, all_logits = [], []
    cur = logits_last
    for step in range(max_new_tokens):
        lg = cur if logits_dtype is None else cur.to(logits_dtype).float()
        all_logits.append(cur)
        nxt = lg.argmax(-1)
        if stop_on_eos:
            nxt = nxt * unfinished + pad_id * (1 - unfinished)
            unfinished = unfinished * (nxt != eos_id).long()
        out.append(nxt)
        if stop_on_eos and unfinished.max() == 0:
            break
        if step == max_new_tokens - 1:
            break
        mask = torch.cat([mask, torch.ones(B, 1, dtype=torch.bool)], dim=1)          # llava_arch.py:195-199
        pos = mask.sum(1, keepdim=True) - 1                                             # llava_arch.py:200
        fed = nxt if forced_tokens is None else forced_tokens[:, step].to(nxt.device)
        emb = sd["model.embed_tokens.weight"][fed][:, None].to(kv[0][0].dtype)
        cur, kv = llama_forward(sd, emb, mask, pos, cfg.llm, past=kv, last_only=True)
        cur = cur[:, -1]
    return torch.stack(out, dim=1), torch.stack(all_logits, dim=1)


def weighted_ce(logits, modified_labels, vocab_weight):
    """LLaVATrainer.compute_loss (train/llava_trainer.py:143-167): class-weighted CE on shifted modified_labels."""
    sl = logits[..., :-1, :].contiguous().view(-1, logits.shape[-1])
    tl = modified_labels[..., 1:].contiguous().view(-1)
    return F.cross_entropy(sl.float(), tl, weight=vocab_weight.float(), ignore_index=IGNORE_INDEX)


def token_weights(freqs: dict, vocab: int):
    """train/train.py:1310-1327: w = 1 / (ln f + 1) for seen tokens, min(w) / 100 for the rest."""
    w = {int(k): 1.0 / (math.log(v) + 1.0) for k, v in freqs.items()}
    lo = min(w.values())
    out = torch.full((vocab,), lo / 100.0)
    for k, v in w.items():
        if 0 <= k < vocab:
            out[k] = v
    return out
