"""Forward with saved activations + backward of the visual side of MM2SG for the fine-tune step: the trainable CLIP layers,
the BERT image pooler and the mlp2x_gelu projector (SURVEY.md 8a rows a1-a6 under training, a12 trainable set:
"mm_projector, image_pooler, last 12 CLIP layers", train.py:1257-1261).

Reference: the same modules as the inference path (clip_encoder.py:29-51 -> HF CLIPEncoderLayer,
multimodal_projector/builder.py:46-53,169-190 -> HF BertLayer) under torch autograd, dropout fixed to 0 (SURVEY.md 7,
"Dropout in training parity"). Every tensor operation here is a kernel of libb200mmor.so called through the C ABI;
torch owns the memory. The frozen CLIP layers run through the inference stage entry point (b200_vit_forward).

The extra-modality tokens of the pooler (builder.py:176-189) are covered by extras_forward / extras_backward: the audio
projection and the seg-mask CNN train (csrc/train_extras.cu); the point-cloud token is computed but PointTransformerV3
stays frozen (its training path -- BatchNorm batch statistics, drop-path -- is not built). The pooler's train-mode
hidden dropout is available (`pooler.dropout`, default 0); its attention-probability dropout is not.
"""
import ctypes

import numpy as np
import torch

from .. import _lib as L
from ..synth import POOLER_GEOMETRY

BF = torch.bfloat16
VIT = "model.vision_tower.vision_tower.vision_model."
POOL = "model.image_pooler.bert."


class _Grads:
    """fp32 gradient store keyed by the reference's parameter names; `accumulate` adds into existing entries."""

    def __init__(self, store, accumulate, device):
        self.g = store if store is not None else {}
        self.accumulate = accumulate
        self.device = device

    def slot(self, name, shape):
        if name not in self.g:
            self.g[name] = torch.zeros(shape, device=self.device, dtype=torch.float32)
            return self.g[name], False
        return self.g[name], self.accumulate


def _lin_bwd(G, x_saved, w, dy, wname, bname=None, need_dx=True, residual=None):
    """Backward of y = x w^T (+ b): dW, db into the store; returns dx (+ residual, fused in the GEMM epilogue)."""
    dw, acc = G.slot(wname, tuple(w.shape))
    L.gemm_ex(dy, x_saved, a_t=True, w_t=True, out=dw, accumulate=acc)
    if bname is not None:
        db, accb = G.slot(bname, (w.shape[0],))
        L.colsum(dy, out=db, accumulate=accb)
    return L.gemm_ex(dy, w, w_t=True, residual=residual) if need_dx else None


def _ln_bwd(G, x_saved, dy, gamma, eps, gname, bname, add=None):
    dg, acc = G.slot(gname, (gamma.numel(),))
    db, _ = G.slot(bname, (gamma.numel(),))
    dx, _, _ = L.norm_backward(x_saved, dy, gamma, eps, rms=False, dgamma=dg, dbeta=db, accumulate=acc, add=add)
    return dx


def _heads(t, n, tokens, heads, hd):
    """(n * tokens, 3 * heads * hd) fused qkv -> three (n, tokens, heads, hd) views."""
    v5 = t.view(n, tokens, 3, heads, hd)
    return v5[:, :, 0], v5[:, :, 1], v5[:, :, 2]


# =====================================================================================================================
# CLIP ViT: frozen prefix through the stage entry point, trainable layers op by op
# =====================================================================================================================
class VitCache:
    __slots__ = ("layers", "n_img", "first")


def vit_forward(tower, pixels, first_trainable):
    """pixels (N, 3, S, S) -> hidden (N, 1 + P, D) after layer n_run (CLS row kept) + cache for vit_backward.
    Layers [0, first_trainable) are frozen and run fused; layers [first_trainable, n_run) save their activations."""
    n_run = tower.n_layers_run()
    first = max(0, min(first_trainable, n_run))
    lib = L.lib()
    pixels = pixels.to(device=tower.device, dtype=BF).contiguous()
    N = pixels.shape[0]
    T, D = tower.num_patches + 1, tower.hidden_size
    H = tower.cfg["num_attention_heads"]
    hd = D // H
    eps = tower.cfg.get("layer_norm_eps", 1e-5)
    w = L.VitWeights.from_buffer_copy(tower._w)
    w.n_layers = first
    x = torch.empty((N, T, D), device=tower.device, dtype=BF)
    nb = lib.b200_vit_workspace_bytes(ctypes.byref(w), N)
    ws = tower._ws.get(nb, tower.device)
    L.check(lib.b200_vit_forward(ctypes.byref(w), L.ptr(pixels), L.ptr(x), N, L.ptr(ws), ws.numel(), L.stream_ptr()),
            "b200_vit_forward")
    x = x.view(N * T, D)
    cache = VitCache()
    cache.layers, cache.n_img, cache.first = [], N, first
    for l in range(first, n_run):
        lt = tower._keep[1][l]
        c = {"x": x}
        c["a"] = L.layernorm(x, lt["ln1_w"], lt["ln1_b"], eps)
        c["qkv"] = L.gemm(c["a"], lt["qkv_w"], bias=lt["qkv_b"])
        q, k, v = _heads(c["qkv"], N, T, H, hd)
        ctx, c["lse"] = L.flash_attention(q, k, v, scale=1.0, return_lse=True)   # q is pre-scaled in qkv_w
        c["ctx"] = ctx.view(N * T, D)
        c["x1"] = L.gemm(c["ctx"], lt["out_w"], bias=lt["out_b"], residual=x)
        c["m"] = L.layernorm(c["x1"], lt["ln2_w"], lt["ln2_b"], eps)
        c["z"] = L.gemm(c["m"], lt["fc1_w"], bias=lt["fc1_b"])
        c["h"] = L.act_forward(c["z"], L.ACT_QUICK_GELU)
        x = L.gemm(c["h"], lt["fc2_w"], bias=lt["fc2_b"], residual=c["x1"])
        cache.layers.append(c)
    return x.view(N, T, D), cache


def vit_backward(tower, cache, d_hidden, grads=None, accumulate=False):
    """d_hidden (N, 1 + P, D) bf16 -> gradients of the trainable CLIP layers under the reference's names."""
    G = _Grads(grads, accumulate, tower.device)
    N = cache.n_img
    T, D = tower.num_patches + 1, tower.hidden_size
    H = tower.cfg["num_attention_heads"]
    hd = D // H
    eps = tower.cfg.get("layer_norm_eps", 1e-5)
    scale = hd ** -0.5
    dx = d_hidden.to(BF).contiguous().view(N * T, D)
    for idx in range(len(cache.layers) - 1, -1, -1):
        l = cache.first + idx
        lt, c = tower._keep[1][l], cache.layers[idx]
        p = VIT + f"encoder.layers.{l}."
        f = "_fused." + p
        dh = _lin_bwd(G, c["h"], lt["fc2_w"], dx, p + "mlp.fc2.weight", p + "mlp.fc2.bias")
        dz = L.act_backward(c["z"], dh, L.ACT_QUICK_GELU)
        dm = _lin_bwd(G, c["m"], lt["fc1_w"], dz, p + "mlp.fc1.weight", p + "mlp.fc1.bias")
        dx1 = _ln_bwd(G, c["x1"], dm, lt["ln2_w"], eps, p + "layer_norm2.weight", p + "layer_norm2.bias", add=dx)
        dctx = _lin_bwd(G, c["ctx"], lt["out_w"], dx1, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias")
        dqkv = torch.empty_like(c["qkv"])
        q, k, v = _heads(c["qkv"], N, T, H, hd)
        L.flash_attention_bwd(q, k, v, c["ctx"].view(N, T, H, hd), dctx.view(N, T, H, hd), c["lse"], scale=1.0,
                              out=_heads(dqkv, N, T, H, hd))
        need_dx = idx > 0
        da = _lin_bwd(G, c["a"], lt["qkv_w"], dqkv, f + "qkv_w", f + "qkv_b", need_dx=True)
        dx = _ln_bwd(G, c["x"], da, lt["ln1_w"], eps, p + "layer_norm1.weight", p + "layer_norm1.bias", add=dx1)
        if not need_dx:
            dx = None
        # the fused qkv weight holds q_proj * head_dim^-0.5: dL/dW_q(reference) = scale * dL/dW_q(fused)
        qkv_w, qkv_b = G.g[f + "qkv_w"], G.g[f + "qkv_b"]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            s = scale if j == 0 else 1.0
            for src, suffix in ((qkv_w, ".weight"), (qkv_b, ".bias")):
                dst, acc = G.slot(p + f"self_attn.{n}" + suffix, tuple(src[j * D:(j + 1) * D].shape))
                part = src[j * D:(j + 1) * D]
                dst.copy_(part * s if not acc else dst + part * s)
        del G.g[f + "qkv_w"], G.g[f + "qkv_b"]
    return G.g


# =====================================================================================================================
# image pooler (BERT, post-LN) and projector
# =====================================================================================================================
def pooler_forward(pooler, hidden, split_sizes):
    """hidden (N_img, 1 + P, D) (CLS rows skipped through the gather map, llava_arch.py:143-170), split_sizes [V_b]
    -> pooled (B, keep, D) + cache."""
    g = pooler.geo
    D, H = g["hidden"], g["heads"]
    hd = D // H
    keep, eps = g["keep"], g["eps"]
    dev = pooler.device
    N, T, _ = hidden.shape
    P = T - 1
    B = len(split_sizes)
    vmax = max(split_sizes)
    S = vmax * P
    gmap = np.full((B, vmax, P), -1, dtype=np.int32)
    img0 = np.concatenate([[0], np.cumsum(split_sizes)[:-1]])
    base = np.arange(P, dtype=np.int32)[None, :] + 1
    for b, vb in enumerate(split_sizes):
        gmap[b, :vb] = (img0[b] + np.arange(vb, dtype=np.int32))[:, None] * T + base
    kv_len = torch.as_tensor(np.asarray(split_sizes, dtype=np.int32) * P).to(dev)
    t, layers_w, _ = pooler._keep
    cache = {"gmap": gmap, "B": B, "S": S, "N": N, "T": T, "kv_len": kv_len, "layers": []}
    gm = torch.as_tensor(gmap.reshape(-1)).to(dev)
    u = L.gather_add_rows(hidden.reshape(N * T, D), gm, t["pos_type"], B * S, period=S)
    cache["u"] = u
    # train-mode hidden dropout of HF BertModel (hidden_dropout_prob 0.1 by default, builder.py:68-79): after the
    # embedding LayerNorm and after the two output dense layers of every block, before the residual add. Off (p = 0)
    # unless `pooler.dropout` is set; the masks are functions of (seed, step, site) and re-created in the backward
    # (b200_dropout). NOT covered: attention_probs_dropout_prob (the probabilities never leave the flash kernel).
    p_drop = float(getattr(pooler, "dropout", 0.0))
    if p_drop > 0.0:
        pooler.rng_step = getattr(pooler, "rng_step", 0) + 1
    seed0 = (int(getattr(pooler, "seed", 0)) * 1000003 + getattr(pooler, "rng_step", 0)) * 64
    cache["p_drop"], cache["seed0"] = p_drop, seed0

    def dense_drop_add(x, w, b, residual, site):
        """dropout(x w^T + b) + residual (BertSelfOutput / BertOutput); one fused GEMM when p = 0."""
        if p_drop == 0.0:
            return L.gemm(x, w, bias=b, residual=residual)
        y = residual.clone()
        return L.dropout(L.gemm(x, w, bias=b), p_drop, seed0 + site, out=y, accumulate=True)

    e = L.layernorm(u, t["emb_ln_w"], t["emb_ln_b"], eps)
    if p_drop > 0.0:
        e = L.dropout(e, p_drop, seed0, out=e)
    for li, lt in enumerate(layers_w):
        c = {"e": e}
        c["qkv"] = L.gemm(e, lt["qkv_w"], bias=lt["qkv_b"])
        q, k, v = _heads(c["qkv"], B, S, H, hd)
        ctx, c["lse"] = L.flash_attention(q, k, v, kv_len=kv_len, return_lse=True)
        c["ctx"] = ctx.view(B * S, D)
        c["y1"] = dense_drop_add(c["ctx"], lt["ao_w"], lt["ao_b"], e, 1 + 2 * li)
        c["e1"] = L.layernorm(c["y1"], lt["ao_ln_w"], lt["ao_ln_b"], eps)
        c["z"] = L.gemm(c["e1"], lt["fc1_w"], bias=lt["fc1_b"])
        c["h"] = L.act_forward(c["z"], L.ACT_GELU)
        c["y2"] = dense_drop_add(c["h"], lt["fc2_w"], lt["fc2_b"], c["e1"], 2 + 2 * li)
        e = L.layernorm(c["y2"], lt["out_ln_w"], lt["out_ln_b"], eps)
        cache["layers"].append(c)
    pooled = e.view(B, S, D)[:, :keep]
    return pooled, cache


def pooler_backward(pooler, cache, d_pooled, grads=None, accumulate=False):
    """d_pooled (B, keep, D) -> (d_hidden (N_img, 1 + P, D) bf16, grads)."""
    G = _Grads(grads, accumulate, pooler.device)
    g = pooler.geo
    D, H = g["hidden"], g["heads"]
    hd = D // H
    keep, eps = g["keep"], g["eps"]
    B, S, N, T = cache["B"], cache["S"], cache["N"], cache["T"]
    t, layers_w, _ = pooler._keep
    de = torch.zeros((B, S, D), device=pooler.device, dtype=BF)
    de[:, :keep] = d_pooled.to(BF)                    # rows >= keep of the last layer are not consumed (builder.py:175)
    de = de.view(B * S, D)
    p_drop, seed0 = cache.get("p_drop", 0.0), cache.get("seed0", 0)
    masked = (lambda gy, site: L.dropout(gy, p_drop, seed0 + site)) if p_drop > 0.0 else (lambda gy, site: gy)
    for i in range(len(layers_w) - 1, -1, -1):
        lt, c = layers_w[i], cache["layers"][i]
        p = POOL + f"encoder.layer.{i}."
        f = "_fused." + p
        dy2 = _ln_bwd(G, c["y2"], de, lt["out_ln_w"], eps, p + "output.LayerNorm.weight", p + "output.LayerNorm.bias")
        dh = _lin_bwd(G, c["h"], lt["fc2_w"], masked(dy2, 2 + 2 * i), p + "output.dense.weight", p + "output.dense.bias")
        dz = L.act_backward(c["z"], dh, L.ACT_GELU)
        de1 = _lin_bwd(G, c["e1"], lt["fc1_w"], dz, p + "intermediate.dense.weight", p + "intermediate.dense.bias",
                       residual=dy2)                                       # + the residual branch y2 = ... + e1
        dy1 = _ln_bwd(G, c["y1"], de1, lt["ao_ln_w"], eps, p + "attention.output.LayerNorm.weight",
                      p + "attention.output.LayerNorm.bias")
        dctx = _lin_bwd(G, c["ctx"], lt["ao_w"], masked(dy1, 1 + 2 * i), p + "attention.output.dense.weight",
                        p + "attention.output.dense.bias")
        dqkv = torch.empty_like(c["qkv"])
        q, k, v = _heads(c["qkv"], B, S, H, hd)
        L.flash_attention_bwd(q, k, v, c["ctx"].view(B, S, H, hd), dctx.view(B, S, H, hd), c["lse"],
                              kv_len=cache["kv_len"], out=_heads(dqkv, B, S, H, hd))
        de = _lin_bwd(G, c["e"], lt["qkv_w"], dqkv, f + "qkv_w", f + "qkv_b", residual=dy1)   # y1 = ... + e
        qkv_w, qkv_b = G.g[f + "qkv_w"], G.g[f + "qkv_b"]
        for j, n in enumerate(("query", "key", "value")):
            for src, suffix in ((qkv_w, ".weight"), (qkv_b, ".bias")):
                part = src[j * D:(j + 1) * D]
                dst, acc = G.slot(p + f"attention.self.{n}" + suffix, tuple(part.shape))
                dst.copy_(part if not acc else dst + part)
        del G.g[f + "qkv_w"], G.g[f + "qkv_b"]
    du = _ln_bwd(G, cache["u"], masked(de, 0), t["emb_ln_w"], eps, POOL + "embeddings.LayerNorm.weight",
                 POOL + "embeddings.LayerNorm.bias")
    # u = gather(hidden) + position[s] + token_type[0]: table gradients are sums over the batch (and over s)
    dpos, acc = G.slot(POOL + "embeddings.position_embeddings.weight", (t["pos_type"].shape[0], D))
    if not acc:
        dpos.zero_()
    L.group_sum(du.view(B, S, D), out=dpos[:S], accumulate=acc)
    dtt, acct = G.slot(POOL + "embeddings.token_type_embeddings.weight", (2, D))
    if not acct:
        dtt.zero_()
    L.colsum(du, out=dtt[0], accumulate=acct)
    # scatter back to the tower's hidden rows (every source row is used at most once; CLS rows get zeros)
    inv = np.full(N * T, -1, dtype=np.int32)
    gm = cache["gmap"].reshape(-1)
    used = gm >= 0
    inv[gm[used]] = np.nonzero(used)[0].astype(np.int32)
    d_hidden = L.embed_rows(torch.as_tensor(inv).to(pooler.device), du, rows=N * T)
    return d_hidden.view(N, T, D), G.g


# =====================================================================================================================
# extra-modality tokens: pc (frozen), audio (Linear 512 -> 1024), seg-masks (5-layer CNN)   builder.py:93-167,176-189
# =====================================================================================================================
POOLX = "model.image_pooler."
_SEG_CH = (8, 64, 128, 256, 512, 1024)
_K_PAD = 64        # the weight-gradient GEMM contracts over the batch: padded with zero rows to one full K tile


def extras_forward(pooler, tokens, keep, pc=None, audio=None, segmasks=None):
    """tokens (B, T, D) bf16 whose [:, :keep] already holds the pooled image tokens: fills the extra tokens in the
    pooler's fixed order pc, audio, seg0..2 (a modality contributes its slot(s) to every sample as soon as its kwarg is
    not None) and returns the cache extras_backward needs."""
    B, T, D = tokens.shape
    dev = tokens.device
    t = keep
    cache = {}
    if pc is not None:
        if pooler.point_transformer is None:
            raise L.B200Error("point clouds were passed but the checkpoint has no point_transformer weights")
        pooler.point_transformer(pc, out=tokens[:, t])
        t += 1
    if audio is not None:                                                  # _encode_audio (builder.py:150-159)
        rows = max(_K_PAD, -(-B // _K_PAD) * _K_PAD)
        feats = torch.zeros((rows, 512), dtype=BF)
        for i, a in enumerate(audio):
            if a is not None:
                feats[i] = a.detach().to("cpu", BF)
        feats = feats.to(dev)
        L.gemm(feats[:B], pooler.audio_w, out=tokens.view(B * T, D)[t::T], bias=pooler.audio_b)
        cache["audio"] = (feats, t)
        t += 1
    if segmasks is not None:                                               # _encode_segmasks (builder.py:161-167)
        if pooler._seg is None:
            raise L.B200Error("seg-mask encoder weights not loaded")
        tokens[:, t:t + 3].zero_()
        maps, rows = [], []
        for i, sm in enumerate(segmasks):
            if sm is not None:
                for j, m in enumerate(sm):
                    if tuple(m.shape) != (32, 32):
                        raise AssertionError(f"Expected input size (batch_size, 32, 32), but got {tuple(m.shape)}")
                    maps.append(m.detach().to("cpu", torch.uint8))
                    rows.append(i * T + t + j)
        if maps:
            cls = torch.stack(maps).contiguous().to(dev)
            rm = torch.as_tensor(np.asarray(rows, dtype=np.int32)).to(dev)
            acts = L.segmask_forward_train(pooler._seg[0], cls, tokens, D, rm)
            cache["seg"] = (cls, rm, acts)
        t += 3
    if t != T:
        raise ValueError(f"token buffer holds {T} tokens but the modalities passed need {t}")
    return cache


def extras_backward(pooler, cache, d_tokens, grads=None, accumulate=False):
    """d_tokens (B, T, D) bf16, gradient w.r.t. the pooler's output tokens -> gradients of project_audio and of the
    seg-mask CNN under the reference's names. A modality that is absent from the batch leaves no entry (its parameters
    are skipped by the optimizer step, like parameters whose .grad is None in torch)."""
    G = _Grads(grads, accumulate, d_tokens.device)
    B, T, D = d_tokens.shape
    if "audio" in cache:
        feats, t = cache["audio"]
        dy = torch.zeros((feats.shape[0], D), device=d_tokens.device, dtype=BF)
        dy[:B] = d_tokens[:, t]
        _lin_bwd(G, feats, pooler.audio_w, dy, POOLX + "project_audio.weight", POOLX + "project_audio.bias",
                 need_dx=False)
    if "seg" in cache:
        cls, rm, acts = cache["seg"]
        p = POOLX + "segmasks_encoder."
        dW, dB = [], []
        acc = False
        for i in range(5):
            w, acc = G.slot(p + f"conv{i + 1}.weight", (_SEG_CH[i + 1], _SEG_CH[i], 3, 3))
            b, _ = G.slot(p + f"conv{i + 1}.bias", (_SEG_CH[i + 1],))
            dW.append(w)
            dB.append(b)
        dE, _ = G.slot(p + "embedding.weight", (30, 8))
        L.segmask_backward(pooler._seg[0], cls, acts, d_tokens.view(B * T, D), D, rm, dW, dB, dE, accumulate=acc)
    return G.g


def projector_forward(proj, tokens):
    """tokens (n, 1024) -> (n, hidden) + cache (mlp2x_gelu, multimodal_projector/builder.py:46-53)."""
    c = {"x": tokens}
    c["z"] = L.gemm(tokens, proj.t["0.weight"], bias=proj.t["0.bias"])
    c["h"] = L.act_forward(c["z"], L.ACT_GELU)
    out = L.gemm(c["h"], proj.t["2.weight"], bias=proj.t["2.bias"])
    return out, c


def projector_backward(proj, cache, d_out, grads=None, accumulate=False, need_dx=True):
    G = _Grads(grads, accumulate, d_out.device)
    dh = _lin_bwd(G, cache["h"], proj.t["2.weight"], d_out, "model.mm_projector.2.weight", "model.mm_projector.2.bias")
    dz = L.act_backward(cache["z"], dh, L.ACT_GELU)
    dx = _lin_bwd(G, cache["x"], proj.t["0.weight"], dz, "model.mm_projector.0.weight", "model.mm_projector.0.bias",
                  need_dx=need_dx)
    return dx, G.g
