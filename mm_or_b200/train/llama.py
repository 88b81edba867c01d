"""Teacher-forced forward + backward of the Llama decoder for the MM2SG fine-tune step.

Reference: LlavaLlamaForCausalLM.forward (LLaVA/llava/model/language_model/llava_llama.py:54-106 -> HF LlamaForCausalLM
with the FlashAttention-2 patch train/llama_flash_attn_monkey_patch.py:15-92) followed by LLaVATrainer.compute_loss
(train/llava_trainer.py:136-174) and torch autograd. Here every tensor operation is a kernel of libb200mmor.so called
through the C ABI (forward kernels of the inference path + the backward operators of train.cu / gemm_sm100.cu /
attention_bwd_sm100.cu); torch only owns the memory. Activations are saved per layer -- x_in, norm(x_in), qkv (q
rotated), K / V in the head-major cache layout, attention context + log-sum-exp, x_mid, norm(x_mid), the SwiGLU
pre-activation and its output -- or, with `recompute=True` (the reference's gradient checkpointing, train.py:1148), only
x_in, the rest being recomputed layer by layer in the backward.

Right-padded batches (train.py:1185-1191): `lengths[b]` real rows, keys beyond them are masked, labels are -100 there.
Gradients are returned in fp32 under the reference's parameter names (q/k/v and gate/up are split back out of the fused
layouts), `accumulate=True` adds into an existing gradient dict (gradient accumulation over micro-batches).
"""
import ctypes

import torch

from .. import _lib as L

BF = torch.bfloat16


class LayerCache:
    __slots__ = ("x_in", "a", "qkv", "kc", "vc", "ctx", "lse", "x_mid", "b", "z", "h", "t")


def _views(qkv, kc, vc, B, Lq, H):
    """(B, L, H, 128) views of q (inside the token-major qkv) and of K / V (head-major caches)."""
    q = qkv.view(B, Lq, 3, H, 128)[:, :, 0]
    k = kc.transpose(1, 2)[:, :Lq]
    v = vc.transpose(1, 2)[:, :Lq]
    return q, k, v


def forward_backward(model, embeds, labels, lengths, vocab_weight=None, grad_scale=1.0, grads=None, accumulate=False,
                     need_input_grad=True, lora=None, train_base=True, recompute=False):
    """embeds (B, L, D) bf16 packed inputs_embeds, labels (B, L) int64 UNshifted modified_labels, lengths (B,) int32.
    lora: a train.lora.LoraState -- adapters on the seven linears of every layer (gradients under `_lora.*` keys);
    train_base = False freezes the base Linear weights (the QLoRA / LoRA recipe), norms and lm_head follow it too.
    recompute: activation recomputation per decoder layer -- the reference trains with gradient checkpointing
    (train.py:1148 `gradient_checkpointing_enable`, recipe README.md:119-165): the forward keeps only every layer's
    input; the backward re-runs that layer's forward (same kernels, same dropout seeds => bit-identical activations and
    gradients) right before differentiating it. What stays alive per layer drops from (a, qkv, K, V, ctx, lse, x_mid,
    b, z, h) -- ~150 KB per token -- to x_in alone (8 KB per token), for one extra forward (+ 1/3 of the step's FLOPs).
    Returns (loss fp32 0-dim, weight sum, grads dict, d_embeds (B, L, D) bf16 or None)."""
    cfg = model.config
    lib = L.lib()
    dev = embeds.device
    B, Lq, D = embeds.shape
    H, F, V = cfg.num_attention_heads, cfg.intermediate_size, cfg.vocab_size
    T = B * Lq
    layers, _, final_norm, rope_cos, rope_sin = model._keep[:5]
    eps = cfg.rms_norm_eps
    lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    labels = labels.to(dev).contiguous()
    x = embeds.to(dev, BF).contiguous().view(T, D)

    def rope_write(qkv, kc, vc):
        L.check(lib.b200_rope_kv_write(L.ptr(qkv), None, L.ptr(rope_cos), L.ptr(rope_sin), cfg.max_position_embeddings,
                                       L.ptr(kc), L.ptr(vc), B, H, Lq, 0, Lq, L.stream_ptr()), "b200_rope_kv_write")

    # LoRA dropout (peft: the adapter branch sees dropout(x), the base branch x; lora_dropout 0.05 in the reference's
    # recipe, train.py:111,1165). The mask is a function of (seed, element): the backward re-creates it. One mask per
    # FUSED projection (q, k, v share it; peft draws one per adapted Linear -- same distribution per projection, but
    # correlated across q / k / v; stated in DESIGN.md).
    p_drop = float(getattr(lora, "dropout", 0.0)) if lora is not None else 0.0
    if p_drop > 0.0:
        lora.rng_step = getattr(lora, "rng_step", 0) + 1
    _fidx = {"qkv_w": 0, "o_w": 1, "gate_up_w": 2, "down_w": 3}

    def drop_seed(i, fname):
        return (int(getattr(lora, "seed", 0)) * 1000003 + lora.rng_step) * 4096 + i * 4 + _fidx[fname]

    def lin_fwd(i, fname, x_in, w, residual=None, keep=None):
        """y = x w^T (+ residual) (+ LoRA: (alpha / r) * (dropout(x) A_cat^T) B_blk^T added in place)."""
        y = L.gemm(x_in, w, residual=residual)
        if lora is not None:
            a_cat, b_blk = lora.fused[i][fname]
            t = L.gemm(L.dropout(x_in, p_drop, drop_seed(i, fname)) if p_drop > 0.0 else x_in, a_cat)
            L.gemm_ex(t, b_blk, out=y, residual=y, scale=lora.scale)
            keep[fname] = t
        return y

    # ---------------------------------------------------------------- forward
    def layer_forward(i, x, want_output=True):
        """One decoder layer; returns (saved activations, output). want_output=False (recomputation in the backward)
        stops after the SwiGLU: the down projection's result is not needed to differentiate the layer."""
        lt = layers[i]
        c = LayerCache()
        c.t = {}
        c.x_in = x
        c.a = L.rmsnorm(x, lt["attn_norm"], eps)
        c.qkv = lin_fwd(i, "qkv_w", c.a, lt["qkv_w"], keep=c.t)
        c.kc = torch.empty((B, H, Lq, 128), device=dev, dtype=BF)
        c.vc = torch.empty((B, H, Lq, 128), device=dev, dtype=BF)
        rope_write(c.qkv, c.kc, c.vc)
        q, k, v = _views(c.qkv, c.kc, c.vc, B, Lq, H)
        ctx, c.lse = L.flash_attention(q, k, v, causal=True, kv_len=lengths, return_lse=True)
        c.ctx = ctx.view(T, D)
        c.x_mid = lin_fwd(i, "o_w", c.ctx, lt["o_w"], residual=x, keep=c.t)
        c.b = L.rmsnorm(c.x_mid, lt["mlp_norm"], eps)
        c.z = lin_fwd(i, "gate_up_w", c.b, lt["gate_up_w"], keep=c.t)   # pre-activation kept for the backward
        c.h = L.swiglu_forward(c.z)
        if lora is not None and not want_output:                        # the adapter's t = dropout(h) A^T is needed
            a_cat, _ = lora.fused[i]["down_w"]
            c.t["down_w"] = L.gemm(L.dropout(c.h, p_drop, drop_seed(i, "down_w")) if p_drop > 0.0 else c.h, a_cat)
        y = lin_fwd(i, "down_w", c.h, lt["down_w"], residual=c.x_mid, keep=c.t) if want_output else None
        return c, y

    caches = []
    for i in range(len(layers)):
        c, y = layer_forward(i, x)
        caches.append(x if recompute else c)           # recompute: keep the layer input only
        x = y
    del c
    x_last = x
    xf = L.rmsnorm(x_last, final_norm, eps)
    logits = L.gemm(xf, model.lm_head)                           # bf16 like the reference's bf16 lm_head
    loss, wsum, dlogits = L.weighted_ce(logits.view(B, Lq, V), labels, vocab_weight, grad_scale=grad_scale,
                                        want_grad=True, inplace=True)
    dlogits = dlogits.view(T, V)

    # ---------------------------------------------------------------- backward
    g = grads if grads is not None else {}

    def slot(name, shape):
        if name not in g:
            # every slot of the decoder is written in full by its first producer (a GEMM or the norm backward with
            # accumulate = False): no memset of the 27 GB gradient set
            g[name] = torch.empty(shape, device=dev, dtype=torch.float32)
            return g[name], False
        return g[name], accumulate

    def lin_bwd(x_saved, w, dy, name, need_dx=True, layer=None, fname=None, t=None):
        dx = L.gemm_ex(dy, w, w_t=True) if need_dx else None
        if train_base or layer is None:
            dw, acc = slot(name, tuple(w.shape))
            L.gemm_ex(dy, x_saved, a_t=True, w_t=True, out=dw, accumulate=acc)
        if lora is not None and layer is not None:
            a_cat, b_blk = lora.fused[layer][fname]
            dt = L.gemm_ex(dy, b_blk, w_t=True, scale=lora.scale)                       # (T, g r)
            db, accb = slot(f"_lora.layers.{layer}.{fname}.B", tuple(b_blk.shape))
            L.gemm_ex(dy, t, a_t=True, w_t=True, out=db, accumulate=accb, scale=lora.scale)
            da, acca = slot(f"_lora.layers.{layer}.{fname}.A", tuple(a_cat.shape))
            if p_drop > 0.0:
                seed = drop_seed(layer, fname)
                L.gemm_ex(dt, L.dropout(x_saved, p_drop, seed), a_t=True, w_t=True, out=da, accumulate=acca)
                if need_dx:                         # dx += mask * (dt A) / (1 - p): the forward's mask, from its seed
                    L.dropout(L.gemm_ex(dt, a_cat, w_t=True), p_drop, seed, out=dx, accumulate=True)
            else:
                L.gemm_ex(dt, x_saved, a_t=True, w_t=True, out=da, accumulate=acca)
                if need_dx:
                    L.gemm_ex(dt, a_cat, w_t=True, out=dx, residual=dx)
        return dx

    def norm_bwd(x_saved, dy, gamma, name, add=None):
        if train_base:
            dg, acc = slot(name, (D,))
        else:                                      # frozen gain: the kernel still needs somewhere to put dgamma
            dg, acc = torch.empty(D, device=dev, dtype=torch.float32), False
        dx, _, _ = L.norm_backward(x_saved, dy, gamma, eps, rms=True, dgamma=dg, accumulate=acc, add=add)
        return dx

    if train_base:
        dxf = lin_bwd(xf, model.lm_head, dlogits, "lm_head.weight")
    else:
        dxf = L.gemm_ex(dlogits, model.lm_head, w_t=True)
    dx = norm_bwd(x_last, dxf, final_norm, "model.norm.weight")
    for i in range(len(layers) - 1, -1, -1):
        lt, c = layers[i], caches[i]
        if recompute:
            c, _ = layer_forward(i, c, want_output=False)
        p = f"_fused.layers.{i}."
        dh = lin_bwd(c.h, lt["down_w"], dx, p + "down_w", layer=i, fname="down_w", t=c.t.get("down_w"))
        dz = L.act_backward(c.z, dh, L.ACT_SWIGLU)
        db = lin_bwd(c.b, lt["gate_up_w"], dz, p + "gate_up_w", layer=i, fname="gate_up_w", t=c.t.get("gate_up_w"))
        dx_mid = norm_bwd(c.x_mid, db, lt["mlp_norm"], f"model.layers.{i}.post_attention_layernorm.weight", add=dx)
        dctx = lin_bwd(c.ctx, lt["o_w"], dx_mid, p + "o_w", layer=i, fname="o_w", t=c.t.get("o_w"))
        dqkv = torch.empty_like(c.qkv)
        dkc, dvc = torch.empty_like(c.kc), torch.empty_like(c.vc)
        q, k, v = _views(c.qkv, c.kc, c.vc, B, Lq, H)
        dq, dk, dv = _views(dqkv, dkc, dvc, B, Lq, H)
        L.flash_attention_bwd(q, k, v, c.ctx.view(B, Lq, H, 128), dctx.view(B, Lq, H, 128), c.lse, causal=True,
                              kv_len=lengths, out=(dq, dk, dv))
        L.check(lib.b200_rope_kv_backward(L.ptr(dqkv), None, L.ptr(rope_cos), L.ptr(rope_sin),
                                          cfg.max_position_embeddings, L.ptr(dkc), L.ptr(dvc), B, H, Lq, Lq,
                                          L.stream_ptr()), "b200_rope_kv_backward")
        da = lin_bwd(c.a, lt["qkv_w"], dqkv, p + "qkv_w", need_dx=True, layer=i, fname="qkv_w", t=c.t.get("qkv_w"))
        need = need_input_grad or i > 0
        dx = norm_bwd(c.x_in, da, lt["attn_norm"], f"model.layers.{i}.input_layernorm.weight", add=dx_mid) if need \
            else None
        caches[i] = c = None      # this layer's saved activations are dead: let the allocator reuse them for gradients
    d_embeds = dx.view(B, Lq, D) if dx is not None else None
    return loss, wsum, g, d_embeds


def unfuse_grads(g, cfg):
    """Gradients of the fused layouts -> the reference's parameter names (q/k/v_proj, gate/up_proj, ...)."""
    D, F = cfg.hidden_size, cfg.intermediate_size
    out = {k: v for k, v in g.items() if not k.startswith("_fused.")}
    for i in range(cfg.num_hidden_layers):
        p, r = f"_fused.layers.{i}.", f"model.layers.{i}."
        if p + "qkv_w" not in g:
            continue
        qkv = g[p + "qkv_w"]
        for j, n in enumerate(("q", "k", "v")):
            out[r + f"self_attn.{n}_proj.weight"] = qkv[j * D:(j + 1) * D]
        out[r + "self_attn.o_proj.weight"] = g[p + "o_w"]
        gu = g[p + "gate_up_w"].view(F, 2, D)
        out[r + "mlp.gate_proj.weight"] = gu[:, 0]
        out[r + "mlp.up_proj.weight"] = gu[:, 1]
        out[r + "mlp.down_proj.weight"] = g[p + "down_w"]
    return out
