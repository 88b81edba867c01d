"""GPU: the NF4 storage kernels (csrc/nf4.cu) through the C ABI against the numpy restatement -- bit-exact (byte /
integer work) -- at a real projection's width, and the QLoRA arm of the fine-tune step (frozen base replaced by its NF4
round trip). Written without GPU access (the same kernel source runs bit-exactly on the CPU emulator, tests/test_nf4.py);
its own late file."""
import numpy as np
import pytest
import torch

from oracle import nf4_oracle as O

pytestmark = pytest.mark.gpu


def bits(t):
    return t.contiguous().cpu().view(torch.int16).numpy().view(np.uint16)


def test_nf4_kernels_bit_exact_on_a_projection_sized_weight():
    from mm_or_b200.train import nf4 as N
    g = torch.Generator(device="cuda").manual_seed(0)
    w = (torch.randn(1024, 4096, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    w[5] = 0                                                            # all-zero blocks
    w[7, ::64] = 3.0                                                    # one outlier per block
    packed, absmax = N.quantize(w)
    ref_p, ref_a = O.quantize(bits(w))
    assert np.array_equal(packed.cpu().numpy(), ref_p) and np.array_equal(absmax.cpu().numpy(), ref_a)
    out = N.dequantize(packed, absmax, w.shape)
    assert np.array_equal(bits(out).reshape(-1), O.dequantize(ref_p, ref_a).reshape(-1))
    # idempotent storage, and the double-quantised block maxima stay within the 8-bit code's resolution
    p2, a2 = N.quantize(out)
    assert torch.equal(p2, packed) and torch.equal(a2, absmax)
    q = N.Nf4Weight(w, double_quant=True)
    bm = q.block_maxima()
    assert float((bm - absmax).abs().max()) < 0.02 * float(absmax.max())
    assert q.nbytes() < 0.26 * w.numel() * 2


def test_qlora_step_trains_against_the_quantised_base():
    import golden_cases as gc
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train.lora import LoraState
    from mm_or_b200.train.step import FineTuner
    torch.set_grad_enabled(False)
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.bf16_round(gc.small_weights(cfg))
    case = gc.make_case(cfg, "train_right")
    args = (case["input_ids"], case["labels"], case["attention_mask"], case["images"])
    losses = {}
    for nf4 in (False, True):
        model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
        lora = LoraState(cfg, r=8, alpha=16, device="cuda", seed=1)
        ft = FineTuner(model, sd, lr=1e-3, max_grad_norm=0.1, first_trainable_clip_layer=1, lora=lora, base_nf4=nf4)
        k = "model.layers.0.self_attn.q_proj.weight"
        if nf4:
            assert ft.nf4_bytes and not torch.equal(ft.sd[k].cpu().float(), sd[k])        # base replaced by Q(W) ...
            assert torch.equal(ft.sd["model.mm_projector.0.weight"].cpu().float(), sd["model.mm_projector.0.weight"])
        loss0, _ = ft.train_step(*args)
        loss1, _ = ft.train_step(*args)
        losses[nf4] = (float(loss0), float(loss1))
        assert np.isfinite(losses[nf4]).all() and losses[nf4][1] != losses[nf4][0]        # ... and the step still updates
    # 4-bit base: a different, slightly worse starting loss (quantisation error), same order of magnitude
    assert losses[True][0] != losses[False][0] and abs(losses[True][0] - losses[False][0]) < 0.5 * losses[False][0]


def test_lora_dropout_masks_are_reproducible_from_the_seed():
    """lora_dropout (0.05 in the recipe; 0.5 here so that it matters): the adapter branch sees dropout(x). The mask is a
    function of (seed, step, layer, projection, element), so the backward re-creates the forward's mask: the same step
    counter gives bit-identical loss and gradients, the next step another mask; p = 0 is the deterministic path."""
    import golden_cases as gc
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train.lora import LoraState
    from mm_or_b200.train.step import FineTuner
    torch.set_grad_enabled(False)
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.bf16_round(gc.small_weights(cfg))
    case = gc.make_case(cfg, "train_right")
    args = (case["input_ids"], case["labels"], case["attention_mask"], case["images"])
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    lora = LoraState(cfg, r=8, alpha=16, device="cuda", seed=1, init_b="randn", dropout=0.5)
    ft = FineTuner(model, sd, lr=1e-3, max_grad_norm=0.1, first_trainable_clip_layer=1, lora=lora)
    runs = []
    for step in (0, 0, 1):
        lora.rng_step = step                                            # forward_backward advances it by one
        loss, _, g = ft.forward_backward(*args)
        runs.append((float(loss), {k: v.clone() for k, v in g.items() if k.startswith("_lora.")}))
    assert runs[0][0] == runs[1][0] and all(torch.equal(runs[0][1][k], runs[1][1][k]) for k in runs[0][1])
    assert runs[2][0] != runs[0][0]
    assert all(bool(torch.isfinite(v).all()) for v in runs[0][1].values())
    assert any(float(v.abs().max()) > 0 for k, v in runs[0][1].items() if k.endswith(".A"))
    lora.dropout = 0.0
    base = [float(ft.forward_backward(*args)[0]) for _ in range(2)]
    assert base[0] == base[1] and base[0] != runs[0][0]


def test_pooler_hidden_dropout_is_reproducible_and_off_by_default():
    """HF BertModel's train-mode hidden dropout in the image pooler (hidden_dropout_prob 0.1; 0.3 here): same step
    counter => bit-identical loss and pooler gradients (the backward re-creates the masks), next step => other masks,
    p = 0 (the default) => the deterministic path the parity tests use."""
    import golden_cases as gc
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.train.step import FineTuner
    torch.set_grad_enabled(False)
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "right"
    sd = gc.bf16_round(gc.small_weights(cfg))
    case = gc.make_case(cfg, "train_right")
    args = (case["input_ids"], case["labels"], case["attention_mask"], case["images"])
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    ft = FineTuner(model, sd, lr=1e-3, max_grad_norm=0.1, first_trainable_clip_layer=1)
    pooler = model.get_image_pooler()
    assert float(getattr(pooler, "dropout", 0.0)) == 0.0
    base = [float(ft.forward_backward(*args)[0]) for _ in range(2)]
    assert base[0] == base[1]
    pooler.dropout, pooler.seed = 0.3, 11
    key = "model.image_pooler.bert.encoder.layer.0.output.dense.weight"
    runs = []
    for step in (0, 0, 1):
        pooler.rng_step = step
        loss, _, g = ft.forward_backward(*args)
        runs.append((float(loss), g[key].clone()))
    assert runs[0][0] == runs[1][0] and torch.equal(runs[0][1], runs[1][1])
    assert runs[2][0] != runs[0][0] and runs[0][0] != base[0]
    assert bool(torch.isfinite(runs[0][1]).all()) and float(runs[0][1].abs().max()) > 0
    pooler.dropout = 0.0
    assert float(ft.forward_backward(*args)[0]) == base[0]


@pytest.mark.parametrize("M,N,K", [(300, 1000, 1024), (128, 256, 64), (3924, 12288, 4096), (777, 4096, 11008)])
def test_fused_nf4_gemm_equals_dequantise_then_gemm(M, N, K):
    """b200_gemm_nf4: the NF4 weight is expanded to bf16 inside the GEMM (four dequantisation warps write the swizzled B
    stage that the TMA would have written) -- bit-identical to b200_gemm_bf16 on b200_nf4_dequantize(codes, absmax):
    same tiles, same MMA order, same epilogue. Ragged N (rows beyond N are zero), both tile widths (128 / 256), a
    residual + scale epilogue, K = 64 (one block)."""
    from mm_or_b200 import _lib as L
    from mm_or_b200.train import nf4 as N4
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    w[N // 2] = 0                                                          # an all-zero block row
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    packed, absmax = N4.quantize(w)
    wq = N4.dequantize(packed, absmax, w.shape)
    ref = L.gemm(x, wq)
    got = L.gemm_nf4(x, packed, absmax, N)
    torch.cuda.synchronize()
    assert got.shape == ref.shape and torch.isfinite(got.float()).all()
    assert torch.equal(got, ref), (got.float() - ref.float()).abs().max().item()
    res = torch.randn(M, N, generator=g, device="cuda").to(torch.bfloat16)
    ref2 = L.gemm_ex(x, wq, residual=res, scale=0.5)
    got2 = L.gemm_nf4(x, packed, absmax, N, residual=res, scale=0.5)
    assert torch.equal(got2, ref2)
    if M >= 3000:                                                           # timing note for the record (not asserted)
        def t(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 10
        print("fused nf4 gemm %dx%dx%d: %.3f ms vs bf16 gemm %.3f ms (+ dequantise %.3f ms)"
              % (M, N, K, t(lambda: L.gemm_nf4(x, packed, absmax, N)), t(lambda: L.gemm(x, wq)),
                 t(lambda: N4.dequantize(packed, absmax, w.shape))))
    with pytest.raises(L.B200Error):
        L.gemm_nf4(x[:, :32].contiguous(), packed[:N * 16], absmax[:N // 2], N)      # K = 32: not a block multiple
