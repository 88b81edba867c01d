"""The decode step's opt-in switches (include/b200_mmor.h: b200_set_option) and the tile widths behind them.

  "pdl"           programmatic dependent launch: the same kernels in the same arithmetic order, only their scheduling
                  overlaps -> tokens and logits must be BIT-identical with the switch on and off, through the eager
                  decode loop (return_logits) and through the captured CUDA graph.
  "decode_tiles"  weight-tile widths 96 / 160 / 224 for the wide projections: per output element the same k order
                  -> bit-identical as well; the new widths are also checked as plain GEMMs against torch.

Validated on B200 in round 2 (46 passed, profiles/r2a_pytest_switches.log), so the file runs with the suite. Timing
(tools/decode_bench.py, profiles/r2_decode_bench.json): "decode_tiles" = 1 is 4 % faster per decode step and became the
default; "pdl" is 4 % slower and stays off. The comparisons here are against both switches OFF (the module fixture
clears them and restores the defaults afterwards). Last file of the GPU suite on purpose."""
import math
import os

import pytest
import torch

import golden_cases as gc
from helpers import rel_err
from mm_or_b200 import _lib as L
from mm_or_b200.synth import synth_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def switches_off():
    saved = {name: int(L.get_option(name)) for name in ("pdl", "decode_tiles")}
    for name in saved:
        L.set_option(name, 0)
    yield
    for name, v in saved.items():
        L.set_option(name, v)


def _model(**cfg_kw):
    torch.set_grad_enabled(False)
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    cfg = gc.small_config(**cfg_kw)
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    model.config.tokenizer_padding_side = "left"
    return cfg, model


@pytest.fixture(scope="module")
def env():
    return _model()


def _with(option, fn, value=1):
    assert not L.get_option(option)
    try:
        L.set_option(option, value)
        assert L.get_option(option) == value
        out = fn()
        torch.cuda.synchronize()
    finally:
        L.set_option(option, False)
    return out


def test_unknown_option_is_an_error():
    with pytest.raises(L.B200Error):
        L.set_option("no_such_switch", 1)


@pytest.mark.parametrize("batch", [2, 40, 130])       # decode GEMM tiles MT = 32 / 64 / 256
def test_decode_step_bit_identical_with_pdl(env, batch):
    cfg, model = env
    b = synth_batch(cfg, batch, 1, 20, seed=60 + batch, jitter=4, image_pos=3)
    kw = dict(images=b["images"], max_new_tokens=12, stop_on_eos=False)
    ref_ids, ref_lg = model.generate(b["input_ids"], return_logits=True, **kw)           # eager loop
    ref_graph = model.generate(b["input_ids"], **kw)                                     # CUDA graph
    ids, lg = _with("pdl", lambda: model.generate(b["input_ids"], return_logits=True, **kw))
    ids_graph = _with("pdl", lambda: model.generate(b["input_ids"], **kw))
    assert torch.equal(ids, ref_ids) and torch.equal(lg, ref_lg)
    assert torch.equal(ids_graph, ref_graph) and torch.equal(ref_graph, ref_ids)


@pytest.mark.parametrize("option,value", [("pdl", 1), ("decode_tiles", 1), ("decode_tiles", 2)])
def test_wide_projection_on_the_tiled_kernel(option, value):
    """lm_head with 12288 rows: 96 weight tiles x 2 > 148 SMs, so stages.cu::linear sends it to the tiled kernel
    (gemm_sm100.cu): with "pdl" that kernel prefetches weights before the grid dependency resolves, with
    "decode_tiles" = 1 it runs 96-column tiles (128 tiles, one wave) instead of 128-column ones, with 2 it runs
    64-column tiles on half-depth rings, two CTAs per SM."""
    cfg, model = _model(vocab_size=12288)
    b = synth_batch(cfg, 3, 1, 16, seed=70, jitter=2, image_pos=2)
    kw = dict(images=b["images"], max_new_tokens=8, stop_on_eos=False)
    ref_ids, ref_lg = model.generate(b["input_ids"], return_logits=True, **kw)
    ids, lg = _with(option, lambda: model.generate(b["input_ids"], return_logits=True, **kw), value)
    ids_graph = _with(option, lambda: model.generate(b["input_ids"], **kw), value)
    assert torch.equal(ids, ref_ids) and torch.equal(lg, ref_lg) and torch.equal(ids_graph, ref_ids)


@pytest.mark.parametrize("bn", [96, 160, 224, -64, -96, -128])      # negative: two CTAs per SM, half-depth rings
@pytest.mark.parametrize("M,N,K", [(64, 22016, 4096), (128, 12288, 4096), (130, 32000, 512), (100, 264, 72),
                                   (1000, 1120, 1024)])
def test_new_tile_widths_as_plain_gemms(bn, M, N, K):
    """C = A W^T with the tile widths added for the decode step, incl. ragged N / K edges and several tiles per CTA,
    against torch fp32 on the same bf16 inputs; and bit-identical to the 128-column tiles."""
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = L.gemm(a, w, bn=bn)
    assert rel_err(out, a.float() @ w.float().t()) < 5e-3
    assert torch.equal(out, L.gemm(a, w, bn=128))


@pytest.mark.parametrize("bn", [96, 160, 224, -64, -96, -128])
def test_new_tile_widths_swiglu_epilogue(bn):
    M, N, K = 48, 22016, 512
    g = torch.Generator(device="cuda").manual_seed(8)
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = L.gemm(a, w, act=L.ACT_SWIGLU, bn=bn)
    assert out.shape == (M, N // 2) and torch.equal(out, L.gemm(a, w, act=L.ACT_SWIGLU, bn=128))


@pytest.mark.parametrize("batch", [3, 100, 200])
def test_all_switches_together_at_wide_shapes(batch):
    """pdl + two CTAs per SM on the shapes that reach the tiled kernel (vocabulary 12288, one and two row tiles)."""
    cfg, model = _model(vocab_size=12288)
    b = synth_batch(cfg, batch, 1, 16, seed=80 + batch, jitter=2, image_pos=2)
    kw = dict(images=b["images"], max_new_tokens=6, stop_on_eos=False)
    ref = model.generate(b["input_ids"], **kw)
    try:
        L.set_option("pdl", 1)
        L.set_option("decode_tiles", 2)
        out = model.generate(b["input_ids"], **kw)
        torch.cuda.synchronize()
    finally:
        L.set_option("pdl", 0)
        L.set_option("decode_tiles", 0)
    assert torch.equal(out, ref)


def test_fused_rope_kv_append_against_oracle_and_separate_kernel():
    """The decode attention kernel rotates q / k and appends k / v itself whenever the step has >= 2 x SMs (row, head)
    pairs (DecodeArgs::rope_k; default on). 96 rows x 4 heads = 384 CTAs take that path: logits against the CPU oracle,
    against the same step with the separate RoPE kernel ("fused_rope" = 0; the only difference is the summation order of
    the one new key's dot product), same greedy ids in the eager and the CUDA-graph loop, and a left-padded batch so
    that the rotary position (slot - kv_start) differs per row."""
    from helpers import TOL_E2E, oracle_cfg
    from oracle import mm2sg_oracle as O
    cfg, model = _model()
    ocfg = oracle_cfg(cfg)
    sd = gc.bf16_round(gc.small_weights(cfg))
    g = torch.Generator().manual_seed(321)
    B, Lt, steps = 96, 20, 6
    ids = torch.zeros(B, Lt, dtype=torch.long)
    for r in range(B):
        n = int(torch.randint(9, Lt + 1, (1,), generator=g))
        ids[r, Lt - n:] = torch.randint(3, cfg.vocab_size, (n,), generator=g)
    assert L.get_option("fused_rope")
    out, lg = model.generate(ids, max_new_tokens=steps, stop_on_eos=False, return_logits=True)
    out_graph = model.generate(ids, max_new_tokens=steps, stop_on_eos=False)
    assert torch.equal(out_graph, out)
    try:
        L.set_option("fused_rope", 0)
        out0, lg0 = model.generate(ids, max_new_tokens=steps, stop_on_eos=False, return_logits=True)
    finally:
        L.set_option("fused_rope", 1)
    assert rel_err(lg, lg0) < 2e-3
    mask = ids.ne(0)
    pos = (mask.long().cumsum(-1) - 1).clamp(min=0)
    emb = sd["model.embed_tokens.weight"][ids] * mask[..., None]
    l0, kv = O.llama_forward(sd, emb, mask, pos, ocfg.llm, last_only=True)
    toks, ref_lg = O.greedy_decode(sd, ocfg, l0[:, -1], kv, mask, steps, stop_on_eos=False, forced_tokens=out[:, Lt:].cpu())
    assert rel_err(lg, ref_lg) < TOL_E2E and rel_err(lg0, ref_lg) < TOL_E2E
    err = (lg.cpu().float() - ref_lg).abs().max().item()
    top2 = ref_lg.topk(2, -1).values
    safe = (top2[..., 0] - top2[..., 1]) > 2 * err
    assert safe.float().mean() > 0.5 and torch.equal(out[:, Lt:].cpu()[safe], toks[safe])
    assert torch.equal(out0[:, Lt:].cpu()[safe], toks[safe])
