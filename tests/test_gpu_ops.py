"""GPU parity tests of the individual kernels, called through the C ABI (ctypes), against fp32 PyTorch restatements
of the same operator evaluated on the same bf16-rounded inputs (floating-point kernels: tolerance TOL_OP, stated in
tests/helpers.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import TOL_OP, max_err, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from mm_or_b200 import _lib
    _lib.lib()
    return _lib


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


# ---------------------------------------------------------------- GEMM ---------------------------------------------
def ref_gemm(a, w, bias=None, residual=None, act=0):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = F.gelu(y)
    elif act == 3:
        y = F.silu(y[:, 0::2]) * y[:, 1::2]
    if residual is not None:
        y = y + residual.float()
    return y


@pytest.mark.parametrize("M,N,K", [(1, 256, 64), (128, 256, 64), (100, 264, 72), (577 * 3, 3072, 1024),
                                   (1000, 4096, 11008), (129, 32000, 512), (3456, 1024, 4096), (64, 22016, 4096)])
@pytest.mark.parametrize("bn", [0, 256, 64, 32])
def test_gemm_shapes(L, M, N, K, bn):
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=1 / math.sqrt(K), seed=2)
    assert rel_err(L.gemm(a, w, bn=bn), ref_gemm(a, w)) < 5e-3


def test_gemm_empty(L):
    a, w = rnd(0, 64), rnd(16, 64)
    assert L.gemm(a, w).shape == (0, 16)


@pytest.mark.parametrize("bn", [256, 128])
def test_gemm_epilogues(L, bn):
    M, N, K = 700, 1024, 512
    a, w = rnd(M, K, seed=3), rnd(N, K, scale=1 / math.sqrt(K), seed=4)
    bias, res = rnd(N, seed=5), rnd(M, N, seed=6)
    assert rel_err(L.gemm(a, w, bias=bias, bn=bn), ref_gemm(a, w, bias)) < 5e-3
    assert rel_err(L.gemm(a, w, bias=bias, act=1, bn=bn), ref_gemm(a, w, bias, act=1)) < 5e-3
    assert rel_err(L.gemm(a, w, bias=bias, act=2, bn=bn), ref_gemm(a, w, bias, act=2)) < 5e-3
    assert rel_err(L.gemm(a, w, act=3, bn=bn), ref_gemm(a, w, act=3)) < 5e-3
    assert rel_err(L.gemm(a, w, bias=bias, residual=res, bn=bn), ref_gemm(a, w, bias, res)) < 5e-3
    assert rel_err(L.gemm(a, w, bias=bias, out_fp32=True, bn=bn), ref_gemm(a, w, bias)) < 1e-5
    x = res.clone()
    L.gemm(a, w, out=x, bias=bias, residual=x, bn=bn)           # in-place residual, as the decoder uses it
    assert rel_err(x, ref_gemm(a, w, bias, res)) < 5e-3
    perm = torch.randperm(M, device="cuda").int()
    perm[::7] = -1
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, w, out=out, bias=bias, row_map=perm, bn=bn)
    ref = torch.zeros(M, N, device="cuda")
    keep = perm >= 0
    ref[perm[keep].long()] = ref_gemm(a, w, bias)[keep]
    assert rel_err(out, ref) < 5e-3


# ---------------------------------------------------------------- decode-step GEMM --------------------------------
@pytest.mark.parametrize("M", [1, 7, 32, 33, 64, 100, 128, 200, 256])
@pytest.mark.parametrize("N,K", [(4096, 4096), (12288, 4096), (4096, 11008), (32000, 4096), (264, 72), (1024, 512)])
def test_gemm_skinny_shapes(L, M, N, K):
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=1 / math.sqrt(K), seed=2)
    assert rel_err(L.gemm_skinny(a, w), ref_gemm(a, w)) < 5e-3


@pytest.mark.parametrize("splits", [0, 1, 2, 4, 8])
def test_gemm_skinny_splits_and_epilogues(L, splits):
    M, N, K = 64, 1024, 2048
    a, w = rnd(M, K, seed=3), rnd(N, K, scale=1 / math.sqrt(K), seed=4)
    bias, res = rnd(N, seed=5), rnd(M, N, seed=6)
    kw = dict(splits=splits)
    assert rel_err(L.gemm_skinny(a, w, bias=bias, **kw), ref_gemm(a, w, bias)) < 5e-3
    assert rel_err(L.gemm_skinny(a, w, bias=bias, act=1, **kw), ref_gemm(a, w, bias, act=1)) < 5e-3
    assert rel_err(L.gemm_skinny(a, w, bias=bias, act=2, **kw), ref_gemm(a, w, bias, act=2)) < 5e-3
    assert rel_err(L.gemm_skinny(a, w, act=3, **kw), ref_gemm(a, w, act=3)) < 5e-3
    assert rel_err(L.gemm_skinny(a, w, bias=bias, residual=res, **kw), ref_gemm(a, w, bias, res)) < 5e-3
    assert rel_err(L.gemm_skinny(a, w, bias=bias, out_fp32=True, **kw), ref_gemm(a, w, bias)) < 1e-5
    x = res.clone()
    L.gemm_skinny(a, w, out=x, bias=bias, residual=x, **kw)        # in-place residual, as the decoder uses it
    assert rel_err(x, ref_gemm(a, w, bias, res)) < 5e-3


def test_gemm_skinny_deterministic(L):
    """Split-K partials are summed in split order through distributed shared memory: repeated launches are
    bit-identical, for every cluster size."""
    M, N, K = 48, 4096, 11008
    a, w = rnd(M, K, seed=7), rnd(N, K, scale=1 / math.sqrt(K), seed=8)
    for splits in (0, 1, 2, 4, 8):
        first = L.gemm_skinny(a, w, splits=splits)
        for _ in range(3):
            assert torch.equal(L.gemm_skinny(a, w, splits=splits), first)
        assert rel_err(first, ref_gemm(a, w)) < 5e-3
        assert rel_err(first, L.gemm(a, w)) < 2e-3     # vs the tiled prefill kernel: fp32 summation order only


def test_gemm_skinny_bad_args(L):
    with pytest.raises(L.B200Error):
        L.gemm_skinny(rnd(300, 64), rnd(128, 64))     # M > 256
    with pytest.raises(L.B200Error):
        L.gemm_skinny(rnd(8, 60), rnd(128, 60))       # K not a multiple of 8


def test_gemm_bad_args(L):
    with pytest.raises(L.B200Error):
        L.gemm(rnd(8, 60), rnd(16, 60))         # K not a multiple of 8


# ---------------------------------------------------------------- norms --------------------------------------------
@pytest.mark.parametrize("D", [512, 1024, 4096])
def test_layernorm_rmsnorm(L, D):
    M = 777
    x, g, b = rnd(M, D, seed=1), rnd(D, scale=0.1, seed=2) + 1, rnd(D, scale=0.1, seed=3)
    ref = F.layer_norm(x.float(), (D,), g.float(), b.float(), 1e-5)
    assert rel_err(L.layernorm(x, g, b, 1e-5), ref) < 5e-3
    xf = x.float()
    ref = g.float() * xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)
    assert rel_err(L.rmsnorm(x, g, 1e-5), ref) < 5e-3


def test_layernorm_gather_add(L):
    D, rows, period = 1024, 300, 50
    x, g, b, add = rnd(200, D, seed=1), rnd(D, seed=2), rnd(D, seed=3), rnd(period, D, seed=4)
    rm = torch.randint(-1, 200, (rows,), device="cuda", dtype=torch.int32)
    src = torch.where(rm[:, None] >= 0, x.float()[rm.clamp(min=0).long()], torch.zeros(1, D, device="cuda"))
    src = src + add.float()[torch.arange(rows, device="cuda") % period]
    ref = F.layer_norm(src, (D,), g.float(), b.float(), 1e-12)
    assert rel_err(L.layernorm(x, g, b, 1e-12, row_map=rm, add=add), ref) < 5e-3


# ---------------------------------------------------------------- attention ----------------------------------------
def ref_attn(q, k, v, causal=False, kv_start=None, kv_len=None):
    B, Lq, H, d = q.shape
    Lk = k.shape[1]
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) / math.sqrt(d)
    kj = torch.arange(Lk, device=q.device)
    vis = torch.ones(B, 1, Lq, Lk, dtype=torch.bool, device=q.device)
    if causal:
        qi = torch.arange(Lq, device=q.device) + (Lk - Lq)
        vis = vis & (kj[None, :] <= qi[:, None])[None, None]
    if kv_start is not None:
        vis = vis & (kj[None, :] >= kv_start[:, None].long())[:, None, None, :]
    if kv_len is not None:
        vis = vis & (kj[None, :] < kv_len[:, None].long())[:, None, None, :]
    s = s.masked_fill(~vis, float("-inf"))
    p = torch.softmax(s, -1).nan_to_num(0.0)       # fully masked rows -> zeros (never consumed downstream)
    return torch.einsum("bhqk,bkhd->bqhd", p, v.float())


@pytest.mark.parametrize("d,H,Lq", [(64, 16, 577), (128, 8, 1152), (128, 4, 61)])
def test_flash_attention_plain_and_keypad(L, d, H, Lq):
    B = 3
    qkv = rnd(B, Lq, 3, H, d, seed=1)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]          # packed, strided views
    assert rel_err(L.flash_attention(q, k, v), ref_attn(q, k, v)) < TOL_OP
    kv_len = torch.tensor([Lq, Lq // 2, 1], device="cuda", dtype=torch.int32)
    assert rel_err(L.flash_attention(q, k, v, kv_len=kv_len), ref_attn(q, k, v, kv_len=kv_len)) < TOL_OP


def test_flash_attention_causal_leftpad(L):
    B, H, Lq, d = 3, 4, 333, 128
    q, k, v = rnd(B, Lq, H, d, seed=1), rnd(B, Lq, H, d, seed=2), rnd(B, Lq, H, d, seed=3)
    start = torch.tensor([0, 70, 200], device="cuda", dtype=torch.int32)
    out = L.flash_attention(q, k, v, causal=True, kv_start=start)
    ref = ref_attn(q, k, v, causal=True, kv_start=start)
    assert torch.isfinite(out.float()).all()                     # pad query rows must not produce NaN/Inf
    for b in range(B):
        s = int(start[b])
        assert rel_err(out[b, s:], ref[b, s:]) < TOL_OP
    kv_len = torch.tensor([333, 100, 5], device="cuda", dtype=torch.int32)
    out = L.flash_attention(q, k, v, causal=True, kv_len=kv_len)
    ref = ref_attn(q, k, v, causal=True, kv_len=kv_len)
    for b in range(B):
        n = int(kv_len[b])
        assert rel_err(out[b, :n], ref[b, :n]) < TOL_OP
    # fewer queries than keys (last-layer pooler shape): Lq < Lk, non-causal
    out = L.flash_attention(q[:, :100], k, v)
    assert rel_err(out, ref_attn(q[:, :100], k, v)) < TOL_OP


def test_flash_attention_rescale_and_long_keys(L):
    """Scores whose running maximum jumps by far more than 2^8 between key tiles (the lazy-rescale path of the
    tcgen05 kernel rescales the TMEM accumulator there), many key tiles, Lq != Lk with the causal offset."""
    B, H, d = 2, 2, 128
    Lq, Lk = 200, 1500
    q, k, v = rnd(B, Lq, H, d, seed=11), rnd(B, Lk, H, d, seed=12), rnd(B, Lk, H, d, seed=13)
    k = k.clone()
    k[:, 700:] *= 6.0                    # later keys dominate: max rises tile after tile
    k[:, 1300:] *= 3.0
    q = (q * 2.0).contiguous()
    assert rel_err(L.flash_attention(q, k, v), ref_attn(q, k, v)) < TOL_OP
    assert rel_err(L.flash_attention(q, k, v, causal=True), ref_attn(q, k, v, causal=True)) < TOL_OP
    d = 64
    q, k, v = rnd(B, 577, H, d, seed=14) * 3, rnd(B, 577, H, d, seed=15) * 3, rnd(B, 577, H, d, seed=16)
    assert rel_err(L.flash_attention(q, k, v), ref_attn(q, k, v)) < TOL_OP


def test_flash_attention_lse(L):
    B, H, Lq, d = 2, 3, 333, 128
    q, k, v = rnd(B, Lq, H, d, seed=31), rnd(B, Lq, H, d, seed=32), rnd(B, Lq, H, d, seed=33)
    start = torch.tensor([0, 100], device="cuda", dtype=torch.int32)
    o, lse = L.flash_attention(q, k, v, causal=True, kv_start=start, return_lse=True)
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) / math.sqrt(d)
    kj = torch.arange(Lq, device="cuda")
    vis = (kj[None, :] <= kj[:, None])[None, None] & (kj[None, None, None, :] >= start[:, None, None, None].long())
    ref = torch.logsumexp(s.masked_fill(~vis, float("-inf")), dim=-1)
    live = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(lse), live)
    assert float((lse[live] - ref[live]).abs().max()) < 2e-2
    assert rel_err(o, L.flash_attention(q, k, v, causal=True, kv_start=start)) == 0.0


def test_flash_attention_kv_cache_layout(L):
    """K / V read from the head-major KV-cache layout [B][H][cap][128] (row stride 128, head stride cap*128), Q and O
    in the token-major qkv layout: the prefill call site of b200_llama_prefill."""
    B, H, Lq, cap, d = 2, 3, 300, 384, 128
    qkv = rnd(B, Lq, 3, H, d, seed=21)
    kc, vc = rnd(B, H, cap, d, seed=22), rnd(B, H, cap, d, seed=23)
    q = qkv[:, :, 0]
    k, v = kc[:, :, :Lq].transpose(1, 2), vc[:, :, :Lq].transpose(1, 2)      # (B, Lq, H, d) views
    start = torch.tensor([0, 37], device="cuda", dtype=torch.int32)
    out = L.flash_attention(q, k, v, causal=True, kv_start=start)
    ref = ref_attn(q, k, v, causal=True, kv_start=start)
    for b in range(B):
        s0 = int(start[b])
        assert rel_err(out[b, s0:], ref[b, s0:]) < TOL_OP
        assert float(out[b, :s0].float().abs().max()) == 0.0 if s0 else True   # fully masked rows: exact zeros


@pytest.mark.parametrize("B,ctx,splits", [(5, 700, 0), (2, 1500, 4), (64, 300, 1), (1, 9, 0)])
def test_decode_attention(L, B, ctx, splits):
    H, cap = 4, 1536
    kc, vc = rnd(B, H, cap, 128, seed=1), rnd(B, H, cap, 128, seed=2)
    q = rnd(B, H * 128, seed=3)
    start = torch.randint(0, max(1, ctx // 3), (B,), device="cuda", dtype=torch.int32)
    out = L.decode_attention(q, kc, vc, ctx, kv_start=start, splits=splits)
    qq = q.view(B, 1, H, 128)
    kk, vv = kc[:, :, :ctx].transpose(1, 2), vc[:, :, :ctx].transpose(1, 2)
    ref = ref_attn(qq, kk, vv, kv_start=start).reshape(B, H * 128)
    assert rel_err(out, ref) < TOL_OP


# ---------------------------------------------------------------- data movement ------------------------------------
def test_rope_kv_write(L):
    B, H, Lq, cap, max_pos = 2, 4, 37, 64, 256
    qkv = rnd(B * Lq, 3 * H * 128, seed=1)
    orig = qkv.clone()
    inv = 1.0 / (10000.0 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
    fr = torch.outer(torch.arange(max_pos, dtype=torch.float32), inv)
    cos, sin = fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()
    kc = torch.zeros(B, H, cap, 128, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    start = torch.tensor([0, 5], device="cuda", dtype=torch.int32)
    slot0 = 3
    L.check(L.lib().b200_rope_kv_write(L.ptr(qkv), L.ptr(start), L.ptr(cos), L.ptr(sin), max_pos, L.ptr(kc), L.ptr(vc),
                                       B, H, Lq, slot0, cap, L.stream_ptr()))
    x = orig.float().view(B, Lq, 3, H, 128)
    pos = (slot0 + torch.arange(Lq, device="cuda"))[None, :] - start[:, None].long()
    pos = pos.clamp(min=0)
    c, s = torch.cat([cos, cos], -1)[pos][:, :, None], torch.cat([sin, sin], -1)[pos][:, :, None]

    def rope(t):
        rot = torch.cat([-t[..., 64:], t[..., :64]], -1)
        return t * c + rot * s

    assert rel_err(qkv.view(B, Lq, 3, H, 128)[:, :, 0], rope(x[:, :, 0])) < 5e-3
    assert rel_err(kc[:, :, slot0:slot0 + Lq].transpose(1, 2), rope(x[:, :, 1])) < 5e-3
    assert torch.equal(vc[:, :, slot0:slot0 + Lq].transpose(1, 2), orig.view(B, Lq, 3, H, 128)[:, :, 2])
    assert torch.equal(qkv.view(B, Lq, 3, H, 128)[:, :, 1:], orig.view(B, Lq, 3, H, 128)[:, :, 1:])
    assert kc[:, :, :slot0].abs().sum() == 0 and kc[:, :, slot0 + Lq:].abs().sum() == 0


def test_embed_rows_and_argmax(L):
    V, D, rows = 300, 512, 50
    table = rnd(V, D, seed=1)
    ids = torch.randint(-2, V, (rows,), device="cuda", dtype=torch.int32)
    out = torch.full((rows, D), 7.0, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib().b200_embed_rows(L.ptr(ids), L.ptr(table), L.ptr(out), D, rows, D, V, L.stream_ptr()))
    for r in range(rows):
        i = int(ids[r])
        exp = table[i] if i >= 0 else (torch.zeros(D, device="cuda") if i == -1 else torch.full((D,), 7.0, device="cuda"))
        assert torch.equal(out[r].float(), exp.float())
    # argmax: bit-exact index work, lowest index wins ties, bf16 and fp32 inputs
    for dt in (torch.bfloat16, torch.float32):
        lg = torch.randn(9, 32000, device="cuda").to(dt)
        lg[3, 100] = lg[3, 20000] = 50.0
        lg[4] = 0.0
        tok = torch.empty(9, device="cuda", dtype=torch.int32)
        L.check(L.lib().b200_argmax(L.ptr(lg), int(dt == torch.float32), 32000, 9, 32000, L.ptr(tok), None, 2, 0,
                                    L.stream_ptr()))
        exp = lg.float().argmax(-1)
        exp[3], exp[4] = 100, 0
        assert torch.equal(tok.long(), exp)
    # finished-row semantics of HF greedy_search
    lg = torch.zeros(3, 16, device="cuda", dtype=torch.bfloat16)
    lg[0, 2] = lg[1, 5] = lg[2, 7] = 1.0
    fin = torch.tensor([0, 0, 1], device="cuda", dtype=torch.int32)
    tok = torch.empty(3, device="cuda", dtype=torch.int32)
    L.check(L.lib().b200_argmax(L.ptr(lg), 0, 16, 3, 16, L.ptr(tok), L.ptr(fin), 2, 0, L.stream_ptr()))
    assert tok.tolist() == [2, 5, 0] and fin.tolist() == [1, 0, 1]


def test_patchify_matches_conv(L):
    N, S, P, D = 3, 336, 14, 64
    px = rnd(N, 3, S, S, seed=1)
    w = rnd(D, 3, P, P, scale=0.05, seed=2)
    kpad = (3 * P * P + 7) // 8 * 8
    cols = torch.empty(N * (S // P) ** 2, kpad, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib().b200_patchify(L.ptr(px), L.ptr(cols), N, 3, S, P, kpad, L.stream_ptr()))
    ref_cols = F.unfold(px.float(), P, stride=P).transpose(1, 2).reshape(-1, 3 * P * P)
    assert torch.equal(cols[:, :3 * P * P].float(), ref_cols)          # pure data movement: bit exact
    assert cols[:, 3 * P * P:].abs().sum() == 0
    wp = torch.zeros(D, kpad, device="cuda", dtype=torch.bfloat16)
    wp[:, :3 * P * P] = w.reshape(D, -1)
    ref = F.conv2d(px.float(), w.float(), stride=P).flatten(2).transpose(1, 2).reshape(-1, D)
    assert rel_err(L.gemm(cols, wp), ref) < 5e-3
