"""CPU: the NF4 storage kernels (mm_or_b200/csrc/nf4.cu, executed through the kernel emulator) against the numpy
restatement of bitsandbytes' published algorithm (oracle/nf4_oracle.py) -- byte / integer work, so the bar is bit-exact
-- plus the known answers the restatement itself can be held to without bitsandbytes (not installed: parity with the
library itself is unpinned, see the oracle's header)."""
import numpy as np
import pytest
import torch

import emu_lib
from mm_or_b200.train import nf4 as N
from oracle import nf4_oracle as O


@pytest.fixture(scope="module")
def ops():
    cdll = emu_lib.lib()

    def ptr(t):
        import ctypes
        assert t is None or not t.is_cuda
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def check(rc, what=""):
        if rc != 0:
            raise RuntimeError(f"{what} failed with code {rc}: {cdll.b200_last_error().decode()}")

    return N.Nf4Ops(cdll, ptr, lambda: None, check)


def bits(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def test_level_table_and_thresholds_are_the_published_ones():
    lv = O.LEVELS
    assert lv.shape == (16,) and lv[0] == -1.0 and lv[7] == 0.0 and lv[15] == 1.0 and (np.diff(lv) > 0).all()
    # QLoRA appendix E: the table is asymmetric (8 positive, 7 negative levels around an exact zero)
    assert (lv > 0).sum() == 8 and (lv < 0).sum() == 7
    # every level is a float32 value (the library stores them as floats)
    assert np.array_equal(lv.astype(np.float32).astype(np.float64), lv)
    # decision-tree thresholds are the midpoints; the literals of the kernel are these float32 values
    want = [-0.8480964004993439, -0.6106329262256622, -0.4599952697753906, -0.33967943489551544, -0.23460740596055984,
            -0.13791173323988914, -0.045525018125772476, 0.03979014977812767, 0.1202552504837513, 0.2035212516784668,
            0.2920137718319893, 0.3893125355243683, 0.5016634166240692, 0.6427869200706482, 0.8614784181118011]
    assert np.array_equal(O.THRESHOLDS, np.asarray(want, dtype=np.float32))
    src = open(emu_lib.SRCS[2]).read()
    for lit in want:
        assert repr(lit).lstrip("-") + "f" in src, lit


def test_kernels_match_the_restatement_bit_for_bit(ops):
    g = torch.Generator().manual_seed(0)
    cases = [torch.randn(64 * 37, generator=g) * 0.02,                        # weight-like
             torch.randn(64 * 5, generator=g) * 1e3,
             torch.zeros(64 * 2),                                              # all-zero blocks: 0 * inf = NaN -> code 0
             torch.cat([torch.zeros(63), torch.tensor([-3.0])]),               # one outlier
             torch.tensor(O.LEVELS, dtype=torch.float32).repeat(4) * 0.5,      # exactly on the grid
             torch.tensor(np.concatenate([np.repeat(O.THRESHOLDS, 4), [1.0, -1.0, 0.0, 0.5]]).astype(np.float32))]
    # (last case: values on / next to the midpoints after bf16 rounding, absmax 1 so that they stay where they are)
    for w in cases:
        w = w.to(torch.bfloat16)
        packed, absmax = N.quantize(w, ops)
        ref_p, ref_a = O.quantize(bits(w))
        assert np.array_equal(packed.numpy(), ref_p) and np.array_equal(absmax.numpy(), ref_a)
        out = N.dequantize(packed, absmax, w.shape, ops=ops)
        assert np.array_equal(bits(out), O.dequantize(ref_p, ref_a))


def test_round_trip_properties(ops):
    g = torch.Generator().manual_seed(1)
    w = (torch.randn(128, 192, generator=g) * 0.03).to(torch.bfloat16)
    packed, absmax = N.quantize(w, ops)
    q = N.dequantize(packed, absmax, w.shape, ops=ops)
    # nearest level: |w - Q(w)| <= half the largest level gap x absmax of the block (+ one bf16 rounding)
    gap = np.diff(O.LEVELS).max() / 2
    err = (w.float() - q.float()).abs().view(-1, 64)
    assert bool((err <= gap * absmax[:, None] * (1 + 2 ** -7)).all())
    # the block maximum itself is reproduced exactly (level +-1), so absmax is stable ...
    assert torch.equal(q.float().abs().view(-1, 64).amax(1), absmax)
    # ... and quantising Q(w) again changes nothing (idempotent storage)
    p2, a2 = N.quantize(q, ops)
    assert torch.equal(p2, packed) and torch.equal(a2, absmax)
    assert torch.equal(N.dequantize(p2, a2, w.shape, ops=ops), q)
    # high nibble = even element
    codes = O.nearest_code((O.bf16_to_f32(bits(w)).reshape(-1, 64) / absmax.numpy()[:, None]).astype(np.float32))
    assert int(packed[0]) >> 4 == int(codes[0, 0]) and int(packed[0]) & 15 == int(codes[0, 1])


def test_double_quantisation_matches_the_restatement():
    g = torch.Generator().manual_seed(2)
    absmax = (torch.rand(1000, generator=g) * 0.05 + 0.01)
    q, a2, off = N.double_quantize(absmax)
    rq, ra2, roff = O.double_quantize(absmax.numpy())
    assert np.array_equal(q.numpy(), rq) and np.allclose(a2.numpy(), ra2, rtol=0, atol=0) and float(off) == float(roff)
    back = N.double_dequantize(q, a2, off)
    assert np.array_equal(back.numpy(), O.double_dequantize(rq, ra2, roff))
    assert float((back - absmax).abs().max()) < 0.02 * float(absmax.max())          # 8-bit dynamic code: ~1 % steps
    book = N.dynamic_map()
    assert book.numel() == 256
    assert np.array_equal(book.numpy(), O.dynamic_map()) and float(book.max()) == 1.0 and bool((book[1:] >= book[:-1]).all())


def test_qlora_base_replaces_exactly_the_quantised_linears(ops):
    g = torch.Generator().manual_seed(3)
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.02).to(torch.bfloat16)
    sd = {"model.layers.0.self_attn.q_proj.weight": mk(64, 64), "model.layers.0.mlp.down_proj.weight": mk(64, 128),
          "lm_head.weight": mk(128, 64), "model.layers.0.input_layernorm.weight": mk(64),
          "model.mm_projector.0.weight": mk(64, 64), "model.embed_tokens.weight": mk(128, 64),
          "model.image_pooler.bert.encoder.layer.0.output.dense.weight": mk(64, 64)}
    before = {k: v.clone() for k, v in sd.items()}
    stored = N.qlora_base_(sd, double_quant=True, ops=ops)
    assert sorted(stored) == ["lm_head.weight", "model.layers.0.mlp.down_proj.weight",
                              "model.layers.0.self_attn.q_proj.weight"]
    for k in sd:
        if k in stored:
            assert not torch.equal(sd[k], before[k]) and sd[k].shape == before[k].shape
            assert torch.equal(stored[k].dequantize(), sd[k])
            assert float((sd[k].float() - before[k].float()).abs().max()) < 0.2 * float(before[k].float().abs().max())
            assert stored[k].nbytes() < 0.3 * before[k].numel() * 2
        else:
            assert torch.equal(sd[k], before[k])                                   # norms, projector, pooler, embeddings


def test_bad_arguments(ops):
    with pytest.raises(ValueError):
        N.quantize(torch.zeros(65, dtype=torch.bfloat16), ops)
    with pytest.raises(TypeError):
        N.quantize(torch.zeros(64), ops)
    assert ops.lib.b200_nf4_quantize(None, 64, None, None, None) == -2
