"""CPU: the C-ABI library builds/loads and exports every symbol include/b200_mmor.h declares (no compute calls)."""
import ctypes
import os
import re

from mm_or_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200_mmor.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        from mm_or_b200.build import build
        build()
    lib = ctypes.CDLL(L.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200_mmor.h but not exported"


def test_binding_covers_header():
    assert sorted(L.EXPORTED_SYMBOLS) == header_symbols()
    lib = L.lib()
    assert lib.b200_abi_version() == 1


def test_struct_sizes_match_header_layout():
    lib = L.lib()
    mirrors = [L.VitLayer, L.VitWeights, L.BertLayer, L.PoolerWeights, L.SegmaskWeights, L.ProjectorWeights,
               L.LlamaLayer, L.LlamaWeights, L.KvCache]
    for i, m in enumerate(mirrors):
        assert lib.b200_sizeof_struct(i) == ctypes.sizeof(m), m.__name__
    assert lib.b200_sizeof_struct(99) == 0


def test_no_cpu_fallback():
    import pytest
    import torch
    with pytest.raises(L.B200Error):
        L.ptr(torch.zeros(4))


def test_decode_step_switches_defaults_and_reject_unknown_names():
    """b200_set_option / b200_get_option (host-side state only; the switches themselves are exercised on the GPU by
    tests/test_gpu_zzzz_switches.py). Defaults as timed on hardware: pdl off, decode_tiles 1."""
    import pytest
    defaults = {"pdl": 0, "decode_tiles": 1, "fused_rope": 1}
    for name in ("pdl", "decode_tiles", "fused_rope"):
        if os.environ.get("B200_" + name.upper()) is None:
            assert int(L.get_option(name)) == defaults[name]
        L.set_option(name, True)
        assert L.get_option(name) == 1
        L.set_option(name, False)
        assert not L.get_option(name)
    L.set_option("decode_tiles", 2)                    # two CTAs per SM
    assert L.get_option("decode_tiles") == 2
    with pytest.raises(L.B200Error):
        L.set_option("decode_tiles", 3)
    L.set_option("decode_tiles", defaults["decode_tiles"])
    L.set_option("fused_rope", defaults["fused_rope"])
    with pytest.raises(L.B200Error):
        L.set_option("no_such_switch", 1)
    with pytest.raises(L.B200Error):
        L.get_option("no_such_switch")


def test_decode_tile_width_policy():
    """stages.cu::decode_bn minimises (waves of weight tiles over the CTA slots) x (tile width). Llama-7B on 148 SMs, one
    CTA per SM: qkv 12288 -> 96 (128 tiles), gate_up 22016 -> 160 (138), lm_head 32000 -> 224 (143); two CTAs per SM
    (296 slots, widths 64 / 96 / 128): 64 (192 tiles), 96 (230), 128 (250): one wave each."""
    w = L.lib().b200_decode_tile_width
    assert [w(64, n, 148, 1) for n in (12288, 22016, 32000)] == [96, 160, 224]
    assert [w(128, n, 148, 1) for n in (12288, 22016, 32000)] == [96, 160, 224]
    assert [w(128, n, 148, 2) for n in (12288, 22016, 32000)] == [64, 96, 128]
    for per_sm, widths in ((1, (96, 128, 160, 224, 256)), (2, (64, 96, 128))):
        for rows in (1, 64, 128, 192, 256):
            for n in (4096, 12288, 22016, 32000, 512):
                bn = w(rows, n, 148, per_sm)
                assert bn in widths
                m_tiles = (rows + 127) // 128
                cost = lambda b: -(-(m_tiles * -(-n // b)) // (148 * per_sm)) * b
                assert cost(bn) == min(cost(b) for b in widths)
    assert w(64, 12288, 0, 1) == 128                   # no device: the default width
