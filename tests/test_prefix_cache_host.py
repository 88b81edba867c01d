"""CPU: host logic of the prefix-KV reuse (mm_or_b200/model/llava_llama.py::PrefixCache) -- when a batch qualifies,
where the prefill may start, and which cache slots of which row receive which prefix tokens. The device side is
tests/test_gpu_zzz_serving.py::test_prefix_kv_reuse_equals_full_prefill."""
import numpy as np
import torch

from mm_or_b200.model.llava_llama import PrefixCache


class _Cache:                                   # the two tensors PrefixCache.copy_into writes
    def __init__(self, layers, batch, heads, cap):
        self.k = torch.zeros(layers, batch, heads, cap, 128)
        self.v = torch.zeros(layers, batch, heads, cap, 128)


def _pc(p=5, layers=2, heads=3):
    ids = torch.arange(10, 10 + p)
    k = torch.arange(layers * heads * p * 128, dtype=torch.float32).view(layers, heads, p, 128)
    return PrefixCache(ids, k, -k, pad_token_id=0)


def test_plan_needs_every_row_to_start_with_the_prefix():
    pc = _pc()
    ids = torch.tensor([[0, 0, 10, 11, 12, 13, 14, -200, 7, 8],
                        [10, 11, 12, 13, 14, -200, 5, 6, 7, 8]])
    mask = ids.ne(0)
    # packed rows: row 0 is 2 shorter; kv_start in PACKED space comes from the pack plan (here 576 visual tokens each)
    assert pc.plan(ids, mask, np.array([2, 0], np.int32), Lq=585) == 5          # min(kv_start) + p
    assert pc.plan(ids, mask, np.array([2, 0], np.int32), Lq=5) == 0            # nothing would be left to compute
    other = ids.clone()
    other[0, 3] = 99                                                               # row 0 deviates inside the prefix
    assert pc.plan(other, other.ne(0), np.array([2, 0], np.int32), Lq=585) == 0
    short = torch.tensor([[10, 11, 12, 13, 14]])                                  # only the prefix: no new token
    assert pc.plan(short, short.ne(0), np.array([0], np.int32), Lq=5) == 0


def test_copy_into_places_the_prefix_behind_each_rows_padding():
    pc = _pc(p=5)
    cache = _Cache(2, 3, 3, 16)
    kv_start = np.array([0, 2, 9], np.int32)           # row 2 starts beyond q0: nothing of it is copied
    q0 = 0 + 5
    pc.copy_into(cache, kv_start, q0)
    assert torch.equal(cache.k[:, 0, :, 0:5], pc.k) and torch.equal(cache.v[:, 0, :, 0:5], pc.v)
    assert torch.equal(cache.k[:, 1, :, 2:5], pc.k[:, :, :3])      # first q0 - kv_start = 3 prefix tokens; the other two
    assert float(cache.k[:, 1, :, 5:].abs().sum()) == 0            # are recomputed inside the prefill block
    assert float(cache.k[:, 1, :, :2].abs().sum()) == 0
    assert float(cache.k[:, 2].abs().sum()) == 0
