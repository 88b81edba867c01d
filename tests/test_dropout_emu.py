"""CPU: the dropout kernel (mm_or_b200/csrc/train_extras.cu::dropout_kernel, executed through the kernel emulator)
against a numpy restatement of the same counter-based generator (Philox4x32-10, Salmon et al. 2011; known-answer vector
of the Random123 distribution), bit for bit, and the properties training relies on: the mask is a pure function of
(seed, index) -- forward and backward agree without storing it --, the keep rate is 1 - p, kept values are scaled by
1 / (1 - p)."""
import ctypes

import numpy as np
import pytest
import torch

import emu_lib
from mm_or_b200 import _lib as L


def philox4x32_10(c0, c1, k0, k1):
    """numpy restatement: counters (c0, c1, 0, 0) as uint32 arrays, key (k0, k1) scalars -> 4 uint32 arrays."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c = [np.asarray(c0, dtype=np.uint32), np.asarray(c1, dtype=np.uint32), np.zeros_like(c0, dtype=np.uint32),
         np.zeros_like(c0, dtype=np.uint32)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ np.uint32(k0)
        n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ np.uint32(k1)
        c = [n0, (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32), n2, (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def keep_mask(n, p, seed):
    g = np.arange((n + 3) // 4, dtype=np.uint64)
    words = philox4x32_10((g & np.uint64(0xFFFFFFFF)).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                          seed & 0xFFFFFFFF, seed >> 32)
    r = np.stack(words, 1).reshape(-1)[:n]
    return r >= np.uint32(int(p * 4294967296.0))


def run(x, p, seed, out=None, accumulate=False):
    cdll = emu_lib.lib()
    if out is None:
        out = torch.empty_like(x)
    fn = cdll.b200_dropout
    fn.restype, fn.argtypes = L._SIGS["b200_dropout"]
    rc = fn(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), x.numel(), p, seed, int(accumulate), None)
    assert rc == 0, cdll.b200_last_error()
    return out


def test_philox_known_answer():
    """Random123 kat_vectors: philox4x32-10 of counter 0 / key 0 and of the all-ones counter / key."""
    z = np.zeros(1, dtype=np.uint32)
    assert [int(w[0]) for w in philox4x32_10(z, z, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


def test_dropout_matches_the_restatement_and_its_own_backward():
    g = torch.Generator().manual_seed(0)
    for n, p, seed in ((4096, 0.05, 1234567890123), (1003, 0.5, 7), (64, 0.0, 99), (5, 0.9, (1 << 63) + 5)):
        x = torch.randn(n, generator=g).to(torch.bfloat16)
        y = run(x, p, seed)
        keep = torch.from_numpy(keep_mask(n, p, seed))
        want = torch.where(keep, (x.float() * np.float32(1.0 / np.float32(1.0 - np.float32(p)))), torch.zeros(n))
        assert torch.equal(y, want.to(torch.bfloat16)), (n, p)
        assert torch.equal(run(x, p, seed), y)                                  # deterministic
        xin = x.clone()
        assert torch.equal(run(xin, p, seed, out=xin), y)                       # in place (out = x, train/encoder.py)
        if n > 8:                                                               # ... and through an unaligned view
            xo = x.clone()
            assert torch.equal(run(xo[1:], p, seed, out=xo[1:]), run(x[1:].clone(), p, seed))
        if 0 < p < 1 and n > 100:
            assert not torch.equal(run(x, p, seed + 1), y)                      # another seed, another mask
        # backward = the same kernel on the gradient, accumulating into dx: the SAME elements survive
        gy = torch.randn(n, generator=g).to(torch.bfloat16)
        dx0 = torch.randn(n, generator=g).to(torch.bfloat16)
        dx = run(gy, p, seed, out=dx0.clone(), accumulate=True)
        add = torch.where(keep, gy.float() * np.float32(1.0 / np.float32(1.0 - np.float32(p))), torch.zeros(n))
        assert torch.equal(dx, (dx0.float() + add).to(torch.bfloat16))
        assert torch.equal((y != 0) | (x == 0), keep | (x == 0))


def test_keep_rate_and_scale():
    n, p = 1 << 16, 0.05
    x = torch.ones(n, dtype=torch.bfloat16)
    y = run(x, p, 42).float()
    kept = (y != 0).float().mean().item()
    assert abs(kept - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5                   # 4 sigma of a binomial
    assert torch.equal(y[y != 0].unique(), torch.tensor([1.0 / (1 - p)]).to(torch.bfloat16).float())
    assert torch.equal(run(x, 0.0, 42), x)                                       # p = 0: identity


def test_bad_arguments():
    cdll = emu_lib.lib()
    fn = cdll.b200_dropout
    fn.restype, fn.argtypes = L._SIGS["b200_dropout"]
    x = torch.zeros(8, dtype=torch.bfloat16)
    for p in (1.0, -0.1, float("nan")):
        assert fn(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(x.data_ptr()), 8, p, 0, 0, None) == -2
    assert fn(None, None, 8, 0.1, 0, 0, None) == -2
