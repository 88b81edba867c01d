"""The kernel emulator (tests/emu/cuda_emu.h) checked on its own: barrier-ordered shared memory, warp shuffles,
multi-dimensional launch indexing. If these fail, no emulated-kernel test means anything."""
import ctypes

import torch

import emu_lib


def p(t):
    return ctypes.c_void_p(t.data_ptr())


def test_block_scan_needs_correct_barriers():
    lib = emu_lib.lib()
    g = torch.Generator().manual_seed(0)
    n = 700
    x = torch.randint(-50, 50, (n,), generator=g, dtype=torch.int32)
    out = torch.zeros(n, dtype=torch.int32)
    assert lib.b200_emu_selftest_scan(p(x), p(out), n) == 0
    ref = torch.cat([x[i:i + 256].cumsum(0) for i in range(0, n, 256)]).to(torch.int32)
    assert torch.equal(out, ref)


def test_warp_shuffles():
    lib = emu_lib.lib()
    n = 128
    x = torch.arange(n, dtype=torch.float32) * 0.5 + 1
    s, l3, dn = torch.zeros(n), torch.zeros(n), torch.zeros(n)
    assert lib.b200_emu_selftest_shuffle(p(x), p(s), p(l3), p(dn), n) == 0
    w = x.view(-1, 32)
    assert torch.equal(s.view(-1, 32), w.sum(1, keepdim=True).expand(-1, 32))
    assert torch.equal(l3.view(-1, 32), w[:, 3:4].expand(-1, 32))
    want = torch.cat([w[:, 5:], w[:, 27:]], dim=1)              # lanes past the end keep their own value
    assert torch.equal(dn.view(-1, 32), want)


def test_grid_and_block_indexing_with_early_exit():
    lib = emu_lib.lib()
    nx, ny, nz = 13, 6, 3
    out = torch.full((nz, ny, nx), -1, dtype=torch.int32)
    assert lib.b200_emu_selftest_index(p(out), nx, ny, nz) == 0
    z, y, x = torch.meshgrid(torch.arange(nz), torch.arange(ny), torch.arange(nx), indexing="ij")
    ref = torch.where(y % 2 == 0, 1000000 * z + 1000 * y + x, torch.full_like(x, -1)).to(torch.int32)
    assert torch.equal(out, ref)
