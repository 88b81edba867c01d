"""CPU: the fine-tune kernels of csrc/train_extras.cu executed through the kernel emulator (tests/emu/) against torch
autograd over the oracle (checks shared with the GPU suite: tests/train_extras_checks.py)."""
import ctypes

import pytest
import torch

import emu_lib
import train_extras_checks as C


@pytest.fixture(scope="module")
def backend():
    cdll = emu_lib.lib()
    return dict(cdll=cdll, device=torch.device("cpu"), stream_fn=lambda: None,
                ptr_fn=lambda t: None if t is None else ctypes.c_void_p(t.data_ptr()))


@pytest.mark.parametrize("check", C.ALL, ids=lambda f: f.__name__[6:])
def test_emulated_train_extras(backend, check):
    check(backend)
