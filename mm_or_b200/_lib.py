"""ctypes binding of libb200mmor.so (C ABI declared in include/b200_mmor.h).

The product path has no CPU or eager-PyTorch fallback: if the shared library is missing or a call fails,
this module raises. torch is used only to own device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200mmor.so")

_lib = None


class B200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the CDLL. Raises B200Error when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} not found: build it with `python -m mm_or_b200.build` (no CPU fallback exists)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.b200_last_error.restype = ctypes.c_char_p
        _declare(_lib)
    return _lib


def _declare(L):
    vp, ci, cf, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
    L.b200_abi_version.restype = ci
    sigs = {
        "b200_gemm_bf16": [vp, ci, vp, ci, vp, ci, ci, ci, ci, vp, vp, ci, vp, ci, ci, ci, vp],
    }
    for name, argtypes in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = ci
    return cf, cl


def check(rc, what=""):
    if rc != 0:
        msg = lib().b200_last_error()
        raise B200Error(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "libb200mmor operates on CUDA tensors only"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


ACT_NONE, ACT_QUICK_GELU, ACT_GELU, ACT_SWIGLU = 0, 1, 2, 3


def gemm(a, w, out=None, bias=None, residual=None, row_map=None, act=ACT_NONE, out_fp32=False, bn=0,
         out_rows=None):
    """out = epilogue(a @ w.T). a: (M,K) bf16 (row stride allowed), w: (N,K) bf16."""
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        rows = M if out_rows is None else out_rows
        out = torch.empty((rows, n_out), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    ldr = residual.stride(0) if residual is not None else 0
    rc = lib().b200_gemm_bf16(ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K,
                              ptr(bias), ptr(residual), ldr, ptr(row_map), act, int(out_fp32), bn, stream_ptr())
    check(rc, "b200_gemm_bf16")
    return out
