"""TEST INFRASTRUCTURE: builds and loads tests/emu/_build/libb200emu.so -- mm_or_b200/csrc/{ptv3,train_extras,nf4}.cu compiled with g++
-DB200_EMU against the CUDA-kernel emulator of tests/emu/cuda_emu.h -- and binds it like mm_or_b200/_lib.py binds the real
library, so that the CPU test-suite executes the SAME kernel source (and the same host orchestration,
mm_or_b200/model/point_transformer.py) against the oracle. Never imported by the product."""
import ctypes
import os
import subprocess

import torch

from mm_or_b200 import _lib as L
from mm_or_b200.model.point_transformer import PcOps

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
CSRC = os.path.join(HERE, "..", "mm_or_b200", "csrc")
SRCS = [os.path.join(CSRC, "ptv3.cu"), os.path.join(CSRC, "train_extras.cu"), os.path.join(CSRC, "nf4.cu")]
_lib = None


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


# fmaf() without -mfma is a libm call per multiply-add: build with the host's FMA instructions when it has them (the
# library name carries the choice, so a build made on another machine is never picked up by mistake)
FMA = _cpu_has_fma()
OUT = os.path.join(EMU, "_build", "libb200emu_fma.so" if FMA else "libb200emu.so")


def build():
    deps = SRCS + [os.path.join(EMU, "cuda_emu.h"), os.path.join(EMU, "emu_common.h"),
                   os.path.join(EMU, "emu_support.cpp"), os.path.join(EMU, "selftest.cpp"),
                   os.path.join(HERE, "..", "include", "b200_mmor.h")]
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2"] + (["-mfma"] if FMA else []) + ["-std=c++17", "-fPIC", "-shared", "-DB200_EMU", "-I" + EMU]
    for src in SRCS + [os.path.join(EMU, "emu_support.cpp"), os.path.join(EMU, "selftest.cpp")]:
        cmd += ["-x", "c++", src]
    tmp = f"{OUT}.{os.getpid()}.tmp"                     # atomic: parallel test workers may build at the same time
    cmd += ["-o", tmp]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + r.stderr)
    os.replace(tmp, OUT)
    return OUT


def lib():
    global _lib
    if _lib is None:
        cdll = ctypes.CDLL(build())
        cdll.b200_last_error.restype = ctypes.c_char_p
        for name in L.EMULATABLE_SYMBOLS:
            fn = getattr(cdll, name)
            fn.restype, fn.argtypes = L._SIGS[name]
        assert cdll.b200_emu_marker() == 1
        _lib = cdll
    return _lib


def ops():
    """PcOps over the emulator: host tensors, host pointers."""
    cdll = lib()

    def ptr(t):
        if t is None:
            return None
        assert not t.is_cuda
        return ctypes.c_void_p(t.data_ptr())

    def check(rc, what=""):
        if rc != 0:
            raise RuntimeError(f"{what} failed with code {rc}: {cdll.b200_last_error().decode()}")

    return PcOps(cdll, torch.device("cpu"), ptr, lambda: None, check)
