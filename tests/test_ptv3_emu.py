"""CPU: the point-cloud KERNELS (mm_or_b200/csrc/ptv3.cu, compiled for the host against the kernel emulator of
tests/emu/) and their host orchestration (mm_or_b200/model/point_transformer.py) against the oracle and the fixtures
recorded from the reference (checks shared with the GPU suite: tests/ptv3_checks.py)."""
import pytest

import emu_lib
import ptv3_checks as C


@pytest.fixture(scope="module")
def ops():
    return emu_lib.ops()


@pytest.mark.parametrize("check", C.ALL, ids=lambda f: f.__name__[6:])
def test_emulated_kernels(ops, check):
    check(ops)


@pytest.mark.parametrize("order", range(4))
def test_codes_match_reference_fixture(ops, order):
    C.check_codes_match_reference_fixture(ops, order)


def test_product_never_uses_the_emulator():
    """The emulator library lives under tests/ and the product binding only ever opens libb200mmor.so."""
    import inspect
    import re
    from mm_or_b200 import _lib as L
    from mm_or_b200.model import point_transformer as PT
    src = inspect.getsource(L)
    assert "libb200emu" not in src and "emu_lib" not in src
    assert re.findall(r"CDLL\((\w+)\)", src) == ["LIB_PATH"]
    assert "tests" not in inspect.getsource(PT.PcOps.cuda)
    assert L.LIB_PATH.endswith("libb200mmor.so")
