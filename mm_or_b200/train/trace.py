"""Phase timing of the fine-tune step for diagnosis (tools/train_bench.py --trace): `with phase("name"):` synchronises
the device on both sides and adds the wall time to a table -- only while tracing is enabled; otherwise it is a no-op
(a traced step is serialised and therefore slower than a real one: read the shares, not the sum)."""
import contextlib
import time

import torch

_table = None


def enable(on=True):
    global _table
    _table = {} if on else None


def collect():
    return {k: round(v, 2) for k, v in (_table or {}).items()}


@contextlib.contextmanager
def phase(name):
    if _table is None:
        yield
        return
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        yield
    finally:
        torch.cuda.synchronize()
        _table[name] = _table.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
