"""GPU timing of the attention kernels on the shapes of the MM2SG path (ViT, pooler, prefill, decode).
Run on a B200:  python tools/gpu_attn_check.py   (torch SDPA is timed beside it as a reference point; the round-1
mma.sync kernel this replaced is recorded in profiles/r1_attention_perf.log)"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mm_or_b200 import _lib as L


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    tag = "tcgen05"
    shapes = [("vit 96 img", 96, 577, 577, 16, 64, False), ("pooler l1 B=16", 16, 3456, 3456, 8, 128, False),
              ("pooler l2 B=16", 16, 576, 3456, 8, 128, False), ("prefill B=16", 16, 831, 831, 32, 128, True)]
    for name, B, Lq, Lk, H, d, causal in shapes:
        q = torch.randn(B, Lq, H, d, device="cuda", dtype=torch.bfloat16)
        k = torch.randn(B, Lk, H, d, device="cuda", dtype=torch.bfloat16)
        v = torch.randn(B, Lk, H, d, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: L.flash_attention(q, k, v, causal=causal))
        fl = 4.0 * B * H * Lq * Lk * d * (0.5 if causal else 1.0)
        ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                               is_causal=causal and Lq == Lk).transpose(1, 2)
        out = L.flash_attention(q, k, v, causal=causal)
        err = ((out.float() - ref.float()).norm() / ref.float().norm()).item()
        ms_ref = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(
            q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal and Lq == Lk))
        print(f"attn {tag:16s} {name:16s}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TFLOP/s  rel err vs torch sdpa {err:.4f}"
              f"   (torch sdpa {ms_ref*1e3:8.1f} us)", flush=True)
    # decode attention, cold KV (B=64/128, ctx ~ 960)
    for B in (64, 128):
        H, cap, ctx = 32, 1088, 960
        kc = torch.randn(B, H, cap, 128, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn(B, H, cap, 128, device="cuda", dtype=torch.bfloat16)
        q = torch.randn(B, H * 128, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: L.decode_attention(q, kc, vc, ctx))
        by = 2.0 * B * H * ctx * 128 * 2
        print(f"decode_attn B={B} ctx={ctx}: {ms*1e3:8.1f} us  {by/ms/1e6:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
