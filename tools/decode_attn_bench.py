"""Decode attention of one layer at the benchmark's geometry (128 rows x 32 heads, KV cache capacity 1088) for several
context lengths and split counts: GB/s of K / V actually read (512 B per key per head) against the measured
copy peak. 4096 CTAs on 148 x 8 resident slots is 3.46 waves: the sweep shows what the tail costs.
  python tools/decode_attn_bench.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mm_or_b200 import _lib as L


def main():
    torch.cuda.set_device(0)
    dev = "cuda"
    B, H, cap = 128, 32, 1088
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(B, H * 128, generator=g, device=dev).to(torch.bfloat16)
    kc = torch.randn(B, H, cap, 128, generator=g, device=dev).to(torch.bfloat16)
    vc = torch.randn(B, H, cap, 128, generator=g, device=dev).to(torch.bfloat16)
    # a second cache: alternate between the two so that no launch finds its K / V in the 126 MB L2
    kc2, vc2 = kc.clone(), vc.clone()
    n = 40
    for ctx in (831, 960, 1087):
        for splits in ((1,) if os.environ.get("SPLITS1") else (1, 2, 3, 4, 6, 8)):
            for _ in range(3):
                L.decode_attention(q, kc, vc, ctx, splits=splits)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                if i & 1:
                    L.decode_attention(q, kc2, vc2, ctx, splits=splits)
                else:
                    L.decode_attention(q, kc, vc, ctx, splits=splits)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            gb = B * H * ctx * 128 * 2 * 2 / 1e9
            print(json.dumps({"ctx": ctx, "splits": splits, "us_per_launch": round(us, 1),
                              "gb_per_s": round(gb / us * 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
