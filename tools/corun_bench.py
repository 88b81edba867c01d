"""Can the HBM-bound decode attention and a tensor-bound prefill GEMM share the SMs? Micro-benchmark on ONE B200:
decode attention of one layer (128 rows x 32 heads, ctx 900: 0.94 GB of K / V per launch) in a loop on a high-priority
stream, the prefill gate_up GEMM (13296 x 22016 x 4096, SwiGLU epilogue) in a loop on a low-priority stream; each alone,
then both at once. together ~ max(alone) => they co-run; together ~ sum(alone) => they time-slice.
  python tools/corun_bench.py            (B200_GEMM_CARVEOUT=1: the GEMM asks for the full 228 KB shared carve-out so that
                                          attention CTAs fit beside it. The attention-side knob B200_ATTN_CARVEOUT that
                                          profiles/r2_corun_attention_vs_gemm.json also shows was removed after the
                                          experiment: it made the attention alone slower and changed nothing together.)
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
    B, H, cap, ctx = 128, 32, 1088, 900
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(B, H * 128, generator=g, device=dev).to(torch.bfloat16)
    kc = torch.randn(B, H, cap, 128, generator=g, device=dev).to(torch.bfloat16)
    vc = torch.randn(B, H, cap, 128, generator=g, device=dev).to(torch.bfloat16)
    M, N, K = 13296, 22016, 4096
    a = (torch.randn(M, K, generator=g, device=dev) * 0.1).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
    out = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
    hi, lo = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)
    n_attn, n_gemm = 120, 16

    def attn_loop():
        with torch.cuda.stream(hi):
            for _ in range(n_attn):
                L.decode_attention(q, kc, vc, ctx, splits=1)

    def gemm_loop(m=M):
        with torch.cuda.stream(lo):
            for _ in range(n_gemm):
                L.gemm(a[:m], w, out=out[:m], act=L.ACT_SWIGLU)

    def timed(fns):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        ends = []
        e0.record()
        hi.wait_event(e0)
        lo.wait_event(e0)
        for fn, st in fns:
            fn()
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            ends.append(e)
        torch.cuda.synchronize()
        return [round(e0.elapsed_time(e), 2) for e in ends]

    attn_loop(); gemm_loop(); torch.cuda.synchronize()                     # warm-up (kernel attributes, tensor maps)
    rec = {"gemm_carveout_env": os.environ.get("B200_GEMM_CARVEOUT", "0"),
           "attn_alone_ms": timed([(attn_loop, hi)])[0], "gemm_alone_ms": timed([(gemm_loop, lo)])[0]}
    t = timed([(gemm_loop, lo), (attn_loop, hi)])
    rec["together_gemm_first"] = {"gemm_done_ms": t[0], "attn_done_ms": t[1]}
    t = timed([(attn_loop, hi), (gemm_loop, lo)])
    rec["together_attn_first"] = {"attn_done_ms": t[0], "gemm_done_ms": t[1]}
    rec["attn_gb_per_launch"] = round(B * H * ctx * 128 * 2 * 2 / 1e9, 3)
    rec["attn_alone_gbs"] = round(rec["attn_gb_per_launch"] * n_attn / rec["attn_alone_ms"] * 1e3, 1)
    rec["gemm_alone_tflops"] = round(2.0 * M * N * K * n_gemm / rec["gemm_alone_ms"] / 1e9, 1)
    # short GEMMs (one prefill sample per launch: 831 rows) as the co-runner
    n_small = n_gemm * 16
    def gemm_small():
        with torch.cuda.stream(lo):
            for _ in range(n_small):
                L.gemm(a[:831], w, out=out[:831], act=L.ACT_SWIGLU)
    gemm_small(); torch.cuda.synchronize()
    rec["gemm831_alone_ms"] = timed([(gemm_small, lo)])[0]
    t = timed([(gemm_small, lo), (attn_loop, hi)])
    rec["together_gemm831_first"] = {"gemm_done_ms": t[0], "attn_done_ms": t[1]}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
