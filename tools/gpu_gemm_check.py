"""GPU bring-up check for the tcgen05 GEMM: correctness across shapes/epilogues, then throughput.
Run on a B200:  python tools/gpu_gemm_check.py [basic|epi|perf]
Writes progress lines to stdout (flush) so a hang still leaves a trail.
"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mm_or_b200 import _lib as L


def log(*a):
    print(*a, flush=True)


def ref_gemm(a, w, bias=None, residual=None, act=0):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    elif act == 3:
        g, u = y[:, 0::2], y[:, 1::2]
        y = torch.nn.functional.silu(g) * u
    if residual is not None:
        y = y + residual.float()
    return y


def check(name, out, ref, tol=2e-2):
    out = out.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    rel = err / scale
    bad = not math.isfinite(err) or rel > tol
    log(f"{'FAIL' if bad else 'ok  '} {name}: max_abs_err={err:.4g} ref_max={scale:.4g} rel={rel:.3g}")
    return not bad


def basic():
    torch.manual_seed(0)
    ok = True
    shapes = [(128, 256, 64), (128, 256, 256), (256, 512, 1024), (384, 1024, 4096), (577 * 3, 3072, 1024),
              (100, 264, 72), (1000, 4096, 11008), (129, 32000, 512)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K)
        ref = ref_gemm(a, w)
        for bn in (0, 256, 128, 64, 32):
            out = L.gemm(a, w, bn=bn)
            torch.cuda.synchronize()
            ok &= check(f"gemm M={M} N={N} K={K} bn={bn}", out, ref, 1e-2)
    return ok


def epi():
    torch.manual_seed(1)
    ok = True
    M, N, K = 700, 1024, 512
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K)
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    for bn in (256, 64):
        ok &= check(f"bias bn={bn}", L.gemm(a, w, bias=bias, bn=bn), ref_gemm(a, w, bias))
        ok &= check(f"bias+quickgelu bn={bn}", L.gemm(a, w, bias=bias, act=1, bn=bn), ref_gemm(a, w, bias, act=1))
        ok &= check(f"bias+gelu bn={bn}", L.gemm(a, w, bias=bias, act=2, bn=bn), ref_gemm(a, w, bias, act=2))
        ok &= check(f"bias+residual bn={bn}", L.gemm(a, w, bias=bias, residual=res, bn=bn),
                    ref_gemm(a, w, bias, residual=res))
        ok &= check(f"swiglu bn={bn}", L.gemm(a, w, act=3, bn=bn), ref_gemm(a, w, act=3))
        ok &= check(f"fp32 out bn={bn}", L.gemm(a, w, bias=bias, out_fp32=True, bn=bn), ref_gemm(a, w, bias), 5e-3)
        # in-place residual
        x = res.clone()
        L.gemm(a, w, out=x, bias=bias, residual=x, bn=bn)
        ok &= check(f"inplace residual bn={bn}", x, ref_gemm(a, w, bias, residual=res))
        # row scatter
        perm = torch.randperm(M, device="cuda").int()
        perm[::7] = -1
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        L.gemm(a, w, out=out, bias=bias, row_map=perm, bn=bn)
        ref = torch.zeros(M, N, device="cuda")
        keep = perm >= 0
        ref[perm[keep].long()] = ref_gemm(a, w, bias)[keep]
        ok &= check(f"row_map bn={bn}", out, ref)
    # strided A (lda > K) and strided output
    big = torch.randn(M, 3 * K, device="cuda", dtype=torch.bfloat16)
    a2 = big[:, K:2 * K]
    outbig = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a2, w, out=outbig[:, N:], bn=128)
    ok &= check("strided A / C", outbig[:, N:], ref_gemm(a2, w))
    ok &= check("strided C untouched half", outbig[:, :N], torch.zeros(M, N, device="cuda"))
    return ok


def perf():
    shapes = [
        ("vit qkv 64img", 64 * 577, 3072, 1024),
        ("vit fc1", 64 * 577, 4096, 1024),
        ("vit fc2", 64 * 577, 1024, 4096),
        ("llama qkv", 16 * 831, 12288, 4096),
        ("llama gateup", 16 * 831, 22016, 4096),
        ("llama down", 16 * 831, 4096, 11008),
        ("square 8192", 8192, 8192, 8192),
        ("decode qkv B=128", 128, 12288, 4096),
        ("decode o B=128", 128, 4096, 4096),
        ("decode gateup B=128", 128, 22016, 4096),
        ("decode down B=128", 128, 4096, 11008),
        ("decode lm_head B=128", 128, 32000, 4096),
    ]
    for name, M, N, K in shapes:
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        bns = (0, 256, 128) if M > 256 else (0, 128, 64, 32)
        for bn in bns:
            for _ in range(3):
                L.gemm(a, w, out=out, bn=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                L.gemm(a, w, out=out, bn=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            tf = 2.0 * M * N * K / ms / 1e9
            gbs = (M * K + N * K + M * N) * 2 / ms / 1e6
            log(f"perf {name:24s} M={M} N={N} K={K} bn={bn}: {ms*1e3:9.1f} us  {tf:8.1f} TFLOP/s  {gbs:8.1f} GB/s")
        # cuBLAS reference point
        for _ in range(3):
            torch.matmul(a, w.t(), out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.matmul(a, w.t(), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        log(f"perf {name:24s} cuBLAS(torch.matmul): {ms*1e3:9.1f} us  {2.0*M*N*K/ms/1e9:8.1f} TFLOP/s")


def skinny(ms=(64, 128)):
    """Decode-step GEMM shapes with COLD weights: every launch reads a different copy out of a pool larger than L2
    (in the decode step each weight matrix is read once per 13 GB pass)."""
    shapes = [("qkv", 12288, 4096), ("o", 4096, 4096), ("gateup", 22016, 4096), ("down", 4096, 11008),
              ("lm_head", 32000, 4096)]
    for M in ms:
        tot_best = 0.0
        for name, N, K in shapes:
            copies = max(2, int(300e6 // (N * K * 2)) + 1)
            ws_ = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K) for _ in range(copies)]
            a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            variants = [("tiled bn=0", lambda w: L.gemm(a, w, out=out)),
                        ("tiled bn=128", lambda w: L.gemm(a, w, out=out, bn=128))]
            for sp in (0, 1, 2, 4, 8):
                variants.append((f"skinny splits={sp}", lambda w, sp=sp: L.gemm_skinny(a, w, out=out, splits=sp)))
            variants.append(("cuBLAS", lambda w: torch.matmul(a, w.t(), out=out)))
            best = 1e9
            for vn, fn in variants:
                for i in range(copies):
                    fn(ws_[i])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 3 * copies
                e0.record()
                for i in range(iters):
                    fn(ws_[i % copies])
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                gbs = (M * K + N * K + M * N) * 2 / ms / 1e6
                if vn.startswith("skinny"):
                    best = min(best, ms)
                log(f"skinny M={M:4d} {name:8s} N={N} K={K} {vn:18s}: {ms*1e3:8.1f} us  {gbs:7.1f} GB/s")
            tot_best += best * (32 if name != "lm_head" else 1)
            del ws_
        log(f"skinny M={M}: best-variant sum over 32 layers + lm_head = {tot_best:.2f} ms/step "
            f"(13.2 GB at 6.45 TB/s = 2.05 ms)")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "basic"
    t0 = time.time()
    log(f"[{which}] device: {torch.cuda.get_device_name(0)}")
    if which == "basic":
        r = basic()
    elif which == "epi":
        r = epi()
    elif which == "skinny":
        skinny()
        r = True
    elif which == "skinny64":
        skinny((64,))
        r = True
    else:
        perf()
        r = True
    log(f"[{which}] done in {time.time()-t0:.1f}s -> {'PASS' if r else 'FAIL'}")
    sys.exit(0 if r else 1)
