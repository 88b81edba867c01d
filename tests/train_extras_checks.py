"""Shared checks of the fine-tune kernels in csrc/train_extras.cu (seg-mask CNN forward-with-activations + backward,
embed_tokens gradient) against torch autograd over the oracle. Run on the CPU through the kernel emulator
(tests/test_train_extras_emu.py) and on the GPU through libb200mmor.so (tests/test_gpu_zz_train_extras.py)."""
import ctypes

import numpy as np
import torch

from mm_or_b200 import _lib as L
from oracle import mm2sg_oracle as O

SEG = O.POOL + "segmasks_encoder."
CH = (8, 64, 128, 256, 512, 1024)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def seg_weights(seed=3):
    """bf16-representable weights under the reference's names (fp32 tensors for the oracle)."""
    g = torch.Generator().manual_seed(seed)
    sd = {SEG + "embedding.weight": torch.randn(30, 8, generator=g)}
    for i in range(5):
        fan = CH[i] * 9
        sd[SEG + f"conv{i + 1}.weight"] = torch.randn(CH[i + 1], CH[i], 3, 3, generator=g) * (2.0 / fan) ** 0.5
        sd[SEG + f"conv{i + 1}.bias"] = torch.randn(CH[i + 1], generator=g) * 0.05
    return {k: v.to(torch.bfloat16).float() for k, v in sd.items()}


def bind(sd, dev):
    keep = {"emb": sd[SEG + "embedding.weight"].to(dev, torch.bfloat16).contiguous()}
    w = L.SegmaskWeights(emb=keep["emb"].data_ptr())
    for i in range(5):
        keep[f"w{i}"] = sd[SEG + f"conv{i + 1}.weight"].to(dev, torch.bfloat16).contiguous()
        keep[f"b{i}"] = sd[SEG + f"conv{i + 1}.bias"].to(dev, torch.bfloat16).contiguous()
        w.conv_w[i] = keep[f"w{i}"].data_ptr()
        w.conv_b[i] = keep[f"b{i}"].data_ptr()
    return w, keep


def check_segmask_forward_backward(be):
    """be: dict(cdll, ptr_fn, stream_fn, device). Tokens and every gradient vs torch autograd over the oracle's
    segmask_features; second call with accumulate doubles the gradients."""
    dev = be["device"]
    kw = dict(cdll=be["cdll"], ptr_fn=be["ptr_fn"], stream_fn=be["stream_fn"])
    sd = seg_weights()
    g = torch.Generator().manual_seed(4)
    n = 3
    cls = torch.randint(0, 30, (n, 32, 32), generator=g, dtype=torch.uint8)
    cls[0, 0, :5] = 200                                   # out-of-range ids clamp to class 29 like the inference kernel
    # token rows live in a (4, 3, 1024) buffer: map n -> row row_map[n]; one map has no consumer (row -1)
    row_map = torch.tensor([7, -1, 2], dtype=torch.int32)
    d_out = (torch.randn(12, 1024, generator=g) * 0.1).to(torch.bfloat16)
    # oracle
    with torch.enable_grad():                          # other tests of the session switch autograd off globally
        leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        feats = O.segmask_features(leaf, cls.clamp(max=29))
        loss = sum((feats[i] * d_out[r].float()).sum() for i, r in enumerate(row_map.tolist()) if r >= 0)
        loss.backward()
    # kernels
    w, keep = bind(sd, dev)
    out = torch.zeros(12, 1024, dtype=torch.bfloat16, device=dev)
    acts = L.segmask_forward_train(w, cls.to(dev), out, 1024, row_map.to(dev), **kw)
    for i, r in enumerate(row_map.tolist()):
        if r >= 0:
            assert rel(out[r], feats[i].detach()) < 4e-3                 # one bf16 rounding of an fp32 pipeline
    assert float(out[0].float().abs().sum()) == 0
    dW = [torch.zeros(CH[i + 1], CH[i], 3, 3, device=dev) for i in range(5)]
    dB = [torch.zeros(CH[i + 1], device=dev) for i in range(5)]
    dE = torch.zeros(30, 8, device=dev)
    for rep in range(2):
        L.segmask_backward(w, cls.to(dev), acts, d_out.to(dev), 1024, row_map.to(dev), dW, dB, dE, accumulate=rep > 0,
                           **kw)
        s = float(rep + 1)
        for i in range(5):
            assert rel(dW[i], s * leaf[SEG + f"conv{i + 1}.weight"].grad) < 1e-4, (rep, i)
            assert rel(dB[i], s * leaf[SEG + f"conv{i + 1}.bias"].grad) < 1e-4, (rep, i)
        assert rel(dE, s * leaf[SEG + "embedding.weight"].grad) < 1e-4


def check_embed_grad(be):
    dev = be["device"]
    kw = dict(cdll=be["cdll"], ptr_fn=be["ptr_fn"], stream_fn=be["stream_fn"])
    g = torch.Generator().manual_seed(6)
    V, D, R = 50, 96, 40
    d_rows = torch.randn(R, D + 8, generator=g).to(torch.bfloat16)[:, :D]       # row stride > D
    tok = torch.randint(0, V, (R,), generator=g)
    tok[::5] = -2                                                                # visual / pad rows carry no token
    tok[1] = tok[2] = tok[3] = 7                                                 # repeated token
    ref = torch.zeros(V, D)
    for r in range(R):
        if tok[r] >= 0:
            ref[tok[r]] += d_rows[r].float()
    table = torch.full((V, D), 5.0, device=dev)
    L.embed_grad(d_rows.to(dev), tok.numpy(), table, accumulate=True, **kw)
    assert rel(table.cpu() - 5.0, ref) < 1e-6
    touched = torch.zeros(V, dtype=torch.bool)
    touched[tok[tok >= 0]] = True
    table2 = torch.full((V, D), 9.0, device=dev)
    L.embed_grad(d_rows.to(dev), tok.numpy(), table2, accumulate=False, **kw)
    assert rel(table2.cpu()[touched], ref[touched]) < 1e-6
    assert bool((table2.cpu()[~touched] == 9.0).all())                          # rows without gradient are not written


ALL = [check_segmask_forward_backward, check_embed_grad]
