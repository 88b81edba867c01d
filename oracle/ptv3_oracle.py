"""CPU oracle for the point-cloud branch of MM2SG's image pooler -- TEST INFRASTRUCTURE ONLY.

Restates, in plain vectorised PyTorch (CPU, fp32), what `ImageEmbeddingPooler._encode_pc`
(multimodal_projector/builder.py:93-148) computes: PointTransformerV3 in cls_mode (multimodal_projector/
pointtransformerv3.py:787-1005, default geometry) over the batched clouds -> per-cloud mean of the 512-d features ->
`project_pc` Linear(512 -> 1024), zero feature rows for missing clouds. Only tests/ may import this module.

Third-party arithmetic that is NOT under /root/reference (restated from the published semantics, call sites cited):
  spconv-cu117 2.x `SubMConv3d` (pointtransformerv3.py:549,768): submanifold cross-correlation at the active voxels,
      centred kernel, weight (out, k, k, k, in); torch_scatter `segment_csr` (pointtransformerv3.py:686-690);
  flash-attn `flash_attn_varlen_qkvpacked_func` (pointtransformerv3.py:479): fp16 inputs, fp32 softmax, fp16 output.
Pinning: tests/golden/make_ptv3_golden.py runs the reference's OWN `_encode_pc` / PointTransformerV3 / serialization
code through oracle/ref_shim.py (naive stand-ins for the three absent packages) and records tests/golden/ptv3_*.pt;
tests/test_ptv3_oracle.py checks this file against them, and the serialization codes (integer work) bit-exactly
against the reference's serialization/{z_order,hilbert}.py outputs recorded in the same fixture. The stand-ins for
spconv / torch_scatter / flash-attn are themselves unpinned (the packages are absent): that part of the parity claim
is "restated from published semantics".

Paths are relative to /root/reference/scene_graph_generation/LLaVA/llava/model/.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

PT = "model.image_pooler.point_transformer."
ORDERS = ("z", "z-trans", "hilbert", "hilbert-trans")  # pointtransformerv3.py:791


@dataclass
class Ptv3Cfg:  # pointtransformerv3.py:788-823 (defaults) with cls_mode=True, project_pc_dim=1024 (builder.py:82-85)
    in_channels: int = 6
    stride: Sequence[int] = (2, 2, 2, 2)
    enc_depths: Sequence[int] = (2, 2, 2, 6, 2)
    enc_channels: Sequence[int] = (32, 64, 128, 256, 512)
    enc_num_head: Sequence[int] = (2, 4, 8, 16, 32)
    patch_size: int = 1024
    mlp_ratio: int = 4
    project_pc_dim: int = 1024
    grid_size: float = 0.01  # builder.py:126
    bn_eps: float = 1e-3     # pointtransformerv3.py:861
    ln_eps: float = 1e-5     # nn.LayerNorm default


# ---------------------------------------------------------------------------------------------------------------------
# serialization (integer work: bit-exact)
# ---------------------------------------------------------------------------------------------------------------------
def z_order_encode(grid: torch.Tensor, depth: int) -> torch.Tensor:
    """serialization/z_order.py:13-61 (LUT form there; the bit loop of KeyLUT.xyz2key gives the same key): bit i of x /
    y / z goes to bit 3i+2 / 3i+1 / 3i."""
    x, y, z = (grid[:, i].long() for i in range(3))
    key = torch.zeros_like(x)
    for i in range(depth):
        m = 1 << i
        key |= ((x & m) << (2 * i + 2)) | ((y & m) << (2 * i + 1)) | ((z & m) << (2 * i))
    return key


def hilbert_encode(grid: torch.Tensor, depth: int) -> torch.Tensor:
    """serialization/hilbert.py:91-193 (Skilling's transform on bit planes there; the same steps on packed integers
    here): for every bit from the top, for every dimension: if the bit is set invert the lower bits of dimension 0, else
    exchange the differing lower bits of dimension 0 and this dimension; interleave (dimension 0 most significant);
    Gray-decode the 3*depth-bit string (prefix xor from the top)."""
    X = [grid[:, i].long().clone() for i in range(3)]
    for bit in range(depth - 1, -1, -1):
        q = 1 << bit
        p = q - 1
        for d in range(3):
            on = (X[d] & q) != 0
            t = (X[0] ^ X[d]) & p
            x0_on = X[0] ^ p
            x0_off = X[0] ^ t
            xd_off = X[d] ^ t
            if d == 0:
                X[0] = torch.where(on, x0_on, X[0])
            else:
                X[d] = torch.where(on, X[d], xd_off)
                X[0] = torch.where(on, x0_on, x0_off)
    g = z_order_encode(torch.stack(X, 1), depth)
    s = 1
    while s < 3 * depth:
        g = g ^ (g >> s)
        s *= 2
    return g


def encode(grid: torch.Tensor, batch: torch.Tensor, depth: int, order: str) -> torch.Tensor:
    """serialization/default.py:9-25"""
    if order in ("z-trans", "hilbert-trans"):
        grid = grid[:, [1, 0, 2]]
    code = z_order_encode(grid, depth) if order.startswith("z") else hilbert_encode(grid, depth)
    return (batch.long() << (depth * 3)) | code


def grid_coords(coord: torch.Tensor, grid_size: float) -> torch.Tensor:
    """pointtransformerv3.py:96-98: fp32 subtraction of the minimum over ALL points of the batch, fp32 division by the
    fp32 grid size, truncation."""
    gs = torch.tensor(grid_size, dtype=torch.float32)
    return torch.div(coord - coord.min(0)[0], gs, rounding_mode="trunc").int()


def _order_inverse(code: torch.Tensor):
    order = torch.argsort(code, dim=1, stable=True)
    inverse = torch.zeros_like(order).scatter_(1, order, torch.arange(code.shape[1]).repeat(code.shape[0], 1))
    return order, inverse


# ---------------------------------------------------------------------------------------------------------------------
# third-party operators, restated
# ---------------------------------------------------------------------------------------------------------------------
def neighbor_table(grid: torch.Tensor, batch: torch.Tensor, ksize: int) -> torch.Tensor:
    """(N, k^3) index of the active voxel at grid + (a-c, b-c, d-c), tap order (a, b, d) row-major, -1 where absent."""
    c = ksize // 2
    g = grid.long()
    side = int(g.max()) + 2 * c + 2
    key = lambda b, p: ((b * side + p[:, 0] + c) * side + p[:, 1] + c) * side + p[:, 2] + c
    keys = key(batch.long(), g)
    skeys, perm = torch.sort(keys)
    assert bool((skeys[1:] != skeys[:-1]).all()), "duplicate voxels: submanifold convolution is undefined"
    taps = []
    for a in range(ksize):
        for b_ in range(ksize):
            for d in range(ksize):
                off = torch.tensor([a - c, b_ - c, d - c])
                q = key(batch.long(), g + off)
                pos = torch.searchsorted(skeys, q).clamp(max=len(skeys) - 1)
                hit = skeys[pos] == q
                taps.append(torch.where(hit, perm[pos], torch.full_like(pos, -1)))
    return torch.stack(taps, 1)


def subm_conv3d(feat, nbr, weight, bias=None):
    """spconv SubMConv3d: out[p] = bias + sum_t W[:, t, :] . feat[nbr[p, t]] over the present taps."""
    co = weight.shape[0]
    w = weight.reshape(co, -1, weight.shape[-1])                         # (out, taps, in)
    out = feat.new_zeros(feat.shape[0], co)
    for t in range(w.shape[1]):
        idx = nbr[:, t]
        ok = idx >= 0
        if bool(ok.any()):
            out[ok] += feat[idx[ok]] @ w[:, t, :].t()
    return out if bias is None else out + bias


def patch_plan(counts: List[int], K: int):
    """SerializedAttention.get_padding_and_inverse (pointtransformerv3.py:385-441): per cloud, patches of K consecutive
    positions of the serialized order; a cloud with n <= K points is one patch of n; when n > K and n % K != 0 the last
    patch is topped up with the K - n % K points that precede it. Returns (pad, unpad, cu_seqlens) as the reference."""
    pad, unpad, cu = [], [], [0]
    off = off_pad = 0
    for n in counts:
        n_pad = n if n <= K else -(-n // K) * K
        loc = torch.arange(n_pad)
        if n_pad != n:
            loc = torch.where(loc >= n, loc - K, loc)
        pad.append(off + loc)
        unpad.append(off_pad + torch.arange(n))
        cu.extend(range(off_pad + min(K, n_pad), off_pad + n_pad + 1, K) if n_pad > 0 else [])
        off += n
        off_pad += n_pad
    return torch.cat(pad), torch.cat(unpad), torch.tensor(cu, dtype=torch.int32)


def varlen_attention_fp16(qkv, cu, H, scale):
    """flash_attn_varlen_qkvpacked_func semantics: qkv (T, 3, H, D) fp16 -> (T, H*D) fp16, non-causal per sequence."""
    T = qkv.shape[0]
    out = torch.empty(T, H, qkv.shape[-1], dtype=torch.float16)
    cu = cu.tolist()
    for s, e in zip(cu[:-1], cu[1:]):
        q, k, v = (qkv[s:e, i].float().permute(1, 0, 2) for i in range(3))
        p = torch.softmax(torch.matmul(q, k.transpose(1, 2)) * scale, dim=-1)
        out[s:e] = torch.matmul(p, v).permute(1, 0, 2).to(torch.float16)
    return out.reshape(T, -1)


# ---------------------------------------------------------------------------------------------------------------------
# the network
# ---------------------------------------------------------------------------------------------------------------------
def _lin(sd, p, x):
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


def _bn(sd, p, x, eps):
    """nn.BatchNorm1d in eval mode (running statistics), pointtransformerv3.py:861."""
    return (x - sd[p + "running_mean"]) / torch.sqrt(sd[p + "running_var"] + eps) * sd[p + "weight"] + sd[p + "bias"]


def _ln(sd, p, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps)


def block(sd, p, feat, nbr3, order, counts, heads, cfg: Ptv3Cfg):
    """Block.forward (pointtransformerv3.py:587-609), pre-norm, drop-path = identity (eval)."""
    C = feat.shape[1]
    t = subm_conv3d(feat, nbr3, sd[p + "cpe.0.weight"], sd[p + "cpe.0.bias"])
    t = _ln(sd, p + "cpe.2.", _lin(sd, p + "cpe.1.", t), cfg.ln_eps)
    feat = feat + t
    # SerializedAttention.forward (pointtransformerv3.py:443-494)
    pad, unpad, cu = patch_plan(counts, cfg.patch_size)
    inverse = torch.empty_like(order)
    inverse[order] = torch.arange(len(order))
    h = _ln(sd, p + "norm1.0.", feat, cfg.ln_eps)
    qkv = _lin(sd, p + "attn.qkv.", h)[order[pad]]
    a = varlen_attention_fp16(qkv.half().reshape(-1, 3, heads, C // heads), cu, heads, (C // heads) ** -0.5)
    a = a.to(qkv.dtype)[unpad[inverse]]
    feat = feat + _lin(sd, p + "attn.proj.", a)
    h = _ln(sd, p + "norm2.0.", feat, cfg.ln_eps)
    h = _lin(sd, p + "mlp.0.fc2.", F.gelu(_lin(sd, p + "mlp.0.fc1.", h)))
    return feat + h


def ptv3_forward(sd, coord, feat, batch, cfg: Ptv3Cfg = Ptv3Cfg(), prefix=PT, perms=None, trace=None):
    """PointTransformerV3.forward (pointtransformerv3.py:982-1002) in cls_mode. `perms`: the order shuffles
    (torch.randperm(4): one in Point.serialization :122-126, one per SerializedPooling :676-680); None draws them from
    torch's global CPU generator in that order, exactly like the reference. Returns (feat (N_last, C_last), batch)."""
    perms = list(perms) if perms is not None else None
    draw = (lambda: perms.pop(0)) if perms is not None else (lambda: torch.randperm(len(ORDERS)))
    grid = grid_coords(coord, cfg.grid_size)
    depth = int(grid.max()).bit_length()                                          # :101-103
    assert depth * 3 + len(batch.bincount()).bit_length() <= 63 and depth <= 16   # :106-111
    code = torch.stack([encode(grid, batch, depth, o) for o in ORDERS])
    order, _ = _order_inverse(code)
    perm = draw()
    code, order = code[perm], order[perm]
    # Embedding (pointtransformerv3.py:748-784): SubMConv3d(6, 32, k=5, no bias) -> BatchNorm1d -> GELU
    feat = subm_conv3d(feat, neighbor_table(grid, batch, 5), sd[prefix + "embedding.stem.conv.weight"])
    feat = F.gelu(_bn(sd, prefix + "embedding.stem.norm.", feat, cfg.bn_eps))
    if trace is not None:
        trace.append(dict(stage="embedding", feat=feat, grid=grid, batch=batch))
    for s, n_blocks in enumerate(cfg.enc_depths):
        p = prefix + f"enc.enc{s}."
        if s > 0:
            # SerializedPooling.forward (pointtransformerv3.py:643-713), stride 2 -> pooling_depth 1
            pd = (cfg.stride[s - 1] - 1).bit_length()
            if pd > depth:
                pd = 0
            c = code >> (pd * 3)
            _, cluster, counts_c = torch.unique(c[0], sorted=True, return_inverse=True, return_counts=True)
            _, indices = torch.sort(cluster, stable=True)
            idx_ptr = torch.cat([counts_c.new_zeros(1), torch.cumsum(counts_c, 0)])
            head = indices[idx_ptr[:-1]]
            proj = _lin(sd, p + "down.proj.", feat)[indices]
            feat = torch.stack([proj[idx_ptr[i]:idx_ptr[i + 1]].max(0).values for i in range(len(head))])
            code = c[:, head]
            order, _ = _order_inverse(code)
            perm = draw()
            code, order = code[perm], order[perm]
            grid, batch, depth = grid[head] >> pd, batch[head], depth - pd
            feat = F.gelu(_bn(sd, p + "down.norm.0.", feat, cfg.bn_eps))
        nbr3 = neighbor_table(grid, batch, 3)
        counts = batch.bincount().tolist()
        for i in range(n_blocks):
            feat = block(sd, p + f"block{i}.", feat, nbr3, order[i % len(ORDERS)], counts, cfg.enc_num_head[s], cfg)
        if trace is not None:
            trace.append(dict(stage=f"enc{s}", feat=feat, grid=grid, batch=batch))
    return feat, batch


def encode_pc(sd, point_clouds: List[Optional[torch.Tensor]], cfg: Ptv3Cfg = Ptv3Cfg(), prefix=PT, perms=None,
              trace=None):
    """ImageEmbeddingPooler._encode_pc (builder.py:93-148): (B, 1024) fp32; missing clouds give project_pc(0) = bias."""
    B = len(point_clouds)
    pooled = torch.zeros(B, cfg.enc_channels[-1])
    valid = [i for i, pc in enumerate(point_clouds) if pc is not None]
    if valid:
        pts = torch.cat([point_clouds[i].float() for i in valid])
        batch = torch.cat([torch.full((point_clouds[i].shape[0],), j, dtype=torch.long) for j, i in enumerate(valid)])
        feat, fb = ptv3_forward(sd, pts[:, :3], pts, batch, cfg, prefix, perms, trace)
        for j, i in enumerate(valid):
            pooled[i] = feat[fb == j].mean(0)                                    # AdaptiveAvgPool1d(1), :139-144
    return _lin(sd, prefix + "project_pc.", pooled)


# ---------------------------------------------------------------------------------------------------------------------
# synthetic inputs / weights shared by the golden generator and the tests
# ---------------------------------------------------------------------------------------------------------------------
def synth_cloud(n: int, seed: int, box=(48, 40, 4), grid_size: float = 0.01) -> torch.Tensor:
    """(n, 6) fp32 xyz (metres) + rgb in [0, 1]: n distinct voxels of a thin slab (dense enough that 3x3x3 / 5x5x5
    neighbourhoods and stride-2 pooling are exercised), jittered inside their voxels."""
    g = torch.Generator().manual_seed(seed)
    cells = box[0] * box[1] * box[2]
    assert n <= cells
    pick = torch.randperm(cells, generator=g)[:n]
    v = torch.stack([pick // (box[1] * box[2]), (pick // box[2]) % box[1], pick % box[2]], 1).float()
    xyz = (v + 0.25 + 0.5 * torch.rand(n, 3, generator=g)) * grid_size + torch.tensor([0.37, -1.21, 0.88])
    return torch.cat([xyz, torch.rand(n, 3, generator=g)], 1)


def dedupe_clouds(clouds: List[Optional[torch.Tensor]], grid_size: float = 0.01):
    """Drops points until every voxel of the batch-wide grid (min over ALL clouds, as the reference computes it) holds
    one point per cloud: spconv's submanifold convolution is undefined for duplicate sites."""
    clouds = list(clouds)
    for _ in range(8):
        valid = [c for c in clouds if c is not None]
        allp = torch.cat(valid)
        lo = allp[:, :3].min(0)[0]
        changed = False
        for i, c in enumerate(clouds):
            if c is None:
                continue
            gcoord = torch.div(c[:, :3] - lo, torch.tensor(grid_size), rounding_mode="trunc").long()
            key = (gcoord[:, 0] * 100003 + gcoord[:, 1]) * 100003 + gcoord[:, 2]
            _, first = torch.unique(key, return_inverse=True)
            keep = torch.zeros(len(key), dtype=torch.bool)
            seen = {}
            for j, k in enumerate(first.tolist()):
                if k not in seen:
                    seen[k] = j
                    keep[j] = True
            if not bool(keep.all()):
                clouds[i] = c[keep]
                changed = True
        if not changed:
            return clouds
    raise RuntimeError("could not make the voxels unique")


def weight_specs(cfg: Ptv3Cfg = Ptv3Cfg(), prefix=PT):
    """(name, shape, kind) of every tensor of PointTransformerV3(cls_mode=True) as the reference's state_dict names them;
    kind: w (weight matrix / conv), b (bias), g (norm gain), m (BN running mean), v (BN running var)."""
    out = []
    c0 = cfg.enc_channels[0]
    out.append((prefix + "embedding.stem.conv.weight", (c0, 5, 5, 5, cfg.in_channels), "w"))
    bn = lambda p, c: [(p + "weight", (c,), "g"), (p + "bias", (c,), "b"), (p + "running_mean", (c,), "m"),
                       (p + "running_var", (c,), "v")]
    lin = lambda p, i, o: [(p + "weight", (o, i), "w"), (p + "bias", (o,), "b")]
    ln = lambda p, c: [(p + "weight", (c,), "g"), (p + "bias", (c,), "b")]
    out += bn(prefix + "embedding.stem.norm.", c0)
    for s, nb in enumerate(cfg.enc_depths):
        C = cfg.enc_channels[s]
        p = prefix + f"enc.enc{s}."
        if s > 0:
            out += lin(p + "down.proj.", cfg.enc_channels[s - 1], C) + bn(p + "down.norm.0.", C)
        for i in range(nb):
            q = p + f"block{i}."
            out += [(q + "cpe.0.weight", (C, 3, 3, 3, C), "w"), (q + "cpe.0.bias", (C,), "b")]
            out += lin(q + "cpe.1.", C, C) + ln(q + "cpe.2.", C) + ln(q + "norm1.0.", C)
            out += lin(q + "attn.qkv.", C, 3 * C) + lin(q + "attn.proj.", C, C) + ln(q + "norm2.0.", C)
            out += lin(q + "mlp.0.fc1.", C, cfg.mlp_ratio * C) + lin(q + "mlp.0.fc2.", cfg.mlp_ratio * C, C)
    out += lin(prefix + "project_pc.", cfg.enc_channels[-1], cfg.project_pc_dim)
    return out


def synth_weights(cfg: Ptv3Cfg = Ptv3Cfg(), seed: int = 11, prefix=PT, bf16_round: bool = True):
    """Seeded random weights with non-trivial BatchNorm statistics. Rounded to bf16 like the loaded model's
    (model/builder.py:166 casts the pooler to bf16 before _encode_pc widens PTv3 back to fp32, builder.py:95)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape, kind in weight_specs(cfg, prefix):
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.0 / fan_in ** 0.5)
        elif kind == "b":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "g":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "m":
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.5 + torch.rand(shape, generator=g)
        sd[name] = t.to(torch.bfloat16).float() if bf16_round else t
    return sd
