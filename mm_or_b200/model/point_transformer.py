"""Point-cloud branch of the image pooler: host orchestration of the b200_pc_* operators (csrc/ptv3.cu).

Mirrors `ImageEmbeddingPooler._encode_pc` (multimodal_projector/builder.py:93-148) and `PointTransformerV3` in cls_mode
(multimodal_projector/pointtransformerv3.py:787-1005, default geometry, eval mode: BatchNorm running statistics,
drop-path off): list of (N_i, 6) xyz+rgb clouds or None -> one 1024-d token per sample, `project_pc(0)` = bias for
samples without a cloud. Like the reference's Python, this file only sequences operators and does the small integer
bookkeeping of cloud sizes; every array operation is a kernel of libb200mmor.so.

Layout (DESIGN.md "Point clouds"): at every level the points are kept physically sorted by their z-order code, so the
"z" serialization is the identity, pooling clusters are contiguous runs and neighbour lookup is a binary search. The
reference's point order inside a level is different (sorted unique of whichever order the shuffle put first); the
network is invariant to it except for fp32 summation order in the final per-cloud mean.

The order shuffles (`torch.randperm(4)`, pointtransformerv3.py:122-126 and :676-680, active in eval mode too) are
drawn from torch's global CPU generator in the reference's sequence: one for the serialization, one per pooling.
"""
import torch

from .. import _lib as L

ORDERS = ("z", "z-trans", "hilbert", "hilbert-trans")   # pointtransformerv3.py:791
GEOMETRY = dict(in_channels=6, stride=(2, 2, 2, 2), enc_depths=(2, 2, 2, 6, 2), enc_channels=(32, 64, 128, 256, 512),
                enc_num_head=(2, 4, 8, 16, 32), patch_size=1024, mlp_ratio=4, project_pc_dim=1024, grid_size=0.01,
                bn_eps=1e-3, ln_eps=1e-5)
ACT_NONE, ACT_GELU = 0, 2


class PcOps:
    """The b200_pc_* entry points bound to one device. The product builds it with `PcOps.cuda(device)` (libb200mmor.so,
    CUDA tensors only); the CPU test-suite builds it over the kernel emulator's library (tests/emu/)."""

    def __init__(self, cdll, device, ptr, stream, check):
        self.lib, self.device, self.ptr, self.stream, self.check = cdll, torch.device(device), ptr, stream, check

    @classmethod
    def cuda(cls, device):
        return cls(L.lib(), device, L.ptr, L.stream_ptr, L.check)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(*args, self.stream()), name)

    # ---- geometry
    def grid_coords(self, pts, grid_size):
        n = pts.shape[0]
        min3, grid, mx = self.empty(3, torch.float32), self.empty((n, 3), torch.int32), self.empty(1, torch.int32)
        self.call("b200_pc_grid_coords", self.ptr(pts), pts.stride(0), n, float(grid_size), self.ptr(min3),
                  self.ptr(grid), self.ptr(mx))
        return grid, mx

    def encode(self, grid, batch, n, depth, order):
        code = self.empty(n, torch.int64)
        self.call("b200_pc_encode", self.ptr(grid), self.ptr(batch), n, depth, order, self.ptr(code))
        return code

    def argsort(self, code, n, bits):
        nb = int(self.lib.b200_pc_argsort_workspace_bytes(n))
        ws = self.empty(max(nb, 256), torch.uint8)
        out, order = self.empty(n, torch.int64), self.empty(n, torch.int32)
        self.call("b200_pc_argsort", self.ptr(code), n, bits, self.ptr(out), self.ptr(order), self.ptr(ws), ws.numel())
        return out, order

    def gather_rows(self, src, idx, n):
        width = src.shape[1] if src.dim() == 2 else 1
        dst = self.empty((n, width) if src.dim() == 2 else (n,), src.dtype)
        assert src.element_size() == 4
        self.call("b200_pc_gather_rows", self.ptr(src), src.stride(0), self.ptr(idx), n, width, self.ptr(dst), width)
        return dst

    def neighbors(self, zcode, grid, batch, n, depth, ksize, dup):
        nbr = self.empty((n, ksize ** 3), torch.int32)
        self.call("b200_pc_neighbors", self.ptr(zcode), self.ptr(grid), self.ptr(batch), n, depth, ksize, self.ptr(nbr),
                  self.ptr(dup))
        return nbr

    def pool_plan(self, zcode, grid, batch, n, pd):
        seg, n_out = self.empty(n + 1, torch.int32), self.empty(1, torch.int32)
        grid_o, batch_o = self.empty((n, 3), torch.int32), self.empty(n, torch.int32)
        self.call("b200_pc_pool_plan", self.ptr(zcode), self.ptr(grid), self.ptr(batch), n, pd, self.ptr(seg),
                  self.ptr(n_out), self.ptr(grid_o), self.ptr(batch_o))
        return seg, n_out, grid_o, batch_o

    def cloud_offsets(self, batch, n, n_clouds):
        off = self.empty(n_clouds + 1, torch.int32)
        self.call("b200_pc_cloud_offsets", self.ptr(batch), n, n_clouds, self.ptr(off))
        return off

    # ---- features
    def gemm(self, a, w, M, K, N, idx=None, taps=1, bias=None, bn=None, act=ACT_NONE, residual=None, out=None,
             ldc=None, out_bf16=False):
        if out is None:
            out = self.empty((M, N), torch.float32)
            ldc = N
        scale, shift = bn if bn is not None else (None, None)
        nb = int(self.lib.b200_pc_gemm_workspace_bytes(M, N, K, taps))      # > 0 when the taps are split over CTAs
        ws = self.empty(nb, torch.uint8) if nb else None
        self.call("b200_pc_gemm_f32", self.ptr(a), a.stride(0), self.ptr(idx), taps, self.ptr(w), self.ptr(bias),
                  self.ptr(scale), self.ptr(shift), act, self.ptr(residual),
                  residual.stride(0) if residual is not None else 0, self.ptr(out), ldc, int(out_bf16), M, N, K,
                  self.ptr(ws), nb)
        return out

    def layernorm(self, x, M, C, gamma, beta, eps, residual=None):
        out = self.empty((M, C), torch.float32)
        self.call("b200_pc_layernorm_f32", self.ptr(x), x.stride(0), self.ptr(gamma), self.ptr(beta), float(eps),
                  self.ptr(residual), residual.stride(0) if residual is not None else 0, self.ptr(out), C, M, C)
        return out

    def patch_attention(self, qkv, order, patches, n_patches, max_q, M, C, heads):
        out = self.empty((M, C), torch.float32)
        self.call("b200_pc_patch_attention", self.ptr(qkv), qkv.stride(0), self.ptr(order), self.ptr(patches), n_patches,
                  max_q, C, heads, float((C // heads) ** -0.5), self.ptr(out), C)
        return out

    def segment_max(self, x, seg, n_seg, C, bn, act):
        out = self.empty((n_seg, C), torch.float32)
        self.call("b200_pc_segment_max", self.ptr(x), x.stride(0), self.ptr(seg), n_seg, C, self.ptr(bn[0]),
                  self.ptr(bn[1]), act, self.ptr(out), C)
        return out

    def cloud_mean(self, x, off, n_clouds, C, row_map, out):
        self.call("b200_pc_cloud_mean", self.ptr(x), x.stride(0), self.ptr(off), n_clouds, C, self.ptr(row_map),
                  self.ptr(out), out.stride(0))


def weight_specs(geometry=None, prefix=""):
    """[(name, shape, kind)] of PointTransformerV3(cls_mode=True) under the reference's state_dict names
    (pointtransformerv3.py:864-919); kind: w weight, b bias, g norm gain, m / v BatchNorm running mean / variance."""
    g = dict(GEOMETRY if geometry is None else geometry)
    bn = lambda p, c: [(p + "weight", (c,), "g"), (p + "bias", (c,), "b"), (p + "running_mean", (c,), "m"),
                       (p + "running_var", (c,), "v")]
    lin = lambda p, i, o: [(p + "weight", (o, i), "w"), (p + "bias", (o,), "b")]
    ln = lambda p, c: [(p + "weight", (c,), "g"), (p + "bias", (c,), "b")]
    c0 = g["enc_channels"][0]
    out = [(prefix + "embedding.stem.conv.weight", (c0, 5, 5, 5, g["in_channels"]), "w")]
    out += bn(prefix + "embedding.stem.norm.", c0)
    for s, nb in enumerate(g["enc_depths"]):
        C = g["enc_channels"][s]
        p = prefix + f"enc.enc{s}."
        if s > 0:
            out += lin(p + "down.proj.", g["enc_channels"][s - 1], C) + bn(p + "down.norm.0.", C)
        for i in range(nb):
            q = p + f"block{i}."
            out += [(q + "cpe.0.weight", (C, 3, 3, 3, C), "w"), (q + "cpe.0.bias", (C,), "b")]
            out += lin(q + "cpe.1.", C, C) + ln(q + "cpe.2.", C) + ln(q + "norm1.0.", C)
            out += lin(q + "attn.qkv.", C, 3 * C) + lin(q + "attn.proj.", C, C) + ln(q + "norm2.0.", C)
            out += lin(q + "mlp.0.fc1.", C, g["mlp_ratio"] * C) + lin(q + "mlp.0.fc2.", g["mlp_ratio"] * C, C)
    return out + lin(prefix + "project_pc.", g["enc_channels"][-1], g["project_pc_dim"])


def patch_descriptors(counts, K):
    """(q_begin, q_len, k_begin, k_len) per patch in positions of a serialized order: the padding rule of
    SerializedAttention.get_padding_and_inverse (pointtransformerv3.py:385-441) without materialising pad / unpad: a
    cloud of n <= K points is one patch; otherwise full patches of K, and a last patch of n % K points that attends to
    the cloud's last K points (the reference tops it up with the K - n % K points before it)."""
    out, off = [], 0
    for n in counts:
        if n <= 0:
            continue
        if n <= K:
            out.append((off, n, off, n))
        else:
            full, r = divmod(n, K)
            out += [(off + j * K, K, off + j * K, K) for j in range(full)]
            if r:
                out.append((off + full * K, r, off + n - K, K))
        off += n
    return out


class _Level:
    __slots__ = ("n", "grid", "batch", "zcode", "depth", "orders", "names", "nbr3", "counts", "seg", "patches",
                 "n_patches", "max_q", "cloud_off")


class PointTransformerV3:
    """Weights of `model.image_pooler.point_transformer.*` (fp32 on the device, linears / convolutions transposed once to
    [K, N], BatchNorm folded to scale / shift) + forward."""

    def __init__(self, geometry=None):
        self.geo = dict(GEOMETRY if geometry is None else geometry)
        self.w = None
        self.ops = None

    # -----------------------------------------------------------------------------------------------------------------
    def load_weights(self, sd, prefix, device, ops=None):
        g = self.geo
        self.ops = ops if ops is not None else PcOps.cuda(device)
        dev = self.ops.device
        # the reference casts the whole pooler (parameters AND BatchNorm buffers) to bf16 at load (model/builder.py:166)
        # and widens PTv3 back to fp32 at call time (multimodal_projector/builder.py:95): values are bf16-rounded
        f32 = lambda k: sd[prefix + k].detach().to(torch.bfloat16).to(torch.float32)
        put = lambda t: t.contiguous().to(dev)

        def lin(p):      # nn.Linear (out, in) -> [in, out]
            return dict(w=put(f32(p + "weight").t()), b=put(f32(p + "bias")))

        def conv(p, bias=True):   # SubMConv3d (out, k, k, k, in) -> [taps * in, out]
            w = f32(p + "weight")
            co, ci = w.shape[0], w.shape[-1]
            d = dict(w=put(w.reshape(co, -1, ci).permute(1, 2, 0).reshape(-1, co)))
            d["b"] = put(f32(p + "bias")) if bias else None
            return d

        def bn(p):       # eval-mode BatchNorm1d folded: y = x * scale + shift
            scale = f32(p + "weight") / torch.sqrt(f32(p + "running_var") + g["bn_eps"])
            return put(scale), put(f32(p + "bias") - f32(p + "running_mean") * scale)

        def ln(p):
            return put(f32(p + "weight")), put(f32(p + "bias"))

        w = dict(stem=conv("embedding.stem.conv.", bias=False), stem_bn=bn("embedding.stem.norm."), stages=[])
        for s, nb in enumerate(g["enc_depths"]):
            p = f"enc.enc{s}."
            st = dict(blocks=[])
            if s > 0:
                st["down"] = lin(p + "down.proj.")
                st["down_bn"] = bn(p + "down.norm.0.")
            for i in range(nb):
                q = p + f"block{i}."
                st["blocks"].append(dict(cpe0=conv(q + "cpe.0."), cpe1=lin(q + "cpe.1."), cpe2=ln(q + "cpe.2."),
                                         norm1=ln(q + "norm1.0."), qkv=lin(q + "attn.qkv."), proj=lin(q + "attn.proj."),
                                         norm2=ln(q + "norm2.0."), fc1=lin(q + "mlp.0.fc1."), fc2=lin(q + "mlp.0.fc2.")))
            w["stages"].append(st)
        w["project_pc"] = lin("project_pc.")
        self.w = w
        return self

    # -----------------------------------------------------------------------------------------------------------------
    def plan(self, pts, batch, n_clouds):
        """Geometry of all levels (depends on coordinates only): z-sorted points, the three other serialized orders,
        neighbour tables, pooling clusters, attention patches. Returns (pts sorted, levels)."""
        ops, g = self.ops, self.geo
        n = pts.shape[0]
        grid, mx = ops.grid_coords(pts, g["grid_size"])
        depth = int(mx.item()).bit_length()                                           # pointtransformerv3.py:101-103
        if depth * 3 + n_clouds.bit_length() > 63 or depth > 16:                      # :106-111
            raise AssertionError(f"point cloud spans 2^{depth} voxels: serialization codes need depth <= 16")
        code = ops.encode(grid, batch, n, depth, 0)
        zcode, order0 = ops.argsort(code, n, 3 * depth + n_clouds.bit_length())
        pts, grid, batch = ops.gather_rows(pts, order0, n), ops.gather_rows(grid, order0, n), \
            ops.gather_rows(batch, order0, n)
        names = [ORDERS[i] for i in torch.randperm(len(ORDERS)).tolist()]             # :122-126
        dup = torch.zeros(1, dtype=torch.int32).to(ops.device)
        nbr5 = ops.neighbors(zcode, grid, batch, n, depth, 5, dup)
        levels = []
        n_stage = len(g["enc_depths"])
        for s in range(n_stage):
            lv = _Level()
            lv.n, lv.grid, lv.batch, lv.zcode, lv.depth, lv.names = n, grid, batch, zcode, depth, names
            bits = 3 * depth + n_clouds.bit_length()
            lv.orders = {0: torch.arange(n, dtype=torch.int32, device=ops.device)}
            for k in (1, 2, 3):
                _, lv.orders[k] = ops.argsort(ops.encode(grid, batch, n, depth, k), n, max(bits, 1))
            lv.nbr3 = ops.neighbors(zcode, grid, batch, n, depth, 3, dup)
            off = ops.cloud_offsets(batch, n, n_clouds)
            lv.cloud_off = off
            lv.seg = None
            if s + 1 < n_stage:
                pd = (g["stride"][s] - 1).bit_length()                                # :644-646
                if pd > depth:
                    pd = 0
                seg, n_out, grid_o, batch_o = ops.pool_plan(zcode, grid, batch, n, pd)
                lv.seg = seg
            host = off.cpu().tolist()                                                 # one host sync per level
            lv.counts = [host[b + 1] - host[b] for b in range(n_clouds)]
            pat = patch_descriptors(lv.counts, g["patch_size"])
            lv.n_patches, lv.max_q = len(pat), max(p[1] for p in pat)
            lv.patches = torch.tensor(pat, dtype=torch.int32).to(ops.device)
            levels.append(lv)
            if s + 1 < n_stage:
                n = int(n_out.item())
                grid, batch, depth = grid_o[:n], batch_o[:n], depth - pd
                zcode = ops.encode(grid, batch, n, depth, 0)                          # born sorted
                perm = torch.randperm(len(ORDERS)).tolist()                           # :676-680
                names = [names[i] for i in perm]
        if int(dup.item()) != 0:
            raise ValueError("point cloud has several points in one 1 cm voxel: submanifold convolution is undefined "
                             "for duplicate sites (voxel-downsample the cloud first)")
        return pts, nbr5, levels

    def forward(self, point_clouds, out=None):
        """point_clouds: list (len B) of (N_i, 6) float tensors or None. Returns (B, project_pc_dim) bf16, or writes the
        rows of `out` (a (B, D) bf16 view with row stride, e.g. the pc token slot of the pooler output)."""
        if self.w is None:
            raise L.B200Error("PointTransformerV3 weights not loaded")
        ops, g, w = self.ops, self.geo, self.w
        B = len(point_clouds)
        C_last = g["enc_channels"][-1]
        pooled = torch.zeros((B, C_last), dtype=torch.float32).to(ops.device)         # builder.py:99
        valid = [i for i, pc in enumerate(point_clouds) if pc is not None]
        if valid:
            for i in valid:
                if point_clouds[i].dim() != 2 or point_clouds[i].shape[1] != g["in_channels"] or \
                        point_clouds[i].shape[0] == 0:
                    raise ValueError(f"point cloud {i}: expected (N > 0, {g['in_channels']}) xyz+rgb, got "
                                     f"{tuple(point_clouds[i].shape)}")
            # clouds may arrive on the host (the reference's loader) or already on the device: one copy either way
            pts = torch.cat([point_clouds[i].detach().to(ops.device, torch.float32) for i in valid]).contiguous()
            batch = torch.cat([torch.full((point_clouds[i].shape[0],), j, dtype=torch.int32)
                               for j, i in enumerate(valid)]).to(ops.device)
            pts, nbr5, levels = self.plan(pts, batch, len(valid))
            feat = self.features(pts, nbr5, levels)
            row_map = torch.tensor(valid, dtype=torch.int32).to(ops.device)
            ops.cloud_mean(feat, levels[-1].cloud_off, len(valid), C_last, row_map, pooled)
        D = g["project_pc_dim"]
        if out is None:
            out = torch.empty((B, D), dtype=torch.bfloat16, device=ops.device)
        assert out.dtype == torch.bfloat16 and out.shape == (B, D) and out.stride(1) == 1
        ops.gemm(pooled, w["project_pc"]["w"], B, C_last, D, bias=w["project_pc"]["b"], out=out, ldc=out.stride(0),
                 out_bf16=True)
        return out

    __call__ = forward

    def features(self, pts, nbr5, levels):
        ops, g, w = self.ops, self.geo, self.w
        C = g["enc_channels"][0]
        n = levels[0].n
        f = ops.gemm(pts, w["stem"]["w"], n, g["in_channels"], C, idx=nbr5, taps=125, bn=w["stem_bn"], act=ACT_GELU)
        for s, lv in enumerate(levels):
            st = w["stages"][s]
            C, H, n = g["enc_channels"][s], g["enc_num_head"][s], lv.n
            if s > 0:
                prev = levels[s - 1]
                p = ops.gemm(f, st["down"]["w"], prev.n, g["enc_channels"][s - 1], C, bias=st["down"]["b"])
                f = ops.segment_max(p, prev.seg, n, C, st["down_bn"], ACT_GELU)
            for i, bw in enumerate(st["blocks"]):
                order = lv.orders[ORDERS.index(lv.names[i % len(ORDERS)])]
                t = ops.gemm(f, bw["cpe0"]["w"], n, C, C, idx=lv.nbr3, taps=27, bias=bw["cpe0"]["b"])
                t = ops.gemm(t, bw["cpe1"]["w"], n, C, C, bias=bw["cpe1"]["b"])
                f = ops.layernorm(t, n, C, *bw["cpe2"], g["ln_eps"], residual=f)
                h = ops.layernorm(f, n, C, *bw["norm1"], g["ln_eps"])
                qkv = ops.gemm(h, bw["qkv"]["w"], n, C, 3 * C, bias=bw["qkv"]["b"])
                a = ops.patch_attention(qkv, order, lv.patches, lv.n_patches, lv.max_q, n, C, H)
                f = ops.gemm(a, bw["proj"]["w"], n, C, C, bias=bw["proj"]["b"], residual=f)
                h = ops.layernorm(f, n, C, *bw["norm2"], g["ln_eps"])
                m = ops.gemm(h, bw["fc1"]["w"], n, C, g["mlp_ratio"] * C, bias=bw["fc1"]["b"], act=ACT_GELU)
                f = ops.gemm(m, bw["fc2"]["w"], n, g["mlp_ratio"] * C, C, bias=bw["fc2"]["b"], residual=f)
        return f
