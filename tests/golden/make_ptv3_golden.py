"""Records golden outputs of the REFERENCE point-cloud branch (ImageEmbeddingPooler._encode_pc ->
PointTransformerV3 -> project_pc, imported unmodified from /root/reference through oracle/ref_shim.py) and of the
reference's own serialization code (serialization/{default,z_order,hilbert}.py, pure torch, runs verbatim).

Run in the build container only:   python tests/golden/make_ptv3_golden.py
Writes tests/golden/ptv3_codes.pt (int64 codes, bit-exact pin) and tests/golden/ptv3_encode.pt (fp32 features).
spconv / torch_scatter / flash-attn are absent here: the shim's naive stand-ins restate their published semantics
(oracle/ref_shim.py::_install_ptv3_stubs), everything else executed is the reference's code.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

from oracle import ptv3_oracle as P
from oracle.ref_shim import load_reference_pooler

OUT = os.path.dirname(os.path.abspath(__file__))
SHUFFLE_SEED = 123


def ptv3_case():
    """3 samples: 2500 points (3 patches, the last one topped up), no cloud, 700 points (one short patch)."""
    clouds = [P.synth_cloud(2500, seed=1), None, P.synth_cloud(700, seed=2, box=(30, 30, 3))]
    return P.dedupe_clouds(clouds)


def main():
    torch.set_grad_enabled(False)
    builder, ser = load_reference_pooler()

    # ---- serialization codes: the reference's functions verbatim
    g = torch.Generator().manual_seed(5)
    codes = {}
    for depth in (1, 3, 7, 9, 12, 16):
        grid = torch.randint(0, 1 << depth, (257, 3), generator=g, dtype=torch.int32)
        if depth >= 3:
            grid[0] = 0
            grid[1] = (1 << depth) - 1
        batch = torch.randint(0, 5, (257,), generator=g)
        codes[depth] = dict(grid=grid, batch=batch,
                            **{o: ser.encode(grid, batch, depth, order=o) for o in P.ORDERS})
    torch.save(codes, os.path.join(OUT, "ptv3_codes.pt"))

    # ---- the whole branch
    pooler = builder.ImageEmbeddingPooler().eval()
    sd = P.synth_weights(prefix="")
    missing, unexpected = pooler.point_transformer.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    clouds = ptv3_case()
    trace = []
    pt = pooler.point_transformer

    def hook(name):
        def f(mod, inp, out):
            trace.append(dict(stage=name, feat=out.feat.clone(), grid=out.grid_coord.clone(), batch=out.batch.clone()))
        return f

    pt.embedding.register_forward_hook(hook("embedding"))
    for s in range(5):
        getattr(pt.enc, f"enc{s}").register_forward_hook(hook(f"enc{s}"))
    torch.manual_seed(SHUFFLE_SEED)
    out = pooler._encode_pc(clouds)
    fx = dict(pc_feats=out.float(), shuffle_seed=SHUFFLE_SEED, n_points=[None if c is None else len(c) for c in clouds])
    for t in trace:                       # canonical row order (batch, x, y, z) so that index order does not matter
        key = ((t["batch"].long() * 4096 + t["grid"][:, 0]) * 4096 + t["grid"][:, 1]) * 4096 + t["grid"][:, 2]
        o = torch.argsort(key)
        fx[t["stage"]] = dict(feat=t["feat"][o].half() if t["stage"] != "enc4" else t["feat"][o], key=key[o])
    torch.save(fx, os.path.join(OUT, "ptv3_encode.pt"))
    print("reference pc_feats", out.shape, float(out.abs().mean()), {k: tuple(v["feat"].shape) for k, v in fx.items()
                                                                  if isinstance(v, dict)})

    # ---- the oracle against what was just recorded
    torch.manual_seed(SHUFFLE_SEED)
    mine = P.encode_pc(P.synth_weights(), clouds)
    err = ((mine - out).norm() / out.norm()).item()
    print("oracle vs reference rel err", err)
    assert err < 1e-4


if __name__ == "__main__":
    main()
