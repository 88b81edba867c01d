"""Runtime shim that imports the UNMODIFIED reference model code from /root/reference (TEST INFRASTRUCTURE ONLY).

Only usable where /root/reference exists (the build container); used by tests/golden/make_golden.py to record
golden outputs of the reference's own LlavaLlamaForCausalLM, and by nothing that runs on the GPU box.

The reference cannot be imported verbatim under transformers 5.5.0:
  * absent third-party modules are stubbed (matplotlib, spconv, torch_scatter, addict, timm, torchinfo);
  * AutoConfig.register("llava", ...) clashes with HF's own "llava" -> registered with exist_ok=True;
  * clip_encoder.py downloads CLIP from the hub -> the three from_pretrained calls build from an in-memory config;
  * multimodal_projector/builder.py hard-codes device 'cuda' and dtype bfloat16 for the audio / seg-mask branches ->
    the module's `torch` global is replaced by a proxy whose torch.device('cuda') returns the CPU device and whose
    bfloat16 is the fp32 working dtype of the golden run (inputs are bf16-representable, so nothing changes).
No reference source is copied or edited.
"""
import sys
import types

REF_ROOT = "/root/reference/scene_graph_generation/LLaVA"


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def _install_ptv3_stubs():
    """Functional stand-ins for the four third-party packages PointTransformerV3 (multimodal_projector/
    pointtransformerv3.py) needs and this container lacks or cannot run on CPU. Each restates the PUBLISHED semantics of
    the one entry point the reference calls -- deliberately written naively (Python dict lookups, per-sequence loops)
    and independently of oracle/ptv3_oracle.py, so that the golden run cross-checks the oracle's vectorised versions:
      spconv.pytorch (spconv-cu117 2.x, README.md:110): SparseConvTensor(features, indices[b,x,y,z], spatial_shape,
        batch_size) / .replace_feature; SubMConv3d(in, out, kernel_size, bias, indice_key, padding): submanifold
        convolution = cross-correlation evaluated at the active sites only, centred kernel whatever `padding` says,
        weight layout (out, k0, k1, k2, in) (spconv 2.x KRSC), bias (out,);
      torch_scatter.segment_csr(src, indptr, reduce): out[i] = reduce(src[indptr[i]:indptr[i+1]]);
      flash_attn.flash_attn_varlen_qkvpacked_func(qkv(total,3,H,D) fp16, cu_seqlens, max_seqlen, dropout_p,
        softmax_scale): per-sequence non-causal softmax(q k^T * scale) v, fp32 accumulation, fp16 result;
      addict.Dict: dict with attribute access whose constructor re-wraps nested dicts in the same class.
    """
    import torch
    import torch.nn as nn

    class SparseConvTensor:
        def __init__(self, features, indices, spatial_shape, batch_size):
            self.features, self.indices = features, indices
            self.spatial_shape, self.batch_size = spatial_shape, batch_size

        def replace_feature(self, feature):
            return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size)

    class SubMConv3d(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size, bias=True, indice_key=None, padding=0, **kw):
            super().__init__()
            k = kernel_size
            self.k = k
            self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))
            nn.init.kaiming_uniform_(self.weight.view(out_channels, -1), a=5 ** 0.5)
            self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

        def forward(self, x):
            idx = x.indices.tolist()
            site = {tuple(v): i for i, v in enumerate(idx)}
            assert len(site) == len(idx), "duplicate voxels: submanifold convolution is undefined"
            k, c = self.k, self.k // 2
            out = x.features.new_zeros(len(idx), self.weight.shape[0])
            for a in range(k):
                for b_ in range(k):
                    for d in range(k):
                        dst, src = [], []
                        for i, (bi, px, py, pz) in enumerate(idx):
                            j = site.get((bi, px + a - c, py + b_ - c, pz + d - c))
                            if j is not None:
                                dst.append(i)
                                src.append(j)
                        if dst:
                            out[dst] += x.features[src] @ self.weight[:, a, b_, d, :].t()
            if self.bias is not None:
                out = out + self.bias
            return x.replace_feature(out)

    sp = _stub("spconv")
    sp.pytorch = _stub("spconv.pytorch", SubMConv3d=SubMConv3d, SparseConvTensor=SparseConvTensor,
                       modules=types.SimpleNamespace(is_spconv_module=lambda m: isinstance(m, SubMConv3d)))

    def segment_csr(src, indptr, reduce="sum"):
        outs = []
        for i in range(len(indptr) - 1):
            seg = src[int(indptr[i]):int(indptr[i + 1])]
            outs.append({"max": lambda s: s.max(0).values, "min": lambda s: s.min(0).values,
                         "mean": lambda s: s.mean(0), "sum": lambda s: s.sum(0)}[reduce](seg))
        return torch.stack(outs)

    _stub("torch_scatter", segment_csr=segment_csr)

    def flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, **kw):
        assert qkv.dtype == torch.float16 and dropout_p == 0
        out = torch.empty_like(qkv[:, 0])
        cu = cu_seqlens.tolist()
        for s, e in zip(cu[:-1], cu[1:]):
            assert e - s <= max_seqlen
            q, k, v = (qkv[s:e, i].float().transpose(0, 1) for i in range(3))       # (H, n, D)
            p = torch.softmax(q @ k.transpose(1, 2) * softmax_scale, dim=-1)
            out[s:e] = (p @ v).transpose(0, 1).to(torch.float16)
        return out

    _stub("flash_attn", flash_attn_varlen_qkvpacked_func=flash_attn_varlen_qkvpacked_func)

    class Dict(dict):
        def __init__(self, *args, **kwargs):
            super().__init__()
            for a in args:
                if a:
                    for k, v in (a.items() if isinstance(a, dict) else a):
                        self[k] = self._hook(v)
            for k, v in kwargs.items():
                self[k] = self._hook(v)

        @classmethod
        def _hook(cls, v):
            return cls(v) if isinstance(v, dict) and not isinstance(v, cls) else v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    _stub("addict", Dict=Dict)


_loaded = None


def load_reference(vit_cfg_kwargs):
    """Returns (LlavaConfig, LlavaLlamaForCausalLM) classes of the reference. vit_cfg_kwargs: CLIPVisionConfig kwargs."""
    global _loaded
    import torch
    import torch.nn as nn
    import transformers  # noqa: F401  (must be imported before the stubs)
    from transformers import AutoConfig, AutoModelForCausalLM, CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel

    if _loaded is None:
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
        _stub("torchinfo", summary=lambda *a, **k: None)
        _install_ptv3_stubs()

        class DropPath(nn.Module):
            def __init__(self, p=0.0):
                super().__init__()

            def forward(self, x):
                return x

        _stub("timm")
        _stub("timm.models")
        _stub("timm.models.layers", DropPath=DropPath)
        _r = AutoConfig.register
        AutoConfig.register = staticmethod(lambda mt, c, exist_ok=False: _r(mt, c, exist_ok=True))
        _r2 = AutoModelForCausalLM.register.__func__
        AutoModelForCausalLM.register = classmethod(lambda cls, c, m, exist_ok=False: _r2(cls, c, m, exist_ok=True))
        sys.path.insert(0, REF_ROOT)
        from llava.model.language_model.llava_llama import LlavaConfig, LlavaLlamaForCausalLM
        from llava.model.multimodal_encoder import clip_encoder
        from llava.model.multimodal_projector import builder as proj_builder

        class _TorchProxy:
            def __getattr__(self, k):
                return getattr(torch, k)

            @staticmethod
            def device(*a, **k):
                return torch.device("cpu")

            # the audio / seg-mask branches allocate bf16 buffers regardless of the module dtype (builder.py:152,163);
            # the golden run is fp32, so inside this one module "bfloat16" means the working dtype
            bfloat16 = torch.float32

        proj_builder.torch = _TorchProxy()
        _loaded = (LlavaConfig, LlavaLlamaForCausalLM, clip_encoder)
    LlavaConfig, LlavaLlamaForCausalLM, clip_encoder = _loaded
    vcfg = CLIPVisionConfig(**vit_cfg_kwargs)
    clip_encoder.CLIPVisionConfig.from_pretrained = classmethod(lambda c, n, **k: vcfg)
    clip_encoder.CLIPVisionModel.from_pretrained = classmethod(lambda c, n, **k: CLIPVisionModel(vcfg))
    s = vit_cfg_kwargs["image_size"]
    clip_encoder.CLIPImageProcessor.from_pretrained = classmethod(
        lambda c, n, **k: CLIPImageProcessor(size={"shortest_edge": s}, crop_size={"height": s, "width": s}))
    return LlavaConfig, LlavaLlamaForCausalLM


def build_reference_model(cfg, state_dict):
    """cfg: mm_or_b200.config.LlavaConfig; state_dict: reference-named fp32 tensors. Returns the reference model (CPU,
    fp32, eval) with the weights loaded; asserts that every tensor on the hot path was consumed."""
    import torch
    vc = cfg.vision_config()
    LlavaConfig, LlavaLlamaForCausalLM = load_reference(dict(
        hidden_size=vc["hidden_size"], intermediate_size=vc["intermediate_size"],
        num_hidden_layers=vc["num_hidden_layers"], num_attention_heads=vc["num_attention_heads"],
        image_size=vc["image_size"], patch_size=vc["patch_size"], projection_dim=768))
    rcfg = LlavaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                       num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                       num_key_value_heads=cfg.num_attention_heads, vocab_size=cfg.vocab_size,
                       max_position_embeddings=cfg.max_position_embeddings, rms_norm_eps=cfg.rms_norm_eps,
                       mm_vision_tower="openai/clip-vit-large-patch14-336", mm_projector_type="mlp2x_gelu",
                       mm_hidden_size=cfg.mm_hidden_size, mm_vision_select_layer=cfg.mm_vision_select_layer,
                       mm_vision_select_feature="patch", mv_type="learned", pad_token_id=0, bos_token_id=1,
                       eos_token_id=2, attn_implementation="eager")
    torch.manual_seed(0)
    model = LlavaLlamaForCausalLM(rcfg).eval()
    model.get_vision_tower().load_model()
    model.float()
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, f"unexpected keys: {unexpected[:5]}"
    bad = [k for k in missing if "point_transformer" not in k and "inv_freq" not in k and "position_ids" not in k
           and "token_type_ids" not in k]
    assert not bad, f"hot-path tensors not provided: {bad[:8]}"
    return model


def load_reference_pooler():
    """Returns (multimodal_projector.builder module, serialization package) of the reference, importable without the
    rest of LLaVA being constructed: used by tests/golden/make_ptv3_golden.py for the point-cloud branch."""
    load_reference(dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                        image_size=28, patch_size=14, projection_dim=32))
    from llava.model.multimodal_projector import builder as proj_builder
    from llava.model.multimodal_projector import serialization
    return proj_builder, serialization
