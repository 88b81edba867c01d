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
        _stub("torch_scatter")
        _stub("torchinfo", summary=lambda *a, **k: None)

        class _SubM(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

        sp = _stub("spconv")
        sp.pytorch = _stub("spconv.pytorch", SubMConv3d=_SubM, SparseConvTensor=object,
                           modules=types.SimpleNamespace(is_spconv_module=lambda m: isinstance(m, _SubM)))
        _stub("addict", Dict=type("Dict", (dict,), {}))

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
