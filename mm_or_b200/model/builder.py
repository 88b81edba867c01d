"""`load_pretrained_model` with the reference's signature and return value.

Reference: LLaVA/llava/model/builder.py:26-184. Same four-tuple `(tokenizer, model, image_processor, context_len)`,
same handling of the three checkpoint layouts MM2SG produces:
  * a full LLaVA checkpoint directory (config.json + weight shards),
  * a LoRA checkpoint (`model_base` = the Vicuna/LLaVA base, `model_path` = adapter dir with adapter_model.*,
    adapter_config.json and non_lora_trainables.bin) -- `'lora' in model_name` (builder.py:51-96),
  * a projector-only checkpoint (`mm_projector.bin` next to config.json, builder.py:97-115).
Key remaps follow builder.py:81-83 (strip `base_model.` / one leading `model.`) and :148-177 (vision tower / image
pooler keys stay under `model.vision_tower.` / `model.image_pooler.`; BERT `position_ids` / `token_type_ids` buffers
are dropped). LoRA adapters are merged on the host in fp32 (W += (alpha / r) * B @ A, what peft's merge_and_unload
does) before the weights are re-laid-out for the fused kernels. 8-bit / 4-bit loading (bitsandbytes) has no
counterpart in the bf16 B200 path and raises NotImplementedError.
"""
import glob
import json
import os
import warnings

import torch

from ..config import LlavaConfig
from ..constants import DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_PATCH_TOKEN
from .llava_llama import LlavaLlamaForCausalLM


def _load_file(path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu", weights_only=True)


def read_checkpoint_dir(path):
    """All tensors of an HF-format directory: sharded or single safetensors / pytorch_model*.bin."""
    sd = {}
    for index in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
        ip = os.path.join(path, index)
        if os.path.exists(ip):
            with open(ip) as f:
                files = sorted(set(json.load(f)["weight_map"].values()))
            for fn in files:
                sd.update(_load_file(os.path.join(path, fn)))
            return sd
    files = sorted(glob.glob(os.path.join(path, "model*.safetensors"))) or \
        sorted(glob.glob(os.path.join(path, "pytorch_model*.bin")))
    if not files:
        raise FileNotFoundError(f"no model*.safetensors / pytorch_model*.bin under {path}")
    for fn in files:
        sd.update(_load_file(fn))
    return sd


def remap_non_lora_trainables(sd):
    """builder.py:81-83: drop the peft wrapper prefixes so keys read `model.…` / `lm_head.…`."""
    sd = {(k[len("base_model."):] if k.startswith("base_model.") else k): v for k, v in sd.items()}
    if any(k.startswith("model.model.") for k in sd):
        sd = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in sd.items()}
    for k in [k for k in sd if k.endswith("bert.embeddings.position_ids") or k.endswith("bert.embeddings.token_type_ids")]:
        sd.pop(k)                                                                   # builder.py:174-175
    return sd


def merge_lora(sd, adapter_sd, adapter_cfg):
    """W += (lora_alpha / r) * lora_B @ lora_A for every adapted Linear (peft merge_and_unload, builder.py:91-94).
    Adapter keys look like `base_model.model.<module path>.lora_A.weight` (peft 0.4) or `...lora_A.default.weight`."""
    scale = float(adapter_cfg["lora_alpha"]) / float(adapter_cfg["r"])
    pairs = {}
    for k, v in adapter_sd.items():
        kk = k.replace(".default", "")
        for tag in ("lora_A", "lora_B"):
            suffix = f".{tag}.weight"
            if kk.endswith(suffix):
                mod = kk[:-len(suffix)]
                for pre in ("base_model.model.", "base_model."):
                    if mod.startswith(pre):
                        mod = mod[len(pre):]
                        break
                pairs.setdefault(mod, {})[tag] = v
    merged = 0
    for mod, ab in pairs.items():
        if "lora_A" not in ab or "lora_B" not in ab:
            raise ValueError(f"incomplete LoRA pair for {mod}")
        key = mod + ".weight"
        if key not in sd:
            raise KeyError(f"LoRA adapter targets {key}, which is not in the base checkpoint")
        w = sd[key].float() + scale * (ab["lora_B"].float() @ ab["lora_A"].float())
        sd[key] = w.to(sd[key].dtype)
        merged += 1
    return merged


VIT_PREFIX = "model.vision_tower.vision_tower."


def ensure_vision_weights(sd, config, overlay=None):
    """The reference's LoRA flow starts from a LLaVA / Vicuna base WITHOUT CLIP weights: CLIPVisionTower.load_model()
    fetches them from `config.mm_vision_tower` (multimodal_encoder/clip_encoder.py:20-24, model/builder.py:148-152) and
    non_lora_trainables.bin then overlays only the unfrozen CLIP layers, the image pooler and the projector
    (builder.py:153-176). Same order here: when `sd` has no vision tower, read it from the LOCAL directory
    `config.mm_vision_tower` (there is no hub access; HF CLIPVisionModel keys `vision_model.*`), then apply `overlay`
    (the remapped non-LoRA trainables). Whatever the route, a checkpoint that leaves the tower, the pooler or the
    projector without weights is an error here -- not a generic 'weights not loaded' at the first encode."""
    probe = VIT_PREFIX + "vision_model.pre_layrnorm.weight"
    if probe not in sd and getattr(config, "mm_vision_tower", None):
        src = str(config.mm_vision_tower)
        if os.path.isdir(src):
            clip = read_checkpoint_dir(src)
            n = 0
            for k, v in clip.items():
                if k.startswith("vision_model."):
                    sd[VIT_PREFIX + k] = v
                    n += 1
            if n == 0:
                raise KeyError(f"{src}: no `vision_model.*` tensors (not a CLIP vision checkpoint)")
    if overlay:
        sd.update(overlay)
    if getattr(config, "mm_vision_tower", None):
        need = {"vision tower": probe, "image pooler": "model.image_pooler.bert.embeddings.LayerNorm.weight",
                "mm_projector": "model.mm_projector.0.weight"}
        missing = [f"{what} ({key.rsplit('.', 2)[0]}.*)" for what, key in need.items() if key not in sd]
        if missing:
            raise KeyError("checkpoint has no weights for: " + ", ".join(missing) + f". The base holds none, "
                           f"`mm_vision_tower` = {config.mm_vision_tower!r} is not a local CLIP directory (the "
                           "reference downloads it from the hub; this path has no network), and "
                           "non_lora_trainables.bin / mm_projector.bin did not provide them.")
    return sd


def _tokenizer(path):
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(path, use_fast=False)


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False, device_map="auto",
                          device="cuda"):
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes 8-bit / NF4 loading is not part of the bf16 B200 path")
    if "llava" not in model_name.lower():
        raise NotImplementedError("plain language-model loading (builder.py:116-139) is outside the MM2SG path")
    if "mpt" in model_name.lower():
        raise NotImplementedError("LlavaMPT is dead code for MM2SG (SURVEY.md 2.1)")
    dev = torch.device(device if device != "cuda" else "cuda")
    extra = None
    if "lora" in model_name.lower() and model_base is None:
        warnings.warn("There is `lora` in model name but no `model_base` is provided. If you are loading a LoRA "
                      "model, please provide the `model_base` argument.")
    if "lora" in model_name.lower() and model_base is not None:
        config = LlavaConfig.from_pretrained(model_path)
        tokenizer = _tokenizer(model_base)
        sd = read_checkpoint_dir(model_base)
        nlt = os.path.join(model_path, "non_lora_trainables.bin")
        if not os.path.exists(nlt):
            raise FileNotFoundError(f"{nlt} not found (the hub download of builder.py:69-79 needs network access)")
        extra = remap_non_lora_trainables(torch.load(nlt, map_location="cpu", weights_only=True))
        ensure_vision_weights(sd, config, overlay=extra)       # CLIP from mm_vision_tower if the base has none
        with open(os.path.join(model_path, "adapter_config.json")) as f:
            acfg = json.load(f)
        adapters = [p for p in (os.path.join(model_path, "adapter_model.safetensors"),
                                os.path.join(model_path, "adapter_model.bin")) if os.path.exists(p)]
        if not adapters:
            raise FileNotFoundError(f"no adapter_model.* under {model_path}")
        merge_lora(sd, _load_file(adapters[0]), acfg)
    elif model_base is not None:
        config = LlavaConfig.from_pretrained(model_path)
        tokenizer = _tokenizer(model_base)
        sd = read_checkpoint_dir(model_base)
        proj = torch.load(os.path.join(model_path, "mm_projector.bin"), map_location="cpu", weights_only=True)
        ensure_vision_weights(sd, config, overlay={k: v.to(torch.bfloat16) for k, v in proj.items()})
    else:
        config = LlavaConfig.from_pretrained(model_path)
        tokenizer = _tokenizer(model_path)
        sd = read_checkpoint_dir(model_path)
        ensure_vision_weights(sd, config)

    model = LlavaLlamaForCausalLM(config)
    if getattr(config, "mm_use_im_patch_token", True):
        tokenizer.add_tokens([DEFAULT_IMAGE_PATCH_TOKEN], special_tokens=True)
    if getattr(config, "mm_use_im_start_end", False):
        tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN], special_tokens=True)
    rows = sd["model.embed_tokens.weight"].shape[0]
    if len(tokenizer) > rows:
        # resize_token_embeddings (builder.py:146): HF appends rows initialised to the mean of the existing ones
        for k in ("model.embed_tokens.weight", "lm_head.weight"):
            w = sd[k]
            pad = w.float().mean(0, keepdim=True).to(w.dtype).expand(len(tokenizer) - rows, -1)
            sd[k] = torch.cat([w, pad], 0)
        config.vocab_size = len(tokenizer)
        model.vocab_size = config.vocab_size
    model.load_state_dict(sd, device=dev)
    tower = model.get_vision_tower()
    if tower is not None and not tower.is_loaded:
        tower.load_model()
    image_processor = tower.get_image_processor() if tower is not None else None
    context_len = getattr(config, "max_sequence_length", 2048)
    return tokenizer, model, image_processor, context_len
