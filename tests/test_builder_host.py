"""Host logic of the loader mirror (mm_or_b200/model/builder.py vs the reference LLaVA/llava/model/builder.py:26-184):
checkpoint-directory reading, non_lora_trainables key remaps and the LoRA merge. The GPU test loads a full synthetic
checkpoint through load_pretrained_model and checks that generation equals the directly-constructed model."""
import json
import os

import pytest
import torch

from mm_or_b200.model import builder as B


def test_read_checkpoint_dir_sharded_safetensors(tmp_path):
    from safetensors.torch import save_file
    a = {"model.embed_tokens.weight": torch.randn(8, 4), "model.norm.weight": torch.ones(4)}
    b = {"lm_head.weight": torch.randn(8, 4)}
    save_file(a, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file(b, str(tmp_path / "model-00002-of-00002.safetensors"))
    wm = {k: "model-00001-of-00002.safetensors" for k in a}
    wm.update({k: "model-00002-of-00002.safetensors" for k in b})
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"weight_map": wm}))
    sd = B.read_checkpoint_dir(str(tmp_path))
    assert set(sd) == set(a) | set(b)
    assert torch.equal(sd["lm_head.weight"], b["lm_head.weight"])


def test_read_checkpoint_dir_bin_and_missing(tmp_path):
    with pytest.raises(FileNotFoundError):
        B.read_checkpoint_dir(str(tmp_path))
    torch.save({"x": torch.arange(3)}, str(tmp_path / "pytorch_model.bin"))
    assert torch.equal(B.read_checkpoint_dir(str(tmp_path))["x"], torch.arange(3))


def test_remap_non_lora_trainables_prefixes():
    raw = {"base_model.model.model.mm_projector.0.weight": torch.zeros(1),
           "base_model.model.model.image_pooler.bert.embeddings.position_ids": torch.zeros(1),
           "base_model.model.model.image_pooler.bert.embeddings.LayerNorm.weight": torch.zeros(1),
           "base_model.model.lm_head.weight": torch.zeros(1)}
    out = B.remap_non_lora_trainables(raw)
    assert set(out) == {"model.mm_projector.0.weight", "model.image_pooler.bert.embeddings.LayerNorm.weight",
                        "lm_head.weight"}


def test_merge_lora_matches_peft_formula():
    torch.manual_seed(0)
    w = torch.randn(16, 12).to(torch.bfloat16)
    A, Bm = torch.randn(4, 12), torch.randn(16, 4)
    sd = {"model.layers.0.self_attn.q_proj.weight": w.clone(), "other": torch.ones(2)}
    ad = {"base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight": A,
          "base_model.model.model.layers.0.self_attn.q_proj.lora_B.default.weight": Bm}
    n = B.merge_lora(sd, ad, {"lora_alpha": 256, "r": 128})
    assert n == 1
    ref = (w.float() + 2.0 * (Bm @ A)).to(torch.bfloat16)
    assert torch.equal(sd["model.layers.0.self_attn.q_proj.weight"], ref)
    with pytest.raises(KeyError):
        B.merge_lora({}, ad, {"lora_alpha": 1, "r": 1})
    with pytest.raises(ValueError):
        B.merge_lora(sd, {"base_model.model.x.lora_A.weight": A}, {"lora_alpha": 1, "r": 1})


def test_quantised_and_foreign_models_rejected():
    with pytest.raises(NotImplementedError):
        B.load_pretrained_model("x", None, "llava-v1.5-7b", load_4bit=True)
    with pytest.raises(NotImplementedError):
        B.load_pretrained_model("x", None, "vicuna-7b")


class _Tok:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def add_tokens(self, toks, special_tokens=False):
        self.n += len(toks)


@pytest.mark.gpu
def test_load_pretrained_model_roundtrip(tmp_path, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    import golden_cases as gc
    from safetensors.torch import save_file
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    torch.set_grad_enabled(False)
    cfg = gc.small_config(mm_use_im_patch_token=False)
    cfg.tokenizer_padding_side = "left"
    sd = {k: v.to(torch.bfloat16).contiguous() for k, v in gc.small_weights(cfg).items()}
    # LoRA layout: base checkpoint without projector/pooler/tower, adapter dir with the rest + a rank-4 adapter
    base_dir, lora_dir = tmp_path / "base", tmp_path / "llava-lora"
    base_dir.mkdir(), lora_dir.mkdir()
    extra_keys = [k for k in sd if k.startswith(("model.mm_projector", "model.image_pooler", "model.vision_tower"))]
    save_file({k: v for k, v in sd.items() if k not in extra_keys}, str(base_dir / "model.safetensors"))
    cfg.save_pretrained(str(lora_dir))
    torch.save({"base_model.model." + k: sd[k] for k in extra_keys}, str(lora_dir / "non_lora_trainables.bin"))
    g = torch.Generator().manual_seed(5)
    key = "model.layers.1.mlp.down_proj"
    A = torch.randn(4, sd[key + ".weight"].shape[1], generator=g) * 0.05
    Bm = torch.randn(sd[key + ".weight"].shape[0], 4, generator=g) * 0.05
    torch.save({f"base_model.model.{key}.lora_A.weight": A, f"base_model.model.{key}.lora_B.weight": Bm},
               str(lora_dir / "adapter_model.bin"))
    (lora_dir / "adapter_config.json").write_text(json.dumps({"lora_alpha": 8, "r": 4}))
    monkeypatch.setattr(B, "_tokenizer", lambda p: _Tok(cfg.vocab_size))
    tok, model, image_processor, ctx = B.load_pretrained_model(str(lora_dir), str(base_dir), "llava-lora-test")
    assert ctx == 2048 and len(tok) == cfg.vocab_size
    merged = dict(sd)
    merged[key + ".weight"] = (sd[key + ".weight"].float() + 2.0 * (Bm @ A)).to(torch.bfloat16)
    ref = LlavaLlamaForCausalLM(cfg).load_state_dict(merged)
    case = gc.make_case(cfg, "infer_left")
    a = model.generate(case["input_ids"], images=case["images"], max_new_tokens=4, stop_on_eos=False)
    b = ref.generate(case["input_ids"], images=case["images"], max_new_tokens=4, stop_on_eos=False)
    assert torch.equal(a, b)


def test_lora_base_without_clip_takes_the_tower_from_mm_vision_tower(tmp_path):
    """The reference's LoRA flow: the base holds no CLIP weights, load_model() fetches them from config.mm_vision_tower
    and non_lora_trainables overlays the trained parts (model/builder.py:148-176). Here the tower comes from the LOCAL
    directory config.mm_vision_tower; nothing is dropped silently."""
    import golden_cases as gc
    from safetensors.torch import save_file
    cfg = gc.small_config()
    full = {k: v.to(torch.bfloat16) for k, v in gc.small_weights(cfg).items()}
    vit = B.VIT_PREFIX
    clip_dir = tmp_path / "clip"
    clip_dir.mkdir()
    save_file({k[len(vit):]: v.contiguous() for k, v in full.items() if k.startswith(vit)},
              str(clip_dir / "model.safetensors"))
    base = {k: v for k, v in full.items() if k.startswith(("model.layers.", "model.embed", "model.norm", "lm_head"))}
    trained_key = vit + "vision_model.encoder.layers.1.mlp.fc1.weight"          # an unfrozen CLIP layer
    overlay = {k: v for k, v in full.items() if k.startswith(("model.image_pooler.", "model.mm_projector."))}
    overlay[trained_key] = full[trained_key] + 1
    cfg.mm_vision_tower = str(clip_dir)
    sd = B.ensure_vision_weights(dict(base), cfg, overlay=overlay)
    assert sorted(sd) == sorted(full)
    assert torch.equal(sd[trained_key], full[trained_key] + 1)                      # overlay wins over the CLIP directory
    assert torch.equal(sd[vit + "vision_model.pre_layrnorm.weight"], full[vit + "vision_model.pre_layrnorm.weight"])
    # no local CLIP directory and an overlay without the tower: a loud error naming what is missing
    cfg.mm_vision_tower = "openai/clip-vit-large-patch14-336"
    with pytest.raises(KeyError, match="vision tower"):
        B.ensure_vision_weights(dict(base), cfg, overlay=overlay)
    with pytest.raises(KeyError, match="image pooler.*mm_projector"):
        B.ensure_vision_weights({k: v for k, v in full.items() if not k.startswith(("model.image_pooler.",
                                                                                    "model.mm_projector."))}, cfg)
    # a full checkpoint passes untouched
    assert sorted(B.ensure_vision_weights(dict(full), cfg)) == sorted(full)
