"""LlavaConfig mirror (reference: LLaVA/llava/model/language_model/llava_llama.py:27-28, a LlamaConfig subclass with
model_type "llava"). Callers mutate attributes after load (mv_type, tokenizer_padding_side,
tokenizer_model_max_length, use_cache ...) and the hot path reads them with getattr at call time
(llava_arch.py:204,303,319), so this is a plain attribute bag with the Llama defaults."""
import json
import os

# CLIP geometries the reference can name through mm_vision_tower (multimodal_encoder/builder.py:6-12 accepts
# openai/* and laion/* names or an absolute path); MM2SG uses openai/clip-vit-large-patch14-336.
_KNOWN_TOWERS = {
    "openai/clip-vit-large-patch14-336": dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                              num_attention_heads=16, image_size=336, patch_size=14,
                                              layer_norm_eps=1e-5),
    "openai/clip-vit-large-patch14": dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                          num_attention_heads=16, image_size=224, patch_size=14, layer_norm_eps=1e-5),
}


class LlavaConfig:
    model_type = "llava"

    def __init__(self, **kw):
        d = dict(
            vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
            num_attention_heads=32, num_key_value_heads=None, max_position_embeddings=4096, rms_norm_eps=1e-5,
            rope_theta=10000.0, pad_token_id=0, bos_token_id=1, eos_token_id=2, use_cache=True,
            pretraining_tp=1, tie_word_embeddings=False,
            mm_vision_tower="openai/clip-vit-large-patch14-336", mm_projector_type="mlp2x_gelu",
            mm_hidden_size=1024, mm_vision_select_layer=-2, mm_vision_select_feature="patch",
            mm_use_im_start_end=False, mm_use_im_patch_token=False, image_aspect_ratio="pad",
            mv_type="learned", tokenizer_padding_side="right", tokenizer_model_max_length=None,
            mm_vision_config=None,  # optional dict overriding the tower geometry (tests use shallow towers)
        )
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads

    def vision_config(self):
        if self.mm_vision_config is not None:
            base = dict(_KNOWN_TOWERS["openai/clip-vit-large-patch14-336"])
            base.update(self.mm_vision_config)
            return base
        if self.mm_vision_tower not in _KNOWN_TOWERS:
            raise ValueError(f"Unknown vision tower: {self.mm_vision_tower}")  # multimodal_encoder/builder.py:12
        return dict(_KNOWN_TOWERS[self.mm_vision_tower])

    def to_dict(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_")}

    @classmethod
    def from_pretrained(cls, path, **kw):
        with open(os.path.join(path, "config.json")) as f:
            d = json.load(f)
        d.update(kw)
        d.pop("model_type", None)
        return cls(**d)

    def save_pretrained(self, path):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(dict(self.to_dict(), model_type=self.model_type), f, indent=1)
