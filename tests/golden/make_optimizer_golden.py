"""Records which parameters the REFERENCE's LLaVATrainer.create_optimizer decays and which take mm_projector_lr
(train/llava_trainer.py:203-233: decay_parameters = get_parameter_names(model, ALL_LAYERNORM_LAYERS) minus names
containing "bias"; projector_parameters = names containing "mm_projector"), evaluated on the reference's own model
instance (oracle/ref_shim.py) -> tests/golden/optimizer_groups.json; and the state_dict key set of the reference's
ImageEmbeddingPooler (what its loader's strict load_state_dict needs) -> tests/golden/pooler_state_keys.json.

Run in the build container only:   python tests/golden/make_optimizer_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_cases as gc
from oracle.ref_shim import build_reference_model


def main():
    cfg = gc.small_config()
    model = build_reference_model(cfg, gc.bf16_round(gc.small_weights(cfg)))
    from transformers.pytorch_utils import ALL_LAYERNORM_LAYERS
    from transformers.trainer_pt_utils import get_parameter_names
    import torch.nn as nn
    norm_types = list(ALL_LAYERNORM_LAYERS)
    rms = type(model.model.norm)                       # transformers 4.31 registers LlamaRMSNorm in ALL_LAYERNORM_LAYERS
    if rms not in norm_types:                          # (modeling_llama.py); newer versions match it by name instead
        norm_types.append(rms)
    assert nn.LayerNorm in norm_types
    decay = get_parameter_names(model, norm_types)
    decay = [n for n in decay if "bias" not in n]       # llava_trainer.py:204-205
    names = [n for n, _ in model.named_parameters()]
    rec = {n: {"decay": n in decay, "projector": "mm_projector" in n} for n in names}
    with open(os.path.join(gc.GOLDEN_DIR, "optimizer_groups.json"), "w") as f:
        json.dump(rec, f, indent=0, sort_keys=True)
    print(len(rec), "parameters;", sum(v["decay"] for v in rec.values()), "decayed;",
          sum(v["projector"] for v in rec.values()), "projector")
    # the key set the reference's loader demands of non_lora_trainables.bin for the image pooler: it loads the stripped
    # `model.image_pooler.*` entries with strict=True (model/builder.py:160-176) -- every parameter AND buffer
    keys = sorted(model.get_image_pooler().state_dict().keys())
    with open(os.path.join(gc.GOLDEN_DIR, "pooler_state_keys.json"), "w") as f:
        json.dump(keys, f, indent=0)
    print(len(keys), "image_pooler state_dict keys")


if __name__ == "__main__":
    main()
