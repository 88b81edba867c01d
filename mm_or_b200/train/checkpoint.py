"""What the reference's training script leaves in `output_dir`, written from a FineTuner -- so that a model trained
here loads through `load_pretrained_model` (this repo's or the reference's) exactly like one trained there.

Reference (LLaVA/llava/train/train.py:1346-1360):
  LoRA recipe      config.json                 model.config.save_pretrained
                   adapter_config.json,        peft `model.save_pretrained(output_dir, state_dict=get_peft_state_maybe_zero_3(...))`
                   adapter_model.bin           keys `base_model.model.<module path>.lora_{A,B}.weight` (peft 0.4 strips the
                                               adapter name on save), train.py:134-156
                   non_lora_trainables.bin     every non-LoRA parameter that requires grad (+ buffers), keys as
                                               PeftModel.named_parameters() spells them (`base_model.model.…`),
                                               train.py:168-178; the loader strips the prefixes (model/builder.py:81-83)
  full fine-tune   config.json + the whole state_dict (safe_save_model_for_hf_trainer, train.py:186-217)
Plus one file the reference gets from HF Trainer / DeepSpeed instead (`trainer.save_state`, checkpoint-*/): the
optimizer state, here `b200_optimizer.pt` (fp32 masters, m, v, step count, learning rates) for `FineTuner.load_optimizer`.

Host-side file I/O only; tensors are copied to the CPU as they are written.
"""
import json
import os

import torch

PEFT_PREFIX = "base_model.model."
LORA_TARGETS = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]   # find_all_linear_names, train.py:187-200


def _cpu(t):
    return t.detach().to("cpu").contiguous()


def adapter_config(r, alpha, base_model_name_or_path=None, dropout=0.05):
    """adapter_config.json as peft 0.4 writes it for train.py:1160-1167 (LoraConfig(r, lora_alpha, target_modules,
    lora_dropout, bias, task_type='CAUSAL_LM'))."""
    return {"peft_type": "LORA", "task_type": "CAUSAL_LM", "base_model_name_or_path": base_model_name_or_path,
            "r": int(r), "lora_alpha": float(alpha), "lora_dropout": float(dropout), "bias": "none",
            "target_modules": list(LORA_TARGETS), "fan_in_fan_out": False, "inference_mode": True,
            "init_lora_weights": True, "modules_to_save": None, "layers_to_transform": None, "layers_pattern": None}


POOLER_PREFIX = "model.image_pooler."


def pooler_checkpoint_tensors(sd):
    """Every `model.image_pooler.*` tensor of a state dict, trained or frozen, plus the BatchNorm step counters.
    The reference marks all image_pooler parameters trainable (llava_arch.py:80-83), saves parameters that require grad
    AND every buffer (train.py:168-178, the `_extended` variant), and its loader feeds the stripped entries to
    `image_pooler.load_state_dict(..., strict=True)` (model/builder.py:160-176): a file that holds only the tensors this
    FineTuner happened to update (no point_transformer.*, no bert.pooler.*, no word_embeddings) would make that load
    raise. `num_batches_tracked` (a buffer of every BatchNorm of PointTransformerV3, absent from this repo's weights
    because inference never reads it) is written as 0: the point-cloud encoder is frozen here, no batch was tracked."""
    out = {k: v for k, v in sd.items() if k.startswith(POOLER_PREFIX)}
    for k in list(out):
        if k.endswith(".running_mean"):
            out.setdefault(k[:-len("running_mean")] + "num_batches_tracked", torch.zeros((), dtype=torch.long))
    return out


def export_lora_checkpoint(out_dir, config, lora_sd, r, alpha, trainable_sd, base_model_name_or_path=None,
                           dropout=0.05):
    """lora_sd: {`model.layers.N.<module>.<proj>.lora_{A,B}.weight`: tensor}; trainable_sd: {reference parameter name:
    tensor} of the non-LoRA trainables (projector, image pooler, unfrozen CLIP layers, ...)."""
    os.makedirs(out_dir, exist_ok=True)
    config.save_pretrained(out_dir)
    with open(os.path.join(out_dir, "adapter_config.json"), "w") as f:
        json.dump(adapter_config(r, alpha, base_model_name_or_path, dropout), f, indent=1)
    bad = [k for k in lora_sd if ".lora_A." not in k and ".lora_B." not in k]
    if bad:
        raise ValueError(f"not LoRA parameters: {bad[:3]}")
    torch.save({PEFT_PREFIX + k: _cpu(v) for k, v in lora_sd.items()}, os.path.join(out_dir, "adapter_model.bin"))
    torch.save({PEFT_PREFIX + k: _cpu(v) for k, v in trainable_sd.items() if "lora_" not in k},
               os.path.join(out_dir, "non_lora_trainables.bin"))
    return out_dir


def export_full_checkpoint(out_dir, config, sd, shard_bytes=5 << 30):
    """config.json + pytorch_model-0000x-of-0000n.bin + pytorch_model.bin.index.json (HF sharded layout, what
    `read_checkpoint_dir` and HF `from_pretrained` read), or a single pytorch_model.bin when it fits one shard."""
    os.makedirs(out_dir, exist_ok=True)
    config.save_pretrained(out_dir)
    shards, cur, size = [], {}, 0
    for k in sorted(sd):
        nb = sd[k].numel() * sd[k].element_size()
        if cur and size + nb > shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = sd[k]
        size += nb
    shards.append(cur)
    if len(shards) == 1:
        torch.save({k: _cpu(v) for k, v in shards[0].items()}, os.path.join(out_dir, "pytorch_model.bin"))
        return out_dir
    weight_map, total = {}, 0
    for i, sh in enumerate(shards):
        fn = "pytorch_model-%05d-of-%05d.bin" % (i + 1, len(shards))
        torch.save({k: _cpu(v) for k, v in sh.items()}, os.path.join(out_dir, fn))
        for k, v in sh.items():
            weight_map[k] = fn
            total += v.numel() * v.element_size()
    with open(os.path.join(out_dir, "pytorch_model.bin.index.json"), "w") as f:
        json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f, indent=1)
    return out_dir


def save_optimizer(path, master, m, v, step_count, lr, proj_lr, extra=None):
    """fp32 masters / first / second moments under the reference's parameter names (a rank's slices when the optimizer
    state is sharded: one file per rank then)."""
    torch.save({"format": "b200_optimizer/1", "step_count": int(step_count), "lr": float(lr), "proj_lr": float(proj_lr),
                "master": {k: _cpu(t) for k, t in master.items()}, "m": {k: _cpu(t) for k, t in m.items()},
                "v": {k: _cpu(t) for k, t in v.items()}, "extra": extra or {}}, path)


def load_optimizer(path, master, m, v):
    """Copies a saved state into the (already allocated) master / m / v tensors; returns (step_count, lr, proj_lr).
    Shapes must match: a state saved with another sharding or trainable set is an error, not a silent partial load."""
    st = torch.load(path, map_location="cpu", weights_only=True)
    if st.get("format") != "b200_optimizer/1":
        raise ValueError(f"{path}: not a b200 optimizer state")
    for name, dst in (("master", master), ("m", m), ("v", v)):
        src = st[name]
        if sorted(src) != sorted(dst):
            missing = sorted(set(dst) - set(src))[:3]
            extra = sorted(set(src) - set(dst))[:3]
            raise KeyError(f"{path}: parameter set differs (missing {missing}, unexpected {extra})")
        for k, t in dst.items():
            if tuple(src[k].shape) != tuple(t.shape):
                raise ValueError(f"{path}: {name}[{k}] has shape {tuple(src[k].shape)}, expected {tuple(t.shape)}")
            t.copy_(src[k].to(t.device, t.dtype))
    return st["step_count"], st["lr"], st["proj_lr"]
