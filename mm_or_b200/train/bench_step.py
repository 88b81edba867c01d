"""Timing of the fine-tune step of BASELINE.json configs[4] at the real model size, shared by bench.py (the `train`
sub-record of the bench line, so that the driver's 1 -> 8 GPU scaling run captures it per N) and tools/train_bench.py.

Workload (cfg 5): per GPU `batch` samples (4 = the reference's per_device_train_batch_size) of 6 views,
prompt 256 ids + 150 answer ids, right padded, labels on the answer only, class-weighted CE; trainable = the whole
decoder + lm_head + projector + image pooler + CLIP layers 12..22 (train.py:1257-1261 with full fine-tuning of the
LLM), fp32 master weights, gradient clipping 0.1, fused AdamW lr 2e-5. Data parallel over `group`: ZeRO-2 (bf16
reduce-scatter of gradients, sliced AdamW, bf16 all-gather of the updated slices, train/zero.py) when world > 1.
Trained tokens/s counts all L = 981 packed positions of every sample. Device-timed (CUDA events), max over ranks.
"""
import json

import torch

from .. import _lib as L
from ..synth import make_state_dict, synth_batch
from .step import FineTuner

PROMPT, ANSWER = 256, 150


def step_flops(cfg, batch, views, L_packed):
    """Algorithmic FLOPs of one fine-tune step: forward = encoder + decoder with logits on all positions,
    backward = 2 x forward for everything that is trained or has a trained ancestor (the 11 frozen CLIP layers and the
    patch embedding only run forward). Recomputation is NOT counted (it is overhead, not model work)."""
    d, f, n_l = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    lin = n_l * (4 * d * d + 3 * d * f)
    dec = 2 * lin * L_packed + n_l * 4 * L_packed * L_packed * d / 2 + 2 * d * cfg.vocab_size * L_packed
    vit_layer = 577 * ((4 * 1024 ** 2 + 2 * 1024 * 4096) * 2 + 4 * 577 * 1024)
    S = views * 576
    pooler = 2 * S * ((4 * 1024 ** 2 + 2 * 1024 * 4096) * 2 + 4 * S * 1024)
    proj = 576 * (1024 * d + d * d) * 2
    fwd = views * (23 * vit_layer + 576 * 1024 * 588 * 2) + pooler + proj + dec
    bwd = 2 * (views * 11 * vit_layer + pooler + proj + dec)
    return batch * (fwd + bwd)


def measure_finetune_step(cfg, dev, group=None, batch=4, views=6, steps=2, warmup=1, zero=2, lora_r=0, nf4=False,
                          accum=1, recompute=False, prof=False, model=None, sd=None, seed=0, trace=False):
    """Builds the FineTuner, runs `warmup` + `steps` optimizer steps, returns the record (a dict; rank 0's is the one
    to print). model / sd: reuse an already loaded model and its reference-named bf16 state dict (bench.py)."""
    import torch.distributed as dist
    from ..model.llava_llama import LlavaLlamaForCausalLM
    world = dist.get_world_size(group) if group is not None else 1
    rank = dist.get_rank(group) if group is not None else 0
    if sd is None:
        sd = make_state_dict(cfg, seed=seed, device=dev, dtype=torch.bfloat16)
    if model is None:
        model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device=dev)
    model.config.tokenizer_padding_side = "right"
    b = synth_batch(cfg, batch, views, PROMPT + ANSWER, seed=3 + rank, jitter=0, image_pos=40, dtype=torch.bfloat16)
    ids = b["input_ids"]
    labels = ids.clone()
    labels[:, :PROMPT] = -100
    labels[ids == -200] = -100
    w = torch.rand(cfg.vocab_size, generator=torch.Generator().manual_seed(1)) + 0.01
    lora = None
    if lora_r > 0:
        from .lora import LoraState
        lora = LoraState(cfg, r=lora_r, alpha=2 * lora_r, device=dev)
    torch.cuda.reset_peak_memory_stats(dev)
    ft = FineTuner(model, sd, lr=2e-5, weight_decay=0.0, max_grad_norm=0.1, first_trainable_clip_layer=12,
                   vocab_weight=w, lora=lora, group=group, shard_optimizer=world > 1 and zero >= 1,
                   shard_gradients=world > 1 and zero >= 2, base_nf4=nf4, recompute_activations=recompute)
    del sd
    n_train = sum(ft.sd[k].numel() for k in ft.names)
    L_packed = PROMPT + ANSWER - 1 + 576
    tokens = batch * L_packed * accum
    mb = dict(input_ids=ids, labels=labels, attention_mask=b["attention_mask"], images=b["images"])

    def one_step():
        if accum > 1:
            return ft.train_step_accumulated([mb] * accum)
        return ft.train_step(ids, labels, b["attention_mask"], b["images"])

    losses = []
    for _ in range(warmup):
        loss, _ = one_step()
        losses.append(float(loss))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier(group)
    n0 = L.launch_count()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record()
    timed_losses = []
    for i in range(steps):
        loss, nsq = one_step()
        marks[i + 1].record()
        timed_losses.append(loss)            # read back after the timed region: no host synchronisation between steps
    torch.cuda.synchronize(dev)
    losses += [float(x) for x in timed_losses]
    # every step is timed on the device; the record quotes the MEDIAN step (a full fine-tune holds ~167 GB of a 180 GB
    # device on one GPU, and a step in which the caching allocator has to return and re-request segments costs up to
    # 100 ms more than its neighbours) and lists all of them
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    mine = sorted(step_ms)[len(step_ms) // 2] if len(step_ms) % 2 else \
        0.5 * (sorted(step_ms)[len(step_ms) // 2 - 1] + sorted(step_ms)[len(step_ms) // 2])
    launches = int((L.launch_count() - n0) / steps)
    per_rank = [mine]
    ms = mine
    if world > 1:
        t = torch.tensor([mine], device=dev)
        allms = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allms, t, group=group)
        per_rank = [round(float(x), 1) for x in allms]
        ms = max(per_rank)                                   # device time, max over ranks
    phases = None
    if trace:
        from . import trace as TR
        TR.enable(True)
        one_step()
        phases = TR.collect()
        TR.enable(False)
    fam = None
    if prof:
        L.prof_enable(True)
        one_step()
        fam = {k: {"ms": round(v["ms"], 2), "launches": v["launches"],
                   "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1) if v["flops"] else None}
               for k, v in L.prof_collect().items() if v["launches"]}
        L.prof_enable(False)
    flops = step_flops(cfg, batch, views, L_packed) * accum if lora_r == 0 else None
    payload = n_train * 2                                    # bf16 bytes reduce-scattered, and again all-gathered
    rec = {"metric": "fine-tune step, trained tokens/s", "value": round(world * tokens / (ms / 1e3), 1),
           "unit": "tokens/s", "n_gpus": world, "scaling": "weak", "ms_per_step": round(ms, 1),
           "per_rank_ms": per_rank, "step_ms_this_rank": [round(x, 1) for x in step_ms], "timing": "median of %d "
           "device-timed steps, max over ranks" % steps, "tokens_per_step_per_gpu": tokens, "trainable_params": n_train,
           "decoder_layers": cfg.num_hidden_layers, "batch_per_gpu": batch, "views": views,
           "gradient_accumulation": accum, "activation_recomputation": bool(recompute),
           "mode": (("qlora (nf4 base) r=%d" if nf4 else "lora r=%d") % lora_r) if lora_r else "full fine-tune",
           "nf4_packed_base_gb": round(ft.nf4_bytes / 1e9, 2) if ft.nf4_bytes else None,
           "data_parallel": ({0: "replicated state, fp32 all-reduce", 1: "ZeRO-1 (sharded fp32 state)",
                              2: "ZeRO-2 (sharded fp32 state, bf16 reduce-scatter + bf16 all-gather)"}[zero]
                             if world > 1 else "none"),
           "collective_bytes_per_rank_per_step": (2 * payload * (world - 1) // world) if world > 1 and zero >= 2 else None,
           "model_tflop_per_step_per_gpu": round(flops / 1e12, 1) if flops else None,
           "tflops_per_gpu": round(flops / (ms / 1e3) / 1e12, 1) if flops else None,
           "losses": [round(x, 4) for x in losses], "grad_norm": round(float(nsq[0]) ** 0.5, 4),
           "gpu_launches_per_step": launches,
           "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1)}
    if fam is not None:
        rec["families"] = fam
    if phases is not None:
        rec["phase_ms_serialised"] = phases
    return rec


if __name__ == "__main__":
    print(json.dumps({"note": "use tools/train_bench.py"}))
