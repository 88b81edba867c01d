"""Fine-tune step of BASELINE.json configs[4] on ONE B200 at the real model size (secondary measurement; the headline
bench is bench.py): per-GPU batch 4 (the reference's per_device_train_batch_size), 6 views, prompt 256 ids + 150 answer
ids, right padded, labels on the answer only, class-weighted CE, full decoder + projector + pooler + CLIP layers 12..22
trainable, fp32 master weights, clip 0.1, AdamW lr 2e-5. Reports trained tokens/s (all L tokens counted).

  python tools/train_bench.py [--layers 32] [--batch 4] [--steps 2] [--lora-r 128] [--accum 4]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py --zero 2   (weak scaling)
  python tools/train_bench.py --cpu-reference        (the same step on the host cores through the CPU oracle)
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mm_or_b200 import _lib as L
from mm_or_b200.config import LlavaConfig
from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
from mm_or_b200.synth import make_state_dict, synth_batch
from mm_or_b200.train.step import FineTuner


def cpu_reference(a):
    """BASELINE configs[4] asks for tokens/s "vs reference CPU": the oracle (the port of the reference's PyTorch path)
    under torch autograd on the host cores, ONE sample of the same shape (6 views, 405 text + 576 visual tokens),
    encoder + 1 and + 2 decoder layers measured in full (forward, weighted CE, backward), the per-layer cost taken from
    the difference and extrapolated to `--layers`; optimizer step not included (it would add to the CPU time)."""
    from oracle import mm2sg_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    times = {}
    for n_layers in (1, 2):
        cfg = LlavaConfig(num_hidden_layers=n_layers, tokenizer_padding_side="right", mv_type="learned")
        sd = make_state_dict(cfg, seed=0, device="cpu", dtype=torch.bfloat16)
        sd = {k: v.float() for k, v in sd.items()}
        vit = "model.vision_tower.vision_tower.vision_model.encoder.layers."
        trainable = [k for k in sd if k.startswith(("model.layers.", "model.mm_projector.", "model.image_pooler.bert.",
                                                    "lm_head.", "model.norm."))
                     or (k.startswith(vit) and 12 <= int(k[len(vit):].split(".")[0]) <= 22)]     # train.py:1257-1261
        b = synth_batch(cfg, 1, a.views, 256 + 150, seed=3, jitter=0, image_pos=40)
        ids = b["input_ids"]
        labels = ids.clone()
        labels[:, :256] = -100
        labels[ids == -200] = -100
        w = torch.rand(cfg.vocab_size, generator=torch.Generator().manual_seed(1)) + 0.01
        best = None
        for rep in range(2):                      # first repetition warms the allocator / thread pool
            with torch.enable_grad():
                params = {k: sd[k].clone().requires_grad_(True) for k in trainable}
                t0 = time.perf_counter()
                ref = O.multimodal_prefill({**sd, **params}, O.cfg_from_llava(cfg), ids, b["attention_mask"],
                                           b["images"], labels=labels, padding_side="right")
                loss = O.weighted_ce(ref["logits"], ref["modified_labels"], w)
                loss.backward()
                dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            del params, ref, loss
        times[n_layers] = best
    per_layer = max(times[2] - times[1], 1e-9)
    total = times[1] + per_layer * (a.layers - 1)
    tokens = 256 + 150 - 1 + 576
    return {"value": round(tokens / total, 2), "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": "1 sample (%d views, %d tokens), fp32, forward + weighted CE + backward through the CPU oracle "
                      "under torch autograd: %.1fs with 1 decoder layer, %.1fs with 2 => %.2fs per layer, extrapolated "
                      "to %d layers (no optimizer step); torch %s, %d threads"
                      % (a.views, tokens, times[1], times[2], per_layer, a.layers, torch.__version__, cores),
            "seconds_per_sample": round(total, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--lora-r", type=int, default=0, help="> 0: the reference's LoRA recipe (r 128, alpha 2r) instead "
                    "of full fine-tuning of the decoder")
    ap.add_argument("--nf4", action="store_true", help="with --lora-r: NF4 round trip of the frozen base (QLoRA, --bits 4)")
    ap.add_argument("--zero", type=int, default=0, choices=[0, 1, 2], help="under torchrun: 0 = replicated optimizer "
                    "state + fp32 all-reduce, 1 = sharded state, 2 = sharded state + bf16 reduce-scatter (zero.py)")
    ap.add_argument("--accum", type=int, default=1, help="gradient accumulation steps (reference recipe: 4)")
    ap.add_argument("--recompute", action="store_true", help="activation recomputation per decoder layer (the "
                    "reference's gradient checkpointing, train.py:1148)")
    ap.add_argument("--trace", action="store_true", help="one extra step with device-synchronised phase timing")
    ap.add_argument("--prof", action="store_true", help="one extra event-instrumented step: device time per kernel family")
    ap.add_argument("--cpu-reference", action="store_true", help="time the CPU oracle's forward + backward instead "
                    "(bounded sample, see cpu_reference)")
    a = ap.parse_args()
    if a.cpu_reference:
        print(json.dumps({"impl": "reference", "metric": "fine-tune step, trained tokens/s", **cpu_reference(a)}),
              flush=True)
        return
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    from mm_or_b200.train.bench_step import measure_finetune_step
    cfg = LlavaConfig(num_hidden_layers=a.layers, tokenizer_padding_side="right", mv_type="learned")
    rec = measure_finetune_step(cfg, dev, group=group, batch=a.batch, views=a.views, steps=a.steps, warmup=a.warmup,
                                zero=a.zero if world > 1 else 0, lora_r=a.lora_r, nf4=a.nf4, accum=a.accum,
                                recompute=a.recompute, prof=a.prof, trace=a.trace)
    if rank == 0:
        rec["metric"] = "fine-tune step, trained tokens/s (%d GPU%s)" % (world, "s" if world > 1 else "")
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
