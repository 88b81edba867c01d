"""Fine-tune step of BASELINE.json configs[4] on ONE B200 at the real model size (secondary measurement; the headline
bench is bench.py): per-GPU batch 4 (the reference's per_device_train_batch_size), 6 views, prompt 256 ids + 150 answer
ids, right padded, labels on the answer only, class-weighted CE, full decoder + projector + pooler + CLIP layers 12..22
trainable, fp32 master weights, clip 0.1, AdamW lr 2e-5. Reports trained tokens/s (all L tokens counted).

  python tools/train_bench.py [--layers 32] [--batch 4] [--steps 2]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--lora-r", type=int, default=0, help="> 0: the reference's LoRA recipe (r 128, alpha 2r) instead "
                    "of full fine-tuning of the decoder")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", 0)
    cfg = LlavaConfig(num_hidden_layers=a.layers, tokenizer_padding_side="right", mv_type="learned")
    sd = make_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16)
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device=dev)
    b = synth_batch(cfg, a.batch, a.views, 256 + 150, seed=3, jitter=0, image_pos=40, dtype=torch.bfloat16)
    ids = b["input_ids"]
    labels = ids.clone()
    labels[:, :256] = -100
    labels[ids == -200] = -100
    g = torch.Generator().manual_seed(1)
    w = torch.rand(cfg.vocab_size, generator=g) + 0.01
    lora = None
    if a.lora_r > 0:
        from mm_or_b200.train.lora import LoraState
        lora = LoraState(cfg, r=a.lora_r, alpha=2 * a.lora_r, device=dev)
    ft = FineTuner(model, sd, lr=2e-5, weight_decay=0.0, max_grad_norm=0.1, first_trainable_clip_layer=12, vocab_weight=w,
                   lora=lora)
    del sd
    n_train = sum(v.numel() for v in ft.master.values())
    tokens = a.batch * (256 + 150 - 1 + 576)
    losses = []
    for _ in range(a.warmup):
        loss, _ = ft.train_step(ids, labels, b["attention_mask"], b["images"])
        losses.append(float(loss))
    torch.cuda.synchronize()
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, nsq = ft.train_step(ids, labels, b["attention_mask"], b["images"])
        losses.append(float(loss))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"metric": "fine-tune step, trained tokens/s (1 GPU)", "value": round(tokens / (ms / 1e3), 1),
                      "unit": "tokens/s", "ms_per_step": round(ms, 1), "tokens_per_step": tokens,
                      "trainable_params": n_train, "decoder_layers": a.layers, "batch": a.batch, "views": a.views,
                      "mode": "lora r=%d" % a.lora_r if a.lora_r else "full fine-tune",
                      "losses": [round(x, 4) for x in losses], "grad_norm": round(float(nsq[0]) ** 0.5, 4),
                      "gpu_launches_per_step": int((L.launch_count() - n0) / a.steps),
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
