"""Decode-step timing on ONE B200 at the real 7B size, for A/B-ing the decode-step switches quickly (a full bench.py
run costs a minute per arm): text-only prompts of the packed length (831), `--batch` rows, the captured CUDA-graph decode
loop; ms per step = (time of a generate with N2 new tokens - time with N1) / (N2 - N1), CUDA events, 3 repetitions,
minimum. Prints one JSON line per arm: none / pdl / tiles1 / tiles2 / pdl+tiles1 / pdl+tiles2 (b200_set_option), with the implied HBM rate of
the weight stream + KV reads per step (13.21 GB of weights + 512 KiB per context token per row).

  python tools/decode_bench.py [--batch 128] [--ctx 831] [--layers 32] [--arms none,pdl,tiles1,tiles2,pdl+tiles2]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mm_or_b200 import _lib as L
from mm_or_b200.config import LlavaConfig
from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
from mm_or_b200.synth import make_state_dict


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None else min(best, t)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--ctx", type=int, default=831)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--n1", type=int, default=4)
    ap.add_argument("--n2", type=int, default=68)
    ap.add_argument("--arms", default="none,pdl,tiles1,tiles2,pdl+tiles1,pdl+tiles2")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    torch.set_grad_enabled(False)
    cfg = LlavaConfig(num_hidden_layers=a.layers, tokenizer_padding_side="left", mv_type="learned")
    sd = make_state_dict(cfg, seed=0, device="cuda", dtype=torch.bfloat16)
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device="cuda")
    del sd
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(3, cfg.vocab_size, (a.batch, a.ctx), generator=g)
    kw = dict(do_sample=False, use_cache=True, stop_on_eos=False)
    weights_gb = 13.21 * a.layers / 32 if a.layers != 32 else 13.21
    for arm in a.arms.split(","):
        pdl, tiles = "pdl" in arm, 2 if "tiles2" in arm else (1 if "tiles1" in arm else 0)
        L.set_option("pdl", pdl)
        L.set_option("decode_tiles", tiles)
        L.set_option("fused_rope", 0 if "nofuse" in arm else 1)      # RoPE + KV append inside the attention kernel
        try:
            model.generate(ids, max_new_tokens=a.n1, **kw)                       # warm-up (lazy kernel attributes)
            t1 = timed(lambda: model.generate(ids, max_new_tokens=a.n1, **kw))
            t2 = timed(lambda: model.generate(ids, max_new_tokens=a.n2, **kw))
            ms = (t2 - t1) / (a.n2 - a.n1)
            mid_ctx = a.ctx + (a.n1 + a.n2) / 2
            gb = weights_gb + a.batch * mid_ctx * a.layers * 16384 / 1e9
            print(json.dumps({"arm": arm, "batch": a.batch, "ctx": a.ctx, "layers": a.layers,
                              "ms_per_decode_step": round(ms, 3), "algorithmic_gb_per_step": round(gb, 2),
                              "implied_gb_per_s": round(gb / (ms / 1e3), 1),
                              "prefill_plus_%d_steps_ms" % a.n1: round(t1, 1)}), flush=True)
        except Exception as e:  # noqa: BLE001 -- an arm that fails must not hide the others
            print(json.dumps({"arm": arm, "error": repr(e)[:300]}), flush=True)
    L.set_option("pdl", 0)
    L.set_option("decode_tiles", 1)
    L.set_option("fused_rope", 1)


if __name__ == "__main__":
    main()
