"""Serial generate() vs pipelined generate_stream() on ONE B200 at the real configs[1] size: K batches of `--batch`
samples (6 views, L = 831, 256 new tokens). Serial = K x generate(); pipelined = encode + prefill of batch n + 1 on a
low-priority stream while batch n decodes. Prints one JSON line per arm (ms per batch, inferences/s) for every chunk
configuration given as --chunks "prefill,vit,pooler;...".

  python tools/overlap_bench.py [--batch 128] [--batches 4] [--chunks "16,96,16;4,24,4;2,12,2"]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mm_or_b200.config import LlavaConfig
from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
from mm_or_b200.synth import make_state_dict, synth_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--new-tokens", type=int, default=256)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--chunks", default="16,96,16;4,24,4;2,12,2")
    ap.add_argument("--no-serial", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    torch.set_grad_enabled(False)
    cfg = LlavaConfig(num_hidden_layers=a.layers, tokenizer_padding_side="left", mv_type="learned")
    sd = make_state_dict(cfg, seed=0, device="cuda", dtype=torch.bfloat16)
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device="cuda")
    del sd
    torch.cuda.empty_cache()
    reqs = []
    for i in range(a.batches):
        b = synth_batch(cfg, a.batch, 6, 256, seed=100 + i, jitter=16, image_pos=40, dtype=torch.bfloat16)
        reqs.append(dict(input_ids=b["input_ids"], images=torch.stack(b["images"]).contiguous().pin_memory()))
    kw = dict(max_new_tokens=a.new_tokens, stop_on_eos=False)

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out

    ref = None
    if not a.no_serial:
        model.generate(**reqs[0], **kw)                                             # warm-up
        ms, ref = timed(lambda: [model.generate(**r, **kw) for r in reqs])
        print(json.dumps({"arm": "serial generate()", "batch": a.batch, "batches": a.batches,
                          "ms_per_batch": round(ms / a.batches, 1),
                          "inferences_per_s": round(a.batch * a.batches / (ms / 1e3), 2)}), flush=True)
    for spec in a.chunks.split(";"):
        pc, vc, oc = (int(x) for x in spec.split(","))
        try:
            ckw = dict(prefill_chunk=pc, vit_chunk=vc, pooler_chunk=oc)
            list(model.generate_stream(reqs[:2], **kw, **ckw))                      # warm-up: slots + graphs
            ms, out = timed(lambda: list(model.generate_stream(reqs, **kw, **ckw)))
            same = None if ref is None else all(torch.equal(x, y) for x, y in zip(out, ref))
            print(json.dumps({"arm": "generate_stream", "prefill_chunk": pc, "vit_chunk": vc, "pooler_chunk": oc,
                              "batch": a.batch, "batches": a.batches, "ms_per_batch": round(ms / a.batches, 1),
                              "inferences_per_s": round(a.batch * a.batches / (ms / 1e3), 2),
                              "ids_equal_serial": same,
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)}), flush=True)
        except Exception as e:  # noqa: BLE001 -- an arm that fails must not hide the others
            print(json.dumps({"arm": "generate_stream", "chunks": spec, "error": repr(e)[:300]}), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
