#!/usr/bin/env python
"""bench.py -- MM2SG scene-graph inference throughput on B200 (BASELINE.json metric, config 2).

One "step" = one batch of B samples per GPU through the whole hot path: 6 x 336^2 RGB views per sample -> CLIP ViT-L
(23 layers) -> BERT image pooler -> mlp2x_gelu projector -> multimodal token pack -> Llama-7B prefill (L ~ 831)
-> 256 greedy decode steps. Synthetic pixels / token ids, random-init weights of the real architecture.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Prints ONE JSON line (rank 0). Keys: see the bench contract in DESIGN.md ("Measurement").
  value     inferences/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public API (model.generate) with pinned HOST inputs: H2D of pixels and D2H of
            the generated ids inside the timed region
  roofline  dominant kernel family of one extra, event-instrumented step (per-launch CUDA events on the launching
            stream, grouped by family inside libb200mmor.so): algorithmic bytes or FLOPs / summed device time
  cpu_baseline  the CPU oracle (oracle/mm2sg_oracle.py, a port of the reference's PyTorch path) timed on this box's
            host cores on a bounded sample of the same workload (rank 0, N = 1 only)
--impl reference: the oracle port alone, on the host cores (the Python reference cannot travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch

METRIC = "scene-graph inferences/sec (6-view, 7B)"
UNIT = "inferences/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=128, help="samples per GPU per step (KV cache: 0.56 GB per sample)")
    p.add_argument("--views", type=int, default=6)
    p.add_argument("--text-len", type=int, default=256)
    p.add_argument("--new-tokens", type=int, default=256)
    p.add_argument("--jitter", type=int, default=16, help="prompt-length jitter (exercises left padding)")
    p.add_argument("--layers", type=int, default=32, help="debug only: fewer decoder layers => NOT the benchmark")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-roofline", action="store_true")
    p.add_argument("--cpu-decode-steps", type=int, default=4)
    p.add_argument("--no-configs", action="store_true", help="skip the configs[2] / configs[3] sub-records")
    p.add_argument("--subrecord-timeout", type=float, default=300.0, help="seconds each sub-record phase "
                   "(other configs, fine-tune step) may take before the headline line is printed without it")
    p.add_argument("--cfg4-batch", type=int, default=64, help="samples per GPU of the configs[3] sub-record")
    p.add_argument("--config-steps", type=int, default=2)
    p.add_argument("--config-batch", type=int, default=64, help="samples per GPU per step of the configs[2] sub-record")
    p.add_argument("--no-pin", action="store_true", help="N > 1: do not pin every rank to its own slice of host cores")
    p.add_argument("--no-train", action="store_true", help="skip the fine-tune-step sub-record (BASELINE configs[4])")
    p.add_argument("--train-steps", type=int, default=5)
    p.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer arm")
    p.add_argument("--no-graph", action="store_true", help="profiling runs only: eager decode loop (every launch "
                   "visible to ncu)")
    p.add_argument("--pdl", action="store_true", help="programmatic dependent launch for the decode step's kernels "
                   "(b200_set_pdl; off by default until timed on hardware)")
    p.add_argument("--decode-tiles", type=int, default=None, choices=[0, 1, 2], help="weight-tile widths of the "
                   "decode step's wide projections (b200_set_option decode_tiles): 1 (the library default since it was "
                   "timed, profiles/r2_decode_bench.json) = 96 / 160 / 224 columns, one CTA per SM; 2 = 64 / 96 / 128 "
                   "columns, two CTAs per SM; 0 = 128 columns")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
                power.append(float(c[3]))
            except ValueError:
                continue
            for n, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
def full_config(layers=32):
    from mm_or_b200.config import LlavaConfig
    return LlavaConfig(num_hidden_layers=layers, tokenizer_padding_side="left", mv_type="learned")


def make_inputs(cfg, args, seed):
    """Host inputs in the shape ModelWrapper.forward hands to generate() (scene_graph_prediction_model.py:117-231)."""
    from mm_or_b200.synth import synth_batch
    b = synth_batch(cfg, args.batch, args.views, args.text_len, seed=seed, jitter=args.jitter, image_pos=40,
                    dtype=torch.bfloat16)
    images = torch.stack(b["images"]).contiguous()           # (B, V, 3, S, S) bf16
    return images, b["input_ids"]


def algorithmic_flops_per_inference(V, L, n_new):
    """SURVEY.md 8(d) formulae."""
    vit = 23 * 577 * ((4 * 1024 ** 2 + 2 * 1024 * 4096) * 2 + 4 * 577 * 1024) + 576 * 1024 * 588 * 2
    S = V * 576
    pooler = 2 * S * ((4 * 1024 ** 2 + 2 * 1024 * 4096) * 2 + 4 * S * 1024)
    proj = 576 * (1024 * 4096 + 4096 ** 2) * 2
    prefill = 2 * 6.476e9 * L + 32 * 4 * L * L * 4096 / 2 + 2 * 4096 * 32000
    decode = sum(2 * (6.476e9 + 0.131e9) + 32 * 4 * (L + t) * 4096 for t in range(n_new))
    return V * vit + pooler + proj + prefill + decode


def run_b200(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    from mm_or_b200 import _lib as L
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from mm_or_b200.synth import make_state_dict

    pinned_cores = None
    if world > 1 and not args.no_pin:
        # one node, N launcher processes: give every rank its own contiguous slice of the host cores so that the ranks'
        # Python launch loops (encode + prefill are ~2300 launches per batch) do not migrate onto each other's cores
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // world
            if per >= 2:
                mine = cores[local * per:(local + 1) * per]
                os.sched_setaffinity(0, mine)
                torch.set_num_threads(max(1, min(per, 8)))
                pinned_cores = per
        except (AttributeError, OSError):
            pass

    if args.pdl:
        L.set_option("pdl", True)
    if args.decode_tiles is not None:
        L.set_option("decode_tiles", args.decode_tiles)
    cfg = full_config(args.layers)
    t0 = time.time()
    sd = make_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16)
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device=dev)
    del sd
    torch.cuda.empty_cache()
    t_init = time.time() - t0

    images_host, ids = make_inputs(cfg, args, seed=100 + rank)
    images_host = images_host.pin_memory()
    images_dev = images_host.to(dev)
    B = args.batch
    gen_kw = dict(do_sample=False, use_cache=True, max_new_tokens=args.new_tokens, stop_on_eos=False,
                  use_cuda_graph=not args.no_graph)

    exchange = None
    if world > 1:
        # samples sharded over ranks, one exchange step (visual tokens, all-gather over NVLink). Preferred:
        # the projector GEMM's epilogue stores into every rank's symmetric buffer; if symmetric memory cannot be set
        # up on this box, the same exchange runs as ncclAllGather (said so in config.exchange).
        exchange = "peer"
        try:
            model.set_process_group(dist.group.WORLD, exchange="peer")
            model.generate(ids[:, :], images=images_dev, do_sample=False, max_new_tokens=2, stop_on_eos=False)
            ok = torch.ones(1, device=dev)
        except Exception as e:  # noqa: BLE001 - any failure of the peer path selects the NCCL path on ALL ranks
            print(f"[bench] rank {rank}: peer-store exchange unavailable ({e!r}); using ncclAllGather", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            exchange = "nccl"
            model.set_process_group(dist.group.WORLD, exchange="nccl")

    def step_resident():
        return model.generate(ids, images=images_dev, **gen_kw)

    def step_e2e():
        out = model.generate(ids, images=images_host, **gen_kw)
        return out.to("cpu", non_blocking=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        launches = L.launch_count() - n0
        barrier()
        if world > 1:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            per_rank_ms.clear()
            per_rank_ms.extend(round(float(x) / steps, 1) for x in every)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches

    per_rank_ms = []                                     # device ms per step of every rank (the max is the headline)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    rank_ms_resident = list(per_rank_ms)
    ms_e2e = float("nan") if args.no_e2e else timed(step_e2e, args.steps, 1)[0]
    n_inf = B * world * args.steps
    value = n_inf / (ms_total / 1e3)
    e2e_value = n_inf / (ms_e2e / 1e3)

    # ---- event-instrumented step: per-family device time (eager decode loop; launches inside a replayed CUDA graph
    #      cannot carry events)
    roof = None
    L_packed = args.text_len - 1 + 576
    if not args.no_roofline and rank == 0:
        # rank 0 alone runs this extra, event-instrumented step: it must not enter the visual-token exchange (a
        # collective), so the model is taken out of the process group for it -- the kernels timed are the same
        model.set_process_group(None)
        L.prof_enable(True)
        model.generate(ids, images=images_dev, **dict(gen_kw, use_cuda_graph=False))
        fam = L.prof_collect()
        L.prof_enable(False)
        pk = peaks()
        lens = (ids != 0).sum(1) - 1 + 576                                   # packed length per sample
        kv_bytes = sum(float((lens + t + 1).sum()) for t in range(args.new_tokens - 1)) * 2 * 32 * 128 * 2 * args.layers
        fam["decode_attn"]["bytes"] = kv_bytes
        total_ms = sum(f["ms"] for f in fam.values())
        name, top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        tensor_bound = name in ("gemm", "flash_attn")
        n_launch = max(1, top["launches"])
        if tensor_bound:
            ach = top["flops"] / (top["ms"] / 1e3) / 1e12
            peak = pk["tf_sustained"]
            roof = {"bound": "tensor", "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s",
                    "frac": round(ach / peak, 4)}
        else:
            ach = top["bytes"] / (top["ms"] / 1e3) / 1e9
            peak = pk["hbm"]
            roof = {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4)}
        # DRAM bytes per launch: the ncu capture's measured / algorithmic ratio (taken at the capture's batch size)
        # applied to this run's algorithmic bytes per launch
        cap = measured_traffic(name)
        traffic = None
        if cap is not None and not tensor_bound:
            traffic = round(cap["ratio"] * top["bytes"] / n_launch)
        elif cap is not None:
            traffic = round(cap["dram_bytes_per_launch"])
        roof.update({"kernel": name, "traffic": traffic, "traffic_capture": cap, "peak_source": pk["source"],
                     "launches": top["launches"], "avg_launch_us": round(top["ms"] * 1e3 / n_launch, 2),
                     "share_of_step": round(top["ms"] / total_ms, 4),
                     "families": {k: {"ms": round(v["ms"], 2), "launches": v["launches"],
                                      "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1) if v["flops"] else None,
                                      "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1) if v["bytes"] else None}
                                  for k, v in fam.items() if v["launches"]}})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_sample(args, layers=args.layers)

    # ---- the headline line is complete here; what follows can only ADD sub-records to it
    line = None
    if rank == 0:
        flops = algorithmic_flops_per_inference(args.views, L_packed, args.new_tokens)
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: 6-view 336x336 RGB -> CLIP ViT-L/14 (23 layers) + BERT pooler + "
                                   "mlp2x_gelu projector + Llama-7B prefill + %d-token greedy decode" % args.new_tokens,
                       "samples_per_gpu_per_step": B, "views": args.views, "text_tokens": args.text_len,
                       "packed_len": L_packed, "new_tokens": args.new_tokens, "decoder_layers": args.layers,
                       "weights": "random-init, seed 0, bf16", "l2": "inputs larger than L2 (13.5 GB of weights + "
                       "KV cache streamed every decode step)", "parallelism": "dp%d" % world,
                       "exchange": {None: "none (1 GPU)", "peer": "visual tokens all-gathered by the projector GEMM "
                                    "epilogue (peer stores over NVLink)", "nccl": "ncclAllGather of visual tokens"}[exchange],
                       "model_tflop_per_inference": round(flops / 1e12, 2),
                       "programmatic_dependent_launch": bool(L.get_option("pdl")),
                       "decode_tile_widths": {0: "128", 1: "96/160/224, one CTA per SM",
                                              2: "64/96/128, two CTAs per SM"}[int(L.get_option("decode_tiles"))]},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(images_host.numel() * 2
                    + ids.numel() * 8), "d2h_bytes_per_step": int(B * (ids.shape[1] + args.new_tokens) * 8),
                    "ms_per_step": round(ms_e2e / args.steps, 2)},
            "gpu_launches": int(launches), "clocks": clocks,
            "tensor_frac_whole_step": round(flops * n_inf / (ms_total / 1e3) / 1e12 / world / peaks()["tf_sustained"], 4),
            "init_s": round(t_init, 1),
        }
        if rank_ms_resident:
            line["per_rank_ms_per_step"] = rank_ms_resident      # weak scaling: the slowest rank sets `value`
            line["host_cores_per_rank"] = pinned_cores
        if roof is not None:
            line["roofline"] = roof
        if cpu is not None:
            line["cpu_baseline"] = cpu

    # Sub-records run under a watchdog: if one of them does not finish within its budget (a hung collective cannot be
    # interrupted from Python), rank 0 prints the headline line as it stands -- with the phase that timed out named in
    # it -- and every rank leaves with exit code 0. The sub-records can never cost the driver its headline number.
    dog = SubRecordWatchdog(rank, line)
    progress = lambda msg: print("[bench %.0fs] rank %d: %s" % (time.time() - t0, rank, msg), file=sys.stderr, flush=True)

    # ---- BASELINE configs[2] and configs[3] at the same N GPUs: short device-timed regions of their own (1 warm-up +
    #      `--config-steps` steps each, max over ranks), reported as sub-records next to the headline configs[1] line
    if not args.no_configs:
        dog.arm("other_configs", args.subrecord_timeout)
        progress("configs[2] / configs[3] sub-records")
        if world > 1:
            model.set_process_group(dist.group.WORLD, exchange=exchange)
        other = other_config_records(args, cfg, model, dev, rank, world, exchange, timed, progress)
        dog.disarm()
        if line is not None:
            line["other_configs"] = other

    # ---- BASELINE configs[4]: the fine-tune step at the same N GPUs, AFTER the timed inference region (its own
    #      device-timed region, max over ranks). The inference model and its buffers are released first.
    if not args.no_train:
        dog.arm("train", args.subrecord_timeout)
        progress("fine-tune sub-record")
        model.set_process_group(None)
        del model, images_dev
        import gc as _gc
        _gc.collect()
        torch.cuda.empty_cache()
        train = finetune_record(args, cfg, dev, dist.group.WORLD if world > 1 else None)
        dog.disarm()
        if line is not None:
            line["train"] = train
    progress("done")

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dog.arm("destroy_process_group", 60, print_line=False)
        dist.destroy_process_group()
        dog.disarm()


class SubRecordWatchdog:
    """Deadline for the optional sub-records of the bench line (see run_b200). arm(phase, seconds) starts the clock;
    when it runs out rank 0 prints the line built so far with "sub_records_timed_out": phase and every rank calls
    os._exit(0) -- the only way out of a hung collective."""

    def __init__(self, rank, line):
        import threading
        self.rank, self.line = rank, line
        self.lock = threading.Lock()
        self.deadline = None
        self.phase = None
        self.print_line = True
        t = threading.Thread(target=self._run, daemon=True)
        t.start()

    def arm(self, phase, seconds, print_line=True):
        with self.lock:
            self.phase, self.deadline, self.print_line = phase, time.time() + float(seconds), print_line

    def disarm(self):
        with self.lock:
            self.deadline = None

    def _run(self):
        while True:
            time.sleep(1.0)
            with self.lock:
                expired = self.deadline is not None and time.time() > self.deadline
                phase, print_line = self.phase, self.print_line
            if expired:
                print("[bench] rank %d: sub-record phase %r exceeded its time budget; leaving" % (self.rank, phase),
                      file=sys.stderr, flush=True)
                if self.rank == 0 and self.line is not None and print_line:
                    self.line["sub_records_timed_out"] = phase
                    print(json.dumps(self.line), flush=True)
                os._exit(0)


def other_config_records(args, cfg, model, dev, rank, world, exchange, timed, progress=lambda m: None):
    """configs[2]: per sample 6 RGB views + 6 depth + 6 seg-mask renderings = 18 encoder passes, an audio embedding and
    three 32x32 class maps -> fused token pack with T_vis = 580, prefill, 256 greedy tokens. Mapping as in SURVEY.md 8d /
    tests/test_gpu_zzz_baseline_configs.py: the reference has no depth / seg-mask IMAGE input, so the 12 extra frames go
    through the tower and, as two more 6-view groups, through the pooler (pure throughput), while the RGB group, the
    audio token and the class-map tokens feed the decoder.
    configs[3]: 64 clips x 8 frames x 6 views over 8 GPUs = 64 samples + 384 images per GPU; the visual tokens are
    all-gathered and every rank decodes the slice its NEIGHBOUR encoded (decode_shift = 1), so the exchange is
    load-bearing. On fewer GPUs the per-GPU geometry is kept (weak scaling)."""
    import torch.distributed as dist
    from mm_or_b200.synth import synth_batch
    out = {}
    steps = max(1, args.config_steps)
    gen_kw = dict(do_sample=False, use_cache=True, max_new_tokens=args.new_tokens, stop_on_eos=False)

    def record(name, fn, samples, note):
        try:
            ms, _ = timed(fn, steps, 1)
        except Exception as e:  # noqa: BLE001 -- a sub-record never takes the headline down
            out[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
            return
        out[name] = {"value": round(samples * world * steps / (ms / 1e3), 3), "unit": UNIT, "n_gpus": world,
                     "samples_per_gpu_per_step": samples, "steps": steps, "ms_per_step": round(ms / steps, 1),
                     "workload": note}

    # configs[2]
    B3 = args.config_batch
    b = synth_batch(cfg, B3, args.views, args.text_len, seed=300 + rank, jitter=args.jitter, image_pos=40, audio=True,
                    segmasks=True, dtype=torch.bfloat16)
    rgb = torch.stack(b["images"]).to(dev)
    extra = synth_batch(cfg, B3, 2 * args.views, 8, seed=400 + rank, dtype=torch.bfloat16)["images"]
    extra = torch.stack(extra).to(dev).flatten(0, 1)                       # (B * 12, 3, S, S): depth + seg renderings

    def cfg3_step():
        model.encode_images_pooled(extra, [args.views] * (2 * B3), None, None, None)
        return model.generate(b["input_ids"], images=rgb, audio=b["audio"], segmasks=b["segmasks"], **gen_kw)

    progress("configs[2]: inputs ready, timing")
    record("configs[2]", cfg3_step, B3,
           "6-view RGB + depth + seg-mask renderings (18 ViT passes, 3 pooler passes per sample) + audio token + 3 "
           "class-map tokens, T_vis = 580, prefill + %d greedy tokens" % args.new_tokens)
    del rgb, extra, b
    torch.cuda.empty_cache()

    # configs[3]: 64 samples + 384 images per GPU, neighbour's slice decoded
    B4 = args.cfg4_batch
    shift = 1 if world > 1 else 0
    if world > 1:
        model.set_process_group(dist.group.WORLD, exchange=exchange, decode_shift=shift)
    mine = synth_batch(cfg, B4, args.views, args.text_len, seed=500 + rank, jitter=0, image_pos=40, dtype=torch.bfloat16)
    owner = (rank + shift) % world
    theirs = mine if owner == rank else synth_batch(cfg, B4, args.views, args.text_len, seed=500 + owner, jitter=0,
                                                    image_pos=40, dtype=torch.bfloat16)
    img4 = torch.stack(mine["images"]).to(dev)
    progress("configs[3]: inputs ready, timing")
    record("configs[3]", lambda: model.generate(theirs["input_ids"], images=img4, **gen_kw), B4,
           "temporal clips: 64 samples (8 clips x 8 frames) x 6 views = 384 images per GPU; visual tokens all-gathered "
           "(%s), every rank decodes the slice rank + %d encoded; %d greedy tokens"
           % (exchange or "single GPU: no exchange", shift, args.new_tokens))
    if world > 1:
        model.set_process_group(dist.group.WORLD, exchange=exchange, decode_shift=0)
    return out


def finetune_record(args, cfg, dev, group):
    """The `train` sub-record: mm_or_b200/train/bench_step.py on this rank's GPU (ZeRO-2 over `group` when N > 1).
    On one GPU the full fine-tune holds 12 B of fp32 state + 4 B of fp32 gradient per parameter (~150 GB of 180): if the
    step with stored activations runs out of memory it is repeated with activation recomputation (the reference's own
    gradient checkpointing). A failure is reported in the record, it never takes the inference line down."""
    from mm_or_b200.train.bench_step import measure_finetune_step
    import gc as _gc
    import torch.distributed as dist
    world = dist.get_world_size(group) if group is not None else 1
    last = None
    for recompute in (False, True):
        ok = torch.ones(1, device=dev)
        rec = None
        try:
            rec = measure_finetune_step(cfg, dev, group=group, batch=4, views=args.views, steps=args.train_steps,
                                        warmup=2, zero=2, recompute=recompute)
        except torch.cuda.OutOfMemoryError as e:
            last = "out of memory with activation_recomputation=%s: %s" % (recompute, str(e)[:120])
            ok.zero_()
        except Exception as e:  # noqa: BLE001 -- reported, not raised: the inference numbers above stand on their own
            return {"error": repr(e)[:300]}
        if group is not None:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all ranks retry together
        if ok.item() == 1:
            pk = peaks()
            if rec.get("tflops_per_gpu"):
                rec["frac_of_sustained_bf16"] = round(rec["tflops_per_gpu"] / pk["tf_sustained"], 4)
            if world > 1 and rec.get("collective_bytes_per_rank_per_step"):
                rec["note"] = ("gradients reduce-scattered and updated slices all-gathered in bf16: %d MB per rank per "
                               "step over NVLink" % (rec["collective_bytes_per_rank_per_step"] // (1 << 20)))
            return rec
        rec = None
        _gc.collect()
        torch.cuda.empty_cache()
    return {"error": last}


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` next to its algorithmic bytes, from the committed `ncu --set full` capture
    (profiles/traffic.json; taken at the capture's batch size, so compare the ratio), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's PyTorch path, on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_weights(cfg):
    """bf16 weights for TIMING the CPU path: every tensor is a window of one N(0, 0.02^2) pool (gains = 1 + pool), so
    the 7B parameter set is materialised at memcpy speed; values do not influence the instruction stream."""
    from mm_or_b200.synth import weight_specs
    g = torch.Generator().manual_seed(0)
    pool = (torch.randn(1 << 26, generator=g) * 0.02).to(torch.bfloat16)
    sd, off = {}, 0
    for name, shape, kind in weight_specs(cfg):
        n = 1
        for s in shape:
            n *= s
        reps = -(-n // pool.numel())
        src = pool if reps == 1 else pool.repeat(reps)
        start = off % max(1, src.numel() - n)
        t = src[start:start + n].clone().view(shape)
        if kind == "g":
            t = t + 1
        sd[name] = t
        off += 7919 * 64
    return sd


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def cpu_reference_sample(args, layers=32):
    """Times ONE inference of the config-2 workload on the CPU oracle: encode (6 views) + pack + prefill measured in
    full, decode measured on `cpu_decode_steps` steps and extrapolated linearly to new_tokens (each step streams the
    same 13.2 GB of weights; KV growth changes a CPU step by < 1 %)."""
    from oracle import mm2sg_oracle as O
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = full_config(layers)
    ocfg = O.cfg_from_llava(cfg)
    sd = cpu_weights(cfg)
    sub = argparse.Namespace(**vars(args))
    sub.batch = 1
    images, ids = make_inputs(cfg, sub, seed=100)
    images = [images[0]]
    t0 = time.perf_counter()
    visual = O.encode_images_pooled(sd, images, ocfg)
    t_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    src, _, mask, pos = O.pack_plan(ids, ids.ne(0), None, visual.shape[1], "left", None)
    emb = O.pack_embeds(sd, src, visual)
    logits, kv = O.llama_forward(sd, emb, mask, pos, ocfg.llm, last_only=True)
    t_pre = time.perf_counter() - t0
    n = max(1, min(args.cpu_decode_steps, args.new_tokens - 1))
    t0 = time.perf_counter()
    O.greedy_decode(sd, ocfg, logits[:, -1], kv, mask, n + 1, stop_on_eos=False, logits_dtype=torch.bfloat16)
    t_dec = (time.perf_counter() - t0) / n
    total = t_enc + t_pre + t_dec * (args.new_tokens - 1)
    full = n == args.new_tokens - 1
    how = ("all %d decode steps measured (%.3fs/step): the whole inference is timed, nothing extrapolated"
           % (n, t_dec)) if full else ("%d decode steps measured (%.3fs/step), extrapolated to %d tokens"
                                       % (n, t_dec, args.new_tokens))
    return {"value": round(1.0 / total, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "1 inference (%d views, L=%d): encode %.1fs + prefill %.1fs measured in full; %s; torch %s CPU "
                      "bf16, %d threads on %s"
                      % (args.views, emb.shape[1], t_enc, t_pre, how, torch.__version__, cores, cpu_model()),
            "extrapolated": not full, "cpu_model": cpu_model(),
            "seconds_per_inference": round(total, 1),
            "stage_seconds": {"vision_tower_pooler_projector": round(t_enc, 2), "pack_prefill": round(t_pre, 2),
                              "decode_per_step": round(t_dec, 4)},
            "decode_tokens_per_s": round(1.0 / t_dec, 3), "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    # one sample per step would take minutes of CPU time; weights are built once, each "step" is one bounded sample
    # The reference arm measures the WHOLE inference (all new_tokens - 1 decode steps, ~40 s on 16 cores); at most two
    # repetitions so that the arm ends within a few minutes whatever --steps says.
    vals, last = [], None
    reps = max(1, min(args.steps, 2))
    args.cpu_decode_steps = args.new_tokens - 1
    for i in range(reps):
        last = cpu_reference_sample(args, layers=args.layers)
        vals.append(last["value"])
    v = max(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": reps,
            "warmup": 0, "ms_per_step": round(1e3 / v, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            # the same workload as the B200 arm's config (one inference per step instead of 128 per GPU)
            "config": {"workload": "configs[1]: 6-view 336x336 RGB -> CLIP ViT-L/14 (23 layers) + BERT pooler + "
                                   "mlp2x_gelu projector + Llama-7B prefill + %d-token greedy decode" % args.new_tokens,
                       "samples_per_step": 1, "views": args.views, "text_tokens": args.text_len,
                       "packed_len": args.text_len - 1 + 576, "new_tokens": args.new_tokens,
                       "decoder_layers": args.layers, "weights": "random-init, bf16",
                       "arm": "host cores, CPU oracle (port of the reference's PyTorch path), whole inference timed, "
                              "see cpu_baseline.sample"},
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse_args()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
