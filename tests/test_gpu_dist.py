"""2-GPU parity test of the sharded inference path: every rank encodes its own samples, the projected
visual tokens are exchanged (fused peer-store epilogue of the projector GEMM, or ncclAllGather), and each rank
decodes its NEIGHBOUR's samples (decode_shift = 1) so that a wrong or missing exchange cannot pass. The result must
equal, token for token and logit for logit, what a single GPU produces for the same samples (same kernels, same
bits). Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_cases as gc

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_case(cfg, rank):
    """Per-rank batch with the SAME geometry on every rank (the gathered buffer is (world, B, T_vis, D))."""
    from mm_or_b200.synth import synth_batch
    return synth_batch(cfg, 2, 2, 24, seed=50 + rank, jitter=0, image_pos=5, audio=False, segmasks=False,
                       dtype=torch.float32)


def _worker(rank, world, port, exchange, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        torch.set_grad_enabled(False)
        from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
        cfg = gc.small_config()
        cfg.tokenizer_padding_side = "left"
        sd = gc.bf16_round(gc.small_weights(cfg))
        model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device=dev)
        mine, theirs = _rank_case(cfg, rank), _rank_case(cfg, (rank + 1) % world)
        # single-GPU result for the neighbour's samples
        ref_ids, ref_lg = model.generate(theirs["input_ids"], images=theirs["images"], max_new_tokens=5,
                                         stop_on_eos=False, return_logits=True)
        model.set_process_group(dist.group.WORLD, exchange=exchange, decode_shift=1)
        for _ in range(2):      # twice: the symmetric buffer is reused across steps
            out_ids, out_lg = model.generate(theirs["input_ids"], images=mine["images"], max_new_tokens=5,
                                             stop_on_eos=False, return_logits=True)
        torch.cuda.synchronize()
        ok = torch.equal(out_ids, ref_ids) and torch.equal(out_lg, ref_lg)
        # another batch geometry on the same model (one sample per rank): the symmetric buffer is re-viewed, not
        # re-allocated (dist.PeerGather.set_shape) -- a second symmetric allocation next to a live one hung the bench's
        # sub-records on 2 and 8 GPUs
        model.set_process_group(None)
        ref1, ref1_lg = model.generate(theirs["input_ids"][:1], images=theirs["images"][:1], max_new_tokens=4,
                                       stop_on_eos=False, return_logits=True)
        model.set_process_group(dist.group.WORLD, exchange=exchange, decode_shift=1)
        out1, out1_lg = model.generate(theirs["input_ids"][:1], images=mine["images"][:1], max_new_tokens=4,
                                       stop_on_eos=False, return_logits=True)
        torch.cuda.synchronize()
        ok = ok and torch.equal(out1, ref1) and torch.equal(out1_lg, ref1_lg)
        q.put((rank, "ok" if ok else "mismatch: max |dlogit| %g" % (out_lg - ref_lg).abs().max().item()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:
        import traceback
        q.put((rank, "error: " + repr(e) + traceback.format_exc()[-1500:]))


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_sharded_generate_matches_single_gpu(exchange):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=280) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def _train_worker(rank, world, port, q):
    """Data-parallel fine-tune step, replicated vs sharded optimizer state: same working weights bit for bit."""
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
        from mm_or_b200.train.step import FineTuner
        cfg = gc.small_config()
        cfg.tokenizer_padding_side = "right"
        sd = gc.bf16_round(gc.small_weights(cfg))
        case = gc.make_case(cfg, "train_extras_right")
        # rank 1's batch carries no audio: its audio-projection gradient must arrive as zeros (align_optional_gradients)
        kw = dict(audio=case["audio"] if rank == 0 else None, segmasks=case["segmasks"])
        results = []
        for shard in (False, True, "zero2"):
            model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device=dev)
            ft = FineTuner(model, sd, lr=1e-3, max_grad_norm=0.1, first_trainable_clip_layer=1,
                           group=dist.group.WORLD, shard_optimizer=bool(shard), shard_gradients=shard == "zero2",
                           grad_comm_dtype=torch.float32)
            for _ in range(2):
                ft.train_step(case["input_ids"], case["labels"], case["attention_mask"], case["images"], **kw)
            torch.cuda.synchronize()
            results.append({k: ft.sd[k].clone() for k in ft.names})
            if shard:
                full = sum(ft.sd[k].numel() for k in ft.names)
                held = sum(t.numel() for t in ft.master.values())
                assert held <= full // world + len(ft.names) * world, (held, full)
        same = all(torch.equal(results[0][k], results[1][k]) for k in results[0])
        # ZeRO-2 with an fp32 wire: the same sums for two ranks, but the clipping norm is assembled from the slices
        # (another summation order), so the weights may differ from the all-reduce path in the last bf16 bit
        for k in results[0]:
            a, b = results[0][k].float(), results[2][k].float()
            same = same and bool(((a - b).abs() <= 2.0 ** -7 * a.abs().clamp_min(1e-3)).all())
        # replicas agree across ranks
        for res in results[1:]:
            probe = res["model.mm_projector.2.weight"].float()
            other = probe.clone()
            dist.all_reduce(other, op=dist.ReduceOp.MAX)
            same = same and torch.equal(other, probe)
        q.put((rank, "ok" if same else "sharded and replicated optimizer states disagree"))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:
        import traceback
        q.put((rank, "error: " + repr(e) + traceback.format_exc()[-1500:]))


def test_sharded_optimizer_matches_replicated_2gpu():
    """Not yet run on hardware (written after the round's GPU budget was spent; host logic covered by the gloo test
    tests/test_zero_host.py)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=280) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
