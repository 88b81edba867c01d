"""GPU: the serving side of the online / temporal mode on the real model path -- token-level
continuous batching (serving/continuous.py) against per-request generate(), and the cross-take scheduler
(serving/scheduler.py: rounds and continuous) against the reference's take-by-take, batch-1 order.

Weight set `chain` (tests/golden_cases.py): greedy continuations walk through distinct ids with top-2 margins > 3.5
logits, so 'equal up to bf16 rounding of different split points' becomes 'token-exact' and the comparisons are strict."""
import pytest
import torch

import golden_cases as gc
from mm_or_b200.synth import chain_successor, synth_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    torch.set_grad_enabled(False)
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    cfg = gc.small_config()
    cfg.tokenizer_padding_side = "left"
    sd = gc.bf16_round(gc.small_weights(cfg, chain=True))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    return cfg, model


def _requests(cfg, n, seed=300):
    """n single-sample requests with different prompt lengths, view counts and (some) audio / class-map tokens."""
    reqs = []
    for i in range(n):
        b = synth_batch(cfg, 1, 1 + i % 3, 14 + 5 * (i % 4), seed=seed + i, jitter=0, image_pos=3,
                        audio=i % 3 == 1, segmasks=i % 3 == 1)
        r = dict(input_ids=b["input_ids"], images=b["images"])
        if i % 3 == 1:
            r.update(audio=b["audio"], segmasks=b["segmasks"])
        reqs.append(r)
    return reqs


def test_continuous_batching_equals_per_request_generate(env):
    """9 requests through 3 decode rows: rows are refilled one by one as answers end (different EOS times: the EOS id is
    chosen inside the chain of some requests), the common position restarts when the batcher runs dry, and every
    request's ids equal a stand-alone generate() of it."""
    from mm_or_b200.serving.continuous import ContinuousBatcher
    cfg, model = env
    reqs = _requests(cfg, 9)
    succ = chain_successor(cfg.vocab_size)
    # EOS = the 4th new token of request 2: requests whose chain passes through it stop early, the others run to 12
    t = int(reqs[2]["input_ids"][0, -1])
    for _ in range(4):
        t = int(succ[t])
    old_eos = cfg.eos_token_id
    cfg.eos_token_id = t
    try:
        want = [model.generate(max_new_tokens=12, **r)[0].cpu() for r in reqs]
        lens = sorted({len(w) - r["input_ids"].shape[1] for w, r in zip(want, reqs)})
        assert lens[0] < 12 and lens[-1] == 12                      # some stop at EOS, some at the length limit
        cb = ContinuousBatcher(model, rows=3, max_prompt_len=700, max_new_tokens=12, window=2, check_every=3)
        tickets = [cb.submit(**r) for r in reqs]
        got = cb.drain()
        assert sorted(got) == sorted(tickets) and cb.prefills == 9
        for tk, w in zip(tickets, want):
            assert torch.equal(got[tk], w), (tk, got[tk][-14:].tolist(), w[-14:].tolist())
        assert cb.steps_run < 9 * 12                                # rows really ran side by side
        # a second wave on the same batcher (graph and cache reused, position restarted)
        tk = cb.submit(**reqs[4])
        assert torch.equal(cb.drain()[tk], want[4])
    finally:
        cfg.eos_token_id = old_eos


def test_continuous_batching_stopping_criterion_and_no_eos(env):
    from mm_or_b200.serving.continuous import ContinuousBatcher
    cfg, model = env
    reqs = _requests(cfg, 4, seed=340)
    want = [model.generate(max_new_tokens=10, stop_on_eos=False, **r)[0].cpu() for r in reqs]
    cb = ContinuousBatcher(model, rows=2, max_prompt_len=700, max_new_tokens=10, check_every=4, stop_on_eos=False)
    target = int(want[1][reqs[1]["input_ids"].shape[1] + 5])        # 6th new token of request 1

    def keyword(ids, scores, **kw):
        return bool(ids[0, -1] == target)

    tks = [cb.submit(stopping_criteria=[keyword] if i == 1 else None, **r) for i, r in enumerate(reqs)]
    got = cb.drain()
    for i, (tk, w) in enumerate(zip(tks, want)):
        n0 = reqs[i]["input_ids"].shape[1]
        assert torch.equal(got[tk], w[:n0 + 6] if i == 1 else w)
    with pytest.raises(ValueError):
        cb.submit(torch.zeros(2, 5, dtype=torch.long))              # one request at a time


def _toy_io(cfg):
    """Tokenizer / detokenizer for the scheduler tests: characters -> ids in [8, vocab), ids -> a parsable scene graph
    whose content depends on every generated id (so the memory string of later frames depends on earlier answers)."""
    V = cfg.vocab_size

    def tokenize(prompt):
        head, tail = prompt.split("<image>\n", 1)
        enc = lambda s: [8 + (ord(ch) * 7 + i) % (V - 8) for i, ch in enumerate(s)]
        return torch.tensor(enc(head) + [-200] + enc(tail), dtype=torch.long)

    def decode(ids):
        ids = [int(i) for i in ids if int(i) != 0]
        trip = ["e%d,e%d,p%d" % (ids[k] % 5, ids[k + 1] % 5, ids[k + 2] % 4) for k in range(0, len(ids) - 2, 3)]
        return "<SG> " + "; ".join(trip) + " </SG>"

    return tokenize, decode


def _takes(cfg, n_takes):
    g = torch.Generator().manual_seed(9)
    S = cfg.vision_config()["image_size"]
    takes = {}
    for t in range(n_takes):
        takes[f"take{t}"] = [dict(frame_id=3 * i + 1, prompt=f"<image>\nEntities: take {t}. Scene graph?",
                                  images=torch.randn(1 + t % 2, 3, S, S, generator=g).to(torch.bfloat16).float())
                             for i in range(2 + t % 2)]
    return takes


def test_online_scheduler_on_the_real_model(env):
    """OnlineScheduler over model.generate on the GPU: batching ACROSS takes (rounds of 3) and continuous batching both
    reproduce, text for text, the reference's order -- one take after the other, batch 1 (main.py:57-60) -- including
    the memory strings that feed each frame's prompt from the earlier answers of its take."""
    from mm_or_b200.serving.continuous import ContinuousBatcher
    from mm_or_b200.serving.scheduler import OnlineScheduler
    cfg, model = env
    tokenize, decode = _toy_io(cfg)
    make = lambda mb: OnlineScheduler(model, tokenize, decode, pad_token_id=0, max_batch=mb, max_new_tokens=9)
    ref = make(1).run(_takes(cfg, 3))                           # the reference's schedule
    assert all(len(r["triplets"]) == 3 for frames in ref.values() for r in frames)
    batched = make(3)
    got = batched.run(_takes(cfg, 3))
    assert got == ref and batched.rounds < sum(len(v) for v in ref.values())
    cb = ContinuousBatcher(model, rows=2, max_prompt_len=800, max_new_tokens=9, check_every=3)
    cont = make(2).run_continuous(_takes(cfg, 3), cb)
    assert cont == ref


def test_prefix_kv_reuse_equals_full_prefill(env):
    """Prefix-KV reuse (make_prefix_cache / generate(prefix_cache=...)): the keys / values of the system text every
    prompt starts with are computed once and copied into each row's cache at its own left-padding offset; the prefill
    then starts behind the prefix (b200_llama_prefill_from). Same greedy ids as the full prefill (chain weights:
    margins of several logits), logits equal up to the bf16 rounding of a differently tiled online softmax; a batch
    whose rows do not start with the prefix falls back to the full prefill."""
    from helpers import rel_err
    cfg, model = env
    g = torch.Generator().manual_seed(5)
    p = 13
    prefix = torch.randint(8, cfg.vocab_size, (p,), generator=g)
    tails = [5, 11, 2]
    Lt = p + 1 + max(tails)
    ids = torch.zeros(len(tails), Lt, dtype=torch.long)
    for r, t in enumerate(tails):
        row = torch.cat([prefix, torch.tensor([-200]), torch.randint(8, cfg.vocab_size, (t,), generator=g)])
        ids[r, Lt - len(row):] = row
    S = cfg.vision_config()["image_size"]
    images = [torch.randn(1 + r % 2, 3, S, S, generator=g).to(torch.bfloat16).float() for r in range(len(tails))]
    pc = model.make_prefix_cache(prefix)
    kw = dict(images=images, max_new_tokens=7, stop_on_eos=False)
    out_a, lg_a = model.generate(ids, return_logits=True, **kw)
    assert model._last_prefill_q0 == 0
    out_b, lg_b = model.generate(ids, return_logits=True, prefix_cache=pc, **kw)
    assert model._last_prefill_q0 == int(ids.eq(0).sum(1).min()) + p                 # the prefix really was skipped
    assert torch.equal(out_a, out_b) and rel_err(lg_b, lg_a) < 5e-3
    assert torch.equal(model.generate(ids, prefix_cache=pc, **kw), out_a)              # CUDA-graph loop
    other = ids.clone()
    other[0, Lt - (p + 1 + tails[0])] += 1                                              # row 0 no longer starts with it
    out_c = model.generate(other, prefix_cache=pc, **kw)
    assert model._last_prefill_q0 == 0 and torch.equal(out_c, model.generate(other, **kw))
    with pytest.raises(ValueError):
        model.make_prefix_cache(torch.tensor([5, -200, 7]))
