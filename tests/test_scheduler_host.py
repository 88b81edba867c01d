"""Online serving loop batched across takes (mm_or_b200/serving/scheduler.py) with a stub model and tokenizer: the text
produced for a frame must not depend on how takes are interleaved (same results as the reference's take-by-take,
batch-1 order), slots are refilled from the waiting queue, modality kwargs follow the reference's None convention."""
import random

import torch

from mm_or_b200.serving.scheduler import OnlineScheduler, left_pad

IMAGE = -200
VOCAB = ["<pad>", "<SG>", "</SG>", ";", ","] + [f"w{i}" for i in range(40)]


def tokenize(prompt):
    """One id per character class of the prompt, image placeholder -> -200 (content-dependent, deterministic)."""
    head, tail = prompt.split("<image>\n", 1)
    ids = [5 + (ord(c) % 40) for c in head] + [IMAGE] + [5 + (ord(c) % 40) for c in tail]
    return torch.tensor(ids, dtype=torch.long)


def decode(ids):
    return " ".join(VOCAB[int(i)] for i in ids if int(i) != 0)


class StubModel:
    """generate(): 'predicts' a scene graph that is a pure function of each row's own prompt ids and image sum, like
    a real model is a function of its inputs; records what it was called with."""

    def __init__(self):
        self.calls = []

    def generate(self, input_ids, images=None, do_sample=False, use_cache=True, max_new_tokens=20,
                 stopping_criteria=None, pc=None, audio=None, segmasks=None):
        self.calls.append(dict(B=input_ids.shape[0], pc=pc, audio=audio, segmasks=segmasks, L=input_ids.shape[1]))
        assert len(images) == input_ids.shape[0] and not do_sample and use_cache
        rows = []
        for r, img in zip(input_ids, images):
            real = r[r != 0]
            h = (int(real[real > 0].sum()) + int(img.sum())) % 1000
            a, b, c = 5 + h % 7, 5 + (h // 7) % 7, 5 + (h // 49) % 5
            rows.append(torch.tensor([1, a, 4, b, 4, c, 3, 2] + [0] * (max_new_tokens - 8)))
        return torch.cat([input_ids, torch.stack(rows)], dim=1)


def make_takes(n_takes, seed=0):
    g = torch.Generator().manual_seed(seed)
    takes = {}
    for t in range(n_takes):
        n_frames = 2 + (t * 3) % 4                           # ragged takes
        takes[f"take{t}"] = [dict(frame_id=5 * i + t, prompt=f"<image>\nEntities: take {t}. Scene graph?",
                                  images=torch.randint(0, 9, (2, 3, 4, 4), generator=g).float(),
                                  audio=torch.ones(512) if (t == 1 and i == 0) else None)
                             for i in range(n_frames)]
    return takes


def run(takes, max_batch):
    model = StubModel()
    sched = OnlineScheduler(model, tokenize, decode, pad_token_id=0, max_batch=max_batch, max_new_tokens=10)
    return sched.run(takes), model, sched


def test_left_pad_matches_reference_trick():
    rows = [torch.tensor([1, 2, 3]), torch.tensor([7]), torch.tensor([4, 5])]
    assert left_pad(rows, 0).tolist() == [[1, 2, 3], [0, 0, 7], [0, 4, 5]]


def test_interleaving_does_not_change_results():
    takes = make_takes(5)
    sequential, m1, s1 = run(takes, max_batch=1)             # the reference's order: one take after the other
    batched, m2, s2 = run(takes, max_batch=3)                # 3 slots, 5 takes: slots are refilled from the queue
    wide, m3, s3 = run(takes, max_batch=64)
    assert sequential == batched == wide
    n_frames = sum(len(v) for v in takes.values())
    assert s1.rounds == n_frames and s3.rounds == max(len(v) for v in takes.values())
    assert s1.rounds > s2.rounds > s3.rounds
    assert max(c["B"] for c in m2.calls) == 3 and max(c["B"] for c in m3.calls) == 5
    # the memory of a take feeds its next prompt: later frames of a take see a longer prompt than its first frame
    first = [c["L"] for c in m1.calls][0]
    assert max(c["L"] for c in m1.calls) > first
    for name, frames in takes.items():
        assert [r["frame_id"] for r in sequential[name]] == [f["frame_id"] for f in frames]
        assert all(len(r["triplets"]) == 1 for r in sequential[name])


def test_modality_kwargs_follow_the_reference_convention():
    takes = make_takes(3)
    _, model, _ = run(takes, max_batch=64)
    first = model.calls[0]                                    # take1's first frame carries audio
    assert first["audio"] is not None and first["audio"][0] is None and first["audio"][1] is not None
    assert first["pc"] is None and first["segmasks"] is None
    assert all(c["audio"] is None for c in model.calls[1:])


def test_per_take_rng_is_independent_of_the_schedule():
    seen = {}

    def factory(name):
        seen[name] = random.Random(name)
        return seen[name]

    takes = make_takes(2)
    model = StubModel()
    OnlineScheduler(model, tokenize, decode, max_batch=2, max_new_tokens=10, rng_factory=factory).run(takes)
    assert sorted(seen) == ["take0", "take1"]
