"""Online (temporal) serving loop of MM2SG batched ACROSS takes.

Reference: in `temporality = PRED` mode the prompt of frame t carries a memory string built from the model's own
predictions for the frames before t of the same take (scene_graph_prediction_model.py:182-195, :309-335), so a take is
strictly sequential; the reference therefore runs batch size 1, one take after the other
(scene_graph_prediction_model.py:302 asserts it; main.py:57-60). Takes do not share state (`take_to_history[take]`),
so different takes CAN be batched: every round this scheduler takes the next frame of each active take (up to
`max_batch`; a finished take hands its slot to a waiting one), splices that take's memory into the prompt, left-pads
the prompts into one batch (the reference's flip / pad_sequence / flip, :203-208) and makes ONE `model.generate` call
with the reference's kwargs (:221-231). The B200 decode step streams the 13.2 GB of weights once per step whatever the
batch, so a round of B takes costs barely more than one frame alone.

Equivalence: the text generated for a frame depends only on its take's history, hence results equal the reference's
take-by-take order -- given one shuffle RNG per take (`rng_factory`; the reference draws the shuffle of
llava_sg_to_surgery_sg from the global `random`, whose consumption order an interleaved schedule cannot reproduce).

Pure host logic over the public model API: no tensors are created here except the padded id batch.
"""
import random
from collections import deque

import torch

from .temporal import TakeMemory


def left_pad(id_rows, pad_id):
    """List of 1-D LongTensors -> (B, Lmax) left-padded batch (scene_graph_prediction_model.py:203-208)."""
    flipped = [torch.flip(r, dims=[0]) for r in id_rows]
    batch = torch.nn.utils.rnn.pad_sequence(flipped, batch_first=True, padding_value=pad_id)
    return torch.flip(batch, dims=[1])


class OnlineScheduler:
    def __init__(self, model, tokenize, decode, pad_token_id=0, max_batch=64, max_new_tokens=300,
                 stopping_criteria_factory=None, rng_factory=None, image_token="<image>", prefix_cache=None):
        """tokenize(prompt) -> 1-D LongTensor with IMAGE_TOKEN_INDEX at the placeholder (the reference's
        tokenizer_image_token); decode(1-D LongTensor of new ids) -> text; stopping_criteria_factory(input_ids) -> list
        (the reference builds KeywordsStoppingCriteria from the batch's input_ids); rng_factory(take_name) -> the
        shuffle RNG of that take (default: random.Random seeded with the take name)."""
        self.model, self.tokenize, self.decode = model, tokenize, decode
        self.pad, self.max_batch, self.max_new = pad_token_id, int(max_batch), int(max_new_tokens)
        self.criteria = stopping_criteria_factory
        self.rng_factory = rng_factory or (lambda name: random.Random(str(name)))
        self.image_token = image_token
        # keys / values of the system text every prompt starts with (model.make_prefix_cache(tokenize(prompt)[:p])):
        # copied into the KV cache of every round instead of being recomputed
        self.prefix_cache = prefix_cache
        self.rounds = 0

    def run(self, takes):
        """takes: {take_name: iterable of frames}, frame = dict(frame_id=int, prompt=str containing '<image>\\n',
        images=(V, 3, S, S) tensor, and optionally pc / audio / segmasks as the reference's loader yields them).
        Returns {take_name: [dict(frame_id, text, triplets)]} in frame order."""
        waiting = deque((name, iter(frames)) for name, frames in takes.items())
        active = []                                         # [name, iterator, TakeMemory]
        results = {name: [] for name in takes}
        while waiting or active:
            while waiting and len(active) < self.max_batch:
                name, it = waiting.popleft()
                active.append([name, it, TakeMemory(rng=self.rng_factory(name))])
            batch, still = [], []
            for slot in active:
                frame = next(slot[1], None)
                if frame is not None:
                    batch.append((slot, frame))
                    still.append(slot)
            active = still
            if not batch:
                continue
            rows = [self.tokenize(slot[2].splice(f["prompt"], int(f["frame_id"]), self.image_token))
                    for slot, f in batch]
            input_ids = left_pad(rows, self.pad)
            kw = {}
            for key in ("pc", "audio", "segmasks"):         # a kwarg is passed only when some sample has it (:228-230)
                vals = [f.get(key) for _, f in batch]
                kw[key] = vals if any(v is not None for v in vals) else None
            criteria = self.criteria(input_ids) if self.criteria is not None else None
            if self.prefix_cache is not None:
                kw["prefix_cache"] = self.prefix_cache
            out = self.model.generate(input_ids, images=[f["images"] for _, f in batch], do_sample=False,
                                      use_cache=True, max_new_tokens=self.max_new, stopping_criteria=criteria, **kw)
            self.rounds += 1
            new = out[:, input_ids.shape[1]:]               # the prompt is echoed back in front (:233-235)
            for (slot, f), ids in zip(batch, new):
                text = self.decode(ids.to("cpu")).strip()
                triplets = slot[2].add_prediction(int(f["frame_id"]), text)
                results[slot[0]].append({"frame_id": int(f["frame_id"]), "text": text, "triplets": triplets})
        return results

    def run_continuous(self, takes, batcher):
        """The same online loop on a serving/continuous.py ContinuousBatcher: a take's next frame is submitted the
        moment its previous answer is complete and joins the decode rows individually -- no round waits for its longest
        answer. `batcher.max_new` bounds the answer length; stopping criteria are built per request. Returns the same
        structure as run(); results are schedule-independent for the same reason (one memory + shuffle RNG per take)."""
        results = {name: [] for name in takes}
        state = {name: [iter(frames), TakeMemory(rng=self.rng_factory(name)), None] for name, frames in takes.items()}
        in_flight = {}                                      # ticket -> (take name, frame, prompt length)

        def feed():
            for name, st in state.items():
                if st[2] is not None or len(in_flight) >= self.max_batch:
                    continue
                frame = next(st[0], None)
                if frame is None:
                    continue
                ids = self.tokenize(st[1].splice(frame["prompt"], int(frame["frame_id"]), self.image_token))[None]
                kw = {k: [frame[k]] for k in ("pc", "audio", "segmasks") if frame.get(k) is not None}
                criteria = self.criteria(ids) if self.criteria is not None else None
                t = batcher.submit(ids, images=[frame["images"]], stopping_criteria=criteria, **kw)
                st[2] = t
                in_flight[t] = (name, frame, ids.shape[1])

        feed()
        while in_flight:
            for ticket, out in batcher.run():
                name, frame, n_prompt = in_flight.pop(ticket)
                text = self.decode(out[n_prompt:].to("cpu")).strip()
                triplets = state[name][1].add_prediction(int(frame["frame_id"]), text)
                results[name].append({"frame_id": int(frame["frame_id"]), "text": text, "triplets": triplets})
                state[name][2] = None
            self.rounds += 1
            feed()
        return results
