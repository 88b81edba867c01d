"""Token-level continuous batching for MM2SG's online mode: a fixed set of decode rows that
requests join and leave individually, instead of whole batches that wait for their slowest row.

Why it exists: in `temporality = PRED` mode a take's frame t can only be submitted after frame t - 1 has been decoded
(scene_graph_prediction_model.py:182-195, 309-335), answers differ a lot in length (a scene graph is 5-60 triplets),
and the reference decodes one frame at a time (main.py:57-60). serving/scheduler.py batches across takes but still in
rounds: every row of a round waits for the longest answer. Here a row that has emitted EOS is handed to the next
pending request at once, while the other rows keep decoding -- the B200 decode step costs the same whatever the
number of live rows (the 13.2 GB of weights are streamed once per step).

Mechanics (all on top of the existing kernels, no new device code):
  * one DecodeSlot (model/llava_llama.py): KV cache [layers][rows][heads][cap][128], tokens, history, finished flags,
    kv_start per row, the step counters -- and ONE captured CUDA graph of a decode step over them, replayed for the
    whole life of the batcher (the buffers never move);
  * rows are right-aligned on a common position P (= the slot's step state): a request whose packed prompt has Lp <= P
    rows is prefilled into its row with kv_start = P - Lp (the same left padding generate() uses; RoPE positions are
    relative to kv_start, so its arithmetic is that of a stand-alone generate() of the request);
  * free and finished rows carry finished = 1: the decode attention skips their cache reads and the argmax emits pad;
  * the host looks at the finished flags / history every `check_every` steps, retires rows (EOS, max_new_tokens or a
    stopping criterion), and admits pending requests into the free rows;
  * P only grows; when P + max_new_tokens would pass the cache capacity, admissions pause until the rows in flight
    have drained, then P restarts at the next request's prompt length.
Results equal per-request generate() up to bf16 rounding of different split points (token-exact where the top-2 margin
allows; exact with the chain weight set, tests/test_gpu_zzz_serving.py).
"""
import ctypes
from collections import deque

import numpy as np
import torch

from .. import _lib as L
from ..model.llava_llama import BF, DecodeSlot


class ContinuousBatcher:
    def __init__(self, model, rows=16, max_prompt_len=1024, max_new_tokens=300, window=2, check_every=8,
                 stop_on_eos=True):
        """rows: decode rows (batch of every decode step). Cache capacity = max_prompt_len + (1 + window) *
        max_new_tokens: `window` is how many generations' worth of steps P may advance before admissions have to wait
        for a drain."""
        if model._w is None:
            raise L.B200Error("weights not loaded")
        c = model.config
        if getattr(c, "tokenizer_padding_side", "right") != "left":
            raise ValueError("continuous batching right-aligns the rows: set config.tokenizer_padding_side = 'left' "
                             "(the reference does for inference, scene_graph_prediction_model.py:54)")
        self.model, self.rows = model, int(rows)
        self.max_new, self.check_every, self.stop_on_eos = int(max_new_tokens), int(check_every), bool(stop_on_eos)
        self.cap = (int(max_prompt_len) + (1 + int(window)) * self.max_new + 7) // 8 * 8
        dev = model.device
        self.slot = DecodeSlot(c, self.rows, self.cap, self.cap, dev)        # history: one column per step
        self.slot.finished.fill_(1)                                           # nothing is live yet
        self.slot.history.fill_(c.pad_token_id if c.pad_token_id is not None else 0)
        self.slot.kv_start.zero_()
        self.slot.tokens.zero_()
        self.pos = 0                      # P: position the next decode step writes (host mirror of slot.state[0])
        self.col = 1                      # history column the next decode step writes (mirror of slot.state[1])
        self.live = {}                    # row -> dict(ticket, start_col, prompt_len, criteria, ids_cpu)
        self.pending = deque()
        self.finished_out = []
        self.tickets = 0
        self.steps_run = 0
        self.prefills = 0
        self._graph = None
        self._ws = L.Workspace()          # own decode workspace: the captured graph holds its address
        self._eos = c.eos_token_id if (self.stop_on_eos and c.eos_token_id is not None) else -1
        self._pad = c.pad_token_id if c.pad_token_id is not None else 0

    # ---- admission ------------------------------------------------------------------------------------------------
    def submit(self, input_ids, images=None, stopping_criteria=None, **modalities):
        """Queue one request (input_ids (1, Lt) or (Lt,), images (V, 3, S, S) or a list with that one tensor, and
        pc / audio / segmasks / vis_descriptor_embs as generate() takes them for a batch of one). Returns its ticket."""
        ids = input_ids.detach().to("cpu")
        if ids.ndim == 1:
            ids = ids[None]
        if ids.shape[0] != 1:
            raise ValueError("submit() takes one request at a time")
        if torch.is_tensor(images) and images.ndim == 4:
            images = [images]
        self.tickets += 1
        self.pending.append(dict(ticket=self.tickets, ids=ids, images=images, criteria=stopping_criteria, kw=modalities))
        return self.tickets

    def idle(self):
        return not self.live and not self.pending

    def _free_rows(self):
        return [r for r in range(self.rows) if r not in self.live]

    def _packed_len(self, req):
        """Packed prompt length of a request without running the encoders (host index arithmetic only)."""
        from ..model.pack import descriptor_row_counts, plan_pack
        model, c, ids = self.model, self.model.config, req["ids"]
        mask = ids.ne(c.pad_token_id) if c.pad_token_id is not None else torch.ones_like(ids, dtype=torch.bool)
        if req["images"] is None or model.get_vision_tower() is None:
            return int(mask.sum())
        pooler, kw = model.get_image_pooler(), req["kw"]
        t_vis = pooler.geo["keep"] + pooler.num_extra_tokens(kw.get("pc"), kw.get("audio"), kw.get("segmasks"))
        desc_rows = None
        if kw.get("vis_descriptor_embs") is not None:
            _, desc_rows = descriptor_row_counts(kw["vis_descriptor_embs"], 1)
        n_blocks = len(req["images"]) if type(req["images"]) is list else int(req["images"].shape[0])
        plan = plan_pack(ids.numpy(), mask.numpy(), None, t_vis, "left",
                         getattr(c, "tokenizer_model_max_length", None), desc_rows=desc_rows, n_blocks=n_blocks)
        return int(plan.lengths[0])

    def _admit(self):
        """Prefill as many pending requests as there are free rows and the position allows."""
        model, c, lib = self.model, self.model.config, L.lib()
        dev, D, V = model.device, c.hidden_size, c.vocab_size
        free = self._free_rows()
        if not self.live and self.pending:
            # nothing in flight: restart the common position at the longest prompt of the requests about to join
            first = list(self.pending)[:len(free)]
            self.pos, self.col = max(self._packed_len(r) for r in first), 1
            self.slot.state.copy_(torch.tensor([self.pos, self.col], dtype=torch.int32), non_blocking=True)
        while self.pending and free:
            req = self.pending[0]
            ids = req["ids"]
            mask = ids.ne(c.pad_token_id) if c.pad_token_id is not None else torch.ones_like(ids, dtype=torch.bool)
            if req["images"] is not None and model.get_vision_tower() is not None:
                (_, _, _, _, embeds, _, plan) = model.prepare_inputs_labels_for_multimodal(
                    ids, None, mask, None, None, req["images"], req["kw"].get("vis_descriptor_embs"),
                    req["kw"].get("pc"), req["kw"].get("audio"), req["kw"].get("segmasks"))
                Lp = int(plan.lengths[0])
                embeds = embeds[:, plan.L - Lp:]                                   # real rows only
            else:
                real = ids[mask][None]
                Lp = real.shape[1]
                embeds = torch.empty((1, Lp, D), device=dev, dtype=BF)
                L.check(lib.b200_embed_rows(L.ptr(real.to(torch.int32).to(dev).contiguous()),
                                            L.ptr(model.model.embed_tokens), L.ptr(embeds), D, Lp, D, V, L.stream_ptr()),
                        "b200_embed_rows")
            if Lp + self.max_new > self.cap:
                self.pending.popleft()          # it can never be served by this batcher: do not block the queue
                raise ValueError(f"request {req['ticket']}: {Lp} packed positions + {self.max_new} new tokens exceed "
                                 f"the cache capacity {self.cap}: raise max_prompt_len")
            if Lp > self.pos or self.pos + self.max_new > self.cap:
                break                       # has to wait: P must grow to its length / the rows in flight must drain
            self.pending.popleft()
            row = free.pop(0)
            P = self.pos
            x = torch.zeros((1, P, D), device=dev, dtype=BF)                       # left padding = zero rows
            x[:, P - Lp:] = embeds
            start = torch.tensor([P - Lp], dtype=torch.int32).to(dev, non_blocking=True)
            self.slot.kv_start[row:row + 1].copy_(start)
            logits = torch.empty((1, V), device=dev, dtype=BF)
            nb = lib.b200_llama_prefill_workspace_bytes(ctypes.byref(model._w), 1, P, 0)
            ws = model._ws_prefill.get(nb, dev)
            cs = self.slot.cache.struct(row)
            L.check(lib.b200_llama_prefill(ctypes.byref(model._w), L.ptr(x), L.ptr(start), None, ctypes.byref(cs), 1, P,
                                           L.ptr(logits), 0, 0, L.ptr(ws), ws.numel(), L.stream_ptr()),
                    "b200_llama_prefill")
            # first token of the answer from the prefill logits -> tokens[row], history[row, col - 1]; the row is live
            self.slot.finished[row:row + 1].zero_()
            L.check(lib.b200_argmax(L.ptr(logits), 0, V, 1, V, L.ptr(self.slot.tokens[row:row + 1]),
                                    L.ptr(self.slot.finished[row:row + 1]), self._eos, self._pad, L.stream_ptr()),
                    "b200_argmax")
            self.slot.history[row, self.col - 1] = self.slot.tokens[row]
            self.live[row] = dict(ticket=req["ticket"], start_col=self.col - 1, prompt_len=Lp, criteria=req["criteria"],
                                  ids_cpu=ids)
            self.prefills += 1

    # ---- decode ---------------------------------------------------------------------------------------------------
    def _step_fn(self):
        model, c, lib, s = self.model, self.model.config, L.lib(), self.slot
        nb = lib.b200_llama_decode_workspace_bytes(ctypes.byref(model._w), self.rows, self.cap)
        ws = self._ws.get(nb, model.device)
        cs = s.cache.struct(0)

        def step():
            L.check(lib.b200_llama_decode_step(ctypes.byref(model._w), L.ptr(s.tokens), L.ptr(s.state), L.ptr(s.kv_start),
                                               ctypes.byref(cs), self.rows, self.cap - 1, L.ptr(s.finished),
                                               self._eos, self._pad, L.ptr(s.history), self.cap, None, L.ptr(ws),
                                               ws.numel(), L.stream_ptr()), "b200_llama_decode_step")
        return step

    def _decode(self, n):
        lib = L.lib()
        step = self._step_fn()
        for _ in range(n):
            if self._graph is None and self.steps_run >= 1:
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                n0 = lib.b200_launch_count()
                with torch.cuda.stream(side):
                    graph.capture_begin(capture_error_mode="thread_local")
                    try:
                        step()
                    finally:
                        graph.capture_end()
                self._graph = (graph, lib.b200_launch_count() - n0)
                L.note_graph_replay(-self._graph[1])
                torch.cuda.current_stream().wait_stream(side)
            if self._graph is not None:
                self._graph[0].replay()
                L.note_graph_replay(self._graph[1])
            else:
                step()                     # the very first step runs eagerly (lazy kernel-attribute setup)
            self.steps_run += 1
            self.pos += 1
            self.col += 1

    def _retire(self):
        """Look at the rows in flight (one device -> host read of the flags and the new history columns)."""
        c = self.model.config
        if not self.live:
            return
        fin = self.slot.finished.to("cpu")
        hist = self.slot.history[:, :self.col].to("cpu", torch.long)
        for row in sorted(self.live):
            st = self.live[row]
            n = self.col - st["start_col"]                      # tokens produced so far (incl. the prefill's)
            new = hist[row, st["start_col"]:st["start_col"] + min(n, self.max_new)]
            stop = n >= self.max_new
            if self._eos >= 0:
                hit = (new == self._eos).nonzero()
                if len(hit):
                    new = new[:int(hit[0]) + 1]                 # HF: the EOS is the last token of the row
                    stop = True
            if not stop and st["criteria"]:
                # HF evaluates stopping criteria after every step: find the first prefix at which one fires
                for k in range(st.get("checked", 0) + 1, len(new) + 1):
                    out_ids = torch.cat([st["ids_cpu"], new[None, :k]], dim=1)
                    fired = False
                    for sc in st["criteria"]:
                        r = sc(out_ids, None)
                        fired = fired or (bool(r.all()) if torch.is_tensor(r) else bool(r))
                    if fired:
                        new, stop = new[:k], True
                        break
                st["checked"] = len(new)
            if stop:
                self.finished_out.append((st["ticket"], torch.cat([st["ids_cpu"][0], new])))
                del self.live[row]
                if int(fin[row]) == 0:
                    self.slot.finished[row:row + 1].fill_(1)   # stopped by length / criterion: the row goes quiet

    def run(self, max_steps=None):
        """Admit what fits, decode `check_every` steps at a time and retire finished rows until something finished or
        `max_steps` decode steps ran. Returns [(ticket, LongTensor prompt ids + new ids)] in completion order."""
        done_steps = 0
        while not self.idle():
            self._admit()
            if not self.live:
                if self.pending:
                    raise RuntimeError("a pending request cannot be admitted into an empty batcher")
                break
            n = self.check_every if max_steps is None else min(self.check_every, max_steps - done_steps)
            self._decode(n)
            done_steps += n
            self._retire()
            if self.finished_out or (max_steps is not None and done_steps >= max_steps):
                break
        out, self.finished_out = self.finished_out, []
        return out

    def drain(self):
        """Run until every submitted request has finished; returns {ticket: ids}."""
        res = {}
        while not self.idle():
            for t, ids in self.run():
                res[t] = ids
        return res
