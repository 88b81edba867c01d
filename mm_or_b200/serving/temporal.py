"""Host logic of MM2SG's online / temporal mode (`temporality = PRED`): parse the decoded text
into triplets, keep a per-take change log of the predicted scene graphs, and render it as the `<memory_start>` string
that is spliced into the next frame's prompt.

Reference (paths under scene_graph_generation/scene_graph_prediction/):
  llava_helpers/scene_graph_converters.py:9-22    collapse_sgs
  llava_helpers/scene_graph_converters.py:52-89   llava_sg_to_surgery_sg (entity_of_interest = None branch, the one the
                                                  model wrapper uses with IRRELEVANT_PREDS = ['closeto', 'closeTo'])
  llava_helpers/scene_graph_converters.py:96-112  parse_llava_sg
  llava_helpers/scene_graph_converters.py:115-174 surgery_sg_to_memory_str (DROP_HISTORY = False)
  scene_graph_helpers/model/scene_graph_prediction_model.py:182-195   memory string into the prompt (5000-char clip)
  scene_graph_helpers/model/scene_graph_prediction_model.py:309-335   output text -> raw triplets -> take history
Pure Python, no tensors. The reference shuffles the modifications of a timepoint with the global `random` module;
`rng` defaults to that module so that, under the same seed and call sequence, the change log is identical.
"""
import random
import re

IRRELEVANT_PREDS = ("closeto", "closeTo")
MEMORY_CLIP = 5000


def parse_scene_graph(text):
    """Decoded answer -> [(sub, pred, obj)]. Grammar '<SG> sub,obj,pred; ... </SG>'; chain-of-thought between triple
    quotes is dropped first (scene_graph_prediction_model.py:311)."""
    text = re.sub(r'""".*?"""', "", text, flags=re.DOTALL)
    if "<SG>" in text and "</SG>" in text and text.index("<SG>") < text.index("</SG>"):
        chunks = text.split("<SG>")[1].split("</SG>")[0].strip().split(";")
    else:
        chunks = text.split(";")
    out = []
    for chunk in chunks:
        chunk = chunk.replace(".", "").replace("</s>", "").replace("<s>", "").strip()
        if not chunk:
            continue
        parts = [p.strip() for p in chunk.split(",")]
        if len(parts) != 3:
            continue
        sub, obj, pred = parts
        out.append((sub, pred, obj))
    return out


def collapse(change_log):
    """Current state implied by a change log [(timepoint, (sub, pred, obj))]: 'not <pred>' entries end a relation."""
    state = {}
    for _, (sub, pred, obj) in change_log:
        if pred.startswith("not "):
            state.pop((sub, obj), None)
        else:
            state[(sub, obj)] = pred
    return state


def to_change_log(history, irrelevant_preds=IRRELEVANT_PREDS, rng=random):
    """history: [{'timepoint_idx': t, 'scene_graph': [(sub, pred, obj)]}] in time order -> change log of additions
    and removals ('not <pred>') relative to the collapsed state so far."""
    log = []
    for entry in history:
        prev = collapse(log)
        if irrelevant_preds is None:
            cur = {(s, o): p for (s, p, o) in entry["scene_graph"] if s != "none" and o != "none"}
        else:
            cur = {(s, o): p for (s, p, o) in entry["scene_graph"]
                   if p not in irrelevant_preds and s != "none" and o != "none"}
        mods = [(entry["timepoint_idx"], (s, p, o)) for (s, o), p in cur.items() if (s, o) not in prev]
        mods += [(entry["timepoint_idx"], (s, f"not {p}", o)) for (s, o), p in prev.items() if (s, o) not in cur]
        rng.shuffle(mods)
        log.extend(mods)
    return log


def memory_string(change_log, style="longshort"):
    """'Long: ' = first occurrence of every (non-'not') action older than the last five changes, 'Short: ' = the last
    five changes verbatim; entries rendered 'sub,obj,pred; '."""
    def fmt(sub, pred, obj):
        return f"{sub},{obj},{pred}; "

    out = ""
    if style in ("long", "longshort"):
        out += "Long: "
        seen = set()
        for _, (sub, pred, obj) in change_log[:-5]:
            if (sub, obj, pred) not in seen and not pred.startswith("not "):
                seen.add((sub, obj, pred))
                out += fmt(sub, pred, obj)
    if style in ("short", "longshort"):
        out += "Short: "
        for _, (sub, pred, obj) in change_log[-5:]:
            out += fmt(sub, pred, obj)
    if style not in ("short", "long", "longshort") or out == "":
        return ""
    return out[:-2]


class TakeMemory:
    """Per-take state of the online mode: one instance replaces `ModelWrapper.take_to_history[take_name]`."""

    def __init__(self, rng=random):
        self.history = []
        self.rng = rng

    def add_prediction(self, timepoint_idx, decoded_text):
        """Record the model's answer for a frame; returns the raw triplets [(sub, pred, obj)]."""
        triplets = parse_scene_graph(decoded_text)
        self.history.append({"timepoint_idx": int(timepoint_idx), "scene_graph": triplets})
        return triplets

    def memory_for(self, timepoint_idx):
        log = to_change_log(self.history, IRRELEVANT_PREDS, self.rng)
        log = [e for e in log if e[0] < int(timepoint_idx)]
        mem = memory_string(log, "longshort")
        if len(mem) > MEMORY_CLIP:
            mem = "..." + mem[-MEMORY_CLIP:]
        return mem

    def splice(self, prompt, timepoint_idx, image_token="<image>"):
        """Insert '<memory_start>: ...<memory_end>.' right after the image placeholder line."""
        mem = self.memory_for(timepoint_idx)
        return prompt.replace(f"{image_token}\n", f"{image_token}\n<memory_start>: {mem}<memory_end>.\n")
