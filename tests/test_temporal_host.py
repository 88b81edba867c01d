"""Online / temporal mode host logic (mm_or_b200/serving/temporal.py) against golden vectors recorded from the
reference's own functions (tests/golden/make_temporal_golden.py -> temporal.json): answer parsing, change log
(including the reference's use of the global `random` shuffle), collapse, and the three memory-string styles."""
import json
import os
import random

import pytest

from mm_or_b200.serving import temporal as T

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "temporal.json")))


def _t3(x):
    return [tuple(t) for t in x]


@pytest.mark.parametrize("case", GOLD, ids=[str(c["seed"]) for c in GOLD])
def test_against_reference_golden(case):
    parsed = [T.parse_scene_graph(a) for a in case["answers"]]
    assert parsed == [_t3(p) for p in case["parsed"]]
    history = [{"timepoint_idx": t, "scene_graph": p} for t, p in zip(case["timepoints"], parsed)]
    random.seed(1000 + case["seed"])                        # the reference shuffles with the global RNG
    log = T.to_change_log(history)
    assert log == [(t, tuple(tr)) for t, tr in case["change_log"]]
    assert sorted(T.collapse(log).items()) == sorted((tuple(k), v) for k, v in case["collapsed"])
    before = [e for e in log if e[0] < case["query_t"]]
    for style, want in case["memory"].items():
        assert T.memory_string(before, style) == want


def test_take_memory_splices_prompt_and_clips():
    tm = T.TakeMemory(rng=random.Random(0))
    assert tm.add_prediction(3, "<SG> nurse,patient,touching; head surgeon,saw,holding; </SG>") == [
        ("nurse", "touching", "patient"), ("head surgeon", "holding", "saw")]
    tm.add_prediction(9, "<SG> head surgeon,saw,holding; </SG>")
    prompt = tm.splice("<image>\nEntities: ...", 12)
    assert prompt.startswith("<image>\n<memory_start>: Long: Short: ") and "<memory_end>.\nEntities: ..." in prompt
    assert "nurse,patient,not touching" in prompt                # the relation ended at t = 9
    assert tm.memory_for(3) == "Long: Short"                     # nothing strictly before t = 3 ([:-2] of 'Short: ')
    big = T.TakeMemory(rng=random.Random(1))
    for t in range(400):
        big.add_prediction(t, f"<SG> e{t},f{t},holding; </SG>")
    m = big.memory_for(10 ** 6)
    assert len(m) == T.MEMORY_CLIP + 3 and m.startswith("...")
