"""Records golden vectors for the online-mode host logic from the reference's OWN functions
(scene_graph_prediction/llava_helpers/scene_graph_converters.py). Run in the build container only (needs
/root/reference); writes tests/golden/temporal.json."""
import importlib.util
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scene_graph_generation/scene_graph_prediction/llava_helpers/scene_graph_converters.py"

spec = importlib.util.spec_from_file_location("ref_converters", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

ENT = ["head surgeon", "assistant surgeon", "nurse", "patient", "instrument table", "operating table", "drill", "saw", "none"]
PRED = ["holding", "cutting", "drilling", "closeTo", "lyingOn", "touching", "assisting", "closeto"]


def random_answer(rng, n):
    parts = []
    for _ in range(n):
        parts.append(f"{rng.choice(ENT)},{rng.choice(ENT)},{rng.choice(PRED)}")
    body = "; ".join(parts)
    style = rng.randrange(4)
    if style == 0:
        return f"<SG> {body}; </SG>"
    if style == 1:
        return f'"""thinking, a, b; c"""<SG> {body} </SG></s>'
    if style == 2:
        return body + "; broken,entry; ."
    return f"<s><SG>{body};</SG>. trailing, text, here"


def main():
    cases = []
    for seed in range(12):
        rng = random.Random(seed)
        answers = [random_answer(rng, rng.randrange(0, 7)) for _ in range(rng.randrange(1, 14))]
        timepoints = sorted(rng.sample(range(0, 400), len(answers)))
        parsed = [ref.parse_llava_sg(__import__("re").sub(r'""".*?"""', "", a, flags=__import__("re").DOTALL)) for a in answers]
        history = [{"timepoint_idx": t, "scene_graph": p} for t, p in zip(timepoints, parsed)]
        query_t = timepoints[-1] + rng.randrange(0, 3)
        random.seed(1000 + seed)
        log = ref.llava_sg_to_surgery_sg(history, entity_of_interest=None, IRRELEVANT_PREDS=["closeto", "closeTo"])
        log_before = [e for e in log if e[0] < query_t]
        cases.append({"seed": seed, "answers": answers, "timepoints": timepoints, "parsed": parsed, "query_t": query_t,
                      "change_log": log, "collapsed": [[list(k), v] for k, v in ref.collapse_sgs(log).items()],
                      "memory": {s: ref.surgery_sg_to_memory_str(log_before, current_timepoint=query_t, TEMPORAL_STYLE=s)
                                 for s in ("short", "long", "longshort")}})
    with open(os.path.join(HERE, "temporal.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
