"""Records the REFERENCE model's outputs for a batch that carries point clouds next to audio and seg-masks: the
reference's own LlavaLlamaForCausalLM.forward(..., pc=, audio=, segmasks=) (imported unmodified through
oracle/ref_shim.py) on the small parity configuration -> tests/golden/pc_left.pt. Pins the token order
[pooled, pc, audio, seg0, seg1, seg2] (multimodal_projector/builder.py:176-189), the fp32 -> pooler-dtype cast of the
point-cloud token (:177) and the greedy continuation with the extra kwargs threaded through every decode step.

Run in the build container only:   python tests/golden/make_pc_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch

import golden_cases as gc
import make_golden as MG
from oracle import mm2sg_oracle as O
from oracle import ptv3_oracle as P
from oracle.ref_shim import build_reference_model

SHUFFLE_SEED = 77
STEPS = 6


def pc_case(cfg):
    case = gc.make_case(cfg, "extras_left")
    case["pc"] = P.dedupe_clouds([P.synth_cloud(1300, seed=21), None, P.synth_cloud(300, seed=22, box=(16, 16, 3))])
    return case


def weights(cfg):
    sd = gc.small_weights(cfg)
    sd.update(P.synth_weights())
    return gc.bf16_round(sd)


def main():
    torch.set_grad_enabled(False)
    cfg = gc.small_config()
    ocfg = MG.oracle_cfg(cfg)
    sd = weights(cfg)
    model = build_reference_model(cfg, sd)
    loaded = model.state_dict()
    for k, v in sd.items():                                   # the point-cloud weights really went in
        if "point_transformer" in k:
            assert torch.equal(loaded[k].float(), v), k
    model.config.tokenizer_padding_side = "left"
    model.config.mv_type = "learned"
    case = pc_case(cfg)
    kw = dict(images=case["images"], pc=case["pc"], audio=case["audio"], segmasks=case["segmasks"])
    concat = torch.cat(case["images"], 0)
    torch.manual_seed(SHUFFLE_SEED)
    visual = model.encode_images_pooled(concat, [im.shape[0] for im in case["images"]], case["pc"], case["audio"],
                                        case["segmasks"])
    assert visual.shape[1] == 576 + 5
    torch.manual_seed(SHUFFLE_SEED)
    ref = model(input_ids=case["input_ids"], attention_mask=case["attention_mask"],
                position_ids=MG.hf_generate_position_ids(case["attention_mask"]), **kw)
    torch.manual_seed(SHUFFLE_SEED)
    orc = O.multimodal_prefill(sd, ocfg, case["input_ids"], case["attention_mask"], case["images"],
                               audio=case["audio"], segmasks=case["segmasks"], pc=case["pc"], padding_side="left")
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-12)).item()
    errs = {"visual": rel(orc["visual"], visual), "pc_token": rel(orc["visual"][:, 576], visual[:, 576]),
            "logits": rel(orc["logits"][orc["mask"]], ref.logits.float()[orc["mask"]])}
    print("oracle-vs-reference rel errors:", errs)
    assert max(errs.values()) < 2e-4, errs
    # greedy continuation: every decode step of the reference re-receives the kwargs (llava_llama.py:108-127) and must
    # not re-encode them (early-exit branch llava_arch.py:192-201): the shuffle RNG is irrelevant after the prefill
    torch.manual_seed(SHUFFLE_SEED)
    toks, lg = MG.reference_greedy(model, case["input_ids"], case["attention_mask"], kw, STEPS)
    otoks, olg = O.greedy_decode(sd, ocfg, orc["logits"][:, -1], orc["kv"], orc["mask"], STEPS, stop_on_eos=False)
    print("greedy ids reference:", toks.tolist(), "oracle:", otoks.tolist(), "logit rel err", rel(olg, lg))
    assert rel(olg, lg) < 2e-4
    torch.save({"shuffle_seed": SHUFFLE_SEED, "visual_tail": visual[:, 570:].clone(),
                "logits_last": ref.logits[:, -1].float().clone(), "greedy_ids": toks, "greedy_logits": lg.half(),
                "n_points": [None if c is None else len(c) for c in case["pc"]]},
               os.path.join(gc.GOLDEN_DIR, "pc_left.pt"))


if __name__ == "__main__":
    main()
