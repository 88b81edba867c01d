"""Records golden outputs of the REFERENCE model code (egeozsoy/MM-OR, imported unmodified from /root/reference
through oracle/ref_shim.py) on the small parity configuration, and checks the CPU oracle against them.

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/<case>.pt (fp16-compressed slices, a few hundred KB each). The GPU box never needs /root/reference:
tests regenerate weights and inputs from their seeds (tests/golden_cases.py) and compare against these files.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import golden_cases as gc
from oracle import mm2sg_oracle as O
from oracle.ref_shim import build_reference_model

GREEDY_STEPS = 12


def oracle_cfg(cfg):
    vc = cfg.vision_config()
    return O.Mm2sgCfg(
        vit=O.VitCfg(hidden=vc["hidden_size"], heads=vc["num_attention_heads"], layers=vc["num_hidden_layers"],
                     ffn=vc["intermediate_size"], image=vc["image_size"], patch=vc["patch_size"],
                     select_layer=cfg.mm_vision_select_layer),
        pooler=O.PoolerCfg(),
        llm=O.LlmCfg(hidden=cfg.hidden_size, heads=cfg.num_attention_heads, layers=cfg.num_hidden_layers,
                     ffn=cfg.intermediate_size, vocab=cfg.vocab_size, eps=cfg.rms_norm_eps,
                     rope_theta=cfg.rope_theta, max_pos=cfg.max_position_embeddings))


def hf_generate_position_ids(mask):
    """What HF 4.31 LlamaForCausalLM.prepare_inputs_for_generation feeds at the prefill step of generate():
    position_ids = attention_mask.cumsum(-1) - 1, 1 on pads. Being non-None, it makes the reference return the
    PACKED position ids (llava_arch.py:309-338, :349-350), i.e. positions restart at 0 on the first real token.
    (With position_ids=None the reference returns None and HF falls back to arange over the padded row; the
    logits are the same -- RoPE is relative -- but the cached keys would not match the decode-step positions.)"""
    m = mask.long()
    return (m.cumsum(-1) - 1).masked_fill(m == 0, 1)


def reference_greedy(model, ids, mask, kwargs, steps):
    """Drives the reference's own forward for the decode steps (llava_arch.py:192-201 early-exit branch). HF 5.x
    hands back a DynamicCache, which that branch indexes like the 4.31 tuple cache; a subscriptable view is added."""
    from transformers import DynamicCache

    class TupleView(DynamicCache):
        def __getitem__(self, i):
            lay = self.layers[i]
            return (lay.keys, lay.values)

    out = model(input_ids=ids, attention_mask=mask, position_ids=hf_generate_position_ids(mask), use_cache=True,
                **kwargs)
    logits = [out.logits[:, -1].float()]
    toks = [logits[-1].argmax(-1)]
    past = out.past_key_values
    past.__class__ = TupleView
    # the mask the reference would carry through generate(): built from the PACKED length
    L = out.logits.shape[1]
    am = torch.zeros(ids.shape[0], L, dtype=torch.long)
    for b in range(ids.shape[0]):
        n = int(mask[b].sum()) - 1 + (out.logits.shape[1] - ids.shape[1] + 1)
        am[b, L - n:] = 1
    for _ in range(steps - 1):
        o = model(input_ids=toks[-1][:, None], attention_mask=am, past_key_values=past, use_cache=True, **kwargs)
        past = o.past_key_values
        am = torch.cat([am, torch.ones(am.shape[0], 1, dtype=am.dtype)], dim=1)
        logits.append(o.logits[:, -1].float())
        toks.append(logits[-1].argmax(-1))
    return torch.stack(toks, 1), torch.stack(logits, 1)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    cfg = gc.small_config()
    ocfg = oracle_cfg(cfg)
    sd = gc.bf16_round(gc.small_weights(cfg))          # bf16-representable weights, shared with the GPU build
    model = build_reference_model(cfg, sd)
    for name in gc.CASES:
        case = gc.make_case(cfg, name)
        model.config.tokenizer_padding_side = case["side"]
        model.config.mv_type = "learned"
        kw = dict(images=case["images"])
        if "audio" in case:
            kw["audio"] = case["audio"]
        if "segmasks" in case:
            kw["segmasks"] = case["segmasks"]
        labels = case.get("labels")
        pos_in = hf_generate_position_ids(case["attention_mask"]) if case["side"] == "left" else None
        ref = model(input_ids=case["input_ids"], attention_mask=case["attention_mask"], position_ids=pos_in,
                    labels=labels, **kw)
        # reference intermediates through its own modules
        concat = torch.cat(case["images"], 0)
        feats = model.get_vision_tower()(concat)
        visual = model.encode_images_pooled(concat, [im.shape[0] for im in case["images"]], None, kw.get("audio"),
                                            kw.get("segmasks"))
        # oracle on the same inputs
        orc = O.multimodal_prefill(sd, ocfg, case["input_ids"], case["attention_mask"], case["images"], labels,
                                   kw.get("audio"), kw.get("segmasks"), padding_side=case["side"])
        ofeats = O.clip_tower_forward(sd, concat, ocfg.vit)

        def rel(a, b):
            return ((a - b).norm() / (b.norm() + 1e-12)).item()

        errs = {"vit": rel(ofeats, feats), "visual": rel(orc["visual"], visual),
                # rows that are padding carry no meaning (their position ids differ: arange vs 0) -> real rows only
                "logits": rel(orc["logits"][orc["mask"]], ref.logits.float()[orc["mask"]])}
        print(name, "oracle-vs-reference rel errors:", errs)
        assert max(errs.values()) < 2e-4, errs
        if labels is not None:
            assert torch.equal(orc["modified_labels"], ref["modified_labels"])
        g = {"logits_last": ref.logits[:, -1].float().clone(),
             "logits_rows": ref.logits[:, ::37].float().half(),           # every 37th position, all vocab
             "vit_slice": feats[:, ::48, ::8].half(), "visual_slice": visual[:, ::24, ::4].half(),
             "visual_tail": visual[:, 570:].half(),                        # last pooled tokens + extra-modality tokens
             "logits_shape": torch.tensor(ref.logits.shape)}
        if labels is not None:
            g["modified_labels"] = ref["modified_labels"].clone()
            g["hf_loss"] = ref.loss.float().clone()
            vw = torch.linspace(0.2, 1.0, cfg.vocab_size)
            sl = ref.logits[..., :-1, :].reshape(-1, cfg.vocab_size).float()
            tl = ref["modified_labels"][..., 1:].reshape(-1)
            g["weighted_loss"] = torch.nn.functional.cross_entropy(sl, tl, weight=vw)   # llava_trainer.py:156-167
            ow = O.weighted_ce(orc["logits"], orc["modified_labels"], vw)
            assert abs(ow.item() - g["weighted_loss"].item()) < 1e-4
        if case["side"] == "left":
            toks, lg = reference_greedy(model, case["input_ids"], case["attention_mask"], kw, GREEDY_STEPS)
            otoks, olg = O.greedy_decode(sd, ocfg, orc["logits"][:, -1], orc["kv"], orc["mask"], GREEDY_STEPS,
                                         stop_on_eos=False)
            print(name, "greedy ids reference:", toks.tolist(), "oracle:", otoks.tolist(),
                  "logit rel err", rel(olg, lg))
            assert rel(olg, lg) < 2e-4
            g["greedy_ids"] = toks
            g["greedy_logits"] = lg.half()
        torch.save(g, os.path.join(gc.GOLDEN_DIR, name + ".pt"))
        print("wrote", name, {k: tuple(v.shape) for k, v in g.items()})
    record_chain(cfg, ocfg)


def record_chain(cfg, ocfg):
    """The exact-token fixture: the reference's own greedy continuation under the `chain` weight set
    (golden_cases.CHAIN), whose ids are all distinct and whose top-2 margins dwarf bf16 rounding."""
    sd = gc.bf16_round(gc.small_weights(cfg, chain=True))
    model = build_reference_model(cfg, sd)
    model.config.tokenizer_padding_side = "left"
    model.config.mv_type = "learned"
    case = gc.make_case(cfg, "infer_left")
    kw = dict(images=case["images"])
    toks, lg = reference_greedy(model, case["input_ids"], case["attention_mask"], kw, gc.CHAIN_STEPS)
    orc = O.multimodal_prefill(sd, ocfg, case["input_ids"], case["attention_mask"], case["images"],
                               padding_side="left")
    otoks, olg = O.greedy_decode(sd, ocfg, orc["logits"][:, -1], orc["kv"], orc["mask"], gc.CHAIN_STEPS,
                                 stop_on_eos=False)
    top2 = lg.topk(2, -1).values
    margin = (top2[..., 0] - top2[..., 1]).min().item()
    err = ((olg - lg).norm() / lg.norm()).item()
    print("chain greedy ids reference:", toks.tolist(), "min top-2 margin", margin, "oracle logit rel err", err)
    assert torch.equal(toks, otoks) and err < 2e-4
    assert all(len(set(r)) == gc.CHAIN_STEPS for r in toks.tolist())
    torch.save({"greedy_ids": toks, "greedy_logits": lg.half(), "min_margin": torch.tensor(margin)},
               os.path.join(gc.GOLDEN_DIR, "infer_left_chain.pt"))


if __name__ == "__main__":
    main()
