"""CPU: the files FineTuner.save_checkpoint writes (mm_or_b200/train/checkpoint.py) follow the reference's output_dir
layout (LLaVA/llava/train/train.py:1346-1360) closely enough that the loaders -- this repo's and the key handling of the
reference's model/builder.py:81-94 -- read back exactly the trained weights."""
import json
import os

import pytest
import torch

import golden_cases as gc
from mm_or_b200.model import builder as B
from mm_or_b200.train import checkpoint as C
from mm_or_b200.train.lora import FUSED, param_name


def _lora_sd(cfg, r, seed=3):
    g = torch.Generator().manual_seed(seed)
    D, F = cfg.hidden_size, cfg.intermediate_size
    shapes = {"q_proj": (D, D), "k_proj": (D, D), "v_proj": (D, D), "o_proj": (D, D), "gate_proj": (F, D),
              "up_proj": (F, D), "down_proj": (D, F)}
    sd = {}
    for i in range(cfg.num_hidden_layers):
        for projs, module in FUSED.values():
            for p in projs:
                out_f, in_f = shapes[p]
                sd[param_name(i, module, p, "A")] = (torch.randn(r, in_f, generator=g) * 0.05).to(torch.bfloat16)
                sd[param_name(i, module, p, "B")] = (torch.randn(out_f, r, generator=g) * 0.05).to(torch.bfloat16)
    return sd


def test_lora_checkpoint_reads_back_through_the_loader(tmp_path):
    cfg = gc.small_config()
    base = {k: v.to(torch.bfloat16) for k, v in gc.small_weights(cfg).items()}
    r, alpha = 8, 16
    lora = _lora_sd(cfg, r)
    trained = {k: (v.float() + 0.01).to(torch.bfloat16) for k, v in base.items()
               if k.startswith(("model.mm_projector.", "model.image_pooler.project_audio."))}
    out = C.export_lora_checkpoint(str(tmp_path / "out"), cfg, lora, r, alpha, trained, "liuhaotian/llava-v1.5-7b")
    assert sorted(os.listdir(out)) == ["adapter_config.json", "adapter_model.bin", "config.json",
                                       "non_lora_trainables.bin"]
    acfg = json.load(open(os.path.join(out, "adapter_config.json")))
    assert acfg["r"] == r and acfg["lora_alpha"] == alpha and acfg["peft_type"] == "LORA"
    assert sorted(acfg["target_modules"]) == sorted(["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj",
                                                     "down_proj"])                       # find_all_linear_names
    # key conventions: peft 0.4 adapter names and PeftModel.named_parameters() names
    adapter = torch.load(os.path.join(out, "adapter_model.bin"), weights_only=True)
    assert "base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight" in adapter
    nlt = torch.load(os.path.join(out, "non_lora_trainables.bin"), weights_only=True)
    assert "base_model.model.model.mm_projector.0.weight" in nlt
    # the reference's own two remap lines (model/builder.py:81-83), restated
    ref = {(k[11:] if k.startswith("base_model.") else k): v for k, v in nlt.items()}
    if any(k.startswith("model.model.") for k in ref):
        ref = {(k[6:] if k.startswith("model.") else k): v for k, v in ref.items()}
    assert sorted(ref) == sorted(trained) and all(torch.equal(ref[k], trained[k]) for k in trained)
    # this repo's loader path: base + non-LoRA trainables + merged adapters
    sd = dict(base)
    sd.update(B.remap_non_lora_trainables(nlt))
    merged = B.merge_lora(sd, adapter, acfg)
    assert merged == 7 * cfg.num_hidden_layers
    for k in trained:
        assert torch.equal(sd[k], trained[k])
    k = "model.layers.1.mlp.down_proj.weight"
    want = base[k].float() + (alpha / r) * (lora[param_name(1, "mlp", "down_proj", "B")].float()
                                            @ lora[param_name(1, "mlp", "down_proj", "A")].float())
    assert torch.equal(sd[k], want.to(torch.bfloat16))
    cfg2 = type(cfg).from_pretrained(out)
    assert cfg2.hidden_size == cfg.hidden_size and cfg2.mm_projector_type == cfg.mm_projector_type
    with pytest.raises(ValueError):
        C.export_lora_checkpoint(str(tmp_path / "bad"), cfg, {"model.norm.weight": base["model.norm.weight"]}, r,
                                 alpha, {})


def test_full_checkpoint_round_trip_single_and_sharded(tmp_path):
    cfg = gc.small_config()
    sd = {k: v.to(torch.bfloat16) for k, v in gc.small_weights(cfg).items()}
    one = C.export_full_checkpoint(str(tmp_path / "one"), cfg, sd)
    assert os.path.exists(os.path.join(one, "pytorch_model.bin"))
    many = C.export_full_checkpoint(str(tmp_path / "many"), cfg, sd, shard_bytes=8 << 20)
    index = json.load(open(os.path.join(many, "pytorch_model.bin.index.json")))
    assert len(set(index["weight_map"].values())) > 2 and sorted(index["weight_map"]) == sorted(sd)
    for path in (one, many):
        back = B.read_checkpoint_dir(path)
        assert sorted(back) == sorted(sd) and all(torch.equal(back[k], sd[k]) for k in sd)


def test_optimizer_state_round_trip_and_mismatch(tmp_path):
    g = torch.Generator().manual_seed(1)
    names = {"a.weight": (4, 3), "b.bias": (5,)}
    mk = lambda: {k: torch.randn(s, generator=g) for k, s in names.items()}
    master, m, v = mk(), mk(), mk()
    path = str(tmp_path / "b200_optimizer.pt")
    C.save_optimizer(path, master, m, v, 17, 2e-5, 1e-4)
    m2, mm2, v2 = ({k: torch.zeros(s) for k, s in names.items()} for _ in range(3))
    assert C.load_optimizer(path, m2, mm2, v2) == (17, 2e-5, 1e-4)
    assert all(torch.equal(m2[k], master[k]) and torch.equal(mm2[k], m[k]) and torch.equal(v2[k], v[k]) for k in names)
    with pytest.raises(KeyError):
        C.load_optimizer(path, {"a.weight": torch.zeros(4, 3)}, mm2, v2)
    with pytest.raises(ValueError):
        C.load_optimizer(path, {"a.weight": torch.zeros(4, 3), "b.bias": torch.zeros(6)}, mm2, v2)


def test_lora_checkpoint_holds_the_whole_image_pooler_state():
    """The reference's loader feeds the stripped `model.image_pooler.*` entries of non_lora_trainables.bin to
    image_pooler.load_state_dict(strict=True) (model/builder.py:160-176), so the file must hold EVERY parameter and
    buffer of the pooler -- frozen ones (point_transformer.*, bert.pooler.*, word_embeddings) included. The required key
    set is recorded from the reference's own ImageEmbeddingPooler (tests/golden/make_optimizer_golden.py)."""
    from oracle import ptv3_oracle as P
    cfg = gc.small_config()
    sd = gc.small_weights(cfg)
    sd.update(P.synth_weights())
    want = set(json.load(open(os.path.join(gc.GOLDEN_DIR, "pooler_state_keys.json"))))
    exported = C.pooler_checkpoint_tensors(sd)
    got = {k[len(C.POOLER_PREFIX):] for k in exported}
    # the two buffers the reference's loader pops before its strict load (model/builder.py:173-174) may be absent
    optional = {"bert.embeddings.position_ids", "bert.embeddings.token_type_ids"}
    assert got - optional == want - optional, (sorted(want - got)[:5], sorted(got - want)[:5])
    assert {"point_transformer.embedding.stem.norm.num_batches_tracked", "bert.pooler.dense.weight",
            "bert.embeddings.word_embeddings.weight"} <= got
    assert exported[C.POOLER_PREFIX + "point_transformer.embedding.stem.norm.num_batches_tracked"].dtype == torch.long
    # and the adapter config carries the dropout that was trained with
    assert C.adapter_config(8, 16, dropout=0.0)["lora_dropout"] == 0.0
