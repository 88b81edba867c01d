"""CPU tests of the fine-tune host logic (mm_or_b200/train): parameter-group rules of the reference's
create_optimizer (LLaVA/llava/train/llava_trainer.py:191-278), the default trainable set of the MM2SG recipe
(train.py:1257-1261) and the un-fusing of gradients from the kernels' fused weight layouts."""
import torch

from mm_or_b200.config import LlavaConfig
from mm_or_b200.train import llama as T
from mm_or_b200.train.step import FineTuner, default_trainable, no_decay

VIT = "model.vision_tower.vision_tower.vision_model."


def test_no_decay_matches_reference_groups():
    # decay = parameters outside LayerNorm-type modules whose name has no "bias" (llava_trainer.py:205-206;
    # LlamaRMSNorm is registered in ALL_LAYERNORM_LAYERS by HF)
    decayed = ["model.layers.3.self_attn.q_proj.weight", "model.layers.0.mlp.down_proj.weight", "lm_head.weight",
               "model.mm_projector.0.weight", "model.image_pooler.bert.embeddings.position_embeddings.weight",
               "model.image_pooler.bert.encoder.layer.1.intermediate.dense.weight",
               VIT + "encoder.layers.20.mlp.fc1.weight"]
    not_decayed = ["model.layers.3.input_layernorm.weight", "model.layers.3.post_attention_layernorm.weight",
                   "model.norm.weight", "model.mm_projector.2.bias", VIT + "encoder.layers.20.layer_norm1.weight",
                   VIT + "encoder.layers.20.self_attn.k_proj.bias", VIT + "pre_layrnorm.weight",
                   "model.image_pooler.bert.embeddings.LayerNorm.weight",
                   "model.image_pooler.bert.encoder.layer.0.attention.output.LayerNorm.bias"]
    assert not any(no_decay(n) for n in decayed)
    assert all(no_decay(n) for n in not_decayed)


def test_default_trainable_set():
    yes = ["model.layers.0.self_attn.v_proj.weight", "model.norm.weight", "lm_head.weight", "model.mm_projector.0.bias",
           "model.image_pooler.bert.encoder.layer.0.output.dense.weight", VIT + "encoder.layers.12.mlp.fc2.bias",
           VIT + "encoder.layers.22.layer_norm2.weight", "model.image_pooler.project_audio.weight",
           "model.image_pooler.segmasks_encoder.conv3.bias", "model.image_pooler.segmasks_encoder.embedding.weight"]
    no = ["model.embed_tokens.weight", VIT + "encoder.layers.11.mlp.fc2.bias", VIT + "embeddings.class_embedding",
          VIT + "pre_layrnorm.weight", "model.image_pooler.bert.embeddings.word_embeddings.weight",
          "model.image_pooler.bert.pooler.dense.weight",
          "model.image_pooler.point_transformer.enc.enc0.block0.attn.qkv.weight"]
    assert all(default_trainable(n, 12) for n in yes)
    assert not any(default_trainable(n, 12) for n in no)
    # layers past the selected hidden state never run; PointTransformerV3 has no training path (stays frozen)
    assert not FineTuner._has_backward(VIT + "encoder.layers.23.mlp.fc1.weight", 23)
    assert FineTuner._has_backward(VIT + "encoder.layers.22.mlp.fc1.weight", 23)
    assert FineTuner._has_backward("model.image_pooler.segmasks_encoder.conv1.weight", 23)
    assert FineTuner._has_backward("model.image_pooler.project_audio.weight", 23)
    assert not FineTuner._has_backward("model.image_pooler.point_transformer.project_pc.weight", 23)


def test_unfuse_grads_inverts_the_fused_layouts():
    cfg = LlavaConfig(hidden_size=16, intermediate_size=24, num_hidden_layers=2, num_attention_heads=2, vocab_size=32)
    D, F = cfg.hidden_size, cfg.intermediate_size
    g = torch.Generator().manual_seed(0)
    ref, fused = {}, {}
    for i in range(cfg.num_hidden_layers):
        r = f"model.layers.{i}."
        for n in ("q", "k", "v", "o"):
            ref[r + f"self_attn.{n}_proj.weight"] = torch.randn(D, D, generator=g)
        ref[r + "mlp.gate_proj.weight"] = torch.randn(F, D, generator=g)
        ref[r + "mlp.up_proj.weight"] = torch.randn(F, D, generator=g)
        ref[r + "mlp.down_proj.weight"] = torch.randn(D, F, generator=g)
        ref[r + "input_layernorm.weight"] = torch.randn(D, generator=g)
        f = f"_fused.layers.{i}."
        # the layouts load_state_dict builds: qkv concatenated, gate/up row-interleaved
        fused[f + "qkv_w"] = torch.cat([ref[r + f"self_attn.{n}_proj.weight"] for n in ("q", "k", "v")])
        fused[f + "o_w"] = ref[r + "self_attn.o_proj.weight"]
        fused[f + "gate_up_w"] = torch.stack([ref[r + "mlp.gate_proj.weight"], ref[r + "mlp.up_proj.weight"]],
                                             dim=1).reshape(2 * F, D)
        fused[f + "down_w"] = ref[r + "mlp.down_proj.weight"]
        fused[r + "input_layernorm.weight"] = ref[r + "input_layernorm.weight"]
    out = T.unfuse_grads(fused, cfg)
    assert set(out) == set(ref)
    for k in ref:
        assert torch.equal(out[k], ref[k]), k


def test_lora_fused_adapters_equal_per_projection_adapters():
    """B_blk A_cat of a fused weight == the per-projection B_j A_j laid out like the fused base weight (q|k|v
    concatenated, gate / up row-interleaved), and unfuse_grads is the inverse mapping of the fusing."""
    from mm_or_b200.train.lora import FUSED, LoraState, param_name
    cfg = LlavaConfig(hidden_size=16, intermediate_size=24, num_hidden_layers=2, num_attention_heads=2, vocab_size=32)
    st = LoraState(cfg, r=8, alpha=16, device="cpu", seed=1, init_b="random")
    assert st.scale == 2.0 and len(st.names()) == 2 * 7 * 2
    D, F, r = cfg.hidden_size, cfg.intermediate_size, 8
    for i in range(2):
        delta = {p: st.sd[param_name(i, m, p, "B")].float() @ st.sd[param_name(i, m, p, "A")].float()
                 for projs, m in FUSED.values() for p in projs}
        a, b = st.fused[i]["qkv_w"]
        assert torch.allclose(b.float() @ a.float(), torch.cat([delta["q_proj"], delta["k_proj"], delta["v_proj"]]))
        a, b = st.fused[i]["gate_up_w"]
        want = torch.stack([delta["gate_proj"], delta["up_proj"]], dim=1).reshape(2 * F, D)
        assert torch.allclose(b.float() @ a.float(), want)
        a, b = st.fused[i]["down_w"]
        assert torch.allclose(b.float() @ a.float(), delta["down_proj"])
    # gradient un-fusing: pretend the fused tensors themselves are the gradients
    g = {}
    for i in range(2):
        for fname in FUSED:
            a, b = st.fused[i][fname]
            g[f"_lora.layers.{i}.{fname}.A"], g[f"_lora.layers.{i}.{fname}.B"] = a.float(), b.float()
    back = st.unfuse_grads(g)
    assert set(back) == set(st.sd)
    for k in st.sd:
        assert torch.equal(back[k], st.sd[k].float()), k
    merged = st.merged_state_dict({f"model.layers.{i}.{m}.{p}.weight": torch.zeros(st.shapes[p])
                                   for i in range(2) for projs, m in FUSED.values() for p in projs})
    k = "model.layers.1.mlp.up_proj.weight"
    assert torch.allclose(merged[k], 2.0 * st.sd[param_name(1, "mlp", "up_proj", "B")].float()
                          @ st.sd[param_name(1, "mlp", "up_proj", "A")].float())


def test_parameter_groups_match_reference_create_optimizer():
    """decay / no-decay and projector-lr membership for every parameter the fine-tune step can train, against what the
    reference's LLaVATrainer.create_optimizer computes on its own model (tests/golden/make_optimizer_golden.py)."""
    import json
    import os
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optimizer_groups.json")))
    checked = 0
    for name, r in ref.items():
        trainable = (default_trainable(name, 0) or name == "model.embed_tokens.weight") and \
            FineTuner._has_backward(name, 10 ** 6)
        if not trainable:
            continue
        assert no_decay(name) == (not r["decay"]), name
        assert name.startswith("model.mm_projector.") == r["projector"], name
        checked += 1
    assert checked > 100


def test_cosine_schedule_matches_transformers():
    """train/schedule.py against transformers.get_cosine_schedule_with_warmup driving a torch optimizer (the scheduler the
    reference's HF Trainer builds for `--lr_scheduler_type cosine --warmup_ratio 0.03`)."""
    import transformers
    from mm_or_b200.train.schedule import cosine_with_warmup, warmup_steps
    for total, ratio in ((100, 0.03), (37, 0.1), (5, 0.0), (1000, 0.03)):
        w = warmup_steps(total, ratio)
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=2e-5)
        sched = transformers.get_cosine_schedule_with_warmup(opt, num_warmup_steps=w, num_training_steps=total)
        for step in range(total):
            assert abs(opt.param_groups[0]["lr"] - 2e-5 * cosine_with_warmup(step, total, w)) < 1e-12, (total, step)
            opt.step()
            sched.step()
    assert warmup_steps(100, 0.03) == 3 and warmup_steps(1000, 0.03) == 30


def test_vocab_weights_match_reference_fixture():
    """train/loss_weights.py on the reference's own token-frequency file against the vector recorded with the
    reference's arithmetic (tests/golden/make_token_weight_golden.py)."""
    import os
    from mm_or_b200.train.loss_weights import vocab_weight_from_frequencies
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "token_weights.pt"))
    vocab = dict(fx["piece_ids"])
    used = set(vocab.values())
    for i in range(32000):
        if i not in used:
            vocab[f"<filler{i}>"] = i
    w = vocab_weight_from_frequencies(fx["frequencies"], vocab)
    assert w.shape == (32000,) and w.dtype == torch.float32
    for i, v in fx["vocab_weight_nonextra"].items():
        assert float(w[i]) == v, i                                   # bit-identical fp32 values
    assert float(w[0]) == fx["extra"] and abs(float(w.double().sum()) - fx["sum"]) < 1e-9
    import pytest
    with pytest.raises(KeyError):
        vocab_weight_from_frequencies({"not a piece": 3}, vocab)
