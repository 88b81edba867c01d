"""GPU parity cases named after BASELINE.json's configs, at the small parity configuration (the headline configs[1] is
what bench.py measures at full size; configs[3] / [4] are covered by tests/test_gpu_dist.py and
tests/test_gpu_train_ops.py):
  configs[0]  single 336x336 frame (list form, one view) -> greedy 32 tokens
  configs[2]  6-view RGB + depth + seg-mask renderings (18 encoder passes per sample) + audio embedding + seg-mask
              class maps, fused multimodal token pack. Mapping: the reference has no depth-image or
              seg-mask-image input, so the 18 frames go through the ViT (checked against the oracle frame by frame),
              the 6 RGB views through the pooler, and audio / 32x32 class maps give the 1 + 3 extra tokens (T_vis = 580).
"""
import pytest
import torch

import golden_cases as gc
from helpers import TOL_E2E, TOL_STAGE, oracle_cfg, rel_err
from mm_or_b200.synth import synth_batch
from oracle import mm2sg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    torch.set_grad_enabled(False)
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    cfg = gc.small_config()
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    model.config.tokenizer_padding_side = "left"
    return cfg, oracle_cfg(cfg), sd, model


def test_config0_single_frame_greedy_32(env):
    cfg, ocfg, sd, model = env
    b = synth_batch(cfg, 1, 1, 64, seed=50, jitter=0, image_pos=40)
    ids, images = b["input_ids"], b["images"]                      # images: list with one (1, 3, 336, 336) tensor
    assert images[0].shape[0] == 1 and int((ids == -200).sum()) == 1
    steps = 32
    out, lg = model.generate(ids, images=images, do_sample=False, use_cache=True, max_new_tokens=steps,
                             stop_on_eos=False, return_logits=True)
    ref = O.multimodal_prefill(sd, ocfg, ids, b["attention_mask"], images, padding_side="left")
    toks, ref_lg = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], steps, stop_on_eos=False)
    assert out.shape == (1, ids.shape[1] + steps) and torch.equal(out[:, :ids.shape[1]].cpu(), ids)
    assert rel_err(lg, ref_lg) < TOL_E2E
    err = (lg.cpu().float() - ref_lg).abs().max().item()
    top2 = ref_lg.topk(2, -1).values
    safe = (top2[..., 0] - top2[..., 1]) > 2 * err
    assert torch.equal(out[:, ids.shape[1]:].cpu()[safe], toks[safe])


def test_config2_rgb_depth_seg_audio_pack(env):
    cfg, ocfg, sd, model = env
    B, V = 2, 6
    b = synth_batch(cfg, B, V, 24, seed=51, jitter=3, image_pos=5, audio=True, segmasks=True)
    extra = synth_batch(cfg, B, 2 * V, 24, seed=52)["images"]      # the depth and seg-mask renderings: 12 more frames
    frames = torch.cat([torch.cat([rgb, ex], 0) for rgb, ex in zip(b["images"], extra)], 0)     # (B * 18, 3, S, S)
    assert frames.shape[0] == B * 18
    feats = model.get_vision_tower()(frames.cuda())
    ref_feats = O.clip_tower_forward(sd, frames, ocfg.vit)
    assert feats.shape == ref_feats.shape == (B * 18, 576, 1024)
    assert rel_err(feats, ref_feats) < TOL_STAGE
    # RGB views + audio + class maps through pooler, projector and pack
    out, lg = model.generate(b["input_ids"], images=b["images"], audio=b["audio"], segmasks=b["segmasks"],
                             max_new_tokens=4, stop_on_eos=False, return_logits=True)
    ref = O.multimodal_prefill(sd, ocfg, b["input_ids"], b["attention_mask"], b["images"], audio=b["audio"],
                               segmasks=b["segmasks"], padding_side="left")
    assert ref["visual"].shape[1] == 580
    toks, ref_lg = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], 4, stop_on_eos=False)
    assert rel_err(lg, ref_lg) < TOL_E2E
    pooled = model.encode_images_pooled(torch.cat(b["images"], 0).cuda(), [V] * B, None, b["audio"], b["segmasks"])
    visual = model.get_model().mm_projector(pooled)
    assert visual.shape == ref["visual"].shape and rel_err(visual, ref["visual"]) < TOL_STAGE


def test_full_width_decoder_layer():
    """configs[1] dimensions where the parity configuration is reduced: ONE decoder layer at the 7B width (hidden 4096,
    32 heads x 128, MLP 11008, vocabulary 32000 -- the tile shapes, split-K cluster sizes and the lm_head / argmax
    extent the benchmark runs) behind the full-width ViT / pooler / projector, greedy 4 tokens, against the oracle.
    (The CPU oracle in bf16 differs from its fp32 self by 1.0e-2 on this case -- the band the tolerance allows for.)"""
    torch.set_grad_enabled(False)
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    cfg = gc.small_config(hidden_size=4096, intermediate_size=11008, num_hidden_layers=1, num_attention_heads=32,
                          vocab_size=32000)
    sd = gc.bf16_round(gc.small_weights(cfg))
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    model.config.tokenizer_padding_side = "left"
    ocfg = oracle_cfg(cfg)
    b = synth_batch(cfg, 2, 1, 48, seed=53, jitter=5, image_pos=7)
    out, lg = model.generate(b["input_ids"], images=b["images"], max_new_tokens=4, stop_on_eos=False,
                             return_logits=True)
    ref = O.multimodal_prefill(sd, ocfg, b["input_ids"], b["attention_mask"], b["images"], padding_side="left")
    toks, ref_lg = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], 4, stop_on_eos=False)
    assert lg.shape == ref_lg.shape == (2, 4, 32000)
    assert rel_err(lg, ref_lg) < TOL_E2E
    err = (lg.cpu().float() - ref_lg).abs().max().item()
    top2 = ref_lg.topk(2, -1).values
    safe = (top2[..., 0] - top2[..., 1]) > 2 * err
    assert torch.equal(out[:, b["input_ids"].shape[1]:].cpu()[safe], toks[safe])


def test_rows_finished_at_eos_skip_their_cache_reads():
    """A row that has emitted EOS is padded from then on (HF greedy) and its decode attention reads no KV cache any more
    (DecodeArgs::finished, attention.cu); the rows still running must not notice. EOS is set to the first token of the
    last row, so that row stops after one step while the others run to max_new_tokens. Checked (i) bit for bit against
    the same model decoding WITHOUT the skip (return_logits=True keeps every row's logits alive, stages.cu) and
    (ii) against the oracle's HF-greedy tokens wherever the oracle's top-2 margin exceeds twice the logit error."""
    torch.set_grad_enabled(False)
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    cfg = gc.small_config()
    ocfg = oracle_cfg(cfg)
    sd = gc.bf16_round(gc.small_weights(cfg, chain=True))
    b = synth_batch(cfg, 3, 2, 24, seed=72, jitter=3, image_pos=5)  # chain weight set: oracle top-2 margins >= 3.5
    steps, Lin = 12, b["input_ids"].shape[1]
    ref = O.multimodal_prefill(sd, ocfg, b["input_ids"], b["attention_mask"], b["images"], padding_side="left")
    free, _ = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], steps, stop_on_eos=False)
    eos = int(free[2, 0])
    assert not (free[:2] == eos).any()                             # rows 0 and 1 never stop
    ref = O.multimodal_prefill(sd, ocfg, b["input_ids"], b["attention_mask"], b["images"], padding_side="left")
    toks, ref_lg = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], steps, eos_id=eos,
                                   stop_on_eos=True)
    assert toks.shape[1] == steps and (toks[2, 1:] == 0).all()      # row 2: EOS, then pad
    cfg.eos_token_id = eos
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd)
    model.config.tokenizer_padding_side = "left"
    full, lg = model.generate(b["input_ids"], images=b["images"], max_new_tokens=steps, return_logits=True)
    skip = model.generate(b["input_ids"], images=b["images"], max_new_tokens=steps)      # graph path, skip active
    assert torch.equal(skip, full)
    gen = skip[:, Lin:].cpu()
    assert gen.shape == toks.shape and (gen[2, 1:] == 0).all() and int(gen[2, 0]) == eos
    err = (lg[:2].cpu().float() - ref_lg[:2]).abs().max().item()
    top2 = ref_lg.topk(2, -1).values
    safe = (top2[..., 0] - top2[..., 1]) > 2 * err
    safe[2] = True                                                   # EOS / pad positions are exact by construction
    assert torch.equal(gen[safe], toks[safe])


def test_vis_descriptor_embeddings_in_the_pack(env):
    """forward / generate with vis_descriptor_embs (llava_llama.py:54-70, llava_arch.py:278-294): descriptor rows are
    spliced in at the VIS_DESCRIPTOR placeholders (index plan pinned bit for bit to the reference on the CPU,
    tests/test_pack_host.py); here the device side -- the third row source of the pack -- against the oracle."""
    from mm_or_b200.constants import VIS_DESCRIPTOR_TOKEN_INDEX
    cfg, ocfg, sd, model = env
    b = synth_batch(cfg, 3, 2, 26, seed=55, jitter=4, image_pos=6)
    ids = b["input_ids"].clone()
    ids[0, -3] = ids[0, -9] = VIS_DESCRIPTOR_TOKEN_INDEX           # two placeholders, one descriptor (2 rows) + dummy
    ids[1, -5] = VIS_DESCRIPTOR_TOKEN_INDEX                          # one placeholder, two descriptors (second unused)
    g = torch.Generator().manual_seed(5)
    D = cfg.hidden_size
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.02).to(torch.bfloat16).float()
    embs = [[mk(2, D)], [mk(D), mk(1, D)], []]
    out, lg = model.generate(ids, images=b["images"], vis_descriptor_embs=embs, max_new_tokens=3, stop_on_eos=False,
                             return_logits=True)
    ref = O.multimodal_prefill(sd, ocfg, ids, b["attention_mask"], b["images"], padding_side="left",
                               vis_descriptor_embs=embs)
    toks, ref_lg = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], 3, stop_on_eos=False)
    assert rel_err(lg, ref_lg) < TOL_E2E
    fw = model(input_ids=ids, attention_mask=b["attention_mask"], images=b["images"], vis_descriptor_embs=embs)
    m = ref["mask"]                                                  # pad rows carry no defined logits
    assert fw.logits.shape == ref["logits"].shape and rel_err(fw.logits.cpu()[m], ref["logits"][m]) < TOL_E2E
    # without the embeddings the text after a placeholder is dropped (the reference's behaviour): a shorter pack
    short = model(input_ids=ids, attention_mask=b["attention_mask"], images=b["images"])
    assert short.logits.shape[1] < fw.logits.shape[1]


def test_two_image_placeholders_in_one_prompt(env):
    """A prompt with two <image> placeholders takes two consecutive entries of `images` (the reference's running
    cur_image_idx, llava_arch.py:239,263-266; index plan pinned bit for bit to the reference on the CPU,
    tests/test_pack_host.py): batch of 2 rows, 3 view groups -- row 0 holds two placeholders, row 1 one."""
    cfg, ocfg, sd, model = env
    b = synth_batch(cfg, 3, 2, 20, seed=77, jitter=0, image_pos=4)
    ids = b["input_ids"][:2].clone()
    ids[0, 11] = -200                                           # second placeholder in row 0
    mask = ids.ne(0)
    out, lg = model.generate(ids, images=b["images"], attention_mask=mask, max_new_tokens=3, stop_on_eos=False,
                             return_logits=True)
    ref = O.multimodal_prefill(sd, ocfg, ids, mask, b["images"], padding_side="left")
    assert ref["mask"].shape[1] == 20 - 2 + 2 * 576 and ref["visual"].shape[0] == 3
    toks, ref_lg = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], 3, stop_on_eos=False)
    assert out.shape == (2, 20 + 3) and rel_err(lg, ref_lg) < TOL_E2E
    fw = model(input_ids=ids, attention_mask=mask, images=b["images"])
    assert rel_err(fw.logits[ref["mask"]], ref["logits"][ref["mask"]]) < TOL_E2E
    with pytest.raises(IndexError):                             # two blocks for three placeholders
        model.generate(ids, images=b["images"][:2], attention_mask=mask, max_new_tokens=1)
