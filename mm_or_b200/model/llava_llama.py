"""Host-side mirror of the reference's multimodal model classes, driving libb200mmor.so.

Reference interface mirrored (paths relative to LLaVA/llava/):
  model/language_model/llava_llama.py:38   LlavaLlamaForCausalLM.forward / prepare_inputs_for_generation
  model/llava_arch.py:94                   LlavaMetaForCausalLM.encode_images_pooled / pad_embeddings /
                                           prepare_inputs_labels_for_multimodal
  model/multimodal_encoder/clip_encoder.py:7   CLIPVisionTower
  model/multimodal_projector/builder.py:61     ImageEmbeddingPooler, :40 build_vision_projector
Same constructor / call signatures, same state_dict key names, same error behaviour; every tensor operation runs
in the CUDA library (no torch math on the hot path, no CPU fallback). torch owns memory, streams and CUDA graphs.
"""
import ctypes
import math

import numpy as np
import torch

from .. import _lib as L
from ..config import LlavaConfig
from ..constants import IGNORE_INDEX, IMAGE_TOKEN_INDEX
from ..synth import POOLER_GEOMETRY
from .pack import DESC_BASE, descriptor_row_counts, plan_pack

VIT = "model.vision_tower.vision_tower.vision_model."
POOL = "model.image_pooler."
BF = torch.bfloat16


class ModelOutput(dict):
    """Attribute + item access like HF CausalLMOutputWithPast (llava_llama.py:105 adds 'modified_labels')."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _dev(t, device):
    return t.detach().to(device=device, dtype=BF).contiguous()


def _i32(a, device):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int32).to(device, non_blocking=True)


# =====================================================================================================================
# CLIP vision tower
# =====================================================================================================================
class CLIPVisionTower:
    """Mirror of CLIPVisionTower (clip_encoder.py:7-103)."""

    def __init__(self, vision_tower, args, delay_load=False):
        self.is_loaded = False
        self.vision_tower_name = vision_tower
        self.select_layer = args.mm_vision_select_layer
        self.select_feature = getattr(args, "mm_vision_select_feature", "patch")
        self.cfg = args.vision_config()
        self._w = None
        self._ws = L.Workspace()
        self.device = torch.device("cuda")
        self.dtype = BF
        self.image_processor = None
        if not delay_load:
            self.load_model()

    def load_model(self):
        # the reference downloads CLIP weights here (clip_encoder.py:22-27); offline the weights arrive through
        # load_state_dict under the same key prefix. The image processor is built lazily (needs transformers).
        self.is_loaded = True

    def get_image_processor(self):
        if self.image_processor is None:
            from transformers import CLIPImageProcessor
            s = self.cfg["image_size"]
            self.image_processor = CLIPImageProcessor(size={"shortest_edge": s}, crop_size={"height": s, "width": s})
        return self.image_processor

    @property
    def hidden_size(self):
        return self.cfg["hidden_size"]

    @property
    def num_patches(self):
        return (self.cfg["image_size"] // self.cfg["patch_size"]) ** 2

    def n_layers_run(self):
        n = self.cfg["num_hidden_layers"]
        sel = self.select_layer
        idx = n + 1 + sel if sel < 0 else sel       # index into the (n + 1)-tuple of hidden states
        if not 0 <= idx <= n:
            raise ValueError(f"mm_vision_select_layer {sel} out of range")
        return idx

    def load_weights(self, sd, device):
        c = self.cfg
        d, P = c["hidden_size"], c["patch_size"]
        hd = d // c["num_attention_heads"]
        f32 = lambda k: sd[VIT + k].detach().to(device=device, dtype=torch.float32)
        k_raw = 3 * P * P
        kpad = (k_raw + 7) // 8 * 8
        pw = torch.zeros(d, kpad, device=device, dtype=torch.float32)
        pw[:, :k_raw] = f32("embeddings.patch_embedding.weight").reshape(d, k_raw)
        pos = f32("embeddings.position_embedding.weight").to(BF).float()
        pos[0] += f32("embeddings.class_embedding").to(BF).float()
        t = {"patch_w": pw.to(BF), "pos_cls": pos.to(BF).contiguous(),
             "pre_ln_w": f32("pre_layrnorm.weight").to(BF), "pre_ln_b": f32("pre_layrnorm.bias").to(BF)}
        n_run = self.n_layers_run()
        layers = (L.VitLayer * max(n_run, 1))()
        keep = []
        scale = hd ** -0.5
        for i in range(n_run):
            p = f"encoder.layers.{i}."
            qw, qb = f32(p + "self_attn.q_proj.weight") * scale, f32(p + "self_attn.q_proj.bias") * scale
            lt = {
                "ln1_w": f32(p + "layer_norm1.weight"), "ln1_b": f32(p + "layer_norm1.bias"),
                "qkv_w": torch.cat([qw, f32(p + "self_attn.k_proj.weight"), f32(p + "self_attn.v_proj.weight")]),
                "qkv_b": torch.cat([qb, f32(p + "self_attn.k_proj.bias"), f32(p + "self_attn.v_proj.bias")]),
                "out_w": f32(p + "self_attn.out_proj.weight"), "out_b": f32(p + "self_attn.out_proj.bias"),
                "ln2_w": f32(p + "layer_norm2.weight"), "ln2_b": f32(p + "layer_norm2.bias"),
                "fc1_w": f32(p + "mlp.fc1.weight"), "fc1_b": f32(p + "mlp.fc1.bias"),
                "fc2_w": f32(p + "mlp.fc2.weight"), "fc2_b": f32(p + "mlp.fc2.bias"),
            }
            lt = {k: v.to(BF).contiguous() for k, v in lt.items()}
            keep.append(lt)
            for k, v in lt.items():
                setattr(layers[i], k, v.data_ptr())
        w = L.VitWeights(hidden=d, heads=c["num_attention_heads"], ffn=c["intermediate_size"],
                         image_size=c["image_size"], patch=P, kpad=kpad, n_layers=n_run,
                         ln_eps=c.get("layer_norm_eps", 1e-5), patch_w=t["patch_w"].data_ptr(),
                         pos_cls=t["pos_cls"].data_ptr(), pre_ln_w=t["pre_ln_w"].data_ptr(),
                         pre_ln_b=t["pre_ln_b"].data_ptr(), layers=layers)
        self._w, self._keep = w, (t, keep, layers)
        self.device = torch.device(device)

    def hidden(self, pixels, chunk=96):
        """pixels (N, 3, S, S) -> selected hidden state (N, 1 + P, D), CLS row included."""
        if self._w is None:
            raise L.B200Error("vision tower weights not loaded")
        pixels = pixels.to(device=self.device, dtype=BF).contiguous()
        N = pixels.shape[0]
        T, D = self.num_patches + 1, self.hidden_size
        out = torch.empty((N, T, D), device=self.device, dtype=BF)
        lib = L.lib()
        for s in range(0, N, chunk):
            n = min(chunk, N - s)
            nb = lib.b200_vit_workspace_bytes(ctypes.byref(self._w), n)
            ws = self._ws.get(nb, self.device)
            L.check(lib.b200_vit_forward(ctypes.byref(self._w), L.ptr(pixels[s:]), L.ptr(out[s:]), n, L.ptr(ws),
                                         ws.numel(), L.stream_ptr()), "b200_vit_forward")
        return out

    def feature_select(self, hidden):
        if self.select_feature == "patch":
            return hidden[:, 1:]
        if self.select_feature == "cls_patch":
            return hidden
        raise ValueError(f"Unexpected select feature: {self.select_feature}")   # clip_encoder.py:37

    def forward(self, images):
        if type(images) is list:
            return [self.feature_select(self.hidden(im.unsqueeze(0))).to(im.dtype) for im in images]
        return self.feature_select(self.hidden(images)).to(images.dtype)

    __call__ = forward

    def to(self, *a, **k):
        return self


# =====================================================================================================================
# image pooler (+ audio / seg-mask tokens) and projector
# =====================================================================================================================
class ImageEmbeddingPooler:
    """Mirror of ImageEmbeddingPooler (multimodal_projector/builder.py:61-190): BERT pooler over the views' tokens plus
    the point-cloud (PointTransformerV3, model/point_transformer.py), audio and seg-mask tokens."""

    def __init__(self):
        self.embedding_dim = POOLER_GEOMETRY["hidden"]
        self.geo = dict(POOLER_GEOMETRY)
        self._w = None
        self._ws = L.Workspace()
        self._ws_seg = L.Workspace()
        self.device = torch.device("cuda")

    def load_weights(self, sd, device):
        g = self.geo
        f32 = lambda k: sd[POOL + k].detach().to(device=device, dtype=torch.float32)
        b = "bert."
        pos_type = (f32(b + "embeddings.position_embeddings.weight") +
                    f32(b + "embeddings.token_type_embeddings.weight")[0]).to(BF).contiguous()
        t = {"pos_type": pos_type, "emb_ln_w": f32(b + "embeddings.LayerNorm.weight").to(BF),
             "emb_ln_b": f32(b + "embeddings.LayerNorm.bias").to(BF)}
        layers = (L.BertLayer * g["layers"])()
        keep = []
        for i in range(g["layers"]):
            p = b + f"encoder.layer.{i}."
            lt = {
                "qkv_w": torch.cat([f32(p + f"attention.self.{n}.weight") for n in ("query", "key", "value")]),
                "qkv_b": torch.cat([f32(p + f"attention.self.{n}.bias") for n in ("query", "key", "value")]),
                "ao_w": f32(p + "attention.output.dense.weight"), "ao_b": f32(p + "attention.output.dense.bias"),
                "ao_ln_w": f32(p + "attention.output.LayerNorm.weight"),
                "ao_ln_b": f32(p + "attention.output.LayerNorm.bias"),
                "fc1_w": f32(p + "intermediate.dense.weight"), "fc1_b": f32(p + "intermediate.dense.bias"),
                "fc2_w": f32(p + "output.dense.weight"), "fc2_b": f32(p + "output.dense.bias"),
                "out_ln_w": f32(p + "output.LayerNorm.weight"), "out_ln_b": f32(p + "output.LayerNorm.bias"),
            }
            lt = {k: v.to(BF).contiguous() for k, v in lt.items()}
            keep.append(lt)
            for k, v in lt.items():
                setattr(layers[i], k, v.data_ptr())
        self._w = L.PoolerWeights(hidden=g["hidden"], heads=g["heads"], ffn=g["ffn"], n_layers=g["layers"],
                                  max_pos=pos_type.shape[0], ln_eps=g["eps"], pos_type=pos_type.data_ptr(),
                                  emb_ln_w=t["emb_ln_w"].data_ptr(), emb_ln_b=t["emb_ln_b"].data_ptr(), layers=layers)
        self.audio_w = _dev(sd[POOL + "project_audio.weight"], device)
        self.audio_b = _dev(sd[POOL + "project_audio.bias"], device)
        self._seg = None
        if POOL + "segmasks_encoder.conv1.weight" in sd:
            s = {"emb": _dev(sd[POOL + "segmasks_encoder.embedding.weight"], device)}
            sw = L.SegmaskWeights(emb=s["emb"].data_ptr())
            for i in range(5):
                s[f"w{i}"] = _dev(sd[POOL + f"segmasks_encoder.conv{i + 1}.weight"], device)
                s[f"b{i}"] = _dev(sd[POOL + f"segmasks_encoder.conv{i + 1}.bias"], device)
                sw.conv_w[i] = s[f"w{i}"].data_ptr()
                sw.conv_b[i] = s[f"b{i}"].data_ptr()
            self._seg = (sw, s)
        self._keep = (t, keep, layers)
        self.device = torch.device(device)
        # point-cloud branch: present when the checkpoint carries it (builder.py:82-85 always constructs it)
        self.point_transformer = None
        if POOL + "point_transformer.project_pc.weight" in sd:
            from .point_transformer import PointTransformerV3
            self.point_transformer = PointTransformerV3().load_weights(sd, POOL + "point_transformer.", device)

    @staticmethod
    def num_extra_tokens(pc, audio, segmasks):
        # a modality contributes its token(s) to EVERY sample as soon as the kwarg is not None (builder.py:176-189)
        return (1 if pc is not None else 0) + (1 if audio is not None else 0) + (3 if segmasks is not None else 0)

    def pooled_tokens(self, src, src_ld, gather_map, kv_len, B, S, pc=None, audio=None, segmasks=None, chunk=16):
        """BERT pooler over gathered rows of `src` + extra modality tokens -> (B, T_vis, hidden)."""
        if self._w is None:
            raise L.B200Error("image pooler weights not loaded")
        if pc is not None and self.point_transformer is None:
            raise L.B200Error("point clouds were passed but the checkpoint has no model.image_pooler.point_transformer.* "
                              "weights")
        if pc is not None and len(pc) != B:
            raise ValueError(f"pc must hold one entry (tensor or None) per sample: got {len(pc)} for batch {B}")
        g = self.geo
        keep, D = g["keep"], g["hidden"]
        if S < keep:
            raise ValueError(f"pooler keeps the first {keep} tokens but only {S} were given")
        T = keep + self.num_extra_tokens(pc, audio, segmasks)
        out = torch.zeros((B, T, D), device=self.device, dtype=BF)
        lib = L.lib()
        for s in range(0, B, chunk):
            n = min(chunk, B - s)
            nb = lib.b200_pooler_workspace_bytes(ctypes.byref(self._w), n, S)
            ws = self._ws.get(nb, self.device)
            L.check(lib.b200_pooler_forward(ctypes.byref(self._w), L.ptr(src), src_ld, L.ptr(gather_map[s * S:]),
                                            L.ptr(kv_len[s:]), n, S, keep, L.ptr(out[s:]), T, L.ptr(ws), ws.numel(),
                                            L.stream_ptr()), "b200_pooler_forward")
        t = keep
        if pc is not None:                                            # _encode_pc (builder.py:93-148), fp32 -> bf16 (:177)
            self.point_transformer(pc, out=out[:, t])
            t += 1
        if audio is not None:                                         # _encode_audio (builder.py:150-159)
            feats = torch.zeros((B, 512), dtype=BF)
            for i, a in enumerate(audio):
                if a is not None:
                    feats[i] = a.detach().to("cpu", BF)
            feats = feats.to(self.device, non_blocking=True)     # pageable source: staged at once, no stream sync
            L.gemm(feats, self.audio_w, out=out.view(B * T, D)[t::T], bias=self.audio_b)
            t += 1
        if segmasks is not None:                                      # _encode_segmasks (builder.py:161-167)
            if self._seg is None:
                raise L.B200Error("seg-mask encoder weights not loaded")
            maps, rows = [], []
            for i, sm in enumerate(segmasks):
                if sm is not None:
                    for j, m in enumerate(sm):
                        if tuple(m.shape) != (32, 32):
                            raise AssertionError(f"Expected input size (batch_size, 32, 32), but got {tuple(m.shape)}")
                        maps.append(m.detach().to("cpu", torch.uint8))
                        rows.append(i * T + t + j)
            if maps:
                cls = torch.stack(maps).contiguous().to(self.device, non_blocking=True)
                rm = _i32(np.asarray(rows, dtype=np.int32), self.device)
                nb = lib.b200_segmask_workspace_bytes(len(maps))
                ws = self._ws_seg.get(nb, self.device)
                L.check(lib.b200_segmask_forward(ctypes.byref(self._seg[0]), L.ptr(cls), len(maps), L.ptr(out), D,
                                                 L.ptr(rm), L.ptr(ws), ws.numel(), L.stream_ptr()),
                        "b200_segmask_forward")
            t += 3
        return out

    def forward(self, embeddings, attention_mask, pc=None, audio=None, segmasks=None):
        """embeddings (B, S, 1024), attention_mask (B, S) with a contiguous run of valid tokens from 0."""
        B, S, D = embeddings.shape
        emb = embeddings.to(device=self.device, dtype=BF).contiguous().view(B * S, D)
        kv_len = attention_mask.to(self.device).sum(1).to(torch.int32)
        gmap = torch.arange(B * S, device=self.device, dtype=torch.int32)
        return self.pooled_tokens(emb, D, gmap, kv_len, B, S, pc, audio, segmasks)

    __call__ = forward


class MMProjector:
    """mlp2x_gelu projector (multimodal_projector/builder.py:46-53)."""

    def __init__(self, config):
        ptype = getattr(config, "mm_projector_type", "linear")
        if ptype != "mlp2x_gelu":
            raise ValueError(f"Unknown projector type: {ptype}")      # builder.py:58 (only the MM2SG type is built)
        self.in_dim, self.hidden = config.mm_hidden_size, config.hidden_size
        self._w = None
        self._ws = L.Workspace()

    def load_weights(self, sd, device):
        self.t = {k: _dev(sd[f"model.mm_projector.{k}"], device) for k in ("0.weight", "0.bias", "2.weight", "2.bias")}
        self._w = L.ProjectorWeights(in_dim=self.in_dim, hidden=self.hidden, w0=self.t["0.weight"].data_ptr(),
                                     b0=self.t["0.bias"].data_ptr(), w2=self.t["2.weight"].data_ptr(),
                                     b2=self.t["2.bias"].data_ptr())
        self.device = torch.device(device)

    def project_pack(self, tokens, row_map, text_ids, embed_table, vocab, embeds, n_rows):
        lib = L.lib()
        n = tokens.shape[0]
        nb = lib.b200_projector_workspace_bytes(ctypes.byref(self._w), n)
        ws = self._ws.get(nb, tokens.device)
        L.check(lib.b200_projector_pack(ctypes.byref(self._w), L.ptr(tokens), n, L.ptr(row_map), L.ptr(text_ids),
                                        L.ptr(embed_table), vocab, L.ptr(embeds), n_rows, L.ptr(ws), ws.numel(),
                                        L.stream_ptr()), "b200_projector_pack")

    def forward(self, x):
        shp = x.shape
        t = x.to(device=self.device, dtype=BF).contiguous().view(-1, shp[-1])
        out = torch.empty((t.shape[0], self.hidden), device=self.device, dtype=BF)
        rm = torch.arange(t.shape[0], device=self.device, dtype=torch.int32)
        self.project_pack(t, rm, None, None, 0, out, t.shape[0])
        return out.view(*shp[:-1], self.hidden)

    __call__ = forward


# =====================================================================================================================
# language model + multimodal glue
# =====================================================================================================================
class KVCache:
    """bf16 KV cache [layers][B][H][cap][128] (replaces HF 4.31's tuple cache grown by torch.cat every step)."""

    def __init__(self, n_layers, batch, heads, cap, device):
        self.n_layers, self.batch, self.heads, self.cap = n_layers, batch, heads, cap
        self.k = torch.empty((n_layers, batch, heads, cap, 128), device=device, dtype=BF)
        self.v = torch.empty((n_layers, batch, heads, cap, 128), device=device, dtype=BF)

    def struct(self, b0=0):
        off = b0 * self.heads * self.cap * 128 * 2
        return L.KvCache(k=self.k.data_ptr() + off, v=self.v.data_ptr() + off,
                         layer_stride=self.batch * self.heads * self.cap * 128, cap=self.cap)


class PrefixCache:
    """Keys / values (layers, heads, p, 128) of a token prefix shared by the prompts of a batch (make_prefix_cache)."""

    def __init__(self, ids, k, v, pad_token_id):
        self.ids, self.k, self.v, self.p = ids.to(torch.long), k, v, int(ids.numel())
        self.pad = pad_token_id

    def plan(self, ids_cpu, attention_mask, kv_start_host, Lq):
        """First cache slot the prefill has to compute (0 = no reuse): every row must start -- after its padding --
        with the prefix ids, and at least one slot must remain to be computed."""
        am = attention_mask.bool()
        for r in range(ids_cpu.shape[0]):
            real = ids_cpu[r][am[r]]
            if real.numel() <= self.p or not torch.equal(real[:self.p], self.ids):
                return 0
        q0 = int(np.min(kv_start_host)) + self.p
        return q0 if 0 < q0 < Lq else 0

    def copy_into(self, cache, kv_start_host, q0):
        """Row b receives the first q0 - kv_start[b] prefix tokens at slots [kv_start[b], q0)."""
        for b, ks in enumerate(np.asarray(kv_start_host).tolist()):
            n = min(self.p, q0 - int(ks))
            if n > 0:
                cache.k[:, b, :, ks:ks + n] = self.k[:, :, :n]
                cache.v[:, b, :, ks:ks + n] = self.v[:, :, :n]


class DecodeSlot:
    """The device buffers one greedy generation lives in -- KV cache, current tokens, step counters, history, finished
    flags, first real slot per row -- and, once captured, the CUDA graph of one decode step over exactly these
    buffers. generate() makes a fresh one per call; generate_stream() ping-pongs two, so the graph of a slot is
    captured once and replayed for every batch that lands in it."""

    def __init__(self, cfg, batch, cap, max_new_tokens, device):
        self.key = (cfg.num_hidden_layers, batch, cfg.num_attention_heads, cap, max_new_tokens)
        self.cache = KVCache(cfg.num_hidden_layers, batch, cfg.num_attention_heads, cap, device)
        self.tokens = torch.empty(batch, device=device, dtype=torch.int32)
        self.history = torch.empty((batch, max_new_tokens), device=device, dtype=torch.int32)
        self.finished = torch.zeros(batch, device=device, dtype=torch.int32)
        self.state = torch.empty(2, device=device, dtype=torch.int32)
        self.kv_start = torch.empty(batch, device=device, dtype=torch.int32)
        self.graph = None              # (CUDAGraph, kernels per replay, stop_on_eos it was captured with)


class LlavaLlamaModel:
    def __init__(self, config):
        self.config = config
        self.vision_tower = CLIPVisionTower(config.mm_vision_tower, config, delay_load=True) \
            if getattr(config, "mm_vision_tower", None) else None
        self.image_pooler = ImageEmbeddingPooler() if self.vision_tower is not None else None
        self.mm_projector = MMProjector(config) if self.vision_tower is not None else None
        self.embed_tokens = None      # (vocab, hidden) bf16 tensor after load

    def get_vision_tower(self):
        return self.vision_tower

    def get_image_pooler(self):
        return self.image_pooler


class LlavaLlamaForCausalLM:
    """Mirror of LlavaLlamaForCausalLM (llava_llama.py:38-127)."""

    config_class = LlavaConfig

    def __init__(self, config):
        self.config = config
        self.model = LlavaLlamaModel(config)
        self.vocab_size = config.vocab_size
        self.lm_head = None
        self.device = torch.device("cuda")
        self.dtype = BF
        self._w = None
        self._ws_prefill = L.Workspace()
        self._ws_decode = L.Workspace()
        self.prefill_chunk = 16          # samples per prefill launch sequence (bounds the activation workspace)
        self.vit_chunk = 96
        self.pooler_chunk = 16
        self.training = False
        self._pg = None
        self._exchange = "peer"
        self._decode_shift = 0
        self._peer = None

    # ---- reference accessors --------------------------------------------------------------------------------
    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.model.get_vision_tower()

    def get_image_pooler(self):
        return self.model.get_image_pooler()

    def eval(self):
        self.training = False
        return self

    def to(self, *a, **k):
        return self

    def set_process_group(self, group, exchange="peer", decode_shift=0):
        """Multi-GPU inference: every rank encodes the views of its own samples; the projected visual
        tokens are all-gathered over NVLink before LLM fusion and each rank decodes the slice of the gathered batch
        that rank (rank + decode_shift) % world encoded (`input_ids` passed to generate()/forward() are that slice's).
        exchange = "peer": the second projector GEMM stores its output tiles into every rank's symmetric buffer
        (fused GEMM + all-gather, b200_projector_gather); "nccl": projector, then ncclAllGather."""
        import torch.distributed as dist
        if exchange not in ("peer", "nccl"):
            raise ValueError(f"unknown exchange {exchange!r}")
        if group is None:                 # back to single-GPU behaviour (keeps the symmetric buffer for a later re-join)
            self._pg = None
            return
        self._pg = group
        self._exchange = exchange
        self._decode_shift = int(decode_shift)
        if self._peer is not None and self._peer.group is not group:
            self._peer = None
        if group is not None and dist.get_world_size(group) > 8:
            raise ValueError("one NVSwitch box: at most 8 ranks")

    def _project_and_gather(self, pooled):
        """pooled (B, T_vis, 1024) of this rank's samples -> projected tokens (B, T_vis, D) of the samples this rank
        decodes, after the all-gather over ranks."""
        import torch.distributed as dist
        from .. import dist as D_
        world, rank = dist.get_world_size(self._pg), dist.get_rank(self._pg)
        owner = D_.decode_owner(rank, world, self._decode_shift)
        B, t_vis, _ = pooled.shape
        Dh = self.config.hidden_size
        proj = self.model.mm_projector
        if self._exchange == "peer":
            if self._peer is None or not self._peer.fits(B, t_vis, Dh):
                grow = 0 if self._peer is None else self._peer.capacity
                self._peer = None                     # release the old buffer BEFORE the new rendezvous
                self._peer = D_.PeerGather(B, t_vis, Dh, self._pg, capacity=grow)
            elif self._peer.shape != (world, B, t_vis, Dh):
                self._peer.set_shape(B, t_vis, Dh)    # same memory, other geometry: no allocation, no collective
            pg = self._peer
            pg.barrier()                              # every reader of the previous step's tokens is done
            lib = L.lib()
            n = B * t_vis
            nb = lib.b200_projector_workspace_bytes(ctypes.byref(proj._w), n)
            ws = proj._ws.get(nb, self.device)
            peers = (ctypes.c_void_p * world)(*pg.peer_ptrs)
            L.check(lib.b200_projector_gather(ctypes.byref(proj._w), L.ptr(pooled.view(n, -1)), n, peers, world,
                                              pg.slot_offset_bytes(), L.ptr(ws), ws.numel(), L.stream_ptr()),
                    "b200_projector_gather")
            pg.barrier()                              # all peers' tiles have landed in this rank's copy
            return pg.buf[owner]
        local = proj(pooled)                                                      # (B, T_vis, D)
        return D_.all_gather_tokens(local, self._pg)[owner]

    def resize_token_embeddings(self, n):
        if n != self.config.vocab_size:
            raise NotImplementedError("vocabulary resize is not supported (MM2SG keeps the 32000-entry vocabulary)")

    # ---- weights ---------------------------------------------------------------------------------------------
    def load_state_dict(self, sd, strict=False, device="cuda"):
        """Accepts the reference's key names; buffers such as rotary inv_freq / position_ids
        are ignored. Linear weights are re-laid-out once for the fused kernels (QKV concat, gate/up interleave)."""
        c = self.config
        D, F, V, H = c.hidden_size, c.intermediate_size, c.vocab_size, c.num_attention_heads
        if D // H != 128:
            raise ValueError("Llama head_dim must be 128")
        if c.num_key_value_heads != H:
            raise NotImplementedError("grouped-query attention is not used by Vicuna/Llama-7B and is not implemented")
        dev = torch.device(device)
        get = lambda k: sd[k].detach().to(device=dev, dtype=BF)
        layers = (L.LlamaLayer * c.num_hidden_layers)()
        keep = []
        for i in range(c.num_hidden_layers):
            p = f"model.layers.{i}."
            lt = {
                "attn_norm": get(p + "input_layernorm.weight").contiguous(),
                "qkv_w": torch.cat([get(p + f"self_attn.{n}_proj.weight") for n in ("q", "k", "v")]).contiguous(),
                "o_w": get(p + "self_attn.o_proj.weight").contiguous(),
                "mlp_norm": get(p + "post_attention_layernorm.weight").contiguous(),
                "gate_up_w": torch.stack([get(p + "mlp.gate_proj.weight"), get(p + "mlp.up_proj.weight")],
                                         dim=1).reshape(2 * F, D).contiguous(),
                "down_w": get(p + "mlp.down_proj.weight").contiguous(),
            }
            keep.append(lt)
            for k, v in lt.items():
                setattr(layers[i], k, v.data_ptr())
        self.model.embed_tokens = get("model.embed_tokens.weight").contiguous()
        self.lm_head = get("lm_head.weight").contiguous()
        final_norm = get("model.norm.weight").contiguous()
        hd = 128
        max_pos = c.max_position_embeddings
        inv = 1.0 / (c.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
        fr = torch.outer(torch.arange(max_pos, dtype=torch.float32), inv)
        rope_cos, rope_sin = fr.cos().contiguous().to(dev), fr.sin().contiguous().to(dev)
        self._w = L.LlamaWeights(hidden=D, heads=H, ffn=F, n_layers=c.num_hidden_layers, vocab=V, max_pos=max_pos,
                                 rms_eps=c.rms_norm_eps, layers=layers, final_norm=final_norm.data_ptr(),
                                 lm_head=self.lm_head.data_ptr(), embed_tokens=self.model.embed_tokens.data_ptr(),
                                 rope_cos=rope_cos.data_ptr(), rope_sin=rope_sin.data_ptr())
        self._keep = (keep, layers, final_norm, rope_cos, rope_sin)
        if self.model.vision_tower is not None and (VIT + "pre_layrnorm.weight") in sd:
            self.model.vision_tower.load_weights(sd, dev)
            self.model.vision_tower.is_loaded = True
            self.model.image_pooler.load_weights(sd, dev)
            self.model.mm_projector.load_weights(sd, dev)
        self.device = dev
        return self

    # ---- multimodal encoding (llava_arch.py:172-183) ------------------------------------------------------------
    def encode_images_pooled(self, images, split_sizes, pc=None, audio=None, segmasks=None):
        """images (sum V_b, 3, S, S); split_sizes [V_b]. Returns pooled tokens BEFORE the projector, (B, T_vis, 1024);
        the projector runs fused with the pack (project_pack) or through mm_projector(...)."""
        tower, pooler = self.get_vision_tower(), self.get_image_pooler()
        hidden = tower.hidden(images, chunk=self.vit_chunk)                       # (N, 1 + P, D) incl. CLS
        N, T, D = hidden.shape
        P = T - 1
        B = len(split_sizes)
        vmax = max(split_sizes)
        # pad_embeddings (llava_arch.py:143-170) as a gather map: row (b, v, p) <- hidden[img(b, v), 1 + p], -1 = zero
        gmap = np.full((B, vmax, P), -1, dtype=np.int32)
        img0 = np.concatenate([[0], np.cumsum(split_sizes)[:-1]])
        base = np.arange(P, dtype=np.int32)[None, :] + 1
        for b, vb in enumerate(split_sizes):
            gmap[b, :vb] = (img0[b] + np.arange(vb, dtype=np.int32))[:, None] * T + base
        kv_len = np.asarray(split_sizes, dtype=np.int32) * P
        return pooler.pooled_tokens(hidden.view(N * T, D), D, _i32(gmap.reshape(-1), self.device),
                                    _i32(kv_len, self.device), B, vmax * P, pc, audio, segmasks, chunk=self.pooler_chunk)

    def _images_to_batch(self, images):
        if type(images) is list or images.ndim == 5:
            if getattr(self.config, "mv_type") == "learned":
                if torch.is_tensor(images):          # (B, V, 3, S, S): one copy instead of B
                    concat = images.to(self.device, BF, non_blocking=True).flatten(0, 1)
                    return concat, [int(images.shape[1])] * int(images.shape[0])
                concat = torch.cat([im.to(self.device, BF, non_blocking=True) for im in images], dim=0)
                return concat, [int(im.shape[0]) for im in images]
            raise NotImplementedError("only mv_type == 'learned' is used by MM2SG")
        raise Exception("SHOULD NOT BE HERE")                                     # llava_arch.py:209

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels,
                                             images, vis_descriptor_embs=None, pc=None, audio=None, segmasks=None):
        """Returns (None, position_ids, attention_mask, past_key_values, inputs_embeds, labels) like
        llava_arch.py:188-353, plus the PackPlan as a 7th element for the native prefill."""
        if getattr(self.config, "tune_mm_mlp_adapter", False) and getattr(self.config, "mm_use_im_start_end", False):
            raise NotImplementedError                                             # llava_arch.py:216
        concat, split = self._images_to_batch(images)
        pooled = self.encode_images_pooled(concat, split, pc, audio, segmasks)   # (blocks, T_vis, 1024)
        NB, t_vis, _ = pooled.shape          # one block of visual tokens per entry of `images` ...
        B = int(input_ids.shape[0])          # ... usually one per row; a row with k <image> placeholders takes k
        D = self.config.hidden_size
        desc_rows = desc_table = None
        if vis_descriptor_embs is not None:
            # llava_arch.py:278-294: descriptor j of a sample replaces its j-th VIS_DESCRIPTOR placeholder. All
            # descriptors of the batch form one (rows, D) table on the device; the plan says which row goes where.
            per_sample, desc_rows = descriptor_row_counts(vis_descriptor_embs, B)
            flat = [e.detach().reshape(-1, D) for per in per_sample for e in per]
            desc_table = (torch.cat(flat).to(self.device, BF) if flat
                          else torch.zeros((1, D), device=self.device, dtype=BF)).contiguous()
        plan = plan_pack(input_ids.cpu().numpy(), None if attention_mask is None else attention_mask.cpu().numpy(),
                         None if labels is None else labels.cpu().numpy(), t_vis,
                         getattr(self.config, "tokenizer_padding_side", "right"),
                         getattr(self.config, "tokenizer_model_max_length", None), desc_rows=desc_rows, n_blocks=NB)
        embeds = torch.empty((B, plan.L, D), device=self.device, dtype=BF)
        if self._pg is None:
            self.model.mm_projector.project_pack(pooled.view(NB * t_vis, -1), _i32(plan.row_map, self.device),
                                                 _i32(plan.src.reshape(-1), self.device), self.model.embed_tokens,
                                                 self.config.vocab_size, embeds, B * plan.L)
        else:
            vis = self._project_and_gather(pooled)                                # (B, T_vis, D), gathered copy
            src = plan.src.reshape(B, plan.L).astype(np.int64)
            # two row gathers into the packed buffer: text / pad rows from embed_tokens, visual rows from `vis`
            text_ids = np.where(src >= -1, src, -2).astype(np.int32)
            lib = L.lib()
            text_dev, vis_dev = _i32(text_ids.reshape(-1), self.device), _i32(plan.vis_ids, self.device)
            L.check(lib.b200_embed_rows(L.ptr(text_dev), L.ptr(self.model.embed_tokens), L.ptr(embeds), D, B * plan.L, D,
                                        self.config.vocab_size, L.stream_ptr()), "b200_embed_rows")
            L.check(lib.b200_embed_rows(L.ptr(vis_dev), L.ptr(vis), L.ptr(embeds), D, B * plan.L, D, NB * t_vis,
                                        L.stream_ptr()), "b200_embed_rows")
        if desc_table is not None and (plan.desc_ids >= 0).any():
            desc_dev = _i32(plan.desc_ids, self.device)         # third row source: descriptor rows; -2 = untouched
            L.check(L.lib().b200_embed_rows(L.ptr(desc_dev), L.ptr(desc_table), L.ptr(embeds), D, B * plan.L, D,
                                            desc_table.shape[0], L.stream_ptr()), "b200_embed_rows")
        new_labels = None if labels is None else torch.from_numpy(plan.labels).to(self.device)
        am = None if attention_mask is None else torch.from_numpy(plan.mask).to(self.device).to(attention_mask.dtype)
        pos = None if position_ids is None else torch.from_numpy(plan.pos).to(self.device)
        return None, pos, am, past_key_values, embeds, new_labels, plan

    # ---- decoder -----------------------------------------------------------------------------------------------
    def _prefill(self, embeds, kv_start, kv_len, cache, all_logits, logits_fp32, q0=0):
        """embeds (B, L, D) (overwritten). Returns logits (B, V) or (B, L, V). q0 > 0: the rows of `embeds` are the
        tokens of cache slots [q0, q0 + L); slots below q0 already hold keys / values (prefix-KV reuse)."""
        B, Lq, D = embeds.shape
        V = self.config.vocab_size
        ldt = torch.float32 if logits_fp32 else BF
        logits = torch.empty((B, Lq, V) if all_logits else (B, V), device=self.device, dtype=ldt)
        lib = L.lib()
        for s in range(0, B, self.prefill_chunk):
            n = min(self.prefill_chunk, B - s)
            nb = lib.b200_llama_prefill_workspace_bytes(ctypes.byref(self._w), n, Lq, int(all_logits))
            ws = self._ws_prefill.get(nb, self.device)
            cs = cache.struct(s)
            L.check(lib.b200_llama_prefill_from(ctypes.byref(self._w), L.ptr(embeds[s:]),
                                                L.ptr(kv_start[s:]) if kv_start is not None else None,
                                                L.ptr(kv_len[s:]) if kv_len is not None else None, ctypes.byref(cs), n,
                                                Lq, int(q0), L.ptr(logits[s:]), int(all_logits), int(logits_fp32),
                                                L.ptr(ws), ws.numel(), L.stream_ptr()), "b200_llama_prefill")
        return logits

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                images=None, return_dict=None, vis_descriptor_embs=None, pc=None, audio=None, segmasks=None):
        """Teacher-forced / prefill forward (llava_llama.py:54-106). Returns logits for all positions (fp32),
        past_key_values (KVCache) and output['modified_labels']. Incremental decoding goes through generate()."""
        if past_key_values is not None:
            raise NotImplementedError("incremental forward with past_key_values: use generate() (native decode loop)")
        if self._w is None:
            raise L.B200Error("weights not loaded")
        plan = None
        if inputs_embeds is None:
            if images is not None and self.get_vision_tower() is not None:
                (_, position_ids, attention_mask, past_key_values, inputs_embeds, labels,
                 plan) = self.prepare_inputs_labels_for_multimodal(input_ids, position_ids, attention_mask,
                                                                   past_key_values, labels, images,
                                                                   vis_descriptor_embs, pc, audio, segmasks)
            else:
                ids = input_ids.to(self.device)
                B, Lq = ids.shape
                inputs_embeds = torch.empty((B, Lq, self.config.hidden_size), device=self.device, dtype=BF)
                L.check(L.lib().b200_embed_rows(L.ptr(ids.to(torch.int32).contiguous()), L.ptr(self.model.embed_tokens),
                                                L.ptr(inputs_embeds), self.config.hidden_size, B * Lq,
                                                self.config.hidden_size, self.config.vocab_size, L.stream_ptr()),
                        "b200_embed_rows")
        else:
            inputs_embeds = inputs_embeds.to(self.device, BF).clone()
        B, Lq, _ = inputs_embeds.shape
        kv_start = kv_len = None
        if plan is not None:
            if getattr(self.config, "tokenizer_padding_side", "right") == "left":
                kv_start = _i32(plan.kv_start, self.device)
            else:
                kv_len = _i32(plan.lengths, self.device)
        elif attention_mask is not None:
            am = attention_mask.to("cpu").bool().numpy()
            first = am.argmax(1).astype(np.int32)
            if first.any():
                kv_start = _i32(first, self.device)
            else:
                kv_len = _i32(am.sum(1).astype(np.int32), self.device)
        cache = KVCache(self.config.num_hidden_layers, B, self.config.num_attention_heads, Lq, self.device)
        logits = self._prefill(inputs_embeds, kv_start, kv_len, cache, all_logits=True, logits_fp32=True)
        loss = None
        if labels is not None:
            # HF LlamaForCausalLM.forward: mean shifted CE over labels != -100 (the trainer then recomputes its
            # class-weighted loss from output['modified_labels'], llava_trainer.py:153-169)
            loss = L.weighted_ce(logits, labels.to(self.device).contiguous(), None)[0]
        out = ModelOutput(loss=loss, logits=logits, past_key_values=cache if use_cache else None,
                          hidden_states=None, attentions=None)
        out["modified_labels"] = labels
        return out

    __call__ = forward

    @torch.no_grad()
    def generate(self, input_ids, images=None, do_sample=False, use_cache=True, max_new_tokens=20,
                 stopping_criteria=None, pc=None, audio=None, segmasks=None, attention_mask=None,
                 stop_on_eos=True, check_every=16, use_cuda_graph=True, return_logits=False,
                 vis_descriptor_embs=None, prefix_cache=None, **unused):
        """Greedy decoding with the call signature the reference uses (scene_graph_prediction_model.py:221-231).
        prefix_cache (make_prefix_cache): keys / values of a token prefix every prompt of the batch starts with (the
        fixed system text of the online mode) are copied into the KV cache instead of being recomputed.
        Returns LongTensor (B, L_in + n_new): the prompt ids (incl. the -200 placeholder) followed by the new tokens.
        HF greedy_search semantics: finished rows emit pad_token_id; stops when every row has produced EOS.
        One call = the three phases below back to back on the current stream; generate_stream() overlaps them across
        successive batches."""
        if do_sample:
            raise NotImplementedError("MM2SG decodes greedily (do_sample=False)")
        job = self._gen_begin(input_ids, images, max_new_tokens, pc, audio, segmasks, attention_mask, stop_on_eos,
                              return_logits, vis_descriptor_embs, prefix_cache=prefix_cache)
        self._gen_decode(job, stopping_criteria, check_every, use_cuda_graph)
        return self._gen_finish(job)

    # ---- the three phases of a greedy generation ------------------------------------------------------------------
    def _gen_begin(self, input_ids, images, max_new_tokens, pc=None, audio=None, segmasks=None, attention_mask=None,
                   stop_on_eos=True, return_logits=False, vis_descriptor_embs=None, slot=None, prefix_cache=None):
        """Phase A (tensor-core bound): encode the views, pack, prefill the KV cache, first token from the prefill
        logits -- everything enqueued on the current stream, no host synchronisation. Returns the job state."""
        if self._w is None:
            raise L.B200Error("weights not loaded")
        c = self.config
        lib = L.lib()
        ids_cpu = input_ids.detach().to("cpu")
        if attention_mask is None:
            # HF infers the mask from the pad token when the caller passes none (scene_graph_prediction_model.py:221)
            # (HF: an all-ones mask when the config has no pad token)
            attention_mask = (ids_cpu.ne(c.pad_token_id) if c.pad_token_id is not None
                              else torch.ones_like(ids_cpu, dtype=torch.bool))
        B = ids_cpu.shape[0]
        if images is not None and self.get_vision_tower() is not None:
            (_, _, _, _, embeds, _, plan) = self.prepare_inputs_labels_for_multimodal(
                ids_cpu, None, attention_mask, None, None, images, vis_descriptor_embs, pc, audio, segmasks)
            Lq = plan.L
            left = getattr(c, "tokenizer_padding_side", "right") == "left"
            kv_start_host = plan.kv_start if left else np.zeros(B, np.int32)
            kv_len = None if left else _i32(plan.lengths, self.device)
            if not left and (plan.lengths != Lq).any():
                raise NotImplementedError("generate() with right-padded prompts of unequal length; the reference sets "
                                          "tokenizer_padding_side='left' for inference")
        else:
            am = attention_mask.bool().numpy()
            Lq = ids_cpu.shape[1]
            ids32 = torch.where(attention_mask.bool(), ids_cpu, torch.full_like(ids_cpu, -1)).to(torch.int32)
            embeds = torch.empty((B, Lq, c.hidden_size), device=self.device, dtype=BF)
            L.check(lib.b200_embed_rows(L.ptr(ids32.to(self.device).contiguous()), L.ptr(self.model.embed_tokens),
                                        L.ptr(embeds), c.hidden_size, B * Lq, c.hidden_size, c.vocab_size,
                                        L.stream_ptr()), "b200_embed_rows")
            kv_start_host = am.argmax(1).astype(np.int32)
            kv_len = None
        cap = (Lq + max_new_tokens + 7) // 8 * 8
        if slot is None or slot.key != (c.num_hidden_layers, B, c.num_attention_heads, cap, max_new_tokens):
            slot = DecodeSlot(c, B, cap, max_new_tokens, self.device)
        # host -> device copies below never block the host on the stream (generate_stream relies on it)
        slot.kv_start.copy_(torch.from_numpy(np.ascontiguousarray(kv_start_host, dtype=np.int32)), non_blocking=True)
        kv_start, cache = slot.kv_start, slot.cache
        # prefix-KV reuse: rows that start with the cached prefix get its keys / values copied in; the prefill then
        # starts at slot q0 = (smallest kv_start) + prefix length (rows with more left padding recompute the tail of
        # their prefix inside the block -- same values)
        q0 = 0
        if prefix_cache is not None and kv_len is None:
            q0 = prefix_cache.plan(ids_cpu, attention_mask, kv_start_host, Lq)
            if q0 > 0:
                prefix_cache.copy_into(cache, kv_start_host, q0)
                embeds = embeds[:, q0:].contiguous()
        self._last_prefill_q0 = q0
        logits = self._prefill(embeds, kv_start, kv_len, cache, all_logits=False, logits_fp32=False, q0=q0)
        del embeds
        eos, pad = c.eos_token_id, c.pad_token_id if c.pad_token_id is not None else 0
        job = ModelOutput(ids_cpu=ids_cpu, B=B, Lq=Lq, cap=cap, cache=cache, kv_start=kv_start, eos=eos, pad=pad,
                          max_new_tokens=max_new_tokens, return_logits=return_logits, n_done=1, slot=slot)
        job.tokens = slot.tokens
        job.history = slot.history.fill_(pad)
        job.finished = slot.finished.zero_() if stop_on_eos else None
        job.state = slot.state
        job.state.copy_(torch.tensor([Lq, 1], dtype=torch.int32), non_blocking=True)
        # first token from the prefill logits
        L.check(lib.b200_argmax(L.ptr(logits), 0, c.vocab_size, B, c.vocab_size, L.ptr(job.tokens), L.ptr(job.finished),
                                eos, pad, L.stream_ptr()), "b200_argmax")
        job.history[:, 0] = job.tokens
        job.step_logits = [logits.float().clone()] if return_logits else None
        return job

    def _gen_decode(self, job, stopping_criteria=None, check_every=16, use_cuda_graph=True, capture_stream=None):
        """Phase B (HBM bound): max_new_tokens - 1 decode steps on the current stream -- the identical launch sequence
        captured once in a CUDA graph and replayed; the host synchronises only for stop checks."""
        c, lib = self.config, L.lib()
        B, cap, max_new_tokens = job.B, job.cap, job.max_new_tokens
        tokens, state, history, finished = job.tokens, job.state, job.history, job.finished
        return_logits = job.return_logits
        nb = lib.b200_llama_decode_workspace_bytes(ctypes.byref(self._w), B, cap)
        ws = self._ws_decode.get(nb, self.device)
        cs = job.cache.struct(0)
        lg_out = torch.empty((B, c.vocab_size), device=self.device, dtype=BF) if return_logits else None

        def step():
            L.check(lib.b200_llama_decode_step(ctypes.byref(self._w), L.ptr(tokens), L.ptr(state), L.ptr(job.kv_start),
                                               ctypes.byref(cs), B, cap - 1, L.ptr(finished), job.eos, job.pad,
                                               L.ptr(history), max_new_tokens, L.ptr(lg_out), L.ptr(ws), ws.numel(),
                                               L.stream_ptr()), "b200_llama_decode_step")

        n_done = 1
        graph = graph_kernels = None
        slot = job.slot
        if use_cuda_graph and not return_logits and slot.graph is not None and slot.graph[2] == (finished is not None):
            graph, graph_kernels, _ = slot.graph          # this slot's decode step was captured by an earlier batch

        def stop_now():
            if stopping_criteria:
                out_ids = torch.cat([job.ids_cpu, history[:, :n_done].to("cpu", torch.long)], dim=1)
                for sc in stopping_criteria:      # HF StoppingCriteriaList: stop if any criterion fires
                    r = sc(out_ids, None)
                    if bool(r.all()) if torch.is_tensor(r) else bool(r):
                        return True
            return finished is not None and bool(finished.min().item() == 1)

        if stopping_criteria:
            # HF evaluates the criteria after EVERY step (mm_utils.py:74-105 KeywordsStoppingCriteria): a keyword stop
            # that is not EOS must not let extra tokens through into the decoded text / the TakeMemory history
            check_every = 1
        while n_done < max_new_tokens:
            if (n_done == 1 or n_done % check_every == 0) and stop_now():
                break
            if graph is None and use_cuda_graph and n_done >= 2 and not return_logits:
                # the first decode step ran eagerly (lazy kernel-attribute setup); capture the launch sequence once
                # (capture_begin / capture_end directly: torch.cuda.graph() would synchronise the whole device first,
                # i.e. wait for the other stream's encode + prefill in generate_stream; step() allocates nothing)
                graph = torch.cuda.CUDAGraph()
                side = capture_stream if capture_stream is not None else torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                n0 = lib.b200_launch_count()
                with torch.cuda.stream(side):
                    graph.capture_begin(capture_error_mode="thread_local")
                    try:
                        step()
                    finally:
                        graph.capture_end()
                graph_kernels = lib.b200_launch_count() - n0
                L.note_graph_replay(-graph_kernels)       # the capture itself executed nothing
                torch.cuda.current_stream().wait_stream(side)
                slot.graph = (graph, graph_kernels, finished is not None)
            if graph is not None:
                graph.replay()
                L.note_graph_replay(graph_kernels)
            else:
                step()
            if return_logits:
                job.step_logits.append(lg_out.float().clone())
            n_done += 1
        job.n_done = n_done
        return job

    def _gen_finish(self, job):
        """Phase C: the generated ids back to the caller's layout (synchronises on the job's stream work)."""
        history, finished = job.history, job.finished
        gen = history[:, :job.n_done].to(torch.long)
        if finished is not None:
            # HF stops right after the step in which the last unfinished row emitted EOS: trim later columns
            is_eos = (gen == job.eos).to("cpu")
            if bool(is_eos.any(1).all()):
                last = int(is_eos.float().argmax(1).max().item()) + 1
                gen = gen[:, :last]
        out = torch.cat([job.ids_cpu.to(self.device), gen], dim=1)
        if job.return_logits:
            return out, torch.stack(job.step_logits[:gen.shape[1]], dim=1)
        return out

    @torch.no_grad()
    def make_prefix_cache(self, prefix_ids):
        """Keys / values of a text prefix (1-D LongTensor of token ids, no placeholders, no padding) for every decoder
        layer, computed once (B = 1 prefill) for generate(..., prefix_cache=...). In MM2SG's online mode every prompt
        starts with the same system text before `<image>` (scene_graph_prediction_model.py:140-199, conversation.py:
        253-263); a left-padded row's prefix sits at slots [kv_start, kv_start + p) and its rotary positions are
        relative to kv_start, so the same keys / values are valid in every row."""
        if self._w is None:
            raise L.B200Error("weights not loaded")
        c, lib = self.config, L.lib()
        ids = prefix_ids.detach().to("cpu").reshape(-1)
        p = int(ids.numel())
        if p == 0 or int(ids.min()) < 0:
            raise ValueError("make_prefix_cache: the prefix must be a non-empty run of real token ids")
        emb = torch.empty((1, p, c.hidden_size), device=self.device, dtype=BF)
        L.check(lib.b200_embed_rows(L.ptr(ids.to(torch.int32).to(self.device).contiguous()), L.ptr(self.model.embed_tokens),
                                    L.ptr(emb), c.hidden_size, p, c.hidden_size, c.vocab_size, L.stream_ptr()),
                "b200_embed_rows")
        cache = KVCache(c.num_hidden_layers, 1, c.num_attention_heads, (p + 7) // 8 * 8, self.device)
        self._prefill(emb, None, None, cache, all_logits=False, logits_fp32=False)
        return PrefixCache(ids, cache.k[:, 0, :, :p].clone(), cache.v[:, 0, :, :p].clone(), c.pad_token_id)

    @torch.no_grad()
    def generate_stream(self, requests, max_new_tokens=20, stop_on_eos=True, stopping_criteria=None, check_every=16,
                        prefill_chunk=None, vit_chunk=None, pooler_chunk=None):
        """Pipelined greedy generation over a sequence of batches: while batch n decodes (HBM bound: 255 passes over the
        weights and the KV cache), the views of batch n + 1 are encoded and its prompt prefilled (tensor-core bound) on a
        second, lower-priority stream, so the tensor pipe and the memory system work at the same time instead of in turn
        (DESIGN.md 8). `requests` yields dicts of generate() keyword arguments (input_ids, images, pc, audio, segmasks,
        attention_mask, vis_descriptor_embs); outputs come back in order, one LongTensor per request, identical to what
        generate() returns for it (same kernels, same arithmetic: only the scheduling differs).
        Two decode slots (KV cache + step buffers + the captured graph of a decode step) are alive at a time and
        ping-pong between the batches (0.56 GB per sample each at L = 831 + 256).
        prefill_chunk / vit_chunk / pooler_chunk: samples / images per launch sequence of phase A while pipelined --
        smaller chunks mean shorter tensor-core kernels, i.e. shorter waits for the decode stream's kernels when both
        compete for the SMs' shared memory."""
        cur_stream = torch.cuda.current_stream()
        hi = torch.cuda.Stream(priority=-1)          # decode: the serial chain of ~260 small kernels per token
        lo = torch.cuda.Stream(priority=0)           # encode + prefill: long tensor-core kernels, fill the gaps
        cap_stream = torch.cuda.Stream(priority=-1)  # graph capture: the nodes inherit its priority
        for s in (hi, lo):
            s.wait_stream(cur_stream)
        saved = (self.prefill_chunk, self.vit_chunk, self.pooler_chunk)
        free_slots = []

        def begin(req):
            kw = dict(req)
            ids = kw.pop("input_ids")
            images = kw.pop("images", None)
            with torch.cuda.stream(lo):
                slot = free_slots.pop() if free_slots else None
                job = self._gen_begin(ids, images, kw.pop("max_new_tokens", max_new_tokens), kw.get("pc"),
                                      kw.get("audio"), kw.get("segmasks"), kw.get("attention_mask"), stop_on_eos,
                                      False, kw.get("vis_descriptor_embs"), slot=slot,
                                      prefix_cache=kw.get("prefix_cache"))
                job.ready = torch.cuda.Event()
                job.ready.record(lo)
            return job

        try:
            if prefill_chunk:
                self.prefill_chunk = int(prefill_chunk)
            if vit_chunk:
                self.vit_chunk = int(vit_chunk)
            if pooler_chunk:
                self.pooler_chunk = int(pooler_chunk)
            it = iter(requests)
            first = next(it, None)
            nxt = begin(first) if first is not None else None
            while nxt is not None:
                cur = nxt
                req = next(it, None)
                nxt = begin(req) if req is not None else None      # phase A of batch n + 1 goes first into its queue
                with torch.cuda.stream(hi):
                    hi.wait_event(cur.ready)
                    self._gen_decode(cur, stopping_criteria, check_every, True, capture_stream=cap_stream)
                    out = self._gen_finish(cur)                    # synchronises with the decode of batch n only
                    done = torch.cuda.Event()
                    done.record(hi)
                done.synchronize()
                free_slots.append(cur.slot)                        # ping-pong: batch n + 2 reuses cache and graph
                cur = None
                yield out
        finally:
            self.prefill_chunk, self.vit_chunk, self.pooler_chunk = saved
            cur_stream.wait_stream(hi)
            cur_stream.wait_stream(lo)
