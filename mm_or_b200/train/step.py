"""One MM2SG fine-tune step on B200 (BASELINE.json configs[4]): multimodal forward, class-weighted shifted CE, backward
through the decoder / projector / image pooler / trainable CLIP layers, gradient-norm clipping and AdamW.

Reference: LLaVATrainer (LLaVA/llava/train/llava_trainer.py:134-278) = HF Trainer.training_step over
LlavaLlamaForCausalLM.forward + compute_loss, torch autograd, clip_grad_norm_(max_grad_norm), torch.optim.AdamW with
the four parameter groups of create_optimizer ({decay, no decay} x {projector, rest}; no-decay = norm weights and
biases), lr / mm_projector_lr from the training arguments (README.md:147-151: lr 2e-5, wd 0, max_grad_norm 0.1).
Everything numeric runs in libb200mmor.so through the C ABI; master weights, m and v are fp32 under the reference's
parameter names, and the model's fused bf16 working copies are rebuilt from them after every update.

Frozen here (documented in DESIGN.md): CLIP embeddings and the CLIP layers below `first_trainable_clip_layer`,
PointTransformerV3 (its token is computed, its training path is not built), and embed_tokens unless
`train_embed_tokens=True` (full fine-tuning; the reference's LoRA recipe keeps it frozen).
"""
import os

import numpy as np
import torch

from .. import _lib as L
from ..model.pack import DESC_BASE, descriptor_row_counts, plan_pack
from . import encoder as E
from . import llama as T

BF = torch.bfloat16


def default_trainable(name, first_clip_layer):
    if name.startswith("model.layers.") or name in ("model.norm.weight", "lm_head.weight"):
        return True
    if name.startswith("model.mm_projector."):
        return True
    if name.startswith("model.image_pooler.bert."):
        # word_embeddings (vocab_size = 1) and the BertPooler head are never used by the path (builder.py:173-175)
        return not ("word_embeddings" in name or ".pooler." in name)
    if name.startswith("model.image_pooler.project_audio.") or name.startswith("model.image_pooler.segmasks_encoder."):
        return True                                          # the whole image_pooler trains (train.py:1257-1261)
    p = E.VIT + "encoder.layers."
    if name.startswith(p):
        return int(name[len(p):].split(".")[0]) >= first_clip_layer
    return False


def no_decay(name):
    """create_optimizer's decay filter: LayerNorm / RMSNorm weights and all biases are not decayed."""
    return name.endswith(".bias") or "norm" in name.lower() or "layrnorm" in name.lower()


class FineTuner:
    def __init__(self, model, state_dict, lr=2e-5, mm_projector_lr=None, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, max_grad_norm=0.1, first_trainable_clip_layer=12, vocab_weight=None,
                 trainable=None, lora=None, train_embed_tokens=False, group=None, shard_optimizer=False,
                 shard_gradients=False, grad_comm_dtype=torch.bfloat16, base_nf4=False, nf4_double_quant=True,
                 recompute_activations=False):
        """lora: a train.lora.LoraState -> the reference's LoRA recipe (train.py:1159-1175): the decoder's base
        weights, norms and lm_head are frozen, the adapters train next to mm_projector / image_pooler / CLIP layers.
        group: data-parallel process group (same as set_process_group). shard_optimizer: keep fp32 master / m / v for
        1 / world of every tensor on each rank (train/zero.py, ZeRO-1); needs `group`. shard_gradients (with
        shard_optimizer): gradients are reduce-scattered in grad_comm_dtype instead of all-reduced in fp32, so a rank
        only receives the averaged gradient of the slices it updates (ZeRO-2, the reference's scripts/zero2.json).
        base_nf4 (LoRA only; the reference's `--bits 4`, train.py:1098-1114): the frozen decoder projections and lm_head
        are replaced by their NF4 round trip Q(W) (train/nf4.py, double quantisation of the block maxima as in the
        reference's default), so the adapters train against the weights a bitsandbytes Linear4bit would multiply with.
        recompute_activations: the reference's gradient checkpointing (train.py:1148) for the decoder -- only every
        layer's input is kept, the layer is re-run in the backward (train/llama.py); gradients are bit-identical."""
        self.model = model
        self.dev = model.device
        # the reference's gradient checkpointing (train.py:1148): decoder activations recomputed layer by layer
        self.recompute = bool(recompute_activations)
        self.first_clip = first_trainable_clip_layer
        self.lr, self.proj_lr = lr, (mm_projector_lr if mm_projector_lr is not None else lr)
        self.betas, self.eps, self.wd, self.max_norm = betas, eps, weight_decay, max_grad_norm
        self.vocab_weight = vocab_weight
        self.lora = lora
        self.train_embed = bool(train_embed_tokens) and lora is None
        pick = trainable if trainable is not None else (
            lambda n: default_trainable(n, first_trainable_clip_layer) or
            (self.train_embed and n == "model.embed_tokens.weight"))
        if lora is not None:
            base_pick = pick
            pick = lambda n: base_pick(n) and not (n.startswith("model.layers.") or n in ("model.norm.weight",
                                                                                         "lm_head.weight"))
        # the full reference-named state dict in bf16 on the device: the source the fused layouts are rebuilt from
        self.sd = {k: v.detach().to(self.dev, BF).contiguous() for k, v in state_dict.items()}
        self.nf4_bytes = None
        if base_nf4:
            if lora is None:
                raise ValueError("base_nf4=True is the QLoRA recipe: it needs lora=LoraState(...) (a 4-bit base is frozen)")
            from .nf4 import qlora_base_
            stored = qlora_base_(self.sd, double_quant=nf4_double_quant)
            self.nf4_bytes = sum(q.nbytes() for q in stored.values())       # what the packed base would occupy
            del stored
            # the kernels read fused copies of the decoder weights: rebuild them from Q(W)
            model._keep = model._w = model.lm_head = None
            model.load_state_dict(self.sd, device=self.dev)
        n_run = model.get_vision_tower().n_layers_run()     # CLIP layers past the selected hidden state are dead
        self.names = sorted(k for k in self.sd if pick(k) and self._has_backward(k, n_run))
        self.zero = None
        self.group = group
        if shard_gradients and not shard_optimizer:
            raise ValueError("shard_gradients=True needs shard_optimizer=True")
        self.shard_gradients, self.grad_comm_dtype = shard_gradients, grad_comm_dtype
        if shard_optimizer:
            if group is None:
                raise ValueError("shard_optimizer=True needs the data-parallel process group")
            from .zero import ShardedAdamW
            source = {k: state_dict[k] for k in self.names}
            if lora is not None:
                for k in lora.names():
                    self.sd[k] = lora.sd[k]
                    source[k] = lora.sd[k]
                self.names = sorted(self.names + lora.names())
            self.zero = ShardedAdamW(self.sd, source, self.names, group, betas=betas, eps=eps)
            self.master, self.m, self.v = self.zero.master, self.zero.m, self.zero.v      # this rank's slices
        else:
            self.master = {k: state_dict[k].detach().to(self.dev, torch.float32).contiguous().clone()
                           for k in self.names}
            if lora is not None:                       # adapters: masters from the LoraState, same optimizer
                for k in lora.names():
                    self.sd[k] = lora.sd[k]
                    self.master[k] = lora.sd[k].float().clone()
                self.names = sorted(self.names + lora.names())
            self.m = {k: torch.zeros_like(v) for k, v in self.master.items()}
            self.v = {k: torch.zeros_like(v) for k, v in self.master.items()}
        self.step_count = 0
        self.last_grads = None

    def set_lr(self, factor):
        """Scale both learning rates by `factor` of their initial values (train/schedule.py: the reference's cosine
        schedule with warm-up multiplies every parameter group's initial lr, mm_projector_lr included)."""
        if not hasattr(self, "_base_lr"):
            self._base_lr = (self.lr, self.proj_lr)
        self.lr, self.proj_lr = self._base_lr[0] * factor, self._base_lr[1] * factor

    def set_process_group(self, group):
        """Data-parallel fine-tuning: every rank runs forward_backward on its own samples; gradients
        are summed over ranks with ncclAllReduce and divided by the world size before clipping / AdamW, which keeps
        the replicas bit-identical (the reference gets the same from DeepSpeed ZeRO-2 / DDP)."""
        self.group = group

    @staticmethod
    def _has_backward(name, n_run):
        if name.startswith("model.image_pooler.point_transformer."):
            return False                                     # PointTransformerV3 stays frozen (no training path)
        p = E.VIT + "encoder.layers."
        if name.startswith(p):
            return int(name[len(p):].split(".")[0]) < n_run
        return True

    # ------------------------------------------------------------------------------------------------------------
    def forward_backward(self, input_ids, labels, attention_mask, images, grads=None, accumulate=False, pc=None,
                         audio=None, segmasks=None, grad_scale=1.0, vis_descriptor_embs=None):
        """Returns (loss, weight sum, grads under the reference's names). pc / audio / segmasks as in
        LlavaLlamaForCausalLM.forward (llava_llama.py:54-70): lists with one entry (or None) per sample.
        grads + accumulate=True add this batch's gradients into an existing store; grad_scale multiplies the loss
        gradient (1 / gradient_accumulation_steps, as HF Trainer scales the loss of every micro-batch)."""
        model, dev = self.model, self.dev
        cfg = model.config
        tower, pooler, proj = model.get_vision_tower(), model.get_image_pooler(), model.get_model().mm_projector
        keep, Dv, D = pooler.geo["keep"], pooler.geo["hidden"], cfg.hidden_size
        concat, split = model._images_to_batch(images)
        B = len(split)
        hidden, vc = E.vit_forward(tower, concat, self.first_clip)
        pooled, pcache = E.pooler_forward(pooler, hidden, split)
        Tv = keep + pooler.num_extra_tokens(pc, audio, segmasks)              # tokens per sample (builder.py:176-189)
        if Tv == keep:
            tokens, xcache = pooled.reshape(B * keep, Dv).contiguous(), {}
        else:
            tok3 = torch.empty((B, Tv, Dv), device=dev, dtype=BF)
            tok3[:, :keep] = pooled
            xcache = E.extras_forward(pooler, tok3, keep, pc, audio, segmasks)
            tokens = tok3.view(B * Tv, Dv)
        vis, prc = E.projector_forward(proj, tokens)
        desc_rows = desc_table = None
        if vis_descriptor_embs is not None:      # inputs only (llava_arch.py:278-294): rows spliced in, no gradient
            per_sample, desc_rows = descriptor_row_counts(vis_descriptor_embs, B)
            flat = [e.detach().reshape(-1, D) for per in per_sample for e in per]
            desc_table = torch.cat(flat).to(dev, BF).contiguous() if flat else None
        plan = plan_pack(input_ids.cpu().numpy(), None if attention_mask is None else attention_mask.cpu().numpy(),
                         None if labels is None else labels.cpu().numpy(), Tv, "right",
                         getattr(cfg, "tokenizer_model_max_length", None), desc_rows=desc_rows, n_blocks=B)
        if plan.src.shape[0] != B:
            raise NotImplementedError("fine-tune step: one row of input_ids per entry of `images`")
        Lq = plan.L
        src = plan.src.reshape(B, Lq).astype(np.int64)
        text_ids = np.where(src >= -1, src, -2).astype(np.int32)
        embeds = torch.empty((B * Lq, D), device=dev, dtype=BF)
        L.embed_rows(torch.as_tensor(text_ids.reshape(-1)).to(dev), model.model.embed_tokens, out=embeds)
        L.embed_rows(torch.as_tensor(plan.vis_ids).to(dev), vis, out=embeds)
        if desc_table is not None and (plan.desc_ids >= 0).any():
            L.embed_rows(torch.as_tensor(plan.desc_ids).to(dev), desc_table, out=embeds)
        loss, wsum, g, d_emb = T.forward_backward(model, embeds.view(B, Lq, D), torch.from_numpy(plan.labels).to(dev),
                                                  torch.from_numpy(plan.lengths).to(dev), self.vocab_weight,
                                                  grad_scale=grad_scale, grads=grads, accumulate=accumulate,
                                                  lora=self.lora, train_base=self.lora is None,
                                                  recompute=self.recompute)
        # pack backward: visual rows of d_embeds back to the projector output (truncated tokens get zeros)
        d_vis = L.embed_rows(torch.as_tensor(plan.row_map).to(dev), d_emb.reshape(B * Lq, D), rows=B * Tv)
        if self.train_embed:                                   # nn.Embedding backward over the text rows
            dst, acc = E._Grads(g, accumulate, dev).slot("model.embed_tokens.weight",
                                                         tuple(model.model.embed_tokens.shape))
            if not acc:
                dst.zero_()
            L.embed_grad(d_emb.reshape(B * Lq, D), text_ids.reshape(-1), dst, accumulate=True)
        dx, g = E.projector_backward(proj, prc, d_vis, grads=g, accumulate=accumulate)
        dx3 = dx.view(B, Tv, Dv)
        if xcache:
            g = E.extras_backward(pooler, xcache, dx3, grads=g, accumulate=accumulate)
        d_hidden, g = E.pooler_backward(pooler, pcache, dx3[:, :keep], grads=g, accumulate=accumulate)
        g = E.vit_backward(tower, vc, d_hidden, grads=g, accumulate=accumulate)
        return loss, wsum, g

    def optimizer_step(self, grads):
        """clip_grad_norm_(max_grad_norm) over the trainable set + AdamW; rebuilds the fused bf16 working weights."""
        if self.group is not None:
            from ..dist import align_optional_gradients, average_gradients
            from .trace import phase as _ph
            optional = {k: tuple(self.sd[k].shape) for k in self.names
                        if k == "model.embed_tokens.weight" or k.startswith(("model.image_pooler.project_audio.",
                                                                             "model.image_pooler.segmasks_encoder."))}
            with _ph("align_optional_gradients"):
                if getattr(self, "_host_group", None) is None:
                    from ..dist import host_side_group
                    self._host_group = host_side_group(self.group)     # collective: every rank reaches it in step 1
                # ranks whose batch lacked a modality contribute zeros
                align_optional_gradients(grads, optional, self.group, host_group=self._host_group)
            if not self.shard_gradients:
                average_gradients(grads, sorted(grads), self.group)    # in place on the (contiguous) fused gradients
        from .trace import phase
        with phase("unfuse_grads"):
            g = T.unfuse_grads(grads, self.model.config)
            if self.lora is not None:
                g.update(self.lora.unfuse_grads(grads))
        self.step_count += 1
        lr_of = lambda k: self.proj_lr if k.startswith("model.mm_projector.") else self.lr
        wd_of = lambda k: 0.0 if no_decay(k) else self.wd
        if self.zero is not None and self.shard_gradients:
            # ZeRO-2: reduce-scatter -> norm from the slices -> update of the local slices -> all-gather of the weights
            out2 = self.zero.step_from_local_grads({k: g[k] for k in self.names if k in g}, self.step_count, lr_of,
                                                   wd_of, max_norm=self.max_norm, comm_dtype=self.grad_comm_dtype)
            return self._after_update(out2)
        out2 = torch.zeros(2, device=self.dev, dtype=torch.float32)
        # parameters that received no gradient this step (a modality absent from the batch) are skipped by the norm and
        # by AdamW, like parameters whose .grad is None in torch
        names = [k for k in self.names if k in g]
        for i, k in enumerate(names):
            L.grad_sq_norm(g[k].contiguous(), out2=out2, accumulate=i > 0, max_norm=self.max_norm)
        clip = out2[1:]
        if self.zero is not None:          # sharded states: update this rank's slices, all-gather the bf16 weights
            self.zero.step({k: g[k] for k in names}, self.step_count, lr_of, wd_of, clip_coef=clip)
        else:
            for k in names:
                L.adamw_step(self.master[k], self.sd[k], g[k].contiguous(), self.m[k], self.v[k], lr_of(k),
                             self.betas[0], self.betas[1], self.eps, wd_of(k), self.step_count, clip_coef=clip)
        return self._after_update(out2)

    def _after_update(self, out2):
        """Rebuild the fused bf16 working weights the kernels read from the updated per-parameter tensors."""
        from .trace import phase
        with phase("refuse_working_weights"):
            return self._after_update_impl(out2)

    def _after_update_impl(self, out2):
        if self.lora is not None:
            self.lora.refuse()                     # adapters changed; the decoder's base weights did not
            self._reload_encoder()
            return out2
        # drop the old fused working copies before rebuilding them (13.5 GB at 7B)
        self.model._keep = self.model._w = self.model.lm_head = None
        self.model.load_state_dict(self.sd, device=self.dev)
        return out2

    def _reload_encoder(self):
        """Rebuild only the vision-side fused weights (tower, pooler, projector) from the updated bf16 tensors."""
        m = self.model.model
        m.vision_tower.load_weights(self.sd, self.dev)
        m.image_pooler.load_weights(self.sd, self.dev)
        m.mm_projector.load_weights(self.sd, self.dev)

    # ------------------------------------------------------------------------------------------------------------
    def save_checkpoint(self, out_dir, base_model_name_or_path=None, optimizer=True):
        """What train.py:1346-1360 leaves in output_dir (train/checkpoint.py): the LoRA recipe writes config.json,
        adapter_config.json, adapter_model.bin and non_lora_trainables.bin (loadable with
        load_pretrained_model(out_dir, model_base, "…lora…")); full fine-tuning writes config.json and the whole
        state_dict. optimizer=True adds b200_optimizer.pt (masters, m, v, step) for load_optimizer; with a sharded
        optimizer every rank writes its own slices (b200_optimizer.rankN.pt) and rank 0 the weights."""
        from . import checkpoint as C
        rank = 0
        if self.group is not None:
            import torch.distributed as dist
            rank = dist.get_rank(self.group)
        if rank == 0:
            if self.lora is not None:
                r = self.lora.r
                trainable = {k: self.sd[k] for k in self.names if "lora_" not in k}
                trainable.update(C.pooler_checkpoint_tensors(self.sd))      # the whole pooler state, frozen parts too
                C.export_lora_checkpoint(out_dir, self.model.config, self.lora.sd, r, self.lora.scale * r, trainable,
                                         base_model_name_or_path, dropout=self.lora.dropout)
            else:
                C.export_full_checkpoint(out_dir, self.model.config, self.sd)
        if optimizer:
            os.makedirs(out_dir, exist_ok=True)
            name = "b200_optimizer.pt" if self.zero is None else "b200_optimizer.rank%d.pt" % rank
            if self.zero is not None or rank == 0:
                C.save_optimizer(os.path.join(out_dir, name), self.master, self.m, self.v, self.step_count,
                                 *getattr(self, "_base_lr", (self.lr, self.proj_lr)))
        return out_dir

    def load_optimizer(self, out_dir):
        """Resume: masters / m / v / step count from save_checkpoint(optimizer=True); the bf16 working weights are
        rebuilt from the masters."""
        from . import checkpoint as C
        name = "b200_optimizer.pt"
        if self.zero is not None:              # sharded state: every rank reads the slices it wrote (same world size)
            import torch.distributed as dist
            name = "b200_optimizer.rank%d.pt" % dist.get_rank(self.group)
        self.step_count, lr, proj_lr = C.load_optimizer(os.path.join(out_dir, name), self.master, self.m, self.v)
        self._base_lr = (lr, proj_lr)
        self.lr, self.proj_lr = lr, proj_lr
        if self.zero is not None:
            self.zero.refresh_from_masters()   # slices -> bf16 -> all-gather
        else:
            for k, mst in self.master.items():
                self.sd[k].copy_(mst.to(BF))
        self._after_update(None)

    def train_step(self, input_ids, labels, attention_mask, images, pc=None, audio=None, segmasks=None,
                   vis_descriptor_embs=None):
        from .trace import phase
        with phase("forward_backward"):
            loss, wsum, grads = self.forward_backward(input_ids, labels, attention_mask, images, pc=pc, audio=audio,
                                                      segmasks=segmasks, vis_descriptor_embs=vis_descriptor_embs)
        self.last_grads = grads
        with phase("optimizer_step (all)"):
            norm_sq = self.optimizer_step(grads)
        return loss, norm_sq

    def train_step_accumulated(self, micro_batches):
        """One optimizer step over several micro-batches (the reference recipe runs gradient_accumulation_steps 4,
        README.md:143): HF Trainer divides every micro-batch loss by the number of micro-batches, sums the gradients and
        steps once. micro_batches: list of dicts with the keyword arguments of train_step. Returns (mean loss, out2)."""
        n = len(micro_batches)
        if n == 0:
            raise ValueError("train_step_accumulated: no micro-batches")
        grads, total = None, 0.0
        for i, mb in enumerate(micro_batches):
            loss, _, grads = self.forward_backward(mb["input_ids"], mb["labels"], mb.get("attention_mask"), mb["images"],
                                                   grads=grads, accumulate=i > 0, pc=mb.get("pc"),
                                                   audio=mb.get("audio"), segmasks=mb.get("segmasks"),
                                                   grad_scale=1.0 / n,
                                                   vis_descriptor_embs=mb.get("vis_descriptor_embs"))
            total = total + loss / n
        self.last_grads = grads
        return total, self.optimizer_step(grads)
