"""LoRA adapters on the seven Llama linears of every decoder layer -- the recipe the MM2SG authors actually trained
(LLaVA/llava/train/train.py:1159-1175: peft LoraConfig(r = 128, lora_alpha = 256, target = every
nn.Linear of the LLM found by find_all_linear_names, train.py:185-200), README.md:120-124).

peft semantics: y = x W^T + (alpha / r) * (x A^T) B^T with A (r, in) Kaiming-uniform and B (out, r) zero at start; the
base weight is frozen. The kernels work on FUSED base weights (q|k|v concatenated, gate/up row-interleaved), so the
adapters of one fused weight are applied with two GEMMs: t = x A_cat^T with A_cat = [A_1; ...; A_g] (g*r, in), then
y += (alpha / r) * t B_blk^T with B_blk (out_fused, g*r) holding B_j in the rows of projection j and the columns of its
own r-slice (zeros elsewhere). Masters are fp32 under peft's parameter names; the fused bf16 working copies are rebuilt
after every optimizer step (they are tiny: 2 * r * (in + out) per projection).
"""
import math

import torch

BF = torch.bfloat16

# fused weight of the kernels -> (reference projection names, module path)
FUSED = {
    "qkv_w": (("q_proj", "k_proj", "v_proj"), "self_attn"),
    "o_w": (("o_proj",), "self_attn"),
    "gate_up_w": (("gate_proj", "up_proj"), "mlp"),
    "down_w": (("down_proj",), "mlp"),
}


def param_name(layer, module, proj, which):
    return f"model.layers.{layer}.{module}.{proj}.lora_{which}.weight"


class LoraState:
    def __init__(self, cfg, r=128, alpha=256, device="cuda", seed=0, init_b="zero", dropout=0.0):
        """dropout: lora_dropout of the adapter branch (0.05 in the reference's recipe, train.py:111; 0 keeps the step
        deterministic, which is what the parity tests compare)."""
        self.cfg, self.r, self.scale = cfg, r, float(alpha) / r
        if not 0.0 <= dropout < 1.0:
            raise ValueError(f"lora dropout {dropout} not in [0, 1)")
        self.dropout, self.seed, self.rng_step = float(dropout), int(seed), 0
        self.device = torch.device(device)
        D, F = cfg.hidden_size, cfg.intermediate_size
        self.shapes = {"q_proj": (D, D), "k_proj": (D, D), "v_proj": (D, D), "o_proj": (D, D), "gate_proj": (F, D),
                       "up_proj": (F, D), "down_proj": (D, F)}
        g = torch.Generator().manual_seed(seed)
        self.sd = {}          # peft name -> bf16 working tensor (reference layout)
        for i in range(cfg.num_hidden_layers):
            for projs, module in FUSED.values():
                for p in projs:
                    out_f, in_f = self.shapes[p]
                    bound = 1.0 / math.sqrt(in_f)       # kaiming_uniform_(a = sqrt(5)) on (r, in): U(-1/sqrt(in), 1/sqrt(in))
                    a = (torch.rand(r, in_f, generator=g) * 2 - 1) * bound
                    b = torch.zeros(out_f, r) if init_b == "zero" else torch.randn(out_f, r, generator=g) * 0.02
                    self.sd[param_name(i, module, p, "A")] = a.to(self.device, BF)
                    self.sd[param_name(i, module, p, "B")] = b.to(self.device, BF)
        self.fused = [dict() for _ in range(cfg.num_hidden_layers)]
        self.refuse()

    def names(self):
        return sorted(self.sd)

    def refuse(self):
        """Rebuild A_cat / B_blk of every fused weight from the reference-layout adapters."""
        r = self.r
        for i in range(self.cfg.num_hidden_layers):
            for fname, (projs, module) in FUSED.items():
                g = len(projs)
                a_cat = torch.cat([self.sd[param_name(i, module, p, "A")] for p in projs], dim=0).contiguous()
                outs = [self.shapes[p][0] for p in projs]
                if fname == "gate_up_w":                       # rows interleaved: 2f = gate, 2f + 1 = up
                    F = outs[0]
                    b_blk = torch.zeros((F, 2, 2 * r), device=self.device, dtype=BF)
                    b_blk[:, 0, :r] = self.sd[param_name(i, module, "gate_proj", "B")]
                    b_blk[:, 1, r:] = self.sd[param_name(i, module, "up_proj", "B")]
                    b_blk = b_blk.view(2 * F, 2 * r)
                else:
                    b_blk = torch.zeros((sum(outs), g * r), device=self.device, dtype=BF)
                    row = 0
                    for j, p in enumerate(projs):
                        b_blk[row:row + outs[j], j * r:(j + 1) * r] = self.sd[param_name(i, module, p, "B")]
                        row += outs[j]
                self.fused[i][fname] = (a_cat, b_blk.contiguous())

    def unfuse_grads(self, g):
        """Gradients of A_cat / B_blk (keys `_lora.layers.{i}.{fused}.A|B`) -> peft parameter names."""
        r = self.r
        out = {}
        for i in range(self.cfg.num_hidden_layers):
            for fname, (projs, module) in FUSED.items():
                ka, kb = f"_lora.layers.{i}.{fname}.A", f"_lora.layers.{i}.{fname}.B"
                if ka not in g:
                    continue
                da, db = g[ka], g[kb]
                outs = [self.shapes[p][0] for p in projs]
                row = 0
                for j, p in enumerate(projs):
                    out[param_name(i, module, p, "A")] = da[j * r:(j + 1) * r]
                    if fname == "gate_up_w":
                        out[param_name(i, module, p, "B")] = db.view(outs[0], 2, 2 * r)[:, j, j * r:(j + 1) * r]
                    else:
                        out[param_name(i, module, p, "B")] = db[row:row + outs[j], j * r:(j + 1) * r]
                    row += outs[j]
        return out

    def merged_state_dict(self, base_sd):
        """W + (alpha / r) B A for every adapted Linear (peft merge_and_unload; what the loader does at inference)."""
        out = dict(base_sd)
        for i in range(self.cfg.num_hidden_layers):
            for projs, module in FUSED.values():
                for p in projs:
                    k = f"model.layers.{i}.{module}.{p}.weight"
                    a, b = self.sd[param_name(i, module, p, "A")].float(), self.sd[param_name(i, module, p, "B")].float()
                    out[k] = (base_sd[k].float().to(a.device) + self.scale * (b @ a)).to(base_sd[k].dtype)
        return out
