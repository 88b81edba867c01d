"""ctypes binding of libb200mmor.so (C ABI declared in include/b200_mmor.h).

The product path has no CPU or eager-PyTorch fallback: if the shared library is missing or a call fails,
this module raises. torch is used only to own device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200mmor.so")

_lib = None

vp, ci, cf, i64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int64, ctypes.c_size_t


class B200Error(RuntimeError):
    pass


# ---- struct mirrors of include/b200_mmor.h -------------------------------------------------------------------
class VitLayer(ctypes.Structure):
    _fields_ = [(n, vp) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_w", "ln2_b",
                                  "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class VitWeights(ctypes.Structure):
    _fields_ = [("hidden", ci), ("heads", ci), ("ffn", ci), ("image_size", ci), ("patch", ci), ("kpad", ci),
                ("n_layers", ci), ("ln_eps", cf), ("patch_w", vp), ("pos_cls", vp), ("pre_ln_w", vp),
                ("pre_ln_b", vp), ("layers", ctypes.POINTER(VitLayer))]


class BertLayer(ctypes.Structure):
    _fields_ = [(n, vp) for n in ("qkv_w", "qkv_b", "ao_w", "ao_b", "ao_ln_w", "ao_ln_b", "fc1_w", "fc1_b",
                                  "fc2_w", "fc2_b", "out_ln_w", "out_ln_b")]


class PoolerWeights(ctypes.Structure):
    _fields_ = [("hidden", ci), ("heads", ci), ("ffn", ci), ("n_layers", ci), ("max_pos", ci), ("ln_eps", cf),
                ("pos_type", vp), ("emb_ln_w", vp), ("emb_ln_b", vp), ("layers", ctypes.POINTER(BertLayer))]


class SegmaskWeights(ctypes.Structure):
    _fields_ = [("emb", vp), ("conv_w", vp * 5), ("conv_b", vp * 5)]


class ProjectorWeights(ctypes.Structure):
    _fields_ = [("in_dim", ci), ("hidden", ci), ("w0", vp), ("b0", vp), ("w2", vp), ("b2", vp)]


class LlamaLayer(ctypes.Structure):
    _fields_ = [(n, vp) for n in ("attn_norm", "qkv_w", "o_w", "mlp_norm", "gate_up_w", "down_w")]


class LlamaWeights(ctypes.Structure):
    _fields_ = [("hidden", ci), ("heads", ci), ("ffn", ci), ("n_layers", ci), ("vocab", ci), ("max_pos", ci),
                ("rms_eps", cf), ("layers", ctypes.POINTER(LlamaLayer)), ("final_norm", vp), ("lm_head", vp),
                ("embed_tokens", vp), ("rope_cos", vp), ("rope_sin", vp)]


class KvCache(ctypes.Structure):
    _fields_ = [("k", vp), ("v", vp), ("layer_stride", i64), ("cap", ci)]


_P = ctypes.POINTER

_SIGS = {
    # name: (restype, argtypes)
    "b200_abi_version": (ci, []),
    "b200_sizeof_struct": (sz, [ci]),
    "b200_launch_count": (ctypes.c_longlong, []),
    "b200_prof_enable": (ci, [ci]),
    "b200_prof_family_count": (ci, []),
    "b200_prof_family_name": (ctypes.c_char_p, [ci]),
    "b200_prof_collect": (ci, [vp, vp, vp, vp]),
    "b200_set_option": (ci, [ctypes.c_char_p, ci]),
    "b200_get_option": (ci, [ctypes.c_char_p]),
    "b200_decode_tile_width": (ci, [ci, ci, ci, ci]),
    "b200_dropout": (ci, [vp, vp, i64, cf, ctypes.c_uint64, ci, vp]),
    "b200_nf4_quantize": (ci, [vp, i64, vp, vp, vp]),
    "b200_nf4_dequantize": (ci, [vp, vp, i64, vp, vp]),
    "b200_gemm_bf16": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, ci, vp, vp, ci, vp, ci, ci, ci, vp]),
    "b200_gemm_bf16_skinny": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, vp]),
    "b200_gemm_bf16_ex": (ci, [vp, ci, ci, vp, ci, ci, vp, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, cf, ci, vp]),
    "b200_colsum_workspace_bytes": (sz, [ci]),
    "b200_colsum": (ci, [vp, i64, ci, ci, ci, vp, vp, sz, vp]),
    "b200_act_backward": (ci, [vp, vp, vp, i64, ci, vp]),
    "b200_swiglu_forward": (ci, [vp, vp, i64, vp]),
    "b200_act_forward": (ci, [vp, vp, i64, ci, vp]),
    "b200_gather_add_rows": (ci, [vp, i64, vp, vp, ci, ci, ci, vp, vp]),
    "b200_group_sum": (ci, [vp, ci, i64, ci, vp, vp]),
    "b200_rope_kv_backward": (ci, [vp, vp, vp, vp, ci, vp, vp, ci, ci, ci, ci, vp]),
    "b200_norm_backward_workspace_bytes": (sz, [ci, ci]),
    "b200_norm_backward": (ci, [vp, vp, vp, cf, ci, ci, ci, vp, vp, vp, vp, ci, vp, sz, vp]),
    "b200_weighted_ce_workspace_bytes": (sz, [ci, ci]),
    "b200_weighted_ce": (ci, [vp, ci, i64, vp, vp, ci, ci, ci, cf, vp, i64, vp, vp, sz, vp]),
    "b200_grad_norm_workspace_bytes": (sz, []),
    "b200_grad_sq_norm": (ci, [vp, ci, i64, ci, cf, vp, vp, sz, vp]),
    "b200_adamw_step": (ci, [vp, vp, vp, ci, vp, vp, i64, cf, cf, cf, cf, cf, ci, vp, vp]),
    "b200_layernorm": (ci, [vp, i64, vp, vp, ci, vp, vp, cf, vp, i64, ci, ci, vp]),
    "b200_rmsnorm": (ci, [vp, i64, vp, cf, vp, i64, ci, ci, vp]),
    "b200_flash_attention": (ci, [vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64,
                                  ci, ci, ci, ci, ci, vp, vp, ci, cf, vp]),
    "b200_flash_attention_lse": (ci, [vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64,
                                      ci, ci, ci, ci, ci, vp, vp, ci, cf, vp, vp]),
    "b200_flash_attention_bwd_workspace_bytes": (sz, [ci, ci, ci]),
    "b200_flash_attention_bwd": (ci, [vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64, vp, vp, i64, i64, i64,
                                      vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp, ci, cf, vp, sz, vp]),
    "b200_decode_attention_workspace_bytes": (sz, [ci, ci, ci]),
    "b200_decode_attention": (ci, [vp, i64, vp, vp, vp, i64, ci, ci, ci, ci, vp, cf, ci, vp, sz, vp]),
    "b200_rope_kv_write": (ci, [vp, vp, vp, vp, ci, vp, vp, ci, ci, ci, ci, ci, vp]),
    "b200_embed_rows": (ci, [vp, vp, vp, i64, ci, ci, ci, vp]),
    "b200_argmax": (ci, [vp, ci, i64, ci, ci, vp, vp, ci, ci, vp]),
    "b200_patchify": (ci, [vp, vp, ci, ci, ci, ci, ci, vp]),
    "b200_preprocess_workspace_bytes": (sz, [ci, ci, ci]),
    "b200_preprocess_images": (ci, [vp, ci, ci, ci, ci, vp, vp, vp, ci, vp, vp, ci, ci, ci, ci, vp, vp, vp, vp, sz, vp]),
    "b200_vit_workspace_bytes": (sz, [_P(VitWeights), ci]),
    "b200_vit_forward": (ci, [_P(VitWeights), vp, vp, ci, vp, sz, vp]),
    "b200_pooler_workspace_bytes": (sz, [_P(PoolerWeights), ci, ci]),
    "b200_pooler_forward": (ci, [_P(PoolerWeights), vp, i64, vp, vp, ci, ci, ci, vp, ci, vp, sz, vp]),
    "b200_segmask_workspace_bytes": (sz, [ci]),
    "b200_segmask_forward": (ci, [_P(SegmaskWeights), vp, ci, vp, i64, vp, vp, sz, vp]),
    "b200_projector_workspace_bytes": (sz, [_P(ProjectorWeights), ci]),
    "b200_projector_pack": (ci, [_P(ProjectorWeights), vp, ci, vp, vp, vp, ci, vp, ci, vp, sz, vp]),
    "b200_projector_gather": (ci, [_P(ProjectorWeights), vp, ci, vp, ci, sz, vp, sz, vp]),
    "b200_gemm_nf4": (ci, [vp, ci, vp, vp, vp, ci, ci, ci, ci, vp, ci, cf, vp]),
    "b200_llama_prefill_workspace_bytes": (sz, [_P(LlamaWeights), ci, ci, ci]),
    "b200_llama_prefill": (ci, [_P(LlamaWeights), vp, vp, vp, _P(KvCache), ci, ci, vp, ci, ci, vp, sz, vp]),
    "b200_llama_prefill_from": (ci, [_P(LlamaWeights), vp, vp, vp, _P(KvCache), ci, ci, ci, vp, ci, ci, vp, sz, vp]),
    "b200_llama_decode_workspace_bytes": (sz, [_P(LlamaWeights), ci, ci]),
    "b200_llama_decode_step": (ci, [_P(LlamaWeights), vp, vp, vp, _P(KvCache), ci, ci, vp, ci, ci, vp, ci, vp, vp,
                                    sz, vp]),
    # point-cloud branch (ptv3.cu)
    "b200_pc_grid_coords": (ci, [vp, ci, ci, cf, vp, vp, vp, vp]),
    "b200_pc_encode": (ci, [vp, vp, ci, ci, ci, vp, vp]),
    "b200_pc_argsort_workspace_bytes": (sz, [ci]),
    "b200_pc_argsort": (ci, [vp, ci, ci, vp, vp, vp, sz, vp]),
    "b200_pc_gather_rows": (ci, [vp, ci, vp, ci, ci, vp, ci, vp]),
    "b200_pc_neighbors": (ci, [vp, vp, vp, ci, ci, ci, vp, vp, vp]),
    "b200_pc_pool_plan": (ci, [vp, vp, vp, ci, ci, vp, vp, vp, vp, vp]),
    "b200_pc_cloud_offsets": (ci, [vp, ci, ci, vp, vp]),
    "b200_pc_gemm_workspace_bytes": (sz, [ci, ci, ci, ci]),
    "b200_pc_gemm_f32": (ci, [vp, ci, vp, ci, vp, vp, vp, vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, vp, sz, vp]),
    "b200_pc_layernorm_f32": (ci, [vp, ci, vp, vp, cf, vp, ci, vp, ci, ci, ci, vp]),
    "b200_pc_patch_attention": (ci, [vp, ci, vp, vp, ci, ci, ci, ci, cf, vp, ci, vp]),
    "b200_pc_segment_max": (ci, [vp, ci, vp, ci, ci, vp, vp, ci, vp, ci, vp]),
    "b200_pc_cloud_mean": (ci, [vp, ci, vp, ci, ci, vp, vp, ci, vp]),
    # fine-tuning of the seg-mask CNN and embed_tokens (train_extras.cu)
    "b200_segmask_train_acts_bytes": (sz, [ci]),
    "b200_segmask_forward_train": (ci, [_P(SegmaskWeights), vp, ci, vp, sz, vp, i64, vp, vp]),
    "b200_segmask_backward_workspace_bytes": (sz, [ci]),
    "b200_segmask_backward": (ci, [_P(SegmaskWeights), vp, ci, vp, vp, i64, vp, _P(vp), _P(vp), vp, ci, vp, sz, vp]),
    "b200_embed_grad": (ci, [vp, i64, vp, vp, vp, ci, ci, vp, ci, vp]),
}

PC_SYMBOLS = sorted(k for k in _SIGS if k.startswith("b200_pc_"))
# entry points whose kernels also build for the host against the test-only kernel emulator (tests/emu/)
EMULATABLE_SYMBOLS = PC_SYMBOLS + ["b200_segmask_train_acts_bytes", "b200_segmask_forward_train",
                                   "b200_segmask_backward_workspace_bytes", "b200_segmask_backward", "b200_embed_grad",
                                   "b200_nf4_quantize", "b200_nf4_dequantize", "b200_dropout"]

EXPORTED_SYMBOLS = ["b200_last_error"] + sorted(_SIGS)


def lib():
    """Load (once) and return the CDLL. Raises B200Error when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} not found: build it with `python -m mm_or_b200.build` (no CPU fallback exists)")
        L = ctypes.CDLL(LIB_PATH)
        L.b200_last_error.restype = ctypes.c_char_p
        L.b200_last_error.argtypes = []
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().b200_last_error()
        raise B200Error(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise B200Error("libb200mmor operates on CUDA tensors only (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Workspace:
    """Grow-only device scratch buffer handed to the stage entry points."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self.buf


ACT_NONE, ACT_QUICK_GELU, ACT_GELU, ACT_SWIGLU = 0, 1, 2, 3


# ---- thin operator wrappers (used by the parity tests and by host code outside the stage entry points) ---------
def gemm(a, w, out=None, bias=None, residual=None, row_map=None, act=ACT_NONE, out_fp32=False, bn=0,
         out_rows=None):
    """out = epilogue(a @ w.T). a: (M,K) bf16 (row stride allowed), w: (N,K) bf16."""
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        rows = M if out_rows is None else out_rows
        out = torch.empty((rows, n_out), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    ldr = residual.stride(0) if residual is not None else 0
    rc = lib().b200_gemm_bf16(ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K,
                              ptr(bias), ptr(residual), ldr, ptr(row_map), act, int(out_fp32), bn, stream_ptr())
    check(rc, "b200_gemm_bf16")
    return out


def gemm_nf4(a, codes, absmax, n, out=None, residual=None, scale=1.0):
    """out = scale * a @ dequant(codes, absmax).T (+ residual). a (M, K) bf16; codes uint8 (n * K / 2,), absmax fp32
    (n * K / 64,): the NF4 storage of an (n, K) weight (train/nf4.py); the weight is expanded to bf16 inside the GEMM."""
    M, K = a.shape
    assert a.stride(1) == 1 and codes.dtype == torch.uint8 and absmax.dtype == torch.float32
    assert codes.numel() * 2 == n * K and absmax.numel() * 64 == n * K
    if out is None:
        out = torch.empty((M, n), device=a.device, dtype=torch.bfloat16)
    check(lib().b200_gemm_nf4(ptr(a), a.stride(0), ptr(codes), ptr(absmax), ptr(out), out.stride(0), M, n, K,
                              ptr(residual), residual.stride(0) if residual is not None else 0, float(scale),
                              stream_ptr()), "b200_gemm_nf4")
    return out


def gemm_skinny(a, w, out=None, bias=None, residual=None, act=ACT_NONE, out_fp32=False, splits=0):
    """Decode-step GEMM (M <= 256): out = epilogue(a @ w.T), weights streamed once, K split over a cluster."""
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    ldr = residual.stride(0) if residual is not None else 0
    rc = lib().b200_gemm_bf16_skinny(ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K,
                                     ptr(bias), ptr(residual), ldr, act, int(out_fp32), splits, stream_ptr())
    check(rc, "b200_gemm_bf16_skinny")
    return out


def gemm_ex(a, w, a_t=False, w_t=False, out=None, out_fp32=False, accumulate=False, bn=0, residual=None, scale=1.0):
    """General GEMM with transposed operands read in place: A is (M, K), or (K, M) when a_t; W is (N, K), or (K, N)
    when w_t. out = A_eff @ W_eff.T, optionally accumulated into an fp32 `out`."""
    M, K = (a.shape[1], a.shape[0]) if a_t else a.shape
    N = w.shape[1] if w_t else w.shape[0]
    assert (w.shape[0] if w_t else w.shape[1]) == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    ldr = residual.stride(0) if residual is not None else 0
    check(lib().b200_gemm_bf16_ex(ptr(a), a.stride(0), int(a_t), ptr(w), w.stride(0), int(w_t), ptr(out), out.stride(0),
                                  M, N, K, None, ptr(residual), ldr, ACT_NONE, int(out.dtype == torch.float32),
                                  int(accumulate), float(scale), bn, stream_ptr()), "b200_gemm_bf16_ex")
    return out


def linear_backward(x, w, dy, dw=None, db=None, accumulate=False, need_dx=True):
    """Backward of y = x @ w.T + b: returns (dx bf16 | None, dw fp32 (N, K), db fp32 (N,) | None)."""
    dx = gemm_ex(dy, w, w_t=True) if need_dx else None                 # dX = dY W
    dw = gemm_ex(dy, x, a_t=True, w_t=True, out=dw, out_fp32=True, accumulate=accumulate)   # dW = dY^T X
    if db is not None or not accumulate:
        N = w.shape[0]
        if db is None:
            db = torch.empty(N, device=x.device, dtype=torch.float32)
        ws = torch.empty(int(lib().b200_colsum_workspace_bytes(N)), dtype=torch.uint8, device=x.device)
        check(lib().b200_colsum(ptr(dy), dy.stride(0), dy.shape[0], N, int(accumulate), ptr(db), ptr(ws), ws.numel(),
                                stream_ptr()), "b200_colsum")
    return dx, dw, db


def swiglu_forward(z):
    """(M, 2F) interleaved (gate, up) pre-activations -> (M, F)."""
    M, F2 = z.shape
    h = torch.empty((M, F2 // 2), device=z.device, dtype=torch.bfloat16)
    check(lib().b200_swiglu_forward(ptr(z), ptr(h), h.numel(), stream_ptr()), "b200_swiglu_forward")
    return h


def gather_add_rows(src, row_map, add, rows, period=None):
    """out[r] = src[row_map[r]] (zero when < 0) + add[r % period]."""
    D = src.shape[1]
    out = torch.empty((rows, D), device=src.device, dtype=torch.bfloat16)
    period = (add.shape[0] if add is not None else 1) if period is None else period
    check(lib().b200_gather_add_rows(ptr(src), src.stride(0), ptr(row_map), ptr(add), period, rows, D, ptr(out),
                                     stream_ptr()), "b200_gather_add_rows")
    return out


def embed_rows(ids, table, out=None, rows=None):
    """out[r] = table[ids[r]] (ids >= 0), zeros (ids == -1), untouched (ids <= -2)."""
    rows = ids.numel() if rows is None else rows
    D = table.shape[1]
    if out is None:
        out = torch.empty((rows, D), device=table.device, dtype=torch.bfloat16)
    check(lib().b200_embed_rows(ptr(ids), ptr(table), ptr(out), out.stride(0), rows, D, table.shape[0], stream_ptr()),
          "b200_embed_rows")
    return out


def colsum(dy, out=None, accumulate=False):
    M, N = dy.shape
    if out is None:
        out = torch.empty(N, device=dy.device, dtype=torch.float32)
    ws = torch.empty(int(lib().b200_colsum_workspace_bytes(N)), dtype=torch.uint8, device=dy.device)
    check(lib().b200_colsum(ptr(dy), dy.stride(0), M, N, int(accumulate), ptr(out), ptr(ws), ws.numel(), stream_ptr()),
          "b200_colsum")
    return out


def act_forward(z, act):
    y = torch.empty_like(z)
    check(lib().b200_act_forward(ptr(z), ptr(y), z.numel(), act, stream_ptr()), "b200_act_forward")
    return y


def group_sum(x, out=None, accumulate=False):
    """x (G, ...) bf16 -> fp32 sum over the leading dimension."""
    G = x.shape[0]
    slab = x[0].numel()
    if out is None:
        out = torch.empty(x.shape[1:], device=x.device, dtype=torch.float32)
    check(lib().b200_group_sum(ptr(x), G, slab, int(accumulate), ptr(out), stream_ptr()), "b200_group_sum")
    return out


def act_backward(z, dy, act):
    dz = torch.empty_like(z)
    check(lib().b200_act_backward(ptr(z), ptr(dy), ptr(dz), dy.numel(), act, stream_ptr()), "b200_act_backward")
    return dz


def norm_backward(x, dy, gamma, eps, rms=False, dgamma=None, dbeta=None, accumulate=False, add=None):
    """dx (+ add: the residual-branch gradient), dgamma, dbeta of LayerNorm / RMSNorm from the saved input x."""
    M, D = x.shape
    dx = torch.empty_like(x)
    if dgamma is None:
        dgamma = torch.empty(D, device=x.device, dtype=torch.float32)
    if dbeta is None and not rms:
        dbeta = torch.empty(D, device=x.device, dtype=torch.float32)
    ws = torch.empty(int(lib().b200_norm_backward_workspace_bytes(M, D)), dtype=torch.uint8, device=x.device)
    check(lib().b200_norm_backward(ptr(x), ptr(dy), ptr(gamma), eps, M, D, int(rms), ptr(add), ptr(dx), ptr(dgamma),
                                   ptr(dbeta),
                                   int(accumulate), ptr(ws), ws.numel(), stream_ptr()), "b200_norm_backward")
    return dx, dgamma, dbeta


def weighted_ce(logits, labels, vocab_weight=None, grad_scale=1.0, want_grad=False, inplace=False):
    """LLaVATrainer.compute_loss. logits (B, L, V) fp32/bf16, labels (B, L) int64 unshifted modified_labels.
    Returns (loss 0-dim fp32 tensor, weight sum, dlogits or None)."""
    B, Lq, V = logits.shape
    assert logits.is_contiguous() and labels.is_contiguous() and labels.dtype == torch.int64
    is32 = logits.dtype == torch.float32
    dl = (logits if inplace else torch.empty_like(logits)) if want_grad else None
    out = torch.empty(2, device=logits.device, dtype=torch.float32)
    ws = torch.empty(int(lib().b200_weighted_ce_workspace_bytes(B, Lq)), dtype=torch.uint8, device=logits.device)
    vw = None if vocab_weight is None else vocab_weight.to(device=logits.device, dtype=torch.float32).contiguous()
    check(lib().b200_weighted_ce(ptr(logits), int(is32), V, ptr(labels), ptr(vw), B, Lq, V, float(grad_scale),
                                 ptr(dl), V, ptr(out), ptr(ws), ws.numel(), stream_ptr()), "b200_weighted_ce")
    return out[0], out[1], dl


def grad_sq_norm(grad, out2=None, accumulate=False, max_norm=0.0):
    """out2[0] (+)= sum(grad^2), out2[1] = clip coefficient for max_norm. grad: contiguous bf16 or fp32 tensor."""
    if out2 is None:
        out2 = torch.zeros(2, device=grad.device, dtype=torch.float32)
    ws = torch.empty(int(lib().b200_grad_norm_workspace_bytes()), dtype=torch.uint8, device=grad.device)
    check(lib().b200_grad_sq_norm(ptr(grad), int(grad.dtype == torch.float32), grad.numel(), int(accumulate),
                                  float(max_norm), ptr(out2), ptr(ws), ws.numel(), stream_ptr()), "b200_grad_sq_norm")
    return out2


def dropout(x, p, seed, out=None, accumulate=False):
    """out (+)= mask(seed) * x / (1 - p) on bf16 (b200_dropout); out=None allocates, out=x works in place."""
    if out is None:
        out = torch.empty_like(x)
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and x.is_contiguous() and out.is_contiguous()
    check(lib().b200_dropout(ptr(x), ptr(out), x.numel(), float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(accumulate),
                             stream_ptr()), "b200_dropout")
    return out


def adamw_step(master, param, grad, m, v, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1,
               clip_coef=None):
    """Fused AdamW on flat buffers: master/m/v fp32, grad bf16 or fp32, param bf16 copy (or None)."""
    check(lib().b200_adamw_step(ptr(master), ptr(param), ptr(grad), int(grad.dtype == torch.float32), ptr(m), ptr(v),
                                master.numel(), lr, beta1, beta2, eps, weight_decay, int(step), ptr(clip_coef),
                                stream_ptr()), "b200_adamw_step")


def layernorm(x, gamma, beta, eps, row_map=None, add=None, out=None):
    M = x.shape[0] if row_map is None else row_map.numel()
    D = x.shape[1]
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=torch.bfloat16)
    period = add.shape[0] if add is not None else 1
    check(lib().b200_layernorm(ptr(x), x.stride(0), ptr(row_map), ptr(add), period, ptr(gamma), ptr(beta), eps,
                               ptr(out), out.stride(0), M, D, stream_ptr()), "b200_layernorm")
    return out


def rmsnorm(x, w, eps, out=None):
    M, D = x.shape
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=torch.bfloat16)
    check(lib().b200_rmsnorm(ptr(x), x.stride(0), ptr(w), eps, ptr(out), out.stride(0), M, D, stream_ptr()),
          "b200_rmsnorm")
    return out


def flash_attention(q, k, v, causal=False, kv_start=None, kv_len=None, scale=None, return_lse=False):
    """q (B, Lq, H, d), k/v (B, Lk, H, d) views with unit stride on d -> (B, Lq, H, d) bf16
    (+ lse (B, H, Lq) fp32 when return_lse)."""
    B, Lq, H, d = q.shape
    Lk = k.shape[1]
    o = torch.empty((B, Lq, H, d), device=q.device, dtype=torch.bfloat16)
    scale = d ** -0.5 if scale is None else scale
    if return_lse:
        lse = torch.empty((B, H, Lq), device=q.device, dtype=torch.float32)
        check(lib().b200_flash_attention_lse(ptr(q), q.stride(0), q.stride(1), q.stride(2), ptr(k), k.stride(0),
                                             k.stride(1), k.stride(2), ptr(v), v.stride(0), v.stride(1), v.stride(2),
                                             ptr(o), o.stride(0), o.stride(1), o.stride(2), B, H, Lq, Lk, d,
                                             ptr(kv_start), ptr(kv_len), int(causal), scale, ptr(lse), stream_ptr()),
              "b200_flash_attention_lse")
        return o, lse
    check(lib().b200_flash_attention(ptr(q), q.stride(0), q.stride(1), q.stride(2), ptr(k), k.stride(0), k.stride(1),
                                     k.stride(2), ptr(v), v.stride(0), v.stride(1), v.stride(2), ptr(o), o.stride(0),
                                     o.stride(1), o.stride(2), B, H, Lq, Lk, d, ptr(kv_start), ptr(kv_len),
                                     int(causal), scale, stream_ptr()), "b200_flash_attention")
    return o


def flash_attention_bwd(q, k, v, o, d_o, lse, causal=False, kv_start=None, kv_len=None, scale=None, out=None):
    """Gradients (dq, dk, dv) of flash_attention; all tensors (B, L, H, d) views with unit stride on d, o and d_o
    share a layout; out = (dq, dk, dv) views with the strides of q / k / v (allocated here when None)."""
    B, Lq, H, d = q.shape
    Lk = k.shape[1]
    assert o.stride() == d_o.stride()
    if out is None:
        out = tuple(torch.empty(t.shape, device=t.device, dtype=t.dtype).as_strided(t.shape, t.stride())
                    if t.is_contiguous() else None for t in (q, k, v))
        assert all(t is not None for t in out), "pass out= for non-contiguous q / k / v views"
    dq, dk, dv = out
    assert dq.stride() == q.stride() and dk.stride() == k.stride() and dv.stride() == v.stride()
    scale = d ** -0.5 if scale is None else scale
    ws = torch.empty(int(lib().b200_flash_attention_bwd_workspace_bytes(B, H, Lq)), dtype=torch.uint8, device=q.device)
    check(lib().b200_flash_attention_bwd(ptr(q), q.stride(0), q.stride(1), q.stride(2), ptr(k), k.stride(0), k.stride(1),
                                         k.stride(2), ptr(v), v.stride(0), v.stride(1), v.stride(2), ptr(o), ptr(d_o),
                                         o.stride(0), o.stride(1), o.stride(2), ptr(lse), ptr(dq), ptr(dk), ptr(dv), B, H,
                                         Lq, Lk, d, ptr(kv_start), ptr(kv_len), int(causal), scale, ptr(ws), ws.numel(),
                                         stream_ptr()), "b200_flash_attention_bwd")
    return dq, dk, dv


def decode_attention(q, k_cache, v_cache, ctx, kv_start=None, splits=0):
    """q (B, H*128); caches (B, H, cap, 128) -> (B, H*128)."""
    B, H, cap, d = k_cache.shape
    o = torch.empty((B, H * d), device=q.device, dtype=torch.bfloat16)
    nb = lib().b200_decode_attention_workspace_bytes(B, H, ctx)
    ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=q.device)
    check(lib().b200_decode_attention(ptr(q), q.stride(0), ptr(k_cache), ptr(v_cache), ptr(o), o.stride(0), B, H, cap,
                                      ctx, ptr(kv_start), d ** -0.5, splits, ptr(ws), ws.numel(), stream_ptr()),
          "b200_decode_attention")
    return o


# ---- launch accounting ------------------------------------------------------------------------------------------
_graph_launches = 0   # kernels executed through CUDA-graph replays (the library only sees the capture)


def note_graph_replay(kernels):
    global _graph_launches
    _graph_launches += int(kernels)


def launch_count():
    """Kernels of this library executed so far in this process (direct launches + graph replays)."""
    return int(lib().b200_launch_count()) + _graph_launches


def set_option(name, value=True):
    """Tuning switches of the decode step (include/b200_mmor.h: b200_set_option): "pdl", "decode_tiles". Call before
    generate(): the decode CUDA graph is captured with whatever is set at that time."""
    check(lib().b200_set_option(name.encode(), int(value)), "b200_set_option")


def get_option(name):
    """Value of a switch: False / True for "pdl" and "fused_rope"; 0 / 1 / 2 for "decode_tiles"."""
    v = int(lib().b200_get_option(name.encode()))
    if v < 0:
        check(v, "b200_get_option")
    return v if name == "decode_tiles" else bool(v)


def set_pdl(on=True):
    set_option("pdl", on)


def pdl_enabled():
    return get_option("pdl")


def prof_enable(on=True):
    check(lib().b200_prof_enable(int(on)), "b200_prof_enable")


def prof_collect():
    """{family: dict(ms, bytes, flops, launches)} for the launches since prof_enable(True). Blocks on the events."""
    n = lib().b200_prof_family_count()
    ms, by, fl = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_double * n)()
    la = (ctypes.c_longlong * n)()
    check(lib().b200_prof_collect(ms, by, fl, la), "b200_prof_collect")
    return {lib().b200_prof_family_name(i).decode(): dict(ms=ms[i], bytes=by[i], flops=fl[i], launches=la[i])
            for i in range(n)}


# ---- fine-tuning of the seg-mask CNN and embed_tokens (train_extras.cu) -------------------------------------------
def segmask_forward_train(weights, cls, out, out_ld, row_map, cdll=None, ptr_fn=None, stream_fn=None):
    """Forward of the seg-mask CNN that keeps its activations. weights: SegmaskWeights; cls (n, 32, 32) uint8.
    Writes the bf16 token rows into `out` and returns the activation buffer for segmask_backward."""
    cdll, ptr_fn, stream_fn = cdll or lib(), ptr_fn or ptr, stream_fn or stream_ptr
    n = cls.shape[0]
    acts = torch.empty(int(cdll.b200_segmask_train_acts_bytes(n)), dtype=torch.uint8, device=cls.device)
    rc = cdll.b200_segmask_forward_train(ctypes.byref(weights), ptr_fn(cls), n, ptr_fn(acts), acts.numel(), ptr_fn(out),
                                         out_ld, ptr_fn(row_map), stream_fn())
    _check_with(cdll, rc, "b200_segmask_forward_train")
    return acts


def segmask_backward(weights, cls, acts, d_out, d_ld, row_map, d_conv_w, d_conv_b, d_emb, accumulate=False, cdll=None,
                     ptr_fn=None, stream_fn=None):
    """d_conv_w / d_conv_b: lists of 5 fp32 tensors, d_emb (30, 8) fp32: overwritten, or added to when accumulate."""
    cdll, ptr_fn, stream_fn = cdll or lib(), ptr_fn or ptr, stream_fn or stream_ptr
    n = cls.shape[0]
    ws = torch.empty(int(cdll.b200_segmask_backward_workspace_bytes(n)), dtype=torch.uint8, device=cls.device)
    pw = (ctypes.c_void_p * 5)(*[ptr_fn(t).value for t in d_conv_w])
    pb = (ctypes.c_void_p * 5)(*[ptr_fn(t).value for t in d_conv_b])
    rc = cdll.b200_segmask_backward(ctypes.byref(weights), ptr_fn(cls), n, ptr_fn(acts), ptr_fn(d_out), d_ld,
                                    ptr_fn(row_map), pw, pb, ptr_fn(d_emb), int(accumulate), ptr_fn(ws), ws.numel(),
                                    stream_fn())
    _check_with(cdll, rc, "b200_segmask_backward")


def embed_grad(d_rows, token_of_row, d_table, accumulate=False, cdll=None, ptr_fn=None, stream_fn=None):
    """nn.Embedding backward: d_table[token] (+)= sum of the bf16 rows of d_rows whose token_of_row equals token
    (token_of_row: host int array, < 0 = row carries no token). Rows are grouped by token on the host (stable), so the
    summation order is fixed."""
    import numpy as np
    cdll, ptr_fn, stream_fn = cdll or lib(), ptr_fn or ptr, stream_fn or stream_ptr
    tok = np.asarray(token_of_row).reshape(-1)
    rows = np.nonzero(tok >= 0)[0]
    if rows.size == 0:
        return d_table
    order = rows[np.argsort(tok[rows], kind="stable")]
    sorted_tok = tok[order]
    heads = np.nonzero(np.concatenate([[True], sorted_tok[1:] != sorted_tok[:-1]]))[0]
    seg_start = np.concatenate([heads, [len(order)]]).astype(np.int32)
    seg_token = sorted_tok[heads].astype(np.int32)
    if int(seg_token.max()) >= d_table.shape[0]:
        raise B200Error("embed_grad: token id outside the embedding table")
    dev = d_rows.device
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    # the index tensors must outlive the launch: a temporary freed inside the call expression hands its block to the
    # next allocation, and all three arrays would alias
    row_list, seg_start_d, seg_token_d = t(order.astype(np.int32)), t(seg_start), t(seg_token)
    rc = cdll.b200_embed_grad(ptr_fn(d_rows), d_rows.stride(0), ptr_fn(row_list), ptr_fn(seg_start_d),
                              ptr_fn(seg_token_d), len(seg_token), d_rows.shape[1], ptr_fn(d_table), int(accumulate),
                              stream_fn())
    _check_with(cdll, rc, "b200_embed_grad")
    return d_table


def _check_with(cdll, rc, what):
    if rc != 0:
        msg = cdll.b200_last_error()
        raise B200Error(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")
