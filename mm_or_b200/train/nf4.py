"""NF4 storage of the frozen base weights -- the `--bits 4` half of the reference's QLoRA recipe (README.md:120-124;
LLaVA/llava/train/train.py:1098-1114: BitsAndBytesConfig(load_in_4bit, bnb_4bit_quant_type='nf4',
bnb_4bit_use_double_quant=True, compute dtype bf16) turns every nn.Linear outside mm_projector / image_pooler -- the
seven projections of every decoder layer and lm_head -- into a bitsandbytes Linear4bit).

What that means numerically: the adapters are trained against Q(W), the 4-bit round trip of the base weight, not against
W; Linear4bit dequantises Q(W) to bf16 in front of every matmul. Here:
  * `quantize` / `dequantize`            the storage kernels (csrc/nf4.cu: b200_nf4_quantize / b200_nf4_dequantize)
  * `double_quantize` / `double_dequantize`  bitsandbytes' second-level 8-bit quantisation of the absmax vector -- load-time
                                         work on n / 64 values, plain torch ops on whatever device the vector is on
  * `Nf4Weight`                          a packed weight (0.516 B per parameter with double quantisation, 0.5625 without)
  * `qlora_base_`                        replaces the named bf16 weights by their round trip Q(W) in place (what the
                                         LoRA step multiplies with) and returns the packed copies.
One B200 holds the bf16 base with room to spare (13.5 GB of 180), so the fine-tune step keeps multiplying with the
dequantised bf16 copy -- the same values Linear4bit would produce on the fly -- and the packed form is what a memory-
bound deployment would keep instead (dequantising one layer ahead of its GEMMs costs 2.6 B of traffic per parameter and
step, < 1 % of the measured LoRA step). bitsandbytes is not available offline: the algorithm is restated from its
published source and cannot be pinned against it here (oracle/nf4_oracle.py says what it IS checked against).
"""
import torch

from .. import _lib as L

BLOCK = 64
ABSMAX_BLOCK = 256


class Nf4Ops:
    """Entry points + pointer plumbing; the CPU tests inject the kernel emulator's library (tests/emu_lib.py)."""

    def __init__(self, lib=None, ptr=None, stream_ptr=None, check=None):
        self.lib = lib if lib is not None else L.lib()
        self.ptr = ptr if ptr is not None else L.ptr
        self.stream_ptr = stream_ptr if stream_ptr is not None else L.stream_ptr
        self.check = check if check is not None else L.check


def quantize(w, ops=None):
    """w: bf16 tensor with numel % 64 == 0 (every Llama projection qualifies) -> (packed uint8 (n / 2,), absmax fp32
    (n / 64,)), blocks running over the flattened (row-major) weight like bitsandbytes' quantize_4bit."""
    ops = ops or Nf4Ops()
    if w.dtype != torch.bfloat16:
        raise TypeError(f"nf4.quantize expects bf16 weights, got {w.dtype}")
    w = w.contiguous()
    n = w.numel()
    if n % BLOCK:
        raise ValueError(f"nf4.quantize: {n} elements are not a multiple of the block size {BLOCK}")
    packed = torch.empty(n // 2, dtype=torch.uint8, device=w.device)
    absmax = torch.empty(n // BLOCK, dtype=torch.float32, device=w.device)
    ops.check(ops.lib.b200_nf4_quantize(ops.ptr(w), n, ops.ptr(packed), ops.ptr(absmax), ops.stream_ptr()),
              "b200_nf4_quantize")
    return packed, absmax


def dequantize(packed, absmax, shape, out=None, ops=None):
    ops = ops or Nf4Ops()
    n = packed.numel() * 2
    if absmax.numel() * BLOCK != n:
        raise ValueError("nf4.dequantize: absmax does not match the packed tensor")
    if out is None:
        out = torch.empty(n, dtype=torch.bfloat16, device=packed.device)
    ops.check(ops.lib.b200_nf4_dequantize(ops.ptr(packed), ops.ptr(absmax.contiguous()), n, ops.ptr(out),
                                          ops.stream_ptr()), "b200_nf4_dequantize")
    return out.view(shape)


def dynamic_map(signed=True, max_exponent_bits=7, total_bits=8):
    """bitsandbytes' functional.create_dynamic_map(): the sorted 256-entry code book of its 8-bit 'dynamic' type --
    per decade 10^-6 .. 10^0 the midpoints of a linear grid over [0.1, 1] (2, 3, 5, ... 65 grid points), both signs,
    plus 0 and 1. Evaluated in double precision and rounded once to fp32 (the library evaluates the grid with a fp32
    torch.linspace, whose last bit depends on the torch build: entries can differ from it by one fp32 ulp)."""
    data = []
    non_sign_bits = total_bits - (1 if signed else 0)
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1

    def midpoints(items):
        grid = [0.1 + 0.9 * k / (items - 1) for k in range(items)]
        return [(a + b) / 2.0 for a, b in zip(grid[:-1], grid[1:])]

    i = 0
    for i in range(max_exponent_bits):
        items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                    else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        scale = 10.0 ** (-(max_exponent_bits - 1) + i)
        data += [scale * m for m in midpoints(items)]
        if signed:
            data += [-scale * m for m in midpoints(items)]
    if additional_items > 0:
        scale = 10.0 ** (-(max_exponent_bits - 1) + i)
        data += [scale * m for m in midpoints(additional_items + 1)]
        if signed:
            data += [-scale * m for m in midpoints(additional_items + 1)]
    data += [0.0, 1.0]
    data += [0.0] * (2 ** total_bits - len(data))
    return torch.tensor(sorted(data), dtype=torch.float64).to(torch.float32)


def double_quantize(absmax):
    """quantize_4bit(compress_statistics=True): offset = mean(absmax); the centred vector in blocks of 256 -> 8-bit
    codes of the dynamic code book (nearest entry, ties to the lower one) + one fp32 scale per block."""
    a = absmax.float()
    offset = a.double().mean().float()      # (summed in fp64: independent of the reduction order of the device)
    c = a - offset
    n = c.numel()
    nb = -(-n // ABSMAX_BLOCK)
    pad = torch.zeros(nb * ABSMAX_BLOCK, dtype=torch.float32, device=a.device)
    pad[:n] = c
    pad = pad.view(nb, ABSMAX_BLOCK)
    absmax2 = pad.abs().amax(1)
    norm = torch.where(absmax2[:, None] > 0, pad / absmax2[:, None], torch.zeros_like(pad))
    book = dynamic_map().to(a.device)
    hi = torch.searchsorted(book, norm.contiguous(), right=False).clamp_(1, 255)
    lo = hi - 1
    q = torch.where((norm - book[lo]).abs() <= (book[hi] - norm).abs(), lo, hi).to(torch.uint8)
    return q.view(-1)[:n].contiguous(), absmax2, offset


def double_dequantize(q, absmax2, offset):
    book = dynamic_map().to(q.device)
    scale = absmax2.repeat_interleave(ABSMAX_BLOCK)[:q.numel()]
    return book[q.long()] * scale + offset


class Nf4Weight:
    """A weight in Linear4bit's storage: packed codes + (double-quantised) block maxima."""

    def __init__(self, w, double_quant=True, ops=None):
        self.shape, self.ops = tuple(w.shape), ops
        self.packed, absmax = quantize(w, ops)
        self.double_quant = double_quant
        if double_quant:
            self.q_absmax, self.absmax2, self.offset = double_quantize(absmax)
            self.absmax = None
        else:
            self.absmax = absmax

    def block_maxima(self):
        return double_dequantize(self.q_absmax, self.absmax2, self.offset) if self.double_quant else self.absmax

    def dequantize(self, out=None):
        return dequantize(self.packed, self.block_maxima(), self.shape, out=out, ops=self.ops)

    def nbytes(self):
        if self.double_quant:
            return self.packed.numel() + self.q_absmax.numel() + 4 * self.absmax2.numel() + 4
        return self.packed.numel() + 4 * self.absmax.numel()


def is_quantized_linear(name):
    """The parameters BitsAndBytesConfig(llm_int8_skip_modules=['mm_projector', 'image_pooler']) quantises in the
    reference's model at load time (train.py:1098-1114; the vision tower is attached afterwards, unquantised): the seven
    projections of every decoder layer and lm_head."""
    if name == "lm_head.weight":
        return True
    return name.startswith("model.layers.") and name.endswith("_proj.weight")


def qlora_base_(sd, names=None, double_quant=True, ops=None):
    """Replace sd[name] (bf16) by the NF4 round trip Q(W) in place for every quantised Linear; returns {name: Nf4Weight}."""
    names = [k for k in sd if is_quantized_linear(k)] if names is None else list(names)
    stored = {}
    for k in names:
        w = sd[k]
        q = Nf4Weight(w, double_quant=double_quant, ops=ops)
        if w.is_contiguous():
            q.dequantize(out=w.view(-1))          # in place: every alias of the tensor sees Q(W)
        else:
            sd[k] = q.dequantize()
        stored[k] = q
    return stored
