"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the 4-bit NormalFloat quantisation the reference's QLoRA recipe gets
from bitsandbytes (call site: LLaVA/llava/train/train.py:1098-1114, BitsAndBytesConfig(load_in_4bit=True,
bnb_4bit_quant_type='nf4', bnb_4bit_use_double_quant=True, bnb_4bit_compute_dtype=bf16)).

PARITY UNPINNED against bitsandbytes itself: bitsandbytes==0.41.0 (scene_graph_generation/pyproject.toml:18) is a
third-party dependency that is neither vendored under /root/reference nor installed in this image, and the reference
holds no golden vectors for it. What is restated here is its published algorithm:
  * functional.py get_4bit_type('nf4'): the 16 levels (QLoRA, Dettmers et al. 2023, appendix E);
  * csrc/kernels.cu kQuantizeBlockwise<T, 64, ..., NF4>: per block of 64 values absmax in fp32, every value multiplied
    by 1 / absmax and mapped by dQuantizeNF4 -- a decision tree whose thresholds are the midpoints between neighbouring
    levels, strict '>' (a value on a midpoint takes the lower level; NaN, from the all-zero block's 0 * inf, takes code
    0) --, two codes per byte with the even element in the high nibble;
  * kDequantizeBlockwise: level[code] * absmax, cast to the compute dtype (bf16);
  * functional.py quantize_4bit(compress_statistics=True) ("double quantisation"): offset = absmax.mean(); the centred
    absmax vector is quantised in blocks of 256 to 8 bits with the "dynamic" code book of create_dynamic_map() and
    dequantised as code_book[q] * absmax2 + offset.
The known-answer checks this module CAN be held to are in tests/test_nf4.py: the published level table, thresholds ==
midpoints, exactly representable inputs, idempotence, the error bound of nearest-level rounding.
"""
import numpy as np

BLOCK = 64
LEVELS = np.array([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                   -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                   0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
                   0.7229568362236023, 1.0], dtype=np.float64)
# thresholds of the decision tree, as float32 like the literals of the kernel
THRESHOLDS = ((LEVELS[:-1] + LEVELS[1:]) / 2).astype(np.float32)


def bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r


def nearest_code(x):
    """x float32 array in [-1, 1] (or NaN) -> uint8 codes: number of thresholds strictly below x."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        return (x[..., None] > THRESHOLDS).sum(-1).astype(np.uint8)


def quantize(w_bf16_bits):
    """w: uint16 array of bf16 bit patterns, size a multiple of 64 -> (packed uint8 (n/2,), absmax float32 (n/64,))."""
    w = bf16_to_f32(np.asarray(w_bf16_bits, dtype=np.uint16).reshape(-1, BLOCK))
    absmax = np.abs(w).max(1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (np.float32(1.0) / absmax).astype(np.float32)
        scaled = (w * inv[:, None]).astype(np.float32)
    codes = nearest_code(scaled)
    packed = ((codes[:, 0::2] << 4) | codes[:, 1::2]).astype(np.uint8).reshape(-1)
    return packed, absmax


def dequantize(packed, absmax):
    """-> uint16 bf16 bit patterns (n,)"""
    p = np.asarray(packed, dtype=np.uint8).reshape(-1, BLOCK // 2)
    codes = np.empty((p.shape[0], BLOCK), dtype=np.uint8)
    codes[:, 0::2] = p >> 4
    codes[:, 1::2] = p & 15
    vals = LEVELS.astype(np.float32)[codes] * np.asarray(absmax, dtype=np.float32)[:, None]
    return f32_to_bf16(vals.astype(np.float32)).reshape(-1)


def dynamic_map(signed=True, max_exponent_bits=7, total_bits=8):
    """bitsandbytes functional.create_dynamic_map: the 8-bit 'dynamic' code book (sorted, 256 entries in [-1, 1]).
    Grid evaluated in float64 and rounded once to float32 (the library's own fp32 torch.linspace is build dependent in
    its last bit)."""
    data = []
    non_sign_bits = total_bits - (1 if signed else 0)
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    i = 0
    for i in range(max_exponent_bits):
        fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                             else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        boundaries = np.linspace(0.1, 1, fraction_items, dtype=np.float64)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10.0 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10.0 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = np.linspace(0.1, 1, additional_items + 1, dtype=np.float64)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10.0 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10.0 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    data.append(0)
    data.append(1.0)
    assert len(data) <= 2 ** total_bits
    data += [0] * (2 ** total_bits - len(data))
    return np.sort(np.asarray(data, dtype=np.float64)).astype(np.float32)


def double_quantize(absmax, block=256):
    """-> (q uint8 (n,), absmax2 float32 (ceil(n / 256),), offset float32): nearest code-book entry of the centred,
    block-normalised absmax (ties to the lower entry)."""
    a = np.asarray(absmax, dtype=np.float32)
    offset = np.float32(a.mean(dtype=np.float64))   # fp64 sum: the library's fp32 tree order is not reproducible here
    c = a - offset
    n = c.size
    nb = -(-n // block)
    pad = np.zeros(nb * block, dtype=np.float32)
    pad[:n] = c
    pad = pad.reshape(nb, block)
    absmax2 = np.abs(pad).max(1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        norm = np.where(absmax2[:, None] > 0, pad / absmax2[:, None], 0).astype(np.float32)
    book = dynamic_map()
    hi = np.clip(np.searchsorted(book, norm, side="left"), 1, 255)
    lo = hi - 1
    q = np.where(np.abs(norm - book[lo]) <= np.abs(book[hi] - norm), lo, hi).astype(np.uint8)
    return q.reshape(-1)[:n], absmax2, offset


def double_dequantize(q, absmax2, offset, block=256):
    book = dynamic_map()
    q = np.asarray(q, dtype=np.uint8)
    scale = np.repeat(np.asarray(absmax2, dtype=np.float32), block)[:q.size]
    return (book[q] * scale + np.float32(offset)).astype(np.float32)
