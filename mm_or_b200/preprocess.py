"""GPU image preprocessing in front of the vision tower, mirror of
`process_images(images, image_processor, model_cfg)` (LLaVA/llava/mm_utils.py:29-40) for image_aspect_ratio 'pad' (the
MM2SG setting) and for the plain CLIPImageProcessor path: expand2square with the CLIP mean colour, PIL-exact bicubic
resize to 336 on the shortest edge, centre crop, rescale, normalise, bf16 -- one call of b200_preprocess_images per
batch of equally sized frames. The host computes the (cached) resampling tables and, by default, decodes the JPEGs
with Pillow like the reference; frames passed as compressed `bytes` are decoded on the GPU by nvJPEG (jpeg.py).

The tables follow Pillow's precompute_coeffs / normalize_coeffs_8bpc (src/libImaging/Resample.c, Pillow is a
dependency of the reference through transformers' CLIPImageProcessor): bicubic a = -0.5, support 2 * max(scale, 1),
double precision, normalised per output pixel, rounded to 22-bit fixed point.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib as L

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_PRECISION_BITS = 32 - 8 - 2
_tables = {}


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1,
                    np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


def resample_tables(in_size, out_size):
    """(bounds int32 [out, 2] = (first input index, taps), coeffs int32 [out, ksize]) for one axis."""
    key = (in_size, out_size)
    if key in _tables:
        return _tables[key]
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = _bicubic((np.arange(xmax, dtype=np.float64) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                      # same accumulation order as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, (-0.5 + kk * (1 << _PRECISION_BITS)).astype(np.int64),
                     (0.5 + kk * (1 << _PRECISION_BITS)).astype(np.int64)).astype(np.int32)
    _tables[key] = (bounds, fixed)
    return _tables[key]


class GpuImageProcessor:
    def __init__(self, size=336, image_mean=CLIP_MEAN, image_std=CLIP_STD, device="cuda"):
        self.size = size
        self.image_mean, self.image_std = tuple(image_mean), tuple(image_std)
        self.device = torch.device(device)
        self._dev_tables = {}
        self._ws = L.Workspace()
        self._jpeg = None

    def _table(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self._dev_tables:
            b, k = resample_tables(in_size, out_size)
            self._dev_tables[key] = (torch.from_numpy(b).to(self.device), torch.from_numpy(k).to(self.device), k.shape[1])
        return self._dev_tables[key]

    def preprocess(self, images, pad=True):
        """images: list of equally sized RGB frames (PIL images, uint8 HWC numpy arrays or uint8 HWC tensors), a list
        of JPEG files as `bytes` (decoded on the GPU with nvJPEG, mm_or_b200/jpeg.py), or one (N, H, W, 3) uint8
        array / tensor. Returns (N, 3, size, size) bf16 on the device."""
        if isinstance(images, (list, tuple)) and len(images) and isinstance(images[0], (bytes, bytearray, memoryview)):
            # compressed JPEG files: decoded on the GPU by nvJPEG (mm_or_b200/jpeg.py; opt-in by passing bytes)
            if self._jpeg is None:
                from .jpeg import GpuJpegDecoder
                self._jpeg = GpuJpegDecoder(self.device)
            batch = self._jpeg.decode_batch(images)
        elif isinstance(images, (list, tuple)):
            frames = []
            for im in images:
                if hasattr(im, "convert"):                              # PIL
                    im = np.asarray(im.convert("RGB"))
                frames.append(torch.as_tensor(np.array(im, copy=True)) if not torch.is_tensor(im) else im)
            batch = torch.stack(frames)
        else:
            batch = images if torch.is_tensor(images) else torch.as_tensor(np.ascontiguousarray(images))
        if batch.dtype != torch.uint8 or batch.ndim != 4 or batch.shape[-1] != 3:
            raise ValueError("expected uint8 RGB frames of shape (N, H, W, 3)")
        batch = batch.to(self.device, non_blocking=True).contiguous()
        n, H, W, _ = batch.shape
        S_h, S_w = (max(H, W), max(H, W)) if pad else (H, W)
        # get_resize_output_image_size(shortest_edge = size), then centre crop size x size
        short, long_ = (S_w, S_h) if S_w <= S_h else (S_h, S_w)
        new_long = int(self.size * long_ / short)
        res_w, res_h = (self.size, new_long) if S_w <= S_h else (new_long, self.size)
        bx = kx = by = ky = None
        ksx = ksy = 0
        if res_w != S_w:
            bx, kx, ksx = self._table(S_w, res_w)
        if res_h != S_h:
            by, ky, ksy = self._table(S_h, res_h)
        out = torch.empty((n, 3, self.size, self.size), device=self.device, dtype=torch.bfloat16)
        lib = L.lib()
        ws = self._ws.get(int(lib.b200_preprocess_workspace_bytes(n, S_h, res_w)), self.device)
        bg = (ctypes.c_uint8 * 3)(*[int(m * 255) for m in self.image_mean])     # mm_utils.py:35
        mean = (ctypes.c_float * 3)(*self.image_mean)
        std = (ctypes.c_float * 3)(*self.image_std)
        L.check(lib.b200_preprocess_images(L.ptr(batch), n, H, W, int(bool(pad)), bg, L.ptr(bx), L.ptr(kx), ksx,
                                           L.ptr(by), L.ptr(ky), ksy, res_h, res_w, self.size, mean, std, L.ptr(out),
                                           L.ptr(ws), ws.numel(), L.stream_ptr()), "b200_preprocess_images")
        return out

    __call__ = preprocess


def process_images(images, image_processor, model_cfg):
    """Drop-in for llava.mm_utils.process_images when `image_processor` is a GpuImageProcessor: returns the stacked
    (N, 3, S, S) tensor for one sample, or (1, N, 3, S, S) for N > 1 like the reference (mm_utils.py:38-39)."""
    pad = getattr(model_cfg, "image_aspect_ratio", None) == "pad"
    out = image_processor.preprocess(list(images), pad=pad)
    return out if out.shape[0] == 1 else out[None]
