"""CPU oracle for the image preprocessing in front of the MM2SG hot path -- TEST INFRASTRUCTURE ONLY.

Reference path: LLaVA/llava/mm_utils.py:14-40 `expand2square` + `process_images` (image_aspect_ratio
== 'pad') -> HF CLIPImageProcessor.preprocess (transformers==4.31.0, un-vendored): resize(shortest_edge = 336,
PIL BICUBIC) -> center_crop(336) -> rescale(1/255) -> normalize(CLIP mean / std) -> float32 CHW; the caller casts to
bf16 (scene_graph_prediction_model.py:119).

The arithmetic lives in third-party code absent from /root/reference:
  * Pillow `ImagingResample` (src/libImaging/Resample.c): separable antialiased bicubic (a = -0.5, support 2 * scale),
    coefficients in double normalised per output pixel, converted to 22-bit fixed point, horizontal pass then vertical
    pass, each rounded and clipped to uint8. Restated below operation by operation; tests/test_preprocess_oracle.py
    pins it bit-exactly against the Pillow installed here on random images and sizes.
  * transformers 4.31 image_transforms.rescale / normalize: uint8 * (1/255) in float64 -> float32, then
    (x - mean) / std in float32.
"""
import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    r1 = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    r2 = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, r1, np.where(x < 2.0, r2, 0.0))


def resample_coeffs(in_size, out_size):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size): returns (bounds int32 [out, 2] =
    (first input index, tap count), coeffs int32 [out, ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        x = np.arange(xmax, dtype=np.float64)
        w = _bicubic((x + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                                   # sequential double accumulation like the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64),
                     (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64)).astype(np.int32)
    return bounds, fixed


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img, out_h, out_w):
    """img (H, W, C) uint8 -> (out_h, out_w, C) uint8 exactly as PIL.Image.resize((out_w, out_h), BICUBIC)."""
    H, W, C = img.shape
    src = img.astype(np.int64)
    if out_w != W:
        bx, kx = resample_coeffs(W, out_w)
        tmp = np.empty((H, out_w, C), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = bx[xx]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, x0:x0 + n, :], kx[xx, :n].astype(np.int64), axes=([1], [0]))
            tmp[:, xx, :] = _clip8(acc)
        src = tmp.astype(np.int64)
    if out_h != H:
        by, ky = resample_coeffs(H, out_h)
        out = np.empty((out_h, src.shape[1], C), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = by[yy]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(ky[yy, :n].astype(np.int64), src[y0:y0 + n], axes=([0], [0]))
            out[yy] = _clip8(acc)
        return out
    return src.astype(np.uint8)


def expand2square(img, mean=CLIP_MEAN):
    """mm_utils.py:14-26 with background tuple(int(x * 255) for x in image_mean)."""
    H, W, C = img.shape
    if W == H:
        return img
    S = max(W, H)
    out = np.empty((S, S, C), dtype=np.uint8)
    out[:] = np.array([int(m * 255) for m in mean], dtype=np.uint8)
    if W > H:
        top = (W - H) // 2
        out[top:top + H] = img
    else:
        left = (H - W) // 2
        out[:, left:left + W] = img
    return out


def clip_preprocess(img, size=336, pad=True, mean=CLIP_MEAN, std=CLIP_STD):
    """uint8 HWC RGB -> float32 (3, size, size): process_images ('pad') + CLIPImageProcessor.preprocess (4.31)."""
    if pad:
        img = expand2square(img, mean)
    H, W, _ = img.shape
    short, long_ = (W, H) if W <= H else (H, W)
    new_short, new_long = size, int(size * long_ / short)          # get_resize_output_image_size(shortest_edge)
    out_w, out_h = (new_short, new_long) if W <= H else (new_long, new_short)
    r = resize_bicubic_u8(img, out_h, out_w)
    top, left = (out_h - size) // 2, (out_w - size) // 2
    r = r[top:top + size, left:left + size]
    x = (r * (1 / 255)).astype(np.float32)                          # rescale: uint8 * python float -> float64 -> float32
    x = (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))
