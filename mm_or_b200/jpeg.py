"""JPEG decoding on the GPU in front of the image preprocessing ("JPEG decode -> pad-to-square ->
bicubic resize -> normalise -> bf16"). The reference decodes with Pillow on the host, one frame at a time in the main
process (`Image.open(path).convert('RGB')`, scene_graph_prediction_model.py:82-113: 7 frames of 2048 x 1536 per sample).

This module binds NVIDIA's nvJPEG (libnvjpeg.so of the CUDA toolkit; a LIBRARY call like cuBLAS would be -- Huffman
decoding is not a kernel this project writes) through ctypes: compressed bytes in, an (H, W, 3) uint8 RGB tensor on the
device out, which `GpuImageProcessor.preprocess` (mm_or_b200/preprocess.py -> b200_preprocess_images) consumes without a
round trip through host memory. Parity: nvJPEG's inverse DCT / chroma upsampling are not bit-identical to libjpeg-turbo
(what Pillow links); the decoded pixels agree within a few grey levels (tests/test_preprocess.py states the bound), so
this path is OPT-IN (`decode="nvjpeg"`); the default keeps Pillow's pixels and the bit-exact bar of the integer stages.
No fallback: a missing library raises.
"""
import ctypes
import os

import torch

from . import _lib as L

_OUTPUT_RGBI = 5                  # nvjpegOutputFormat_t: interleaved RGB in channel[0]
_STATUS = {1: "NOT_INITIALIZED", 2: "INVALID_PARAMETER", 3: "BAD_JPEG", 4: "JPEG_NOT_SUPPORTED", 5: "ALLOCATOR_FAILURE",
           6: "EXECUTION_FAILED", 7: "ARCH_MISMATCH", 8: "INTERNAL_ERROR", 9: "IMPLEMENTATION_NOT_SUPPORTED",
           10: "INCOMPLETE_BITSTREAM"}


class _NvjpegImage(ctypes.Structure):             # nvjpegImage_t (nvjpeg.h)
    _fields_ = [("channel", ctypes.c_void_p * 4), ("pitch", ctypes.c_size_t * 4)]


def _load():
    names = ["libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12",
             "/usr/local/cuda/lib64/libnvjpeg.so"]
    try:
        import nvidia.nvjpeg as _pkg             # the pip wheel torch / torchvision depend on, when present
        names.insert(0, os.path.join(os.path.dirname(_pkg.__file__), "lib", "libnvjpeg.so.12"))
    except Exception:  # noqa: BLE001
        pass
    last = None
    for n in names:
        try:
            return ctypes.CDLL(n)
        except OSError as e:
            last = e
    raise L.B200Error(f"libnvjpeg not found ({last}); GPU JPEG decoding has no fallback -- decode on the host instead")


class GpuJpegDecoder:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.lib = _load()
        self.handle, self.state = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self.lib.nvjpegCreateSimple(ctypes.byref(self.handle)), "nvjpegCreateSimple")
        self._check(self.lib.nvjpegJpegStateCreate(self.handle, ctypes.byref(self.state)), "nvjpegJpegStateCreate")
        self.lib.nvjpegGetImageInfo.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t,
                                                ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                                ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        self.lib.nvjpegDecode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t,
                                          ctypes.c_int, ctypes.POINTER(_NvjpegImage), ctypes.c_void_p]

    @staticmethod
    def _check(rc, what):
        if rc != 0:
            raise L.B200Error(f"{what} failed: nvjpeg status {rc} ({_STATUS.get(rc, '?')})")

    def image_size(self, data):
        comps, css = ctypes.c_int(), ctypes.c_int()
        w, h = (ctypes.c_int * 4)(), (ctypes.c_int * 4)()
        self._check(self.lib.nvjpegGetImageInfo(self.handle, data, len(data), ctypes.byref(comps), ctypes.byref(css),
                                                w, h), "nvjpegGetImageInfo")
        return int(h[0]), int(w[0])

    def decode(self, data, out=None):
        """data: bytes of one JPEG file -> (H, W, 3) uint8 RGB tensor on the device (written into `out` if given).
        Enqueued on the current stream."""
        data = bytes(data)
        H, W = self.image_size(data)
        if out is None:
            out = torch.empty((H, W, 3), dtype=torch.uint8, device=self.device)
        elif tuple(out.shape) != (H, W, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
            raise ValueError(f"decode: out must be a contiguous uint8 ({H}, {W}, 3) tensor")
        img = _NvjpegImage()
        img.channel[0] = out.data_ptr()
        img.pitch[0] = W * 3
        self._check(self.lib.nvjpegDecode(self.handle, self.state, data, len(data), _OUTPUT_RGBI, ctypes.byref(img),
                                          L.stream_ptr()), "nvjpegDecode")
        return out

    def decode_batch(self, files):
        """Equally sized JPEG files (the views of a sample) -> (N, H, W, 3) uint8 on the device."""
        files = [bytes(f) for f in files]
        if not files:
            raise ValueError("decode_batch: no files")
        H, W = self.image_size(files[0])
        out = torch.empty((len(files), H, W, 3), dtype=torch.uint8, device=self.device)
        for i, f in enumerate(files):
            if self.image_size(f) != (H, W):
                raise ValueError("decode_batch: frames of one batch must have the same size")
            self.decode(f, out=out[i])
        return out

    def __del__(self):
        try:
            if self.state:
                self.lib.nvjpegJpegStateDestroy(self.state)
            if self.handle:
                self.lib.nvjpegDestroy(self.handle)
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass
