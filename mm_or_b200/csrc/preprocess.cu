// Image preprocessing in front of the vision tower: pad-to-square with the CLIP mean colour,
// antialiased bicubic resize to 336 on the shortest edge, centre crop, rescale, normalise, cast to bf16 -- on the GPU,
// bit-exact with the reference's host path.
//
// Reference: LLaVA/llava/mm_utils.py:14-40 (expand2square + process_images, image_aspect_ratio == 'pad') -> HF
// CLIPImageProcessor.preprocess (transformers 4.31: PIL resize BICUBIC -> center_crop -> rescale 1/255 -> normalize)
// -> `.to(bfloat16)` in scene_graph_prediction_model.py:119. In the reference this is PIL work on the single main
// process: seven 2048 x 1536 frames per sample, ~20-30 ms each.
//
// The resize is Pillow's ImagingResample (Resample.c), restated: separable passes (horizontal, then vertical), per
// output pixel a window of 22-bit fixed-point coefficients (precomputed on the host exactly like precompute_coeffs /
// normalize_coeffs_8bpc: mm_or_b200/preprocess.py, pinned against Pillow in the CPU tests), integer accumulation
// starting from 2^21, arithmetic shift by 22 and clipping to uint8 after EACH pass. Integer work: results are
// bit-identical to Pillow. The padding is never materialised: the horizontal pass reads the source image through a
// virtual square canvas. The float tail follows numpy's types: uint8 * (1/255) in double -> float32, then
// (x - mean) / std in float32 (IEEE subtract and divide, no FMA contraction), round-to-nearest-even to bf16.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

static constexpr int kPrecisionBits = 32 - 8 - 2;

struct PreArgs {
  const uint8_t* img;  // [n, H, W, 3]
  int n, H, W;
  int S_h, S_w;        // canvas (= padded) size
  int top, left;       // where the image sits on the canvas
  uint8_t bg[3];
  const int* bx;       // [res_w, 2] (first canvas column, taps)
  const int* kx;       // [res_w, ksize_x]
  int ksize_x;
  const int* by;       // [res_h, 2]
  const int* ky;       // [res_h, ksize_y]
  int ksize_y;
  int res_h, res_w;    // size after the resize
  int crop_top, crop_left, out;  // centre crop window (out x out)
  float mean[3], stdv[3];
  uint8_t* tmp;        // [n, S_h, res_w, 3] after the horizontal pass
  bf16* dst;           // [n, 3, out, out]
};

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: one thread per (canvas row, output column), three channels
__global__ void __launch_bounds__(256) resize_h_kernel(const PreArgs a) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int img_i = blockIdx.z;
  if (xx >= a.res_w) return;
  uint8_t* out = a.tmp + ((static_cast<size_t>(img_i) * a.S_h + y) * a.res_w + xx) * 3;
  const int yi = y - a.top;
  if (a.S_w == a.res_w) {  // no horizontal resampling (Pillow skips the pass): copy through the canvas
    const int xi = xx - a.left;
    const bool in = yi >= 0 && yi < a.H && xi >= 0 && xi < a.W;
    const uint8_t* p = a.img + ((static_cast<size_t>(img_i) * a.H + yi) * a.W + xi) * 3;
    for (int c = 0; c < 3; ++c) out[c] = in ? p[c] : a.bg[c];
    return;
  }
  const int x0 = a.bx[2 * xx], nt = a.bx[2 * xx + 1];
  const int* k = a.kx + static_cast<size_t>(xx) * a.ksize_x;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const bool row_in = yi >= 0 && yi < a.H;
  const uint8_t* row = a.img + (static_cast<size_t>(img_i) * a.H + (row_in ? yi : 0)) * a.W * 3;
  for (int t = 0; t < nt; ++t) {
    const int xi = x0 + t - a.left;
    const int kv = k[t];
    if (row_in && xi >= 0 && xi < a.W) {
      const uint8_t* p = row + static_cast<size_t>(xi) * 3;
      s0 += p[0] * kv;
      s1 += p[1] * kv;
      s2 += p[2] * kv;
    } else {
      s0 += a.bg[0] * kv;
      s1 += a.bg[1] * kv;
      s2 += a.bg[2] * kv;
    }
  }
  out[0] = static_cast<uint8_t>(clip8(s0));
  out[1] = static_cast<uint8_t>(clip8(s1));
  out[2] = static_cast<uint8_t>(clip8(s2));
}

// vertical pass over the crop window + rescale + normalise + bf16, CHW output
__global__ void __launch_bounds__(256) resize_v_norm_kernel(const PreArgs a) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x;
  const int yo = blockIdx.y;
  const int img_i = blockIdx.z;
  if (xo >= a.out) return;
  const int yy = yo + a.crop_top, xx = xo + a.crop_left;
  const uint8_t* col = a.tmp + (static_cast<size_t>(img_i) * a.S_h * a.res_w + xx) * 3;
  int v[3];
  if (a.S_h == a.res_h) {
    const uint8_t* p = col + static_cast<size_t>(yy) * a.res_w * 3;
    v[0] = p[0];
    v[1] = p[1];
    v[2] = p[2];
  } else {
    const int y0 = a.by[2 * yy], nt = a.by[2 * yy + 1];
    const int* k = a.ky + static_cast<size_t>(yy) * a.ksize_y;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < nt; ++t) {
      const uint8_t* p = col + static_cast<size_t>(y0 + t) * a.res_w * 3;
      const int kv = k[t];
      s0 += p[0] * kv;
      s1 += p[1] * kv;
      s2 += p[2] * kv;
    }
    v[0] = clip8(s0);
    v[1] = clip8(s1);
    v[2] = clip8(s2);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = static_cast<float>(static_cast<double>(v[c]) * (1.0 / 255.0));  // numpy: uint8 * float -> f64 -> f32
    const float y = __fdiv_rn(__fsub_rn(x, a.mean[c]), a.stdv[c]);                  // float32 (x - mean) / std
    a.dst[((static_cast<size_t>(img_i) * 3 + c) * a.out + yo) * a.out + xo] = __float2bfloat16_rn(y);
  }
}

size_t preprocess_workspace_bytes(int n, int canvas_h, int res_w) {
  return static_cast<size_t>(n) * canvas_h * res_w * 3;
}

int preprocess_images(const uint8_t* img, int n, int H, int W, int pad, const uint8_t* bg, const int* bx, const int* kx,
                      int ksize_x, const int* by, const int* ky, int ksize_y, int res_h, int res_w, int out,
                      const float* mean, const float* stdv, bf16* dst, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
  if (n <= 0) return 0;
  if (H <= 0 || W <= 0 || out <= 0 || res_h < out || res_w < out)
    return fail(-2, "preprocess: bad geometry (H %d W %d resized %d x %d crop %d)", H, W, res_h, res_w, out);
  PreArgs a;
  a.img = img;
  a.n = n;
  a.H = H;
  a.W = W;
  const int S = H > W ? H : W;
  a.S_h = pad ? S : H;
  a.S_w = pad ? S : W;
  a.top = pad && W > H ? (W - H) / 2 : 0;   // expand2square pastes at ((w - h) // 2) rows / ((h - w) // 2) columns
  a.left = pad && H > W ? (H - W) / 2 : 0;
  for (int c = 0; c < 3; ++c) {
    a.bg[c] = bg ? bg[c] : 0;
    a.mean[c] = mean[c];
    a.stdv[c] = stdv[c];
  }
  a.bx = bx;
  a.kx = kx;
  a.ksize_x = ksize_x;
  a.by = by;
  a.ky = ky;
  a.ksize_y = ksize_y;
  a.res_h = res_h;
  a.res_w = res_w;
  a.crop_top = (res_h - out) / 2;
  a.crop_left = (res_w - out) / 2;
  a.out = out;
  if ((a.S_w != res_w && (bx == nullptr || kx == nullptr)) || (a.S_h != res_h && (by == nullptr || ky == nullptr)))
    return fail(-2, "preprocess: resampling coefficients missing");
  if (workspace == nullptr || workspace_bytes < preprocess_workspace_bytes(n, a.S_h, res_w))
    return fail(-2, "preprocess: workspace too small");
  a.tmp = static_cast<uint8_t*>(workspace);
  a.dst = dst;
  LaunchScope scope(kFamPatchify, stream, 3.0 * n * H * W + 2.0 * n * 3 * out * out, 0.0, 2);
  resize_h_kernel<<<dim3((res_w + 255) / 256, a.S_h, n), 256, 0, stream>>>(a);
  resize_v_norm_kernel<<<dim3((out + 255) / 256, out, n), 256, 0, stream>>>(a);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200
