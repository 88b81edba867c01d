// train_extras.cu -- fine-tune kernels of the small modality encoders and the token embedding (SURVEY.md 8a rows a5 /
// a12: `image_pooler.segmasks_encoder.*` is in the reference's trainable set, train.py:1257-1261; `embed_tokens` under
// full fine-tuning):
//   * SegmentationMapFeatureExtractor (model/multimodal_projector/segmentation_map_feature_extractor.py:53-75:
//     Embedding(30, 8) -> 5 x [Conv2d 3x3, stride 2, pad 1 + ReLU], 32x32 -> 1x1x1024) forward with saved activations,
//     and its backward (what torch autograd does for the reference): weight / bias / embedding gradients in fp32;
//   * gradient of `embed_tokens` (nn.Embedding backward = sum of the output-gradient rows of every occurrence of a
//     token), deterministic: the host sorts the text rows by token id, one block sums the rows of one token in order.
// Tiny FLOP counts (40 MFLOP per map forward): plain SIMT, fixed summation order (bit-deterministic gradients).
// Like ptv3.cu the file also compiles with g++ -DB200_EMU against tests/emu/cuda_emu.h, so the CPU test-suite runs these
// kernels against torch autograd over the oracle.
#ifdef B200_EMU
#include "cuda_emu.h"
#include "emu_common.h"
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#else
#include "../../include/b200_mmor.h"
#include "common.h"
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace b200 {
namespace tx {

__device__ __forceinline__ float bf16_to_f32(uint16_t u) {
  uint32_t w = (uint32_t)u << 16;
  float f;
  memcpy(&f, &w, 4);
  return f;
}
__device__ __forceinline__ uint16_t f32_to_bf16(float x) {  // round to nearest even (finite inputs)
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

constexpr int kSegChan[6] = {8, 64, 128, 256, 512, 1024};
constexpr int kSegSide[6] = {32, 16, 8, 4, 2, 1};
static inline long long seg_act_offset(int l, int n_maps) {  // floats before activation l (host only)
  long long o = 0;
  for (int i = 0; i < l; ++i) o += (long long)n_maps * kSegChan[i] * kSegSide[i] * kSegSide[i];
  return o;
}

// x0[n][e][pix] = emb[min(cls, 29)][e]
__global__ void seg_embed_kernel(const uint8_t* __restrict__ cls, const uint16_t* __restrict__ emb,
                                 float* __restrict__ out, int n_maps) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_maps * 8 * 1024) return;
  int pix = (int)(i % 1024), e = (int)((i / 1024) % 8), n = (int)(i / 8192);
  int c = cls[(size_t)n * 1024 + pix];
  c = c < 30 ? c : 29;
  out[i] = bf16_to_f32(emb[c * 8 + e]);
}

// y[n][co][oy][ox] = relu(b[co] + sum_{ci,ky,kx} w[co][ci][ky][kx] * x[n][ci][2oy-1+ky][2ox-1+kx]); thread per output
__global__ void seg_conv_fwd_kernel(const float* __restrict__ x, const uint16_t* __restrict__ w,
                                    const uint16_t* __restrict__ bias, float* __restrict__ y, int n_maps, int Cin,
                                    int Cout, int Hin) {
  int Hout = Hin / 2;
  long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (long long)n_maps * Cout * Hout * Hout) return;
  int ox = (int)(o % Hout), oy = (int)((o / Hout) % Hout), co = (int)((o / (Hout * Hout)) % Cout);
  int n = (int)(o / ((long long)Hout * Hout * Cout));
  const float* xn = x + (size_t)n * Cin * Hin * Hin;
  const uint16_t* wr = w + (size_t)co * Cin * 9;
  float acc = 0.f;
  for (int ci = 0; ci < Cin; ++ci)
    for (int k = 0; k < 9; ++k) {
      int iy = oy * 2 - 1 + k / 3, ix = ox * 2 - 1 + k % 3;
      if (iy >= 0 && iy < Hin && ix >= 0 && ix < Hin)
        acc = fmaf(xn[((size_t)ci * Hin + iy) * Hin + ix], bf16_to_f32(wr[ci * 9 + k]), acc);
    }
  y[o] = fmaxf(acc + bf16_to_f32(bias[co]), 0.f);
}

// token rows: out[row_map[n] or n][0:1024] = bf16(y5[n][0:1024])
__global__ void seg_store_tokens_kernel(const float* __restrict__ y5, int n_maps, uint16_t* __restrict__ out,
                                        long long out_ld, const int* __restrict__ row_map) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_maps * 1024) return;
  int n = i / 1024, c = i % 1024;
  int r = row_map ? row_map[n] : n;
  if (r >= 0) out[(size_t)r * out_ld + c] = f32_to_bf16(y5[i]);
}

// dz5[n][c] = d_out[row][c] * (y5 > 0)
__global__ void seg_dtok_kernel(const uint16_t* __restrict__ d_out, long long d_ld, const int* __restrict__ row_map,
                                const float* __restrict__ y5, int n_maps, float* __restrict__ dz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_maps * 1024) return;
  int n = i / 1024, c = i % 1024;
  int r = row_map ? row_map[n] : n;
  float g = r >= 0 ? bf16_to_f32(d_out[(size_t)r * d_ld + c]) : 0.f;
  dz[i] = y5[i] > 0.f ? g : 0.f;
}

// dW[co][ci][k] (+)= sum_{n,oy,ox} dz[n][co][oy][ox] * x[n][ci][2oy-1+ky][2ox-1+kx]; the last Cout threads do db[co]
__global__ void seg_conv_dw_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ dW,
                                   float* __restrict__ db, int n_maps, int Cin, int Cout, int Hin, int accumulate) {
  int Hout = Hin / 2;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nw = (long long)Cout * Cin * 9;
  if (t >= nw + Cout) return;
  float acc = 0.f;
  if (t < nw) {
    int k = (int)(t % 9), ci = (int)((t / 9) % Cin), co = (int)(t / (9LL * Cin));
    for (int n = 0; n < n_maps; ++n)
      for (int oy = 0; oy < Hout; ++oy) {
        int iy = oy * 2 - 1 + k / 3;
        if (iy < 0 || iy >= Hin) continue;
        for (int ox = 0; ox < Hout; ++ox) {
          int ix = ox * 2 - 1 + k % 3;
          if (ix < 0 || ix >= Hin) continue;
          acc = fmaf(dz[(((size_t)n * Cout + co) * Hout + oy) * Hout + ox],
                     x[(((size_t)n * Cin + ci) * Hin + iy) * Hin + ix], acc);
        }
      }
    dW[t] = accumulate ? dW[t] + acc : acc;
  } else {
    int co = (int)(t - nw);
    for (int n = 0; n < n_maps; ++n)
      for (int p = 0; p < Hout * Hout; ++p) acc += dz[((size_t)n * Cout + co) * Hout * Hout + p];
    db[co] = accumulate ? db[co] + acc : acc;
  }
}

// dz_prev[n][ci][iy][ix] = (x > 0 or first layer) * sum_{co,ky,kx : 2oy-1+ky = iy, 2ox-1+kx = ix} dz[n][co][oy][ox] w[co][ci][ky][kx]
__global__ void seg_conv_dx_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                   const uint16_t* __restrict__ w, float* __restrict__ dx, int n_maps, int Cin, int Cout,
                                   int Hin, int relu_mask) {
  int Hout = Hin / 2;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_maps * Cin * Hin * Hin) return;
  int ix = (int)(t % Hin), iy = (int)((t / Hin) % Hin), ci = (int)((t / (Hin * Hin)) % Cin);
  int n = (int)(t / ((long long)Hin * Hin * Cin));
  float acc = 0.f;
  if (!relu_mask || x[t] > 0.f) {
    for (int ky = 0; ky < 3; ++ky) {
      int ny = iy + 1 - ky;
      if (ny < 0 || (ny & 1)) continue;
      int oy = ny >> 1;
      if (oy >= Hout) continue;
      for (int kx = 0; kx < 3; ++kx) {
        int nx = ix + 1 - kx;
        if (nx < 0 || (nx & 1)) continue;
        int ox = nx >> 1;
        if (ox >= Hout) continue;
        for (int co = 0; co < Cout; ++co)
          acc = fmaf(dz[(((size_t)n * Cout + co) * Hout + oy) * Hout + ox],
                     bf16_to_f32(w[((size_t)co * Cin + ci) * 9 + ky * 3 + kx]), acc);
      }
    }
  }
  dx[t] = acc;
}

// dE[c][e] (+)= sum over (n, pix) with min(cls, 29) == c of dx0[n][e][pix]
__global__ void seg_embed_grad_kernel(const uint8_t* __restrict__ cls, const float* __restrict__ dx0, int n_maps,
                                      float* __restrict__ dE, int accumulate) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 30 * 8) return;
  int c = t / 8, e = t % 8;
  float acc = 0.f;
  for (int n = 0; n < n_maps; ++n)
    for (int pix = 0; pix < 1024; ++pix) {
      int k = cls[(size_t)n * 1024 + pix];
      k = k < 30 ? k : 29;
      if (k == c) acc += dx0[((size_t)n * 8 + e) * 1024 + pix];
    }
  dE[t] = accumulate ? dE[t] + acc : acc;
}

// d_table[token[s]][:] (+)= sum over j in [seg_start[s], seg_start[s+1]) of d_rows[row_list[j]][:]   (bf16 rows, fp32 sum)
__global__ void embed_grad_kernel(const uint16_t* __restrict__ d_rows, long long ld, const int* __restrict__ row_list,
                                  const int* __restrict__ seg_start, const int* __restrict__ seg_token, int D,
                                  float* __restrict__ d_table, int accumulate) {
  int s = blockIdx.x;
  int j0 = seg_start[s], j1 = seg_start[s + 1];
  float* dst = d_table + (size_t)seg_token[s] * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int j = j0; j < j1; ++j) acc += bf16_to_f32(d_rows[(size_t)row_list[j] * ld + c]);
    dst[c] = accumulate ? dst[c] + acc : acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Inverted dropout with a counter-based generator (Philox4x32-10): element i is dropped iff word (i mod 4) of
// Philox(counter = i / 4, key = seed) is below floor(p * 2^32); kept elements are multiplied by 1 / (1 - p). The mask
// is a pure function of (seed, i): the backward pass re-creates it from the seed instead of storing it, and the same
// kernel applied to the output gradient is the backward (accumulate = 1 adds into y: dx += mask * g / (1 - p)).
// (The reference's dropout masks come from torch's own Philox stream and are not reproducible outside torch; what is
// kept is the distribution: Bernoulli(1 - p) per element, scaled by 1 / (1 - p). LoRA dropout 0.05, train.py:111.)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  uint32_t c[4] = {c0, c1, 0u, 0u};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = (uint32_t)p1;
    c[2] = n2;
    c[3] = (uint32_t)p0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0];
  out[1] = c[1];
  out[2] = c[2];
  out[3] = c[3];
}

// one thread per group of 4 consecutive elements (one Philox call, 8 bytes in, 8 bytes out)
// x and y may be the same buffer (in-place dropout, train/encoder.py): no __restrict__; every thread reads its own
// elements before it writes them and touches no other thread's
__global__ void dropout_kernel(const uint16_t* x, uint16_t* y, long long n, uint32_t threshold,
                               float scale, uint32_t k0, uint32_t k1, int accumulate) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g * 4 >= n) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)g, (uint32_t)((unsigned long long)g >> 32), k0, k1, r);
  const long long i0 = g * 4;
  if (i0 + 4 <= n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 7) == 0) {
    // whole group, 8-byte aligned buffers: one 8-byte load and store per thread (a warp moves 256 contiguous bytes)
    const uint2 in = *reinterpret_cast<const uint2*>(x + i0);
    uint2 old = make_uint2(0u, 0u);
    if (accumulate) old = *reinterpret_cast<const uint2*>(y + i0);
    const uint32_t xi[2] = {in.x, in.y}, yi[2] = {old.x, old.y};
    uint32_t out[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v0 = r[2 * h] < threshold ? 0.f : bf16_to_f32((uint16_t)(xi[h] & 0xffffu)) * scale;
      float v1 = r[2 * h + 1] < threshold ? 0.f : bf16_to_f32((uint16_t)(xi[h] >> 16)) * scale;
      if (accumulate) {
        v0 += bf16_to_f32((uint16_t)(yi[h] & 0xffffu));
        v1 += bf16_to_f32((uint16_t)(yi[h] >> 16));
      }
      out[h] = (uint32_t)f32_to_bf16(v0) | ((uint32_t)f32_to_bf16(v1) << 16);
    }
    *reinterpret_cast<uint2*>(y + i0) = make_uint2(out[0], out[1]);
    return;
  }
  for (int j = 0; j < 4; ++j) {  // ragged tail or unaligned views
    const long long i = i0 + j;
    if (i >= n) break;
    float v = r[j] < threshold ? 0.f : bf16_to_f32(x[i]) * scale;
    if (accumulate) v += bf16_to_f32(y[i]);
    y[i] = f32_to_bf16(v);
  }
}

}  // namespace tx
}  // namespace b200

using namespace b200;
using namespace b200::tx;

#define TX_CHECK_LAUNCH(name)                                                                  \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) return fail(-5, "%s launch failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

static inline unsigned tx_blocks(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }

extern "C" {

size_t b200_segmask_train_acts_bytes(int n_maps) {
  return n_maps > 0 ? (size_t)seg_act_offset(6, n_maps) * sizeof(float) : 0;
}

int b200_segmask_forward_train(const b200_segmask_weights* w, const uint8_t* cls, int n_maps, void* acts,
                               size_t acts_bytes, void* out, int64_t out_ld, const int32_t* out_row_map,
                               b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!w || !cls || !acts || !out || n_maps <= 0 || out_ld < 1024)
    return fail(-2, "b200_segmask_forward_train: bad argument (n_maps=%d)", n_maps);
  if (acts_bytes < b200_segmask_train_acts_bytes(n_maps))
    return fail(-2, "b200_segmask_forward_train: activation buffer too small");
  float* a = reinterpret_cast<float*>(acts);
  LaunchScope ls(kFamTrain, stream, 0.0, 0.0, 7);
  B200_LAUNCH(seg_embed_kernel, dim3(tx_blocks((long long)n_maps * 8192, 256)), dim3(256), 0, stream, cls,
              reinterpret_cast<const uint16_t*>(w->emb), a, n_maps);
  for (int l = 0; l < 5; ++l) {
    long long total = (long long)n_maps * kSegChan[l + 1] * kSegSide[l + 1] * kSegSide[l + 1];
    B200_LAUNCH(seg_conv_fwd_kernel, dim3(tx_blocks(total, 128)), dim3(128), 0, stream, a + seg_act_offset(l, n_maps),
                reinterpret_cast<const uint16_t*>(w->conv_w[l]), reinterpret_cast<const uint16_t*>(w->conv_b[l]),
                a + seg_act_offset(l + 1, n_maps), n_maps, kSegChan[l], kSegChan[l + 1], kSegSide[l]);
  }
  B200_LAUNCH(seg_store_tokens_kernel, dim3(tx_blocks((long long)n_maps * 1024, 256)), dim3(256), 0, stream,
              a + seg_act_offset(5, n_maps), n_maps, reinterpret_cast<uint16_t*>(out), (long long)out_ld, out_row_map);
  TX_CHECK_LAUNCH("b200_segmask_forward_train");
  return 0;
}

size_t b200_segmask_backward_workspace_bytes(int n_maps) {
  // two gradient buffers, each as large as the largest activation (64 x 16 x 16 per map)
  return n_maps > 0 ? (size_t)n_maps * 2 * 16384 * sizeof(float) : 0;
}

int b200_segmask_backward(const b200_segmask_weights* w, const uint8_t* cls, int n_maps, const void* acts,
                          const void* d_out, int64_t d_ld, const int32_t* row_map, float* const* d_conv_w,
                          float* const* d_conv_b, float* d_emb, int accumulate, void* workspace, size_t workspace_bytes,
                          b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!w || !cls || !acts || !d_out || !d_conv_w || !d_conv_b || !d_emb || !workspace || n_maps <= 0 || d_ld < 1024)
    return fail(-2, "b200_segmask_backward: bad argument (n_maps=%d)", n_maps);
  if (workspace_bytes < b200_segmask_backward_workspace_bytes(n_maps))
    return fail(-2, "b200_segmask_backward: workspace too small");
  const float* a = reinterpret_cast<const float*>(acts);
  float* g0 = reinterpret_cast<float*>(workspace);
  float* g1 = g0 + (size_t)n_maps * 16384;
  LaunchScope ls(kFamTrain, stream, 0.0, 0.0, 12);
  B200_LAUNCH(seg_dtok_kernel, dim3(tx_blocks((long long)n_maps * 1024, 256)), dim3(256), 0, stream,
              reinterpret_cast<const uint16_t*>(d_out), (long long)d_ld, row_map, a + seg_act_offset(5, n_maps), n_maps,
              g0);
  float* dz = g0;
  float* dx = g1;
  for (int l = 4; l >= 0; --l) {
    int Cin = kSegChan[l], Cout = kSegChan[l + 1], Hin = kSegSide[l];
    const float* x = a + seg_act_offset(l, n_maps);
    B200_LAUNCH(seg_conv_dw_kernel, dim3(tx_blocks((long long)Cout * Cin * 9 + Cout, 128)), dim3(128), 0, stream, x, dz,
                d_conv_w[l], d_conv_b[l], n_maps, Cin, Cout, Hin, accumulate);
    // gradient w.r.t. this layer's input, masked by the ReLU that produced it (layer 0's input is the embedding)
    B200_LAUNCH(seg_conv_dx_kernel, dim3(tx_blocks((long long)n_maps * Cin * Hin * Hin, 128)), dim3(128), 0, stream, x,
                dz, reinterpret_cast<const uint16_t*>(w->conv_w[l]), dx, n_maps, Cin, Cout, Hin, l > 0 ? 1 : 0);
    float* t = dz;
    dz = dx;
    dx = t;
  }
  B200_LAUNCH(seg_embed_grad_kernel, dim3(1), dim3(256), 0, stream, cls, dz, n_maps, d_emb, accumulate);
  TX_CHECK_LAUNCH("b200_segmask_backward");
  return 0;
}

int b200_embed_grad(const void* d_rows, int64_t ld, const int32_t* row_list, const int32_t* seg_start,
                    const int32_t* seg_token, int n_seg, int D, float* d_table, int accumulate, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_rows || !row_list || !seg_start || !seg_token || !d_table || n_seg <= 0 || D <= 0 || ld < D)
    return fail(-2, "b200_embed_grad: bad argument (n_seg=%d D=%d)", n_seg, D);
  LaunchScope ls(kFamTrain, stream, 0.0, 0.0, 1);
  B200_LAUNCH(embed_grad_kernel, dim3(n_seg), dim3(256), 0, stream, reinterpret_cast<const uint16_t*>(d_rows),
              (long long)ld, row_list, seg_start, seg_token, D, d_table, accumulate);
  TX_CHECK_LAUNCH("b200_embed_grad");
  return 0;
}

int b200_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, int accumulate, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) return 0;
  if (!x || !y || n < 0 || !(p >= 0.f) || !(p < 1.f))
    return fail(-2, "b200_dropout: bad argument (n=%lld, p=%g must be in [0, 1))", (long long)n, (double)p);
  const uint32_t threshold = (uint32_t)((double)p * 4294967296.0);  // floor(p * 2^32)
  const long long groups = (n + 3) / 4;
  LaunchScope ls(kFamTrain, stream, (accumulate ? 6.0 : 4.0) * (double)n, 0.0, 1);
  B200_LAUNCH(dropout_kernel, dim3(tx_blocks(groups, 256)), dim3(256), 0, stream, reinterpret_cast<const uint16_t*>(x),
              reinterpret_cast<uint16_t*>(y), (long long)n, threshold, 1.0f / (1.0f - p), (uint32_t)seed,
              (uint32_t)(seed >> 32), accumulate);
  TX_CHECK_LAUNCH("b200_dropout");
  return 0;
}

}  // extern "C"
