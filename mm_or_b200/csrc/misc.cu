// Data-movement and small elementwise kernels of the MM2SG hot path (all HBM-bound, vectorised):
//   patchify_im2col   CLIP patch_embedding conv (k = s = 14, no bias) rewritten as an im2col so the contraction runs
//                     on the tcgen05 GEMM (HF CLIPVisionEmbeddings, call site clip_encoder.py:48)
//   rope_kv_write     rotary embedding on q/k (theta 1e4, rotate_half) + KV-cache append
//                     (HF LlamaAttention / apply_rotary_pos_emb, call site llava_llama.py:93; replaces the per-step
//                      torch.cat cache growth of HF 4.31)
//   embed_rows        embed_tokens gather + zero pad rows of the multimodal token pack (llava_arch.py:235-338)
//   argmax_rows       greedy next-token choice (HF greedy_search argmax, lowest index wins ties)
//   segmask_*         Embedding(30,8) + 5 x [conv3x3 s2 p1 + ReLU] (segmentation_map_feature_extractor.py:53-75)
#include "common.h"
#include "ptx.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------
// patchify: pixels [N, C, S, S] bf16 -> cols [N * G * G, Kpad], K index = (c * P + ky) * P + kx, zero padded
// ---------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const bf16* __restrict__ px, bf16* __restrict__ out, int N, int C, int S, int P,
                                int G, int Kpad) {
  const long long total = static_cast<long long>(N) * G * G * C * P;  // one thread per (patch, c, ky) segment
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int ky = i % P;
  const int c = (i / P) % C;
  const long long patch = i / (P * C);
  const int gx = patch % G;
  const int gy = (patch / G) % G;
  const int n = patch / (G * G);
  const bf16* src = px + ((static_cast<long long>(n) * C + c) * S + gy * P + ky) * S + gx * P;
  bf16* dst = out + patch * Kpad + (c * P + ky) * P;
  if ((P & 1) == 0) {
    for (int k = 0; k < P; k += 2)
      *reinterpret_cast<uint32_t*>(dst + k) = *reinterpret_cast<const uint32_t*>(src + k);
  } else {
    for (int k = 0; k < P; ++k) dst[k] = src[k];
  }
  if (c == 0 && ky == 0) {
    for (int k = C * P * P; k < Kpad; ++k) out[patch * Kpad + k] = __float2bfloat16(0.f);
  }
}

int patchify_im2col(const bf16* pixels, bf16* cols, int N, int C, int S, int P, int Kpad, cudaStream_t stream) {
  if (N <= 0) return 0;
  if (S % P != 0 || Kpad < C * P * P) return fail(-2, "patchify: bad geometry S=%d P=%d Kpad=%d", S, P, Kpad);
  const int G = S / P;
  const long long total = static_cast<long long>(N) * G * G * C * P;
  LaunchScope scope(kFamPatchify, stream, 2.0 * N * G * G * (C * P * P + Kpad), 0.0);
  patchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(pixels, cols, N, C, S, P, G, Kpad);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// RoPE + KV-cache write. qkv: [T, 3 * H * 128] (q | k | v), T = B * Lq tokens, token (b, l) at row b * Lq + l.
// q is rotated in place; k (rotated) and v go to cache[b][h][slot][128] with slot = slot0 + l, where slot0 is an
// immediate (prefill) or read from device memory (decode under a CUDA graph). The rotary position is
// slot - kv_start[b] (left padding: positions restart at 0 on the first real token, llava_arch.py:319-327 and
// :200 `position_ids = attention_mask.sum(1) - 1`), clamped to 0 on pad rows.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rope_kv_kernel(bf16* __restrict__ qkv, const int* __restrict__ kv_start,
                                                      const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                      bf16* __restrict__ kc, bf16* __restrict__ vc, int T, int H, int Lq,
                                                      int slot0_imm, const int* __restrict__ slot0_dev, int cap,
                                                      int max_pos) {
  griddep_wait();    // qkv (previous GEMM) and *slot0_dev (previous step) must be complete
  griddep_launch();
  const long long w = blockIdx.x * 4ll + (threadIdx.x >> 5);  // one warp per (token, head)
  if (w >= static_cast<long long>(T) * H) return;
  const int lane = threadIdx.x & 31;
  const int h = w % H;
  const int tok = w / H;
  const int b = tok / Lq, l = tok % Lq;
  const int slot = (slot0_dev ? *slot0_dev : slot0_imm) + l;
  if (slot >= cap) return;
  int p = slot - (kv_start ? kv_start[b] : 0);
  p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
  const float2 cs = *reinterpret_cast<const float2*>(cos_t + static_cast<long long>(p) * 64 + lane * 2);
  const float2 sn = *reinterpret_cast<const float2*>(sin_t + static_cast<long long>(p) * 64 + lane * 2);
  bf16* row = qkv + static_cast<long long>(tok) * 3 * H * 128;
  bf16* cache_row_k = kc + ((static_cast<long long>(b) * H + h) * cap + slot) * 128;
  bf16* cache_row_v = vc + ((static_cast<long long>(b) * H + h) * cap + slot) * 128;
#pragma unroll
  for (int which = 0; which < 2; ++which) {  // 0 = q (in place), 1 = k (to cache)
    bf16* src = row + which * H * 128 + h * 128;
    const uint32_t lo = *reinterpret_cast<const uint32_t*>(src + lane * 2);
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(src + 64 + lane * 2);
    const float x0 = bf16lo(lo), x1 = bf16hi(lo), y0 = bf16lo(hi), y1 = bf16hi(hi);
    const uint32_t olo = pack_bf16x2(x0 * cs.x - y0 * sn.x, x1 * cs.y - y1 * sn.y);
    const uint32_t ohi = pack_bf16x2(y0 * cs.x + x0 * sn.x, y1 * cs.y + x1 * sn.y);
    bf16* dst = which == 0 ? src : cache_row_k;
    *reinterpret_cast<uint32_t*>(dst + lane * 2) = olo;
    *reinterpret_cast<uint32_t*>(dst + 64 + lane * 2) = ohi;
  }
  const bf16* vsrc = row + 2 * H * 128 + h * 128;
  *reinterpret_cast<uint2*>(cache_row_v + lane * 4) = *reinterpret_cast<const uint2*>(vsrc + lane * 4);
}

int rope_kv_write(bf16* qkv, const int* kv_start, const float* cos_t, const float* sin_t, int max_pos, bf16* kc,
                  bf16* vc, int B, int H, int Lq, int slot0, const int* slot0_dev, int cap, cudaStream_t stream) {
  const long long warps = static_cast<long long>(B) * Lq * H;
  if (warps <= 0) return 0;
  if (slot0_dev == nullptr && slot0 + Lq > cap)
    return fail(-2, "rope_kv_write: slot %d + %d exceeds cache capacity %d", slot0, Lq, cap);
  LaunchScope scope(kFamRope, stream, static_cast<double>(warps) * 128 * 2 * 5, 0.0);
  B200_CUDA_OK(launch_ex(rope_kv_kernel, dim3(static_cast<unsigned>((warps + 3) / 4)), dim3(128), 0, stream, 0, true, qkv,
                         kv_start, cos_t, sin_t, kc, vc, B * Lq, H, Lq, slot0, slot0_dev, cap, max_pos));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// embed_rows: out[r] = table[ids[r]] if ids[r] >= 0; zeros if ids[r] == -1; untouched if ids[r] <= -2
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) embed_rows_kernel(const int* __restrict__ ids, const bf16* __restrict__ table,
                                                         bf16* __restrict__ out, long long ldo, int rows, int D,
                                                         int vocab) {
  griddep_wait();    // ids are the previous step's argmax output
  griddep_launch();
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int id = ids[r];
  if (id <= -2) return;
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<long long>(r) * ldo);
  const int nvec = D / 8;
  if (id < 0 || id >= vocab) {
    for (int i = lane; i < nvec; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
  } else {
    const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<long long>(id) * D);
    for (int i = lane; i < nvec; i += 32) dst[i] = __ldg(src + i);
  }
}

int embed_rows(const int* ids, const bf16* table, bf16* out, long long ldo, int rows, int D, int vocab,
               cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (D % 8) return fail(-2, "embed_rows: D must be a multiple of 8");
  LaunchScope scope(kFamEmbed, stream, 4.0 * rows * D, 0.0);
  B200_CUDA_OK(launch_ex(embed_rows_kernel, dim3((rows + 3) / 4), dim3(128), 0, stream, 0, true, ids, table, out, ldo,
                         rows, D, vocab));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// argmax over the vocabulary, lowest index wins ties (torch.argmax / HF greedy semantics on equal logits).
// Optional `finished` handling mirrors HF greedy_search: rows already finished emit pad_id; a row finishes
// when it emits eos_id.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256) argmax_kernel(const T* __restrict__ logits, long long ld, int V,
                                                     int* __restrict__ out_tok, int* __restrict__ finished, int eos_id,
                                                     int pad_id, int* __restrict__ history, int hist_ld, int step_imm,
                                                     const int* __restrict__ step_dev) {
  griddep_wait();
  griddep_launch();
  const int r = blockIdx.x;
  const T* row = logits + static_cast<long long>(r) * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += 256) {
    const float v = to_f<T>(row[i]);
    if (v > best || (v == best && i < bi)) {
      best = v;
      bi = i;
    }
  }
  __shared__ float sv[256];
  __shared__ int si[256];
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float ov = sv[threadIdx.x + s];
      const int oi = si[threadIdx.x + s];
      if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
        sv[threadIdx.x] = ov;
        si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int tok = si[0] == 0x7fffffff ? 0 : si[0];
    if (finished != nullptr) {
      if (finished[r]) tok = pad_id;
      else if (tok == eos_id) finished[r] = 1;
    }
    out_tok[r] = tok;
    const int step = step_dev ? *step_dev : step_imm;
    if (history != nullptr && step < hist_ld) history[static_cast<long long>(r) * hist_ld + step] = tok;
  }
}

int argmax_rows(const void* logits, int is_fp32, long long ld, int rows, int V, int* out_tok, int* finished, int eos_id,
                int pad_id, int* history, int hist_ld, int step, const int* step_dev, cudaStream_t stream) {
  if (rows <= 0) return 0;
  LaunchScope scope(kFamArgmax, stream, static_cast<double>(rows) * V * (is_fp32 ? 4 : 2), 0.0);
  if (is_fp32)
    B200_CUDA_OK(launch_ex(argmax_kernel<float>, dim3(rows), dim3(256), 0, stream, 0, true,
                           static_cast<const float*>(logits), ld, V, out_tok, finished, eos_id, pad_id, history, hist_ld,
                           step, step_dev));
  else
    B200_CUDA_OK(launch_ex(argmax_kernel<bf16>, dim3(rows), dim3(256), 0, stream, 0, true,
                           static_cast<const bf16*>(logits), ld, V, out_tok, finished, eos_id, pad_id, history, hist_ld,
                           step, step_dev));
  return 0;
}

// decode-loop device counters: state[0] = KV slots in use, state[1] = decode step index
__global__ void bump_counters_kernel(int* state) {
  griddep_wait();  // the step's kernels read state[]: all of them are complete once the argmax before this one is
  state[0] += 1;
  state[1] += 1;
}
int bump_counters(int* state, cudaStream_t stream) {
  LaunchScope scope(kFamMisc, stream);
  B200_CUDA_OK(launch_ex(bump_counters_kernel, dim3(1), dim3(1), 0, stream, 0, true, state));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// seg-mask feature extractor
// ---------------------------------------------------------------------------------------------------
__global__ void segmask_embed_kernel(const uint8_t* __restrict__ cls, const bf16* __restrict__ emb,
                                     float* __restrict__ out, int n_maps, int E, int HW, int num_classes) {
  // out[n][e][pix] = emb[cls[n][pix]][e]
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n_maps) * E * HW) return;
  const int pix = i % HW;
  const int e = (i / HW) % E;
  const int n = i / (static_cast<long long>(HW) * E);
  int c = cls[static_cast<long long>(n) * HW + pix];
  c = c < num_classes ? c : num_classes - 1;
  out[i] = __bfloat162float(emb[c * E + e]);
}

// conv 3x3, stride 2, pad 1, + bias + ReLU.  in [n][Cin][Hin][Hin] fp32, w [Cout][Cin][3][3] bf16, out fp32 (or bf16
// rows for the last layer). One warp per output element; lanes split the Cin * 9 reduction (coalesced weight reads).
__global__ void __launch_bounds__(128) conv3x3s2_relu_kernel(const float* __restrict__ in, const bf16* __restrict__ w,
                                                             const bf16* __restrict__ bias, float* __restrict__ out,
                                                             bf16* __restrict__ out_bf16, long long out_bf16_ld,
                                                             const int* __restrict__ out_row_map, int n_maps, int Cin,
                                                             int Cout, int Hin) {
  const int Hout = Hin / 2;
  const long long o = blockIdx.x * 4ll + (threadIdx.x >> 5);
  const long long total = static_cast<long long>(n_maps) * Cout * Hout * Hout;
  if (o >= total) return;
  const int lane = threadIdx.x & 31;
  const int ox = o % Hout;
  const int oy = (o / Hout) % Hout;
  const int co = (o / (Hout * Hout)) % Cout;
  const int n = o / (static_cast<long long>(Hout) * Hout * Cout);
  const bf16* wr = w + static_cast<long long>(co) * Cin * 9;
  const float* inn = in + static_cast<long long>(n) * Cin * Hin * Hin;
  float acc = 0.f;
  for (int r = lane; r < Cin * 9; r += 32) {
    const int ci = r / 9, k = r % 9;
    const int iy = oy * 2 - 1 + k / 3, ix = ox * 2 - 1 + k % 3;
    if (iy >= 0 && iy < Hin && ix >= 0 && ix < Hin)
      acc += inn[(static_cast<long long>(ci) * Hin + iy) * Hin + ix] * __bfloat162float(wr[r]);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    const float y = fmaxf(acc + __bfloat162float(bias[co]), 0.f);
    if (out_bf16 != nullptr) {  // final layer: Hout == 1, write token rows
      const int orow = out_row_map ? out_row_map[n] : n;
      if (orow >= 0) out_bf16[static_cast<long long>(orow) * out_bf16_ld + co] = __float2bfloat16(y);
    } else {
      out[o] = y;
    }
  }
}

size_t segmask_workspace_bytes(int n_maps) {
  // embedded input 8x32x32 + ping-pong activations (largest: 64x16x16)
  return static_cast<size_t>(n_maps) * (8 * 1024 + 64 * 256 + 128 * 64) * sizeof(float);
}

// weights: emb [30, 8]; conv i weight [Cout, Cin, 3, 3], bias [Cout] for the channel chain 8-64-128-256-512-1024.
// cls: [n_maps, 32, 32] uint8. Writes token rows out[out_row_map[n] or n][0..1024) with leading dimension out_ld.
int segmask_forward(const uint8_t* cls, int n_maps, const bf16* emb, const bf16* const* conv_w,
                    const bf16* const* conv_b, bf16* out, long long out_ld, const int* out_row_map, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream) {
  if (n_maps <= 0) return 0;
  if (workspace_bytes < segmask_workspace_bytes(n_maps)) return fail(-2, "segmask_forward: workspace too small");
  float* x0 = static_cast<float*>(workspace);
  float* x1 = x0 + static_cast<size_t>(n_maps) * 8 * 1024;
  float* x2 = x1 + static_cast<size_t>(n_maps) * 64 * 256;
  const long long tot = static_cast<long long>(n_maps) * 8 * 1024;
  LaunchScope scope(kFamSegmask, stream, 0.0, 0.0, 6);
  segmask_embed_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, stream>>>(cls, emb, x0, n_maps, 8, 1024, 30);
  B200_CUDA_OK(cudaGetLastError());
  const int chans[6] = {8, 64, 128, 256, 512, 1024};
  float* bufs[2] = {x1, x2};
  const float* cur = x0;
  int H = 32;
  for (int l = 0; l < 5; ++l) {
    const int Cin = chans[l], Cout = chans[l + 1];
    const long long total = static_cast<long long>(n_maps) * Cout * (H / 2) * (H / 2);
    float* dst = bufs[l & 1];
    const bool last = l == 4;
    conv3x3s2_relu_kernel<<<static_cast<unsigned>((total + 3) / 4), 128, 0, stream>>>(
        cur, conv_w[l], conv_b[l], dst, last ? out : nullptr, out_ld, last ? out_row_map : nullptr, n_maps, Cin, Cout, H);
    B200_CUDA_OK(cudaGetLastError());
    cur = dst;
    H /= 2;
  }
  return 0;
}

}  // namespace b200
