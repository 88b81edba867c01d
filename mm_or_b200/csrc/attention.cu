// Attention entry points of the MM2SG hot path.
//
//  flash_attn            softmax(Q K^T * scale + mask) V for the three "many queries" sites (CLIP ViT-L, BERT image pooler,
//      Llama prefill / training): forwards to the tcgen05 kernel in attention_sm100.cu.
//
//  decode_attn_kernel    one new token per sequence against the bf16 KV cache (llava_arch.py:192-201 +
//      HF LlamaAttention with past_key_values). One query row per (sequence, head): a pure HBM-bound
//      stream over K and V with 16-byte vector loads; no tensor cores needed.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace b200 {

int flash_attn_tc(const AttnArgs& a, int head_dim, cudaStream_t stream);  // attention_sm100.cu

// prefill / ViT / pooler attention: the tcgen05 kernel of attention_sm100.cu (this wrapper only does the accounting)
int flash_attn(const AttnArgs& a, int head_dim, cudaStream_t stream) {
  if (a.B <= 0 || a.H <= 0 || a.Lq <= 0) return 0;
  const double qk = static_cast<double>(a.B) * a.H * a.Lq * a.Lk * (a.causal ? 0.5 : 1.0);
  LaunchScope scope(kFamFlashAttn, stream, 2.0 * a.B * a.H * head_dim * (2.0 * a.Lq + 2.0 * a.Lk), 4.0 * qk * head_dim);
  return flash_attn_tc(a, head_dim, stream);
}

// ---------------------------------------------------------------------------------------------------
// decode attention: one query token per (sequence, head), head_dim 128, KV cache [B][H][ctx_cap][128]
// ---------------------------------------------------------------------------------------------------
static constexpr int kDecThreads = 256;
static constexpr int kDecWarps = kDecThreads / 32;
static constexpr int kDecMaxCtx = 8192;  // per split chunk, scores staged in smem

// Register-bounded for 4 resident CTAs per SM (<= 64 registers): 6.25-6.5 TB/s at the benchmark's geometry; bounding it
// for 6 / 8 CTAs (40 / 32 registers, small spills) measured 5.6-5.9 TB/s (profiles/r2_decode_attn_occupancy.json).
// ROPE: fused RoPE + KV append of the decode step (DecodeArgs::rope_k, common.h), splits == 1.
template <bool ROPE>
__global__ void __launch_bounds__(kDecThreads, 4) decode_attn_kernel(const DecodeArgs a) {
  extern __shared__ float dec_smem[];
  float* sc = dec_smem;                       // [chunk] scores / probabilities
  __shared__ float red[kDecWarps];
  __shared__ float red_o[kDecWarps][128];
  __shared__ float q_s[128];
  __shared__ float kn_s[ROPE ? 128 : 1];      // the step's new key (rotated, bf16-rounded) and value
  __shared__ float vn_s[ROPE ? 128 : 1];

  griddep_wait();  // q / the new cache slot (rope_kv_kernel) and *ctx_dev (previous step) must be complete
  const int bh = blockIdx.x;
  const int b = bh / a.H, h = bh % a.H;
  const int split = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int start = a.kv_start ? a.kv_start[b] : 0;
  const int ctx = a.ctx_dev ? min(*a.ctx_dev + a.ctx_add, a.cap) : a.ctx;
  const bool skip = a.finished != nullptr && a.finished[b] != 0;  // block-uniform: the row stopped at EOS earlier
  const int total = skip ? 0 : max(0, ctx - start);
  const int per = (total + a.splits - 1) / a.splits;
  const int k0 = start + split * per;
  const int k1 = min(ctx, k0 + per);
  const int n = max(0, k1 - k0);

  if (!ROPE) {
    if (threadIdx.x < 128) q_s[threadIdx.x] = __bfloat162float(a.q[b * a.q_rs + h * 128 + threadIdx.x]);
  } else {
    // the token of this step lives at slot ctx - 1; rotary position = slot - kv_start (clamped like rope_kv_kernel)
    const int slot = ctx - 1;
    int p = slot - start;
    p = p < 0 ? 0 : (p >= a.max_pos ? a.max_pos - 1 : p);
    const long long cache_row = ((static_cast<long long>(b) * a.H + h) * a.cap + slot) * 128;
    const int t = threadIdx.x;
    if (t < 128) {
      // threads 0..63 rotate q, 64..127 the new key: pair (i, i + 64), same arithmetic as rope_kv_kernel, results
      // rounded to bf16 like the values that kernel stores
      const int i = t & 63;
      const bf16* src = (t < 64 ? a.q : a.rope_k) + b * a.q_rs + h * 128;
      const float x = __bfloat162float(src[i]), y = __bfloat162float(src[i + 64]);
      const float cs = a.cos_t[static_cast<long long>(p) * 64 + i], sn = a.sin_t[static_cast<long long>(p) * 64 + i];
      const bf16 lo = __float2bfloat16(x * cs - y * sn), hi = __float2bfloat16(y * cs + x * sn);
      if (t < 64) {
        q_s[i] = __bfloat162float(lo);
        q_s[i + 64] = __bfloat162float(hi);
      } else {
        kn_s[i] = __bfloat162float(lo);
        kn_s[i + 64] = __bfloat162float(hi);
        if (slot >= 0) {
          a.kc_w[cache_row + i] = lo;
          a.kc_w[cache_row + i + 64] = hi;
        }
      }
    } else {
      const int i = t - 128;
      const bf16 v = a.rope_v[b * a.q_rs + h * 128 + i];
      vn_s[i] = __bfloat162float(v);
      if (slot >= 0) a.vc_w[cache_row + i] = v;
    }
  }
  __syncthreads();
  // with the fusion the last key of the context (the one appended above) is scored from shared memory
  const int ng = (ROPE && n > 0) ? n - 1 : n;

  const bf16* kbase = a.kc + (static_cast<long long>(b) * a.H + h) * a.cap * 128;
  const bf16* vbase = a.vc + (static_cast<long long>(b) * a.H + h) * a.cap * 128;

  // ---- phase 1: scores. 8 lanes per key (16 dims = 2 x 16 B each), 4 keys per warp-iteration
  const int sub = lane & 7, kq = lane >> 3;
  float qreg[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) qreg[j] = q_s[sub * 16 + j];
  float lmax = -INFINITY;
  for (int base = warp * 4; base < ng; base += kDecWarps * 4) {  // warp-uniform trip count (shuffles inside)
    const int i = base + kq;
    const bool ok = i < ng;
    float acc = 0.f;
    if (ok) {
      const uint4* kp = reinterpret_cast<const uint4*>(kbase + static_cast<long long>(k0 + i) * 128 + sub * 16);
      const uint4 u0 = __ldg(kp), u1 = __ldg(kp + 1);
      acc = qreg[0] * bf16lo(u0.x) + qreg[1] * bf16hi(u0.x) + qreg[2] * bf16lo(u0.y) + qreg[3] * bf16hi(u0.y) +
            qreg[4] * bf16lo(u0.z) + qreg[5] * bf16hi(u0.z) + qreg[6] * bf16lo(u0.w) + qreg[7] * bf16hi(u0.w) +
            qreg[8] * bf16lo(u1.x) + qreg[9] * bf16hi(u1.x) + qreg[10] * bf16lo(u1.y) + qreg[11] * bf16hi(u1.y) +
            qreg[12] * bf16lo(u1.z) + qreg[13] * bf16hi(u1.z) + qreg[14] * bf16lo(u1.w) + qreg[15] * bf16hi(u1.w);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (ok) {
      acc *= a.scale_log2;
      if (sub == 0) sc[i] = acc;
      lmax = fmaxf(lmax, acc);
    }
  }
  if (ROPE && n > 0 && warp == 0) {  // the new key: q . k from shared memory (4 dims per lane)
    float acc = q_s[lane * 4] * kn_s[lane * 4] + q_s[lane * 4 + 1] * kn_s[lane * 4 + 1] +
                q_s[lane * 4 + 2] * kn_s[lane * 4 + 2] + q_s[lane * 4 + 3] * kn_s[lane * 4 + 3];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    acc *= a.scale_log2;
    if (lane == 0) sc[n - 1] = acc;
    lmax = fmaxf(lmax, acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float gmax = red[0];
#pragma unroll
  for (int w = 1; w < kDecWarps; ++w) gmax = fmaxf(gmax, red[w]);
  const float msafe = (gmax == -INFINITY) ? 0.f : gmax;
  __syncthreads();

  // ---- phase 2: probabilities + sum
  float lsum = 0.f;
  for (int i = threadIdx.x; i < n; i += kDecThreads) {
    const float p = exp2f(sc[i] - msafe);
    sc[i] = p;
    lsum += p;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  float gsum = 0.f;
#pragma unroll
  for (int w = 0; w < kDecWarps; ++w) gsum += red[w];

  // ---- phase 3: o = sum_t p[t] V[t]. 16 lanes per key (8 dims = 16 B each), 2 keys per warp-iteration
  const int vsub = lane & 15, vk = lane >> 4;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int i = warp * 2 + vk; i < ng; i += kDecWarps * 2) {
    const float p = sc[i];
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(vbase + static_cast<long long>(k0 + i) * 128 + vsub * 8));
    acc[0] += p * bf16lo(u.x);
    acc[1] += p * bf16hi(u.x);
    acc[2] += p * bf16lo(u.y);
    acc[3] += p * bf16hi(u.y);
    acc[4] += p * bf16lo(u.z);
    acc[5] += p * bf16hi(u.z);
    acc[6] += p * bf16lo(u.w);
    acc[7] += p * bf16hi(u.w);
  }
  // late programmatic-launch trigger: this CTA has streamed its K and V (4096 CTAs run in ~7 waves; successors parked
  // in griddepcontrol.wait must not take CTA slots from the later waves, see gemm_skinny.cu)
  griddep_launch();
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
  if (lane < 16) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red_o[warp][lane * 8 + j] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) v += red_o[w][threadIdx.x];
    if (ROPE && n > 0) v += sc[n - 1] * vn_s[threadIdx.x];  // the appended token's own value
    if (a.splits == 1) {
      const float inv = gsum > 0.f ? 1.f / gsum : 0.f;
      a.o[b * a.o_rs + h * 128 + threadIdx.x] = __float2bfloat16(v * inv);
    } else {
      const long long pidx = static_cast<long long>(bh) * a.splits + split;
      a.part_o[pidx * 128 + threadIdx.x] = v;
      if (threadIdx.x == 0) {
        a.part_ml[pidx * 2] = gmax;
        a.part_ml[pidx * 2 + 1] = gsum;
      }
    }
  }
}

__global__ void decode_combine_kernel(const DecodeArgs a) {
  griddep_wait();
  griddep_launch();
  const int bh = blockIdx.x;
  const int b = bh / a.H, h = bh % a.H;
  const int d = threadIdx.x;  // 128 threads
  float gmax = -INFINITY;
  for (int s = 0; s < a.splits; ++s) gmax = fmaxf(gmax, a.part_ml[(static_cast<long long>(bh) * a.splits + s) * 2]);
  const float msafe = (gmax == -INFINITY) ? 0.f : gmax;
  float num = 0.f, den = 0.f;
  for (int s = 0; s < a.splits; ++s) {
    const long long pidx = static_cast<long long>(bh) * a.splits + s;
    const float w = exp2f(a.part_ml[pidx * 2] - msafe);  // -inf -> 0
    num += w * a.part_o[pidx * 128 + d];
    den += w * a.part_ml[pidx * 2 + 1];
  }
  a.o[b * a.o_rs + h * 128 + d] = __float2bfloat16(den > 0.f ? num / den : 0.f);
}

size_t decode_attn_workspace_bytes(int B, int H, int splits) {
  return splits > 1 ? static_cast<size_t>(B) * H * splits * (128 + 2) * sizeof(float) : 0;
}

int decode_attn_pick_splits(int B, int H, int ctx) {
  const int ctas = B * H;
  int splits = 1;
  const int target = 2 * num_sms();
  while (ctas * splits < target && ctx / (splits * 2) >= 256 && splits < 16) splits *= 2;
  return splits;
}

int decode_attn(DecodeArgs a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (a.B <= 0) return 0;
  if (a.splits <= 0) a.splits = decode_attn_pick_splits(a.B, a.H, a.ctx);
  const int chunk = (a.ctx + a.splits - 1) / a.splits;
  if (chunk > kDecMaxCtx) return fail(-2, "decode_attn: context chunk %d exceeds %d", chunk, kDecMaxCtx);
  if (a.splits > 1) {
    const size_t need = decode_attn_workspace_bytes(a.B, a.H, a.splits);
    if (workspace == nullptr || workspace_bytes < need)
      return fail(-2, "decode_attn: workspace too small (%zu < %zu)", workspace_bytes, need);
    a.part_o = static_cast<float*>(workspace);
    a.part_ml = a.part_o + static_cast<size_t>(a.B) * a.H * a.splits * 128;
  }
  const int smem = chunk * sizeof(float);
  const bool rope = a.rope_k != nullptr;
  if (rope && (a.splits != 1 || a.rope_v == nullptr || a.kc_w == nullptr || a.vc_w == nullptr || a.cos_t == nullptr ||
               a.sin_t == nullptr || a.max_pos <= 0))
    return fail(-2, "decode_attn: the fused RoPE + KV append needs splits == 1 and all of rope_v / kc_w / vc_w / tables");
  static bool cfg = false;
  if (!cfg) {
    const int bytes = kDecMaxCtx * (int)sizeof(float);
    B200_CUDA_OK(cudaFuncSetAttribute(decode_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    B200_CUDA_OK(cudaFuncSetAttribute(decode_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    cfg = true;
  }
  LaunchScope scope(kFamDecodeAttn, stream, 0.0, 0.0, a.splits > 1 ? 2 : 1);  // bytes depend on the device-side ctx
  const dim3 grid(a.B * a.H, a.splits), block(kDecThreads);
  if (rope)
    B200_CUDA_OK(launch_ex(decode_attn_kernel<true>, grid, block, smem, stream, 0, true, a));
  else
    B200_CUDA_OK(launch_ex(decode_attn_kernel<false>, grid, block, smem, stream, 0, true, a));
  if (a.splits > 1)
    B200_CUDA_OK(launch_ex(decode_combine_kernel, dim3(a.B * a.H), dim3(128), 0, stream, 0, true, a));
  return 0;
}

}  // namespace b200
