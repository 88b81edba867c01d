// LayerNorm / RMSNorm kernels (HBM-bound; one warp per row, 16-byte vector loads, fp32 statistics).
//
//  layernorm: y = LN(gather(x)[row] + add_table[row % period]) * gamma + beta
//     - CLIP pre_layrnorm over [CLS ; patches] + position_embedding, LayerNorm1/2 (eps 1e-5)  (HF CLIPVisionTransformer,
//       call site clip_encoder.py:48)
//     - BERT pooler embeddings LN(x + token_type[0] + position[s]) and the post-LN after each sub-block (eps 1e-12)
//       (HF BertEmbeddings/BertSelfOutput/BertOutput, call site multimodal_projector/builder.py:173); the optional
//       row gather builds the zero-padded (B, Vmax*576, 1024) pooler input of llava_arch.py:143-170 on the fly.
//  rmsnorm: y = w * x * rsqrt(mean(x^2) + eps)    (HF LlamaRMSNorm, fp32 variance; call site llava_llama.py:93)
#include "common.h"
#include "ptx.cuh"

namespace b200 {

struct LnArgs {
  const bf16* x;
  long long ldx;
  const int* row_map;  // optional: source row per output row, < 0 = all-zero source row
  const bf16* add;     // optional [period, D]
  int period;
  const bf16* gamma;
  const bf16* beta;  // null for RMSNorm
  float eps;
  bf16* out;
  long long ldo;
  int M, D;
  int out_group;         // 0: output row = row; else output row = (row / out_group) * out_group_stride + row % out_group
  int out_group_stride;  // (writes e.g. the first 576 pooled tokens of every sample into a (B, T_vis, D) buffer)
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x);
  f[1] = bf16hi(u.x);
  f[2] = bf16lo(u.y);
  f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z);
  f[5] = bf16hi(u.z);
  f[6] = bf16lo(u.w);
  f[7] = bf16hi(u.w);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// kVec = number of 16-byte vectors each lane holds (D = kVec * 256); values stay in registers across passes.
template <int kVec, bool kRms>
__global__ void __launch_bounds__(128) norm_kernel(const LnArgs a) {
  griddep_wait();
  griddep_launch();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= a.M) return;
  const int lane = threadIdx.x & 31;
  long long src = row;
  bool zero_src = false;
  if (a.row_map != nullptr) {
    const int r = a.row_map[row];
    zero_src = r < 0;
    src = zero_src ? 0 : r;
  }
  const uint4* xp = reinterpret_cast<const uint4*>(a.x + src * a.ldx);
  const uint4* ap = a.add ? reinterpret_cast<const uint4*>(a.add + static_cast<long long>(row % a.period) * a.D) : nullptr;
  float v[kVec][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int idx = i * 32 + lane;
    if (zero_src) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    } else {
      unpack8(xp[idx], v[i]);
    }
    if (ap != nullptr) {
      float t[8];
      unpack8(__ldg(ap + idx), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += kRms ? v[i][j] * v[i][j] : v[i][j];
  }
  sum = warp_sum(sum);
  float mean = 0.f, rstd;
  if (kRms) {
    rstd = rsqrtf(sum / a.D + a.eps);
  } else {
    mean = sum / a.D;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        var += d * d;
      }
    var = warp_sum(var);
    rstd = rsqrtf(var / a.D + a.eps);
  }
  const uint4* gp = reinterpret_cast<const uint4*>(a.gamma);
  const uint4* bp = reinterpret_cast<const uint4*>(a.beta);
  const long long orow =
      a.out_group > 0 ? static_cast<long long>(row / a.out_group) * a.out_group_stride + row % a.out_group : row;
  uint4* op = reinterpret_cast<uint4*>(a.out + orow * a.ldo);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int idx = i * 32 + lane;
    float gm[8], bt[8];
    unpack8(__ldg(gp + idx), gm);
    if (!kRms) unpack8(__ldg(bp + idx), bt);
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = kRms ? v[i][j] * rstd * gm[j] : (v[i][j] - mean) * rstd * gm[j] + bt[j];
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]);
    o.w = pack_bf16x2(y[6], y[7]);
    op[idx] = o;
  }
}

template <bool kRms>
static int launch_norm(const LnArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return 0;
  if (a.D % 256 != 0) return fail(-2, "norm: hidden size %d must be a multiple of 256", a.D);
  if ((a.ldx % 8) || (a.ldo % 8)) return fail(-2, "norm: row strides must be multiples of 8 elements");
  const int grid = (a.M + 3) / 4;
  LaunchScope scope(kFamNorm, stream, 4.0 * a.M * a.D, 0.0);
  switch (a.D / 256) {
    case 1: B200_CUDA_OK(launch_ex(norm_kernel<1, kRms>, dim3(grid), dim3(128), 0, stream, 0, true, a)); break;
    case 2: B200_CUDA_OK(launch_ex(norm_kernel<2, kRms>, dim3(grid), dim3(128), 0, stream, 0, true, a)); break;
    case 4: B200_CUDA_OK(launch_ex(norm_kernel<4, kRms>, dim3(grid), dim3(128), 0, stream, 0, true, a)); break;
    case 8: B200_CUDA_OK(launch_ex(norm_kernel<8, kRms>, dim3(grid), dim3(128), 0, stream, 0, true, a)); break;
    case 16: B200_CUDA_OK(launch_ex(norm_kernel<16, kRms>, dim3(grid), dim3(128), 0, stream, 0, true, a)); break;
    default: return fail(-2, "norm: hidden size %d not supported (256/512/1024/2048/4096)", a.D);
  }
  return 0;
}

int layernorm(const bf16* x, long long ldx, const int* row_map, const bf16* add, int period, const bf16* gamma,
              const bf16* beta, float eps, bf16* out, long long ldo, int M, int D, int out_group,
              int out_group_stride, cudaStream_t stream) {
  LnArgs a{x, ldx, row_map, add, period > 0 ? period : 1, gamma, beta, eps, out, ldo, M, D, out_group, out_group_stride};
  return launch_norm<false>(a, stream);
}

int rmsnorm(const bf16* x, long long ldx, const bf16* w, float eps, bf16* out, long long ldo, int M, int D,
            cudaStream_t stream) {
  LnArgs a{x, ldx, nullptr, nullptr, 1, w, nullptr, eps, out, ldo, M, D, 0, 0};
  return launch_norm<true>(a, stream);
}

}  // namespace b200
