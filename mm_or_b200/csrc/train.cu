// Training-side kernels of the MM2SG fine-tune step.
//
//  weighted_ce     LLaVATrainer.compute_loss (train/llava_trainer.py:136-174): shifted cross-entropy over
//                  `modified_labels` with per-class weights, nn.CrossEntropyLoss(weight=vocab_weight) semantics:
//                      loss = sum_i w[y_i] * nll_i / sum_i w[y_i]        over positions with y_i != -100
//                  (vocab_weight[v] = 1 / (ln f_v + 1) for the listed pieces, min_w / 100 otherwise: train.py:1316-1322).
//                  The reference materialises a contiguous shifted copy of the (B, L, 32000) logits and runs log_softmax
//                  + nll_loss + the backward of both as separate kernels. Here one pass per supervised row computes
//                  max / log-sum-exp with 16-byte loads, the weighted nll, and (optionally) writes
//                  dlogits = w[y] / W * (softmax - onehot) * grad_scale in place of the logits; rows that carry no
//                  label (prompt, visual tokens, padding, the last position) are only zero-filled. The shift is index
//                  arithmetic (row (b, l) reads label (b, l + 1)). Reductions are two-stage in a fixed order, so the
//                  loss is bit-deterministic.
//  adamw / clip    LLaVATrainer.create_optimizer (train/llava_trainer.py:191-278) -> torch.optim.AdamW + HF Trainer
//                  clip_grad_norm_(max_grad_norm = 0.1) (README.md:147-151): the reference runs for-each kernels with
//                  >= 4 passes over params / grads / m / v. Here: one two-stage sum of squares over the flat gradient
//                  buffer, then ONE pass that applies the clip coefficient (read from device memory, no host sync),
//                  decoupled weight decay, the Adam update on fp32 master weights and writes the bf16 working copy.
//                  HBM bound: 2 (grad) + 4 + 4 + 4 (read p, m, v) + 4 + 4 + 4 + 2 (write) = 28 B per parameter.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

static constexpr int kCeThreads = 256;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, v, off);
    v = is_max ? fmaxf(v, o) : v + o;
  }
  __syncthreads();  // red may still be read from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < kCeThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// stage 1 of the weight normaliser: W = sum of w[label] over supervised positions (fixed-order two-stage sum)
__global__ void __launch_bounds__(kCeThreads) ce_weight_sum_kernel(const long long* __restrict__ labels,
                                                                   const float* __restrict__ vocab_w, int Bn, int L,
                                                                   int V, float* __restrict__ wsum) {
  __shared__ float red[kCeThreads / 32];
  float acc = 0.f;
  const long long T = static_cast<long long>(Bn) * L;
  for (long long i = threadIdx.x; i < T; i += kCeThreads) {  // single CTA: order independent of the grid
    const int l = static_cast<int>(i % L);
    if (l == L - 1) continue;
    const long long y = labels[i + 1];
    if (y >= 0 && y < V) acc += vocab_w ? vocab_w[y] : 1.f;
  }
  const float tot = block_reduce(acc, red, false);
  if (threadIdx.x == 0) *wsum = tot;
}

template <typename T>
__device__ __forceinline__ float ld_logit(const T* p, long long i);
template <>
__device__ __forceinline__ float ld_logit<float>(const float* p, long long i) {
  return p[i];
}
template <>
__device__ __forceinline__ float ld_logit<bf16>(const bf16* p, long long i) {
  return __bfloat162float(p[i]);
}
template <typename T>
__device__ __forceinline__ void st_logit(T* p, long long i, float v);
template <>
__device__ __forceinline__ void st_logit<float>(float* p, long long i, float v) {
  p[i] = v;
}
template <>
__device__ __forceinline__ void st_logit<bf16>(bf16* p, long long i, float v) {
  p[i] = __float2bfloat16(v);
}

// one CTA per (b, l) row. row_loss[row] = w[y] * nll (0 for unsupervised rows).
template <typename T>
__global__ void __launch_bounds__(kCeThreads)
    ce_row_kernel(const T* logits /* may alias dlogits: no __restrict__ */, long long ld,
                  const long long* __restrict__ labels,
                  const float* __restrict__ vocab_w, int L, int V, const float* __restrict__ wsum, float grad_scale,
                  T* dlogits, long long ldd, float* __restrict__ row_loss) {
  __shared__ float red[kCeThreads / 32];
  const long long row = blockIdx.x;
  const int l = static_cast<int>(row % L);
  const long long y = (l == L - 1) ? -100 : labels[row + 1];
  const bool live = y >= 0 && y < V;
  T* drow = dlogits ? dlogits + row * ldd : nullptr;
  if (!live) {
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    if (drow != nullptr)
      for (int i = threadIdx.x; i < V; i += kCeThreads) st_logit<T>(drow, i, 0.f);
    return;
  }
  const T* xr = logits + row * ld;
  // dlogits may alias logits (in-place gradient): every thread reads the label's logit BEFORE the barriers of the
  // reductions below, i.e. before any thread of the block can reach the store loop that overwrites it
  const float x_y = ld_logit<T>(xr, y);
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += kCeThreads) mx = fmaxf(mx, ld_logit<T>(xr, i));
  mx = block_reduce(mx, red, true);
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += kCeThreads) se += __expf(ld_logit<T>(xr, i) - mx);
  se = block_reduce(se, red, false);
  const float lse = mx + __logf(se);
  const float w = vocab_w ? vocab_w[y] : 1.f;
  if (threadIdx.x == 0) row_loss[row] = w * (lse - x_y);
  if (drow != nullptr) {
    const float W = *wsum;
    const float coef = W > 0.f ? grad_scale * w / W : 0.f;
    for (int i = threadIdx.x; i < V; i += kCeThreads) {
      const float p = __expf(ld_logit<T>(xr, i) - lse);
      st_logit<T>(drow, i, coef * (p - (i == y ? 1.f : 0.f)));
    }
  }
}

__global__ void __launch_bounds__(kCeThreads) ce_finish_kernel(const float* __restrict__ row_loss, long long rows,
                                                               const float* __restrict__ wsum,
                                                               float* __restrict__ loss) {
  __shared__ float red[kCeThreads / 32];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < rows; i += kCeThreads) acc += row_loss[i];
  const float tot = block_reduce(acc, red, false);
  if (threadIdx.x == 0) {
    const float W = *wsum;
    loss[0] = W > 0.f ? tot / W : 0.f;  // no supervised token: torch returns nan; we return 0 and report W
    loss[1] = W;
  }
}

size_t weighted_ce_workspace_bytes(int Bn, int L) { return (static_cast<size_t>(Bn) * L + 64) * sizeof(float); }

// logits [Bn*L, V] (row stride ld), labels [Bn, L] int64 (UNshifted modified_labels), vocab_w [V] fp32 or null.
// loss_out[2] = {loss, sum of weights}. dlogits may be null (forward only) or alias logits.
int weighted_ce(const void* logits, int is_fp32, long long ld, const long long* labels, const float* vocab_w, int Bn,
                int L, int V, float grad_scale, void* dlogits, long long ldd, float* loss_out, void* workspace,
                size_t workspace_bytes, cudaStream_t stream) {
  if (Bn <= 0 || L <= 0) return fail(-2, "weighted_ce: empty batch");
  if (workspace == nullptr || workspace_bytes < weighted_ce_workspace_bytes(Bn, L))
    return fail(-2, "weighted_ce: workspace too small");
  float* wsum = static_cast<float*>(workspace);
  float* row_loss = wsum + 64;
  const long long rows = static_cast<long long>(Bn) * L;
  LaunchScope scope(kFamTrain, stream, (dlogits ? 3.0 : 2.0) * rows * V * (is_fp32 ? 4 : 2), 0.0, 3);
  ce_weight_sum_kernel<<<1, kCeThreads, 0, stream>>>(labels, vocab_w, Bn, L, V, wsum);
  if (is_fp32)
    ce_row_kernel<float><<<static_cast<unsigned>(rows), kCeThreads, 0, stream>>>(
        static_cast<const float*>(logits), ld, labels, vocab_w, L, V, wsum, grad_scale, static_cast<float*>(dlogits),
        ldd, row_loss);
  else
    ce_row_kernel<bf16><<<static_cast<unsigned>(rows), kCeThreads, 0, stream>>>(
        static_cast<const bf16*>(logits), ld, labels, vocab_w, L, V, wsum, grad_scale, static_cast<bf16*>(dlogits),
        ldd, row_loss);
  ce_finish_kernel<<<1, kCeThreads, 0, stream>>>(row_loss, rows, wsum, loss_out);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// gradient norm + fused AdamW
// ---------------------------------------------------------------------------------------------------
static constexpr int kOptThreads = 256;
static constexpr int kNormBlocks = 1184;  // 8 per SM; fixed so that the two-stage sum has a fixed order

__device__ __forceinline__ float grad_at(const bf16* g, long long i) { return __bfloat162float(g[i]); }
__device__ __forceinline__ float grad_at(const float* g, long long i) { return g[i]; }

template <typename GT>
__global__ void __launch_bounds__(kOptThreads) grad_sq_partial_kernel(const GT* __restrict__ g, long long n,
                                                                      float* __restrict__ partial) {
  __shared__ float red[kOptThreads / 32];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * kOptThreads + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * kOptThreads) {
    const float f = grad_at(g, i);
    acc += f * f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// out[0] += sum(partial) (accumulate = 1 chains several gradient buffers), out[1] = clip coefficient for max_norm
__global__ void __launch_bounds__(kOptThreads) grad_norm_finish_kernel(const float* __restrict__ partial, int nblocks,
                                                                       int accumulate, float max_norm,
                                                                       float* __restrict__ out) {
  __shared__ float red[kOptThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += kOptThreads) acc += partial[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = accumulate ? out[0] : 0.f;
    for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
    out[0] = t;
    const float norm = sqrtf(t);
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
    out[1] = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
  }
}

size_t grad_norm_workspace_bytes() { return kNormBlocks * sizeof(float); }

int grad_sq_norm(const void* grad, int grad_fp32, long long n, int accumulate, float max_norm, float* out2,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < grad_norm_workspace_bytes())
    return fail(-2, "grad_sq_norm: workspace too small");
  LaunchScope scope(kFamTrain, stream, (grad_fp32 ? 4.0 : 2.0) * n, 0.0, 2);
  float* partial = static_cast<float*>(workspace);
  if (grad_fp32)
    grad_sq_partial_kernel<float><<<kNormBlocks, kOptThreads, 0, stream>>>(static_cast<const float*>(grad), n, partial);
  else
    grad_sq_partial_kernel<bf16><<<kNormBlocks, kOptThreads, 0, stream>>>(static_cast<const bf16*>(grad), n, partial);
  grad_norm_finish_kernel<<<1, kOptThreads, 0, stream>>>(partial, kNormBlocks, accumulate, max_norm, out2);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename GT>
struct AdamArgs {
  float* master;    // fp32 master weights
  bf16* param;      // bf16 working copy (may be null)
  const GT* grad;
  float* m;
  float* v;
  long long n;
  float lr, beta1, beta2, eps, weight_decay;
  float bc1, bc2_sqrt;       // 1 - beta1^t, sqrt(1 - beta2^t)
  const float* clip;         // device pointer to the clip coefficient (out2 + 1 of grad_sq_norm) or null
};

// one element of the update (torch.optim.AdamW arithmetic); shared by the scalar and the 4-wide path so that both
// produce the same bits
__device__ __forceinline__ void adam_elem(float g, float clip, float decay, float step_size, float beta1, float beta2,
                                          float bc2_sqrt, float eps, float& p, float& m, float& v) {
  g *= clip;
  p *= decay;                                                    // decoupled weight decay
  m = beta1 * m + (1.f - beta1) * g;
  v = beta2 * v + (1.f - beta2) * g * g;
  p -= step_size * m / (sqrtf(v) / bc2_sqrt + eps);
}

__device__ __forceinline__ void load_grad4(const float* g, long long i4, float (&o)[4]) {
  const float4 t = reinterpret_cast<const float4*>(g)[i4];
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void load_grad4(const bf16* g, long long i4, float (&o)[4]) {
  const uint2 t = reinterpret_cast<const uint2*>(g)[i4];
  o[0] = bf16lo(t.x); o[1] = bf16hi(t.x); o[2] = bf16lo(t.y); o[3] = bf16hi(t.y);
}

// VEC: n % 4 == 0 and every pointer is 16-byte aligned (8 for the bf16 ones): 16-byte loads / stores, four elements per
// thread and iteration (the kernel is a pure HBM stream: 28 B per parameter). Same per-element arithmetic.
template <typename GT, bool VEC>
__global__ void __launch_bounds__(kOptThreads) adamw_kernel(const AdamArgs<GT> a) {
  const float clip = a.clip ? *a.clip : 1.f;
  const float step_size = a.lr / a.bc1;
  const float decay = 1.f - a.lr * a.weight_decay;
  if (VEC) {
    const long long n4 = a.n >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * kOptThreads + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * kOptThreads) {
      float g[4];
      load_grad4(a.grad, i, g);
      float4 p = reinterpret_cast<float4*>(a.master)[i];
      float4 m = reinterpret_cast<float4*>(a.m)[i];
      float4 v = reinterpret_cast<float4*>(a.v)[i];
      adam_elem(g[0], clip, decay, step_size, a.beta1, a.beta2, a.bc2_sqrt, a.eps, p.x, m.x, v.x);
      adam_elem(g[1], clip, decay, step_size, a.beta1, a.beta2, a.bc2_sqrt, a.eps, p.y, m.y, v.y);
      adam_elem(g[2], clip, decay, step_size, a.beta1, a.beta2, a.bc2_sqrt, a.eps, p.z, m.z, v.z);
      adam_elem(g[3], clip, decay, step_size, a.beta1, a.beta2, a.bc2_sqrt, a.eps, p.w, m.w, v.w);
      reinterpret_cast<float4*>(a.master)[i] = p;
      reinterpret_cast<float4*>(a.m)[i] = m;
      reinterpret_cast<float4*>(a.v)[i] = v;
      if (a.param != nullptr)
        reinterpret_cast<uint2*>(a.param)[i] = make_uint2(pack_bf16x2(p.x, p.y), pack_bf16x2(p.z, p.w));
    }
    return;
  }
  for (long long i = static_cast<long long>(blockIdx.x) * kOptThreads + threadIdx.x; i < a.n;
       i += static_cast<long long>(gridDim.x) * kOptThreads) {
    float p = a.master[i], m = a.m[i], v = a.v[i];
    adam_elem(grad_at(a.grad, i), clip, decay, step_size, a.beta1, a.beta2, a.bc2_sqrt, a.eps, p, m, v);
    a.master[i] = p;
    a.m[i] = m;
    a.v[i] = v;
    if (a.param != nullptr) a.param[i] = __float2bfloat16(p);
  }
}

template <typename GT>
static int adamw_launch(float* master, bf16* param, const GT* grad, float* m, float* v, long long n, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int step, const float* clip_coef,
                        cudaStream_t stream) {
  AdamArgs<GT> a;
  a.master = master;
  a.param = param;
  a.grad = grad;
  a.m = m;
  a.v = v;
  a.n = n;
  a.lr = lr;
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = eps;
  a.weight_decay = weight_decay;
  a.bc1 = 1.f - powf(beta1, static_cast<float>(step));
  a.bc2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  a.clip = clip_coef;
  auto al = [](const void* p, uintptr_t k) { return (reinterpret_cast<uintptr_t>(p) & (k - 1)) == 0; };
  const bool vec = (n & 3) == 0 && al(master, 16) && al(m, 16) && al(v, 16) && al(grad, 4 * sizeof(GT)) &&
                   (param == nullptr || al(param, 8));
  const long long items = vec ? n / 4 : n;
  const long long want = (items + kOptThreads - 1) / kOptThreads;
  const int grid = static_cast<int>(want < 1 ? 1 : (want > 16LL * num_sms() ? 16LL * num_sms() : want));
  LaunchScope scope(kFamTrain, stream, (26.0 + sizeof(GT)) * n, 0.0);
  if (vec)
    adamw_kernel<GT, true><<<grid, kOptThreads, 0, stream>>>(a);
  else
    adamw_kernel<GT, false><<<grid, kOptThreads, 0, stream>>>(a);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

int adamw_step(float* master, bf16* param, const void* grad, int grad_fp32, float* m, float* v, long long n, float lr,
               float beta1, float beta2, float eps, float weight_decay, int step, const float* clip_coef,
               cudaStream_t stream) {
  if (n <= 0) return 0;
  if (step < 1) return fail(-2, "adamw_step: step counts from 1");
  if (grad_fp32)
    return adamw_launch<float>(master, param, static_cast<const float*>(grad), m, v, n, lr, beta1, beta2, eps,
                               weight_decay, step, clip_coef, stream);
  return adamw_launch<bf16>(master, param, static_cast<const bf16*>(grad), m, v, n, lr, beta1, beta2, eps, weight_decay,
                            step, clip_coef, stream);
}

// ---------------------------------------------------------------------------------------------------
// backward building blocks (bias gradient, activation / norm backward). Forward formulas: gemm_sm100.cu act_apply,
// norm.cu; references: HF CLIPMLP quick_gelu, BertIntermediate gelu(erf), LlamaMLP silu(gate) * up, nn.LayerNorm,
// LlamaRMSNorm (fp32 statistics).
// ---------------------------------------------------------------------------------------------------
static constexpr int kColSplit = 64;

// db[n] (+)= sum_m dY[m, n]: two-stage, fixed order. stage 1: grid (ceil(N/256), kColSplit)
__global__ void __launch_bounds__(256) colsum_partial_kernel(const bf16* __restrict__ dy, long long ld, int M, int N,
                                                             float* __restrict__ partial) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const int rows = (M + kColSplit - 1) / kColSplit;
  const int m0 = blockIdx.y * rows, m1 = min(M, m0 + rows);
  float acc = 0.f;
  for (int m = m0; m < m1; ++m) acc += __bfloat162float(dy[static_cast<long long>(m) * ld + n]);
  partial[static_cast<long long>(blockIdx.y) * N + n] = acc;
}
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ partial, int N, int accumulate,
                                                            float* __restrict__ out) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float acc = accumulate ? out[n] : 0.f;
  for (int s = 0; s < kColSplit; ++s) acc += partial[static_cast<long long>(s) * N + n];
  out[n] = acc;
}

size_t colsum_workspace_bytes(int N) { return static_cast<size_t>(kColSplit) * N * sizeof(float); }

int colsum(const bf16* dy, long long ld, int M, int N, int accumulate, float* out, void* workspace,
           size_t workspace_bytes, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (workspace == nullptr || workspace_bytes < colsum_workspace_bytes(N)) return fail(-2, "colsum: workspace too small");
  LaunchScope scope(kFamTrain, stream, 2.0 * M * N, 0.0, 2);
  float* partial = static_cast<float*>(workspace);
  colsum_partial_kernel<<<dim3((N + 255) / 256, kColSplit), 256, 0, stream>>>(dy, ld, M, N, partial);
  colsum_finish_kernel<<<(N + 255) / 256, 256, 0, stream>>>(partial, N, accumulate, out);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// dz = dy * act'(z) for the pre-activation z saved by the forward GEMM (act: kActQuickGelu / kActGeluErf), or for
// SwiGLU: z = interleaved (gate, up) pairs [M, 2F], dy [M, F] -> dz [M, 2F] interleaved.
__global__ void __launch_bounds__(256) act_backward_kernel(const bf16* __restrict__ z, const bf16* __restrict__ dy,
                                                           bf16* __restrict__ dz, long long n_out, int act) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;  // one output element of dy each
  if (i >= n_out) return;
  const float g = __bfloat162float(dy[i]);
  if (act == kActSwiGLU) {
    const float gate = __bfloat162float(z[2 * i]), up = __bfloat162float(z[2 * i + 1]);
    const float sg = 1.f / (1.f + __expf(-gate));
    const float silu = gate * sg;
    dz[2 * i] = __float2bfloat16(g * up * (sg + silu * (1.f - sg)));   // d silu(x) = s + x s (1 - s)
    dz[2 * i + 1] = __float2bfloat16(g * silu);
    return;
  }
  const float x = __bfloat162float(z[i]);
  float d;
  if (act == kActQuickGelu) {
    const float sg = 1.f / (1.f + __expf(-1.702f * x));
    d = sg + 1.702f * x * sg * (1.f - sg);
  } else {  // gelu(erf): Phi(x) + x phi(x)
    d = 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
  }
  dz[i] = __float2bfloat16(g * d);
}

int act_backward(const bf16* z, const bf16* dy, bf16* dz, long long n_out, int act, cudaStream_t stream) {
  if (n_out <= 0) return 0;
  if (act != kActQuickGelu && act != kActGeluErf && act != kActSwiGLU) return fail(-2, "act_backward: unknown act %d", act);
  LaunchScope scope(kFamTrain, stream, (act == kActSwiGLU ? 10.0 : 6.0) * n_out, 0.0);
  act_backward_kernel<<<static_cast<unsigned>((n_out + 255) / 256), 256, 0, stream>>>(z, dy, dz, n_out, act);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// h[m, j] = silu(z[m, 2j]) * z[m, 2j+1]: the SwiGLU of the training forward, which keeps the pre-activation z for
// act_backward (the inference path fuses the same formula into the GEMM epilogue and never stores z)
__global__ void __launch_bounds__(256) swiglu_forward_kernel(const bf16* __restrict__ z, bf16* __restrict__ h,
                                                             long long n_out) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n_out) return;
  const uint32_t gu = reinterpret_cast<const uint32_t*>(z)[i];
  const float gate = bf16lo(gu), up = bf16hi(gu);
  h[i] = __float2bfloat16(gate / (1.f + __expf(-gate)) * up);
}

int swiglu_forward(const bf16* z, bf16* h, long long n_out, cudaStream_t stream) {
  if (n_out <= 0) return 0;
  LaunchScope scope(kFamTrain, stream, 6.0 * n_out, 0.0);
  swiglu_forward_kernel<<<static_cast<unsigned>((n_out + 255) / 256), 256, 0, stream>>>(z, h, n_out);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// y = act(z) elementwise (quick_gelu / gelu-erf) from the stored pre-activation: training forward of the CLIP MLP, the
// BERT pooler FFN and the mm_projector (the inference path fuses the activation into the GEMM epilogue)
__global__ void __launch_bounds__(256) act_forward_kernel(const bf16* __restrict__ z, bf16* __restrict__ y, long long n,
                                                          int act) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float x = __bfloat162float(z[i]);
  const float r = act == kActQuickGelu ? x / (1.f + __expf(-1.702f * x))
                                       : 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  y[i] = __float2bfloat16(r);
}

int act_forward(const bf16* z, bf16* y, long long n, int act, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (act != kActQuickGelu && act != kActGeluErf) return fail(-2, "act_forward: unknown act %d", act);
  LaunchScope scope(kFamTrain, stream, 4.0 * n, 0.0);
  act_forward_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(z, y, n, act);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// out[r, :] (+)= sum_g x[g, r, :] over `groups` slabs of [rows, D]: gradient of a table broadcast over the batch
// (BERT position + token-type embeddings of the image pooler). Fixed summation order.
__global__ void __launch_bounds__(256) group_sum_kernel(const bf16* __restrict__ x, int groups, long long slab,
                                                        int accumulate, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= slab) return;
  float acc = accumulate ? out[i] : 0.f;
  for (int g = 0; g < groups; ++g) acc += __bfloat162float(x[g * slab + i]);
  out[i] = acc;
}

int group_sum(const bf16* x, int groups, long long slab, int accumulate, float* out, cudaStream_t stream) {
  if (groups <= 0 || slab <= 0) return 0;
  LaunchScope scope(kFamTrain, stream, 2.0 * groups * slab, 0.0);
  group_sum_kernel<<<static_cast<unsigned>((slab + 255) / 256), 256, 0, stream>>>(x, groups, slab, accumulate, out);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// out[r] = src[row_map[r]] (zero row when row_map[r] < 0) + add[r % period]: the pooler's pre-LayerNorm embedding sum
// (gathered patch features + position / token-type embeddings), materialised for the backward of its LayerNorm
// (the inference path fuses the same gather + add into the LayerNorm kernel, norm.cu)
__global__ void __launch_bounds__(128) gather_add_rows_kernel(const bf16* __restrict__ src, long long ld_src,
                                                              const int* __restrict__ row_map,
                                                              const bf16* __restrict__ add, int period, int rows, int D,
                                                              bf16* __restrict__ out) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int sr = row_map ? row_map[r] : r;
  const uint4* sp = sr >= 0 ? reinterpret_cast<const uint4*>(src + static_cast<long long>(sr) * ld_src) : nullptr;
  const uint4* ap = add ? reinterpret_cast<const uint4*>(add + static_cast<long long>(r % period) * D) : nullptr;
  uint4* op = reinterpret_cast<uint4*>(out + static_cast<long long>(r) * D);
  for (int i = lane; i < D / 8; i += 32) {
    const uint4 u = sp ? sp[i] : make_uint4(0, 0, 0, 0);
    const uint4 a = ap ? __ldg(ap + i) : make_uint4(0, 0, 0, 0);
    uint4 o;
    o.x = pack_bf16x2(bf16lo(u.x) + bf16lo(a.x), bf16hi(u.x) + bf16hi(a.x));
    o.y = pack_bf16x2(bf16lo(u.y) + bf16lo(a.y), bf16hi(u.y) + bf16hi(a.y));
    o.z = pack_bf16x2(bf16lo(u.z) + bf16lo(a.z), bf16hi(u.z) + bf16hi(a.z));
    o.w = pack_bf16x2(bf16lo(u.w) + bf16lo(a.w), bf16hi(u.w) + bf16hi(a.w));
    op[i] = o;
  }
}

int gather_add_rows(const bf16* src, long long ld_src, const int* row_map, const bf16* add, int period, int rows, int D,
                    bf16* out, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (D % 8 != 0 || ld_src % 8 != 0) return fail(-2, "gather_add_rows: D and the source stride must be multiples of 8");
  LaunchScope scope(kFamTrain, stream, 4.0 * rows * D, 0.0);
  gather_add_rows_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(src, ld_src, row_map, add, period > 0 ? period : 1, rows,
                                                             D, out);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// Backward of rope_kv_write (misc.cu): gathers dK / dV from the head-major cache layout [B][H][cap][128] back into the
// token-major dqkv [B*L, 3*H*128] and applies the inverse rotation (angle -theta) to dq (in place) and dk.
__global__ void __launch_bounds__(128) rope_kv_backward_kernel(bf16* __restrict__ dqkv, const int* __restrict__ kv_start,
                                                               const float* __restrict__ cos_t,
                                                               const float* __restrict__ sin_t,
                                                               const bf16* __restrict__ dkc, const bf16* __restrict__ dvc,
                                                               int T, int H, int Lq, int cap, int max_pos) {
  const long long w = blockIdx.x * 4ll + (threadIdx.x >> 5);  // one warp per (token, head)
  if (w >= static_cast<long long>(T) * H) return;
  const int lane = threadIdx.x & 31;
  const int h = w % H;
  const int tok = w / H;
  const int b = tok / Lq, l = tok % Lq;
  int p = l - (kv_start ? kv_start[b] : 0);
  p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
  const float2 cs = *reinterpret_cast<const float2*>(cos_t + static_cast<long long>(p) * 64 + lane * 2);
  const float2 sn = *reinterpret_cast<const float2*>(sin_t + static_cast<long long>(p) * 64 + lane * 2);
  bf16* row = dqkv + static_cast<long long>(tok) * 3 * H * 128;
  const bf16* crow_k = dkc + ((static_cast<long long>(b) * H + h) * cap + l) * 128;
  const bf16* crow_v = dvc + ((static_cast<long long>(b) * H + h) * cap + l) * 128;
#pragma unroll
  for (int which = 0; which < 2; ++which) {  // 0 = dq (in place), 1 = dk (from the cache layout)
    bf16* dst = row + which * H * 128 + h * 128;
    const bf16* src = which == 0 ? dst : crow_k;
    const uint32_t lo = *reinterpret_cast<const uint32_t*>(src + lane * 2);
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(src + 64 + lane * 2);
    const float x0 = bf16lo(lo), x1 = bf16hi(lo), y0 = bf16lo(hi), y1 = bf16hi(hi);
    *reinterpret_cast<uint32_t*>(dst + lane * 2) = pack_bf16x2(x0 * cs.x + y0 * sn.x, x1 * cs.y + y1 * sn.y);
    *reinterpret_cast<uint32_t*>(dst + 64 + lane * 2) = pack_bf16x2(y0 * cs.x - x0 * sn.x, y1 * cs.y - x1 * sn.y);
  }
  *reinterpret_cast<uint2*>(row + 2 * H * 128 + h * 128 + lane * 4) = *reinterpret_cast<const uint2*>(crow_v + lane * 4);
}

int rope_kv_backward(bf16* dqkv, const int* kv_start, const float* cos_t, const float* sin_t, int max_pos,
                     const bf16* dk_cache, const bf16* dv_cache, int Bn, int H, int Lq, int cap, cudaStream_t stream) {
  if (Bn <= 0 || Lq <= 0) return 0;
  if (Lq > cap) return fail(-2, "rope_kv_backward: L %d exceeds the cache capacity %d", Lq, cap);
  const long long warps = static_cast<long long>(Bn) * Lq * H;
  LaunchScope scope(kFamTrain, stream, 10.0 * warps * 128, 0.0);
  rope_kv_backward_kernel<<<static_cast<unsigned>((warps + 3) / 4), 128, 0, stream>>>(
      dqkv, kv_start, cos_t, sin_t, dk_cache, dv_cache, Bn * Lq, H, Lq, cap, max_pos);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// LayerNorm / RMSNorm backward. One CTA walks a block of rows; every thread owns 8 * kVec fixed columns
// (D = kThreads * 8 * kVec), so the dgamma / dbeta sums of the CTA's rows stay in registers; fp32 statistics are
// recomputed from x with two block reductions per row:
//   LN : xhat = (x - mean) rstd;  dx = rstd (dyg - mean(dyg) - xhat mean(dyg xhat)),  dyg = dy * gamma
//   RMS: xhat = x rstd;           dx = rstd (dyg - xhat mean(dyg xhat))
// Per-CTA dgamma / dbeta partials are summed in block order by norm_param_finish_kernel (deterministic).
template <int kThreads>
__device__ __forceinline__ void block_sum2(float& a, float& b, float (*red)[2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
  }
  __syncthreads();
  if (lane == 0) {
    red[warp][0] = a;
    red[warp][1] = b;
  }
  __syncthreads();
  a = b = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    a += red[w][0];
    b += red[w][1];
  }
}

template <int kThreads, int kVec, bool kRms>
__global__ void __launch_bounds__(kThreads) norm_backward_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                                 const bf16* __restrict__ gamma, float eps, int M,
                                                                 int D, const bf16* __restrict__ add,
                                                                 bf16* __restrict__ dx,
                                                                 float* __restrict__ pgamma, float* __restrict__ pbeta,
                                                                 int rows_per_block) {
  __shared__ float red[kThreads / 32][2];
  float gm[kVec][8], dg[kVec][8], db[kVec][8];
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(gamma) + i * kThreads + threadIdx.x);
    const float t[8] = {bf16lo(g4.x), bf16hi(g4.x), bf16lo(g4.y), bf16hi(g4.y),
                        bf16lo(g4.z), bf16hi(g4.z), bf16lo(g4.w), bf16hi(g4.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gm[i][j] = t[j];
      dg[i][j] = db[i][j] = 0.f;
    }
  }
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  for (int row = r0; row < r1; ++row) {
    const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * D);
    const uint4* gp = reinterpret_cast<const uint4*>(dy + static_cast<long long>(row) * D);
    float xv[kVec][8], gv[kVec][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const uint4 u = xp[i * kThreads + threadIdx.x], w = gp[i * kThreads + threadIdx.x];
      const float a[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y),
                          bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
      const float b[8] = {bf16lo(w.x), bf16hi(w.x), bf16lo(w.y), bf16hi(w.y),
                          bf16lo(w.z), bf16hi(w.z), bf16lo(w.w), bf16hi(w.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xv[i][j] = a[j];
        gv[i][j] = b[j];
        s1 += a[j];
        s2 += a[j] * a[j];
      }
    }
    block_sum2<kThreads>(s1, s2, red);
    float mean = 0.f, rstd;
    if (kRms) {
      rstd = rsqrtf(s2 / D + eps);
    } else {
      mean = s1 / D;
      float var = 0.f, dummy = 0.f;
#pragma unroll
      for (int i = 0; i < kVec; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = xv[i][j] - mean;
          var += d * d;
        }
      block_sum2<kThreads>(var, dummy, red);
      rstd = rsqrtf(var / D + eps);
    }
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xhat = (xv[i][j] - mean) * rstd;
        dg[i][j] += gv[i][j] * xhat;
        db[i][j] += gv[i][j];
        const float dyg = gv[i][j] * gm[i][j];
        xv[i][j] = xhat;
        gv[i][j] = dyg;
        c1 += dyg;
        c2 += dyg * xhat;
      }
    block_sum2<kThreads>(c1, c2, red);
    c1 = kRms ? 0.f : c1 / D;
    c2 /= D;
    uint4* op = reinterpret_cast<uint4*>(dx + static_cast<long long>(row) * D);
    const uint4* ap = add ? reinterpret_cast<const uint4*>(add + static_cast<long long>(row) * D) : nullptr;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = rstd * (gv[i][j] - c1 - xv[i][j] * c2);
      if (ap != nullptr) {  // gradient arriving through the residual branch around the norm (may alias dx)
        const uint4 u = ap[i * kThreads + threadIdx.x];
        y[0] += bf16lo(u.x);
        y[1] += bf16hi(u.x);
        y[2] += bf16lo(u.y);
        y[3] += bf16hi(u.y);
        y[4] += bf16lo(u.z);
        y[5] += bf16hi(u.z);
        y[6] += bf16lo(u.w);
        y[7] += bf16hi(u.w);
      }
      uint4 o;
      o.x = pack_bf16x2(y[0], y[1]);
      o.y = pack_bf16x2(y[2], y[3]);
      o.z = pack_bf16x2(y[4], y[5]);
      o.w = pack_bf16x2(y[6], y[7]);
      op[i * kThreads + threadIdx.x] = o;
    }
  }
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    float* pgd = pgamma + static_cast<long long>(blockIdx.x) * D + (i * kThreads + threadIdx.x) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) pgd[j] = dg[i][j];
    if (!kRms && pbeta != nullptr) {
      float* pbd = pbeta + static_cast<long long>(blockIdx.x) * D + (i * kThreads + threadIdx.x) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) pbd[j] = db[i][j];
    }
  }
}

__global__ void __launch_bounds__(256) norm_param_finish_kernel(const float* __restrict__ partial, int blocks, int D,
                                                                int accumulate, float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= D) return;
  float acc = accumulate ? out[c] : 0.f;
  for (int b = 0; b < blocks; ++b) acc += partial[static_cast<long long>(b) * D + c];
  out[c] = acc;
}

static constexpr int kNormBwdRows = 16;  // rows per CTA

size_t norm_backward_workspace_bytes(int M, int D) {
  const size_t blocks = (M + kNormBwdRows - 1) / kNormBwdRows;
  return 2 * blocks * D * sizeof(float);
}

// dgamma / dbeta: fp32 [D], (+)= when accumulate. beta == rms -> dbeta ignored (pass nullptr).
int norm_backward(const bf16* x, const bf16* dy, const bf16* gamma, float eps, int M, int D, int rms, const bf16* add,
                  bf16* dx, float* dgamma, float* dbeta, int accumulate, void* workspace, size_t workspace_bytes,
                  cudaStream_t stream) {
  if (M <= 0) return 0;
  if (workspace == nullptr || workspace_bytes < norm_backward_workspace_bytes(M, D))
    return fail(-2, "norm_backward: workspace too small");
  const int blocks = (M + kNormBwdRows - 1) / kNormBwdRows;
  float* pg = static_cast<float*>(workspace);
  float* pb = pg + static_cast<size_t>(blocks) * D;
  LaunchScope scope(kFamTrain, stream, 6.0 * M * D, 0.0, 3);
#define B200_NB(T, V)                                                                                          \
  if (rms)                                                                                                     \
    norm_backward_kernel<T, V, true><<<blocks, T, 0, stream>>>(x, dy, gamma, eps, M, D, add, dx, pg, nullptr, \
                                                               kNormBwdRows);                                 \
  else                                                                                                         \
    norm_backward_kernel<T, V, false><<<blocks, T, 0, stream>>>(x, dy, gamma, eps, M, D, add, dx, pg, pb,     \
                                                                kNormBwdRows);
  switch (D) {
    case 512: B200_NB(64, 1) break;
    case 1024: B200_NB(128, 1) break;
    case 4096: B200_NB(256, 2) break;
    default: return fail(-2, "norm_backward: hidden size %d not supported (512 / 1024 / 4096)", D);
  }
#undef B200_NB
  if (dgamma != nullptr) norm_param_finish_kernel<<<(D + 255) / 256, 256, 0, stream>>>(pg, blocks, D, accumulate, dgamma);
  if (!rms && dbeta != nullptr)
    norm_param_finish_kernel<<<(D + 255) / 256, 256, 0, stream>>>(pb, blocks, D, accumulate, dbeta);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200
