// Attention kernels of the MM2SG hot path.
//
//  flash_attn_kernel<D>  fused softmax(Q K^T * scale + mask) V for the three "many queries" sites:
//      - CLIP ViT-L self-attention, S = 577, 16 heads x 64, no mask     (HF CLIPAttention via clip_encoder.py:48)
//      - BERT image-pooler self-attention, S = V*576, 8 heads x 128, key-padding mask
//                                                                       (multimodal_projector/builder.py:173)
//      - Llama prefill self-attention, 32 heads x 128, causal + left-pad mask
//                                                 (HF LlamaAttention via llava_llama.py:93; training uses the
//                                                  varlen FlashAttention-2 patch llama_flash_attn_monkey_patch.py:78)
//    The reference materialises the S x S score matrix in HBM; here scores never leave registers
//    (online softmax, fp32 statistics). Round-1 implementation: bf16 mma.sync.m16n8k16 with cp.async
//    double-buffered K/V tiles; the tcgen05 version is the planned replacement (DESIGN.md "next").
//
//  decode_attn_kernel    one new token per sequence against the bf16 KV cache (llava_arch.py:192-201 +
//      HF LlamaAttention with past_key_values). One query row per (sequence, head): a pure HBM-bound
//      stream over K and V with 16-byte vector loads; no tensor cores needed.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = smem_u32(smem);
  const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

static constexpr int kFaBM = 64;  // query rows per CTA (16 per warp)
static constexpr int kFaBN = 64;  // keys per tile
static constexpr int kFaThreads = 128;

// smem tile [rows][D] bf16 with the 16-byte chunk index XOR-swizzled by (row & 7): conflict-free ldmatrix
template <int D>
__device__ __forceinline__ bf16* tile_ptr(bf16* base, int row, int chunk) {
  return base + row * D + ((chunk ^ (row & 7)) << 3);
}

template <int D>
__device__ __forceinline__ void load_tile_async(bf16* smem, const bf16* g, long long row_stride, int row0,
                                                int rows_valid_end) {
  // 64 rows x D elements, 16 B per cp.async
  constexpr int kChunks = D / 8;
  for (int i = threadIdx.x; i < kFaBN * kChunks; i += kFaThreads) {
    const int r = i / kChunks, c = i % kChunks;
    const int gr = row0 + r;
    const bool ok = gr < rows_valid_end;
    const bf16* src = g + static_cast<long long>(ok ? gr : 0) * row_stride + c * 8;
    cp_async16(tile_ptr<D>(smem, r, c), src, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(kFaThreads) flash_attn_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t fa_smem[];
  bf16* sK = reinterpret_cast<bf16*>(fa_smem);             // [2][64][D]
  bf16* sV = sK + 2 * kFaBN * D;                           // [2][64][D]
  bf16* sQ = sV + 2 * kFaBN * D;                           // [64][D]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * kFaBM;
  const int h = blockIdx.y, b = blockIdx.z;

  const bf16* qg = a.q + b * a.q_bs + h * a.q_hs;
  const bf16* kg = a.k + b * a.k_bs + h * a.k_hs;
  const bf16* vg = a.v + b * a.v_bs + h * a.v_hs;

  int key_begin = a.kv_start ? a.kv_start[b] : 0;
  int key_end = a.kv_len ? min(a.kv_len[b], a.Lk) : a.Lk;
  const int causal_off = a.Lk - a.Lq;
  if (a.causal) key_end = min(key_end, q0 + kFaBM - 1 + causal_off + 1);
  const int tile_begin = key_begin / kFaBN;
  const int tile_end = (key_end + kFaBN - 1) / kFaBN;  // exclusive

  // ---- load Q tile, first K/V tile
  load_tile_async<D>(sQ, qg, a.q_rs, q0, a.Lq);
  if (tile_begin < tile_end) {
    load_tile_async<D>(sK, kg, a.k_rs, tile_begin * kFaBN, key_end);
    load_tile_async<D>(sV, vg, a.v_rs, tile_begin * kFaBN, key_end);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // Q fragments: 16 rows x D per warp
  uint32_t qf[D / 16][4];
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
    const int c = ks * 2 + (lane >> 4);
    ldmatrix_x4(qf[ks], tile_ptr<D>(sQ, r, c));
  }

  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int qi0 = q0 + warp * 16 + g, qi1 = qi0 + 8;

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int buf = (tile - tile_begin) & 1;
    bf16* cK = sK + buf * kFaBN * D;
    bf16* cV = sV + buf * kFaBN * D;
    // prefetch next tile into the other buffer
    if (tile + 1 < tile_end) {
      load_tile_async<D>(sK + (buf ^ 1) * kFaBN * D, kg, a.k_rs, (tile + 1) * kFaBN, key_end);
      load_tile_async<D>(sV + (buf ^ 1) * kFaBN * D, vg, a.v_rs, (tile + 1) * kFaBN, key_end);
    }
    cp_async_commit();

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[kFaBN / 8][4];
#pragma unroll
    for (int i = 0; i < kFaBN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < kFaBN / 16; ++np) {
        uint32_t kf[4];
        const int r = np * 16 + (lane & 7) + 8 * (lane >> 4);
        const int c = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(kf, tile_ptr<D>(cK, r, c));
        mma_bf16_16816(s[np * 2], qf[ks], kf[0], kf[1]);
        mma_bf16_16816(s[np * 2 + 1], qf[ks], kf[2], kf[3]);
      }
    }

    // ---- mask + online softmax
    const int kbase = tile * kFaBN;
    const bool need_mask = (kbase < key_begin) || (kbase + kFaBN > key_end) ||
                           (a.causal && (kbase + kFaBN - 1 > q0 + warp * 16 + causal_off));
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int i = 0; i < kFaBN / 8; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float val = s[i][j] * a.scale_log2;
        if (need_mask) {
          const int kj = kbase + i * 8 + 2 * t + (j & 1);
          const int qi = (j < 2) ? qi0 : qi1;
          const bool vis = kj >= key_begin && kj < key_end && (!a.causal || kj <= qi + causal_off);
          if (!vis) val = -INFINITY;
        }
        s[i][j] = val;
      }
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float ms0 = (mx0 == -INFINITY) ? 0.f : mx0;  // fully-masked rows must not produce NaN
    const float ms1 = (mx1 == -INFINITY) ? 0.f : mx1;
    const float corr0 = exp2f(m0 - ms0), corr1 = exp2f(m1 - ms1);  // m = -inf -> 0
    m0 = mx0;
    m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int i = 0; i < kFaBN / 8; ++i) {
      s[i][0] = exp2f(s[i][0] - ms0);
      s[i][1] = exp2f(s[i][1] - ms0);
      s[i][2] = exp2f(s[i][2] - ms1);
      s[i][3] = exp2f(s[i][3] - ms1);
      rs0 += s[i][0] + s[i][1];
      rs1 += s[i][2] + s[i][3];
    }
    l0 = l0 * corr0 + rs0;
    l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o[i][0] *= corr0;
      o[i][1] *= corr0;
      o[i][2] *= corr1;
      o[i][3] *= corr1;
    }

    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < kFaBN / 16; ++ks) {
      uint32_t pf[4];
      pf[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
      pf[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
      pf[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pf[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t vf[4];
        const int r = ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int c = np * 2 + (lane >> 4);
        ldmatrix_x4_trans(vf, tile_ptr<D>(cV, r, c));
        mma_bf16_16816(o[np * 2], pf, vf[0], vf[1]);
        mma_bf16_16816(o[np * 2 + 1], pf, vf[2], vf[3]);
      }
    }
    cp_async_wait<0>();
    __syncthreads();
  }

  // ---- finalize
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f;
  const float inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  bf16* og = a.o + b * a.o_bs + h * a.o_hs;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const int col = i * 8 + 2 * t;
    if (qi0 < a.Lq)
      *reinterpret_cast<uint32_t*>(og + qi0 * a.o_rs + col) = pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0);
    if (qi1 < a.Lq)
      *reinterpret_cast<uint32_t*>(og + qi1 * a.o_rs + col) = pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1);
  }
}

int flash_attn_tc(const AttnArgs& a, int head_dim, cudaStream_t stream);  // attention_sm100.cu

int flash_attn(const AttnArgs& a, int head_dim, cudaStream_t stream) {
  if (a.B <= 0 || a.H <= 0 || a.Lq <= 0) return 0;
  dim3 grid((a.Lq + kFaBM - 1) / kFaBM, a.H, a.B);
  const double qk = static_cast<double>(a.B) * a.H * a.Lq * a.Lk * (a.causal ? 0.5 : 1.0);
  LaunchScope scope(kFamFlashAttn, stream,
                    2.0 * a.B * a.H * head_dim * (2.0 * a.Lq + 2.0 * a.Lk), 4.0 * qk * head_dim);
  // the tcgen05 kernel is the product path; the mma.sync kernel below stays only as an A/B reference for bring-up
  static const bool legacy = getenv("B200_FA_LEGACY") != nullptr;
  if (!legacy) return flash_attn_tc(a, head_dim, stream);
  if (head_dim == 64) {
    constexpr int smem = (4 * kFaBN + kFaBM) * 64 * 2;
    static bool cfg = false;
    if (!cfg) {
      B200_CUDA_OK(cudaFuncSetAttribute(flash_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    flash_attn_kernel<64><<<grid, kFaThreads, smem, stream>>>(a);
  } else if (head_dim == 128) {
    constexpr int smem = (4 * kFaBN + kFaBM) * 128 * 2;
    static bool cfg = false;
    if (!cfg) {
      B200_CUDA_OK(cudaFuncSetAttribute(flash_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cfg = true;
    }
    flash_attn_kernel<128><<<grid, kFaThreads, smem, stream>>>(a);
  } else {
    return fail(-2, "flash_attn: head_dim %d not supported (64 or 128)", head_dim);
  }
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// decode attention: one query token per (sequence, head), head_dim 128, KV cache [B][H][ctx_cap][128]
// ---------------------------------------------------------------------------------------------------
static constexpr int kDecThreads = 256;
static constexpr int kDecWarps = kDecThreads / 32;
static constexpr int kDecMaxCtx = 8192;  // per split chunk, scores staged in smem

__global__ void __launch_bounds__(kDecThreads) decode_attn_kernel(const DecodeArgs a) {
  extern __shared__ float dec_smem[];
  float* sc = dec_smem;                       // [chunk] scores / probabilities
  __shared__ float red[kDecWarps];
  __shared__ float red_o[kDecWarps][128];
  __shared__ float q_s[128];

  const int bh = blockIdx.x;
  const int b = bh / a.H, h = bh % a.H;
  const int split = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int start = a.kv_start ? a.kv_start[b] : 0;
  const int ctx = a.ctx_dev ? min(*a.ctx_dev + a.ctx_add, a.cap) : a.ctx;
  const int total = max(0, ctx - start);
  const int per = (total + a.splits - 1) / a.splits;
  const int k0 = start + split * per;
  const int k1 = min(ctx, k0 + per);
  const int n = max(0, k1 - k0);

  if (threadIdx.x < 128) q_s[threadIdx.x] = __bfloat162float(a.q[b * a.q_rs + h * 128 + threadIdx.x]);
  __syncthreads();

  const bf16* kbase = a.kc + (static_cast<long long>(b) * a.H + h) * a.cap * 128;
  const bf16* vbase = a.vc + (static_cast<long long>(b) * a.H + h) * a.cap * 128;

  // ---- phase 1: scores. 8 lanes per key (16 dims = 2 x 16 B each), 4 keys per warp-iteration
  const int sub = lane & 7, kq = lane >> 3;
  float qreg[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) qreg[j] = q_s[sub * 16 + j];
  float lmax = -INFINITY;
  for (int base = warp * 4; base < n; base += kDecWarps * 4) {  // warp-uniform trip count (shuffles inside)
    const int i = base + kq;
    const bool ok = i < n;
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
  for (int i = warp * 2 + vk; i < n; i += kDecWarps * 2) {
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
  static bool cfg = false;
  if (!cfg) {
    B200_CUDA_OK(cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kDecMaxCtx * (int)sizeof(float)));
    cfg = true;
  }
  LaunchScope scope(kFamDecodeAttn, stream, 0.0, 0.0, a.splits > 1 ? 2 : 1);  // bytes depend on the device-side ctx
  decode_attn_kernel<<<dim3(a.B * a.H, a.splits), kDecThreads, smem, stream>>>(a);
  B200_CUDA_OK(cudaGetLastError());
  if (a.splits > 1) {
    decode_combine_kernel<<<a.B * a.H, 128, 0, stream>>>(a);
    B200_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace b200
