// tcgen05 flash attention for sm_100a:  O = softmax(Q K^T * scale + mask) V, scores never leave the SM.
//
// The three "many queries" attention sites of the MM2SG hot path:
//   - CLIP ViT-L self-attention, S = 577, 16 heads x 64, no mask           (HF CLIPAttention via clip_encoder.py:48)
//   - BERT image-pooler self-attention, S = V*576, 8 heads x 128, key-padding mask
//                                                                          (multimodal_projector/builder.py:173)
//   - Llama prefill self-attention, 32 heads x 128, causal + left-pad / right-pad mask
//         (HF LlamaAttention via llava_llama.py:93; the training path's varlen FlashAttention-2 patch
//          train/llama_flash_attn_monkey_patch.py:78-89 is the same math on right-padded batches)
// The reference materialises the S x S score matrix in HBM (eager bmm -> softmax(fp32) -> bmm).
//
// One CTA = one (sample, head, 128-query tile); six warps:
//   warp 0      TMA producer: Q tile once, then K / V tiles of 64 keys into a 2..3-stage ring (4-D tensor maps over
//               (d, token, head, sample) so any row / head / batch stride works; 128B swizzle; OOB rows read zero)
//   warp 1      MMA issuer: S[128 x 64] = Q K^T (tcgen05.mma, both operands K-major) into TMEM columns [0, 64);
//               O[128 x D] += P V with P (bf16, written by the softmax warps into swizzled smem) as the K-major A
//               operand and the V tile as an MN-major B operand; O lives in TMEM columns [64, 64 + D).
//               QK^T of tile j+1 is issued as soon as the softmax warps have pulled S(j) into registers, so the
//               tensor pipe works on it while the SFUs do the exponentials of tile j.
//   warps 2..5  softmax: one query row per thread (tcgen05.ld of its TMEM lane), fp32 running max / sum in the
//               log2 domain (exp2f, scale folded in). The running max is only raised when the tile max exceeds it
//               by more than 2^8 (FlashAttention-4's lazy rescale): P stays <= 256 in bf16, and the O accumulator
//               in TMEM is rescaled (tcgen05.ld -> mul -> tcgen05.st) only on those rare tiles, per warp.
//               Fully masked rows (left-padding queries) produce exact zeros, never NaN.
// Two CTAs are resident per SM (<= 113 KB smem, <= 256 TMEM columns each), so one CTA's exponentials overlap the
// other's MMAs as well.
#include "attention_common.cuh"

namespace b200 {

static constexpr int kTcBM = 128;  // query rows per CTA
static constexpr int kTcBN = 64;   // keys per tile
static constexpr int kTcThreads = 192;
static constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct TcAttnArgs {
  bf16* o;
  long long o_bs, o_rs, o_hs;
  int B, H, Lq, Lk;
  const int* kv_start;
  const int* kv_len;
  int causal;
  float scale_log2;
  float* lse;
  // tensor-map dimension (1..3) that carries the token / head / sample index, per tensor (dims are sorted by stride)
  int q_dim[3], k_dim[3], v_dim[3];
};

template <int D>
struct TcCfg {
  static constexpr int kPanels = D / 64;
  static constexpr int kStages = (D == 64) ? 3 : 2;
  static constexpr int kQBytes = kTcBM * D * 2;
  static constexpr int kKBytes = kTcBN * D * 2;
  static constexpr int kVBytes = kTcBN * D * 2;
  // P (128 x 64 bf16 = 16 KB): for D = 128 it overwrites the K tile it was computed from (same size, dead once
  // S = Q K^T has completed), which keeps two CTAs per SM; for D = 64 it has its own buffer
  static constexpr bool kPAliasesK = (kKBytes == kTcBM * kTcBN * 2);
  static constexpr int kPBytes = kPAliasesK ? 0 : kTcBM * kTcBN * 2;
  static constexpr int kTmemCols = (kTcBN + D <= 128) ? 128 : 256;
  static constexpr int kSmemBytes = kQBytes + kStages * (kKBytes + kVBytes) + kPBytes + 1024 + 256;
};

template <int D>
__global__ void __launch_bounds__(kTcThreads, 2)
    flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const TcAttnArgs a) {
  using Cfg = TcCfg<D>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kPanels = Cfg::kPanels;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                   // kPanels x [128 rows][128 B]
  uint8_t* sK = sQ + Cfg::kQBytes;                      // stages x kPanels x [64 rows][128 B]
  uint8_t* sV = sK + kStages * Cfg::kKBytes;            // stages x kPanels x [64 keys][128 B]
  uint8_t* sP0 = sV + kStages * Cfg::kVBytes;           // [128 rows][128 B] (D = 64 only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP0 + Cfg::kPBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kStages;
  uint64_t* s_full = kv_empty + kStages;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;
  uint64_t* p_free = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = a.causal ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);  // heavy first
  const int q0 = qt * kTcBM;
  const int h = blockIdx.y, b = blockIdx.z;

  int key_begin = a.kv_start ? a.kv_start[b] : 0;
  int key_end = a.kv_len ? min(a.kv_len[b], a.Lk) : a.Lk;
  const int causal_off = a.Lk - a.Lq;
  if (a.causal) key_end = min(key_end, q0 + kTcBM + causal_off);
  const int tile_begin = key_begin / kTcBN;
  const int tile_end = key_end > key_begin ? (key_end + kTcBN - 1) / kTcBN : tile_begin;
  const int n_tiles = tile_end - tile_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(p_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;           // columns [0, 64)
  const uint32_t tmem_O = tmem_base + kTcBN;   // columns [64, 64 + D)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(q_full, Cfg::kQBytes);
      for (int p = 0; p < kPanels; ++p) tc_load(sQ + p * (kTcBM * 128), &tmQ, q_full, a.q_dim, p * 64, q0, h, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = tile_begin; t < tile_end; ++t) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        mbar_expect_tx(&kv_full[stage], Cfg::kKBytes + Cfg::kVBytes);
        for (int p = 0; p < kPanels; ++p) {
          tc_load(sK + stage * Cfg::kKBytes + p * (kTcBN * 128), &tmK, &kv_full[stage], a.k_dim, p * 64, t * kTcBN, h,
                  b);
          tc_load(sV + stage * Cfg::kVBytes + p * (kTcBN * 128), &tmV, &kv_full[stage], a.v_dim, p * 64, t * kTcBN, h,
                  b);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kTcBM, kTcBN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kTcBM, D, 0, 1);  // B (= V) is MN-major
      mbar_wait(q_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      int prev_stage = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_full[stage], phase);
        if (j > 0) mbar_wait(s_free, (j - 1) & 1u);  // softmax has pulled S(j-1) out of TMEM
        tc_fence_after_sync();
#pragma unroll
        for (int p = 0; p < kPanels; ++p) {
          const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ + p * (kTcBM * 128)));
          const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(sK + stage * Cfg::kKBytes + p * (kTcBN * 128)));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_qk, (p | k) != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        if (j > 0) {
          // O += P(j-1) V(j-1)
          mbar_wait(p_full, (j - 1) & 1u);
          tc_fence_after_sync();
          const uint64_t dp =
              umma_desc_kmajor_sw128(smem_u32(Cfg::kPAliasesK ? sK + prev_stage * Cfg::kKBytes : sP0));
          const uint64_t dv = umma_desc_mnmajor_sw128(smem_u32(sV + prev_stage * Cfg::kVBytes), kTcBN * 128);
#pragma unroll
          for (int k = 0; k < kTcBN / 16; ++k)
            umma_bf16_ss(tmem_O, dp + 2 * k, dv + 128 * k, idesc_pv, (j > 1 || k != 0) ? 1u : 0u);
          umma_commit(&kv_empty[prev_stage]);
          umma_commit(p_free);
        }
        prev_stage = stage;
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      {
        const int j = n_tiles;  // tail: PV of the last tile
        mbar_wait(p_full, (j - 1) & 1u);
        tc_fence_after_sync();
        const uint64_t dp =
            umma_desc_kmajor_sw128(smem_u32(Cfg::kPAliasesK ? sK + prev_stage * Cfg::kKBytes : sP0));
        const uint64_t dv = umma_desc_mnmajor_sw128(smem_u32(sV + prev_stage * Cfg::kVBytes), kTcBN * 128);
#pragma unroll
        for (int k = 0; k < kTcBN / 16; ++k)
          umma_bf16_ss(tmem_O, dp + 2 * k, dv + 128 * k, idesc_pv, (j > 1 || k != 0) ? 1u : 0u);
        umma_commit(&kv_empty[prev_stage]);
        umma_commit(p_free);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 2..5) =====================
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // query row inside the tile = TMEM lane
    const int qi = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1u);
      tc_fence_after_sync();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_S + lane_off, v0);
      tmem_ld_32x32(tmem_S + lane_off + 32, v1);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);

      const int kbase = (tile_begin + j) * kTcBN;
      const bool need_mask = (kbase < key_begin) || (kbase + kTcBN > key_end) ||
                             (a.causal && (kbase + kTcBN - 1 > q0 + qd * 32 + causal_off));
      // scores stay unscaled in registers; the softmax scale is folded into one FFMA per element:
      // p = 2^(s * scale_log2 - m)
      float s[64];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        s[c] = __uint_as_float(v0[c]);
        s[32 + c] = __uint_as_float(v1[c]);
      }
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const int kj = kbase + c;
          const bool vis = kj >= key_begin && kj < key_end && (!a.causal || kj <= qi + causal_off);
          if (!vis) s[c] = -INFINITY;
        }
      }
      // four independent chains (a 64-deep dependent FMNMX chain would be latency bound)
      float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]), mx2 = fmaxf(s[4], s[5]), mx3 = fmaxf(s[6], s[7]);
#pragma unroll
      for (int c = 8; c < 64; c += 8) {
        mx0 = fmaxf(mx0, fmaxf(s[c], s[c + 1]));
        mx1 = fmaxf(mx1, fmaxf(s[c + 2], s[c + 3]));
        mx2 = fmaxf(mx2, fmaxf(s[c + 4], s[c + 5]));
        mx3 = fmaxf(mx3, fmaxf(s[c + 6], s[c + 7]));
      }
      const float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * a.scale_log2;  // scale > 0: max commutes
      // lazy rescale: raise the running max only when it would otherwise let P exceed 2^8
      const bool raise = tmax > m_run + kRescaleThreshold;  // also true for the first visible key (m_run = -inf)
      const float m_new = raise ? tmax : m_run;
      const float msafe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = raise ? fast_exp2(m_run - msafe) : 1.f;  // m_run = -inf -> 0
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float p0 = fast_exp2(fmaf(s[2 * c], a.scale_log2, -msafe));
        const float p1 = fast_exp2(fmaf(s[2 * c + 1], a.scale_log2, -msafe));
        const float p2 = fast_exp2(fmaf(s[2 * c + 2], a.scale_log2, -msafe));
        const float p3 = fast_exp2(fmaf(s[2 * c + 3], a.scale_log2, -msafe));
        rs0 += p0;
        rs1 += p1;
        rs2 += p2;
        rs3 += p3;
        pk[c] = pack_bf16x2(p0, p1);
        pk[c + 1] = pack_bf16x2(p2, p3);
      }
      l_run = l_run * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m_run = m_new;

      if (j > 0) {
        mbar_wait(p_free, (j - 1) & 1u);  // PV(j-1) done: sP reusable, O stable
        tc_fence_after_sync();
        if (__any_sync(0xffffffffu, raise)) {
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32(tmem_O + lane_off + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32(tmem_O + lane_off + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      // P row -> swizzled smem (K-major, 128 B per row, 16-byte chunk index XOR (row & 7))
      uint8_t* prow = (Cfg::kPAliasesK ? sK + (j % kStages) * Cfg::kKBytes : sP0) + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) =
            make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 (+ the row's log-sum-exp for the backward pass)
    if (a.lse != nullptr && qi < a.Lq)
      a.lse[(static_cast<long long>(b) * a.H + h) * a.Lq + qi] =
          l_run > 0.f ? (m_run + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
    bf16* orow = a.o + b * a.o_bs + static_cast<long long>(qi) * a.o_rs + h * a.o_hs;
    if (n_tiles > 0) {
      mbar_wait(p_free, (n_tiles - 1) & 1u);
      tc_fence_after_sync();
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32(tmem_O + lane_off + c * 32, ov);
        tmem_ld_wait();
        if (qi < a.Lq) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(ov[8 * i + 0]) * inv, __uint_as_float(ov[8 * i + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(ov[8 * i + 2]) * inv, __uint_as_float(ov[8 * i + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(ov[8 * i + 4]) * inv, __uint_as_float(ov[8 * i + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(ov[8 * i + 6]) * inv, __uint_as_float(ov[8 * i + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 32 + 8 * i) = o;
          }
        }
      }
    } else if (qi < a.Lq) {
      for (int c = 0; c < D / 8; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(0, 0, 0, 0);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <int D>
static int launch_tc(const AttnArgs& a, cudaStream_t stream) {
  using Cfg = TcCfg<D>;
  static bool cfg = false;
  if (!cfg) {
    B200_CUDA_OK(cudaFuncSetAttribute(flash_attn_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
    cfg = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  TcAttnArgs t;
  B200_TRY(make_tmap_attn(&tmQ, a.q, D, a.Lq, a.H, a.B, a.q_rs, a.q_hs, a.q_bs, kTcBM, t.q_dim));
  B200_TRY(make_tmap_attn(&tmK, a.k, D, a.Lk, a.H, a.B, a.k_rs, a.k_hs, a.k_bs, kTcBN, t.k_dim));
  B200_TRY(make_tmap_attn(&tmV, a.v, D, a.Lk, a.H, a.B, a.v_rs, a.v_hs, a.v_bs, kTcBN, t.v_dim));
  t.o = a.o;
  t.o_bs = a.o_bs;
  t.o_rs = a.o_rs;
  t.o_hs = a.o_hs;
  t.B = a.B;
  t.H = a.H;
  t.Lq = a.Lq;
  t.Lk = a.Lk;
  t.kv_start = a.kv_start;
  t.kv_len = a.kv_len;
  t.causal = a.causal;
  t.scale_log2 = a.scale_log2;
  t.lse = a.lse;
  dim3 grid((a.Lq + kTcBM - 1) / kTcBM, a.H, a.B);
  flash_attn_tc_kernel<D><<<grid, kTcThreads, Cfg::kSmemBytes, stream>>>(tmQ, tmK, tmV, t);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

int flash_attn_tc(const AttnArgs& a, int head_dim, cudaStream_t stream) {
  if ((a.o_rs % 8) != 0 || (a.o_hs % 8) != 0 || (a.o_bs % 8) != 0 || (reinterpret_cast<uintptr_t>(a.o) & 15) != 0)
    return fail(-2, "flash_attn: output pointer and strides must be 16-byte aligned");
  if (head_dim == 64) return launch_tc<64>(a, stream);
  if (head_dim == 128) return launch_tc<128>(a, stream);
  return fail(-2, "flash_attn: head_dim %d not supported (64 or 128)", head_dim);
}

}  // namespace b200
