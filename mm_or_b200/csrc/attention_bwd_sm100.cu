// tcgen05 flash-attention BACKWARD for sm_100a (training path).
//
// Reference: the fine-tune step runs FlashAttention-2's varlen backward behind
// train/llama_flash_attn_monkey_patch.py:78-89 (flash_attn_varlen_qkvpacked_func under autograd) for the Llama layers
// and torch autograd through the eager bmm / softmax / bmm of HF CLIPAttention and BertSelfAttention for the trainable
// CLIP layers and the image pooler. Same math everywhere, with P recomputed from the saved row log-sum-exp:
//     P  = exp(S * scale - LSE)          S = Q K^T
//     dV = P^T dO          dP = dO V^T          Delta_i = sum_d dO_id O_id
//     dS = P o (dP - Delta) * scale      dQ = dS K            dK = dS^T Q
// Two kernels, both built from the operand patterns of the forward kernel (attention_sm100.cu): K-major operands
// straight from TMA tiles, thread-written 128B-swizzled bf16 tiles as K-major A operands, TMA tiles re-read as
// MN-major B operands, fp32 accumulators in TMEM. No atomics: every output element is owned by exactly one CTA, so
// gradients are bit-deterministic (FlashAttention-2 accumulates dQ with atomicAdd).
//
//   flash_bwd_dkdv_kernel   one CTA per (sample, head, 128-key tile); loops over 64-query tiles. Works on the
//       TRANSPOSED score tile so that a thread owns a KEY row:  S^T = K Q^T and dP^T = V dO^T (M = 128 keys, N = 64
//       queries) land in TMEM; the compute warps form P^T and dS^T row-wise, park them in shared memory as bf16, and
//       dV += P^T dO, dK += dS^T Q are accumulated in TMEM over the whole query loop ([128 x D] fp32 each).
//   flash_bwd_dq_kernel     one CTA per (sample, head, 128-query tile); loops over 64-key tiles like the forward:
//       S and dP in TMEM, a thread owns a QUERY row (its LSE / Delta are scalars), dQ += dS K accumulates in TMEM.
//   attn_delta_kernel       Delta = rowsum(dO o O), one warp per (sample, head, query).
// Recomputing S and dP in both kernels costs 7 tile MMAs per (query tile, key tile) instead of FA-2's 5, and buys
// determinism and a TMEM budget that fits (384 / 256 columns).
#include "attention_common.cuh"

namespace b200 {

static constexpr int kBwThreads = 192;
static constexpr float kLog2e = 1.4426950408889634f;

struct BwdArgs {
  // outputs (same stride convention as the inputs: batch, row, head; unit stride on d)
  bf16* dq;
  bf16* dk;
  bf16* dv;
  long long dq_bs, dq_rs, dq_hs;
  long long dk_bs, dk_rs, dk_hs;
  long long dv_bs, dv_rs, dv_hs;
  const float* lse;    // [B, H, Lq] natural log (forward output)
  const float* delta;  // [B, H, Lq]
  int B, H, Lq, Lk;
  const int* kv_start;
  const int* kv_len;
  int causal;
  float scale, scale_log2;
  int q_dim[3], k_dim[3], v_dim[3], do_dim[3];
};

__device__ __forceinline__ void bw_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// write 64 bf16 (packed in 32 regs) as row r of a [128][64] K-major 128B-swizzled tile
__device__ __forceinline__ void store_row_sw128(uint8_t* tile, int r, const uint32_t (&pk)[32]) {
  uint8_t* row = tile + r * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

// ---------------------------------------------------------------------------------------------------
// dK / dV
// ---------------------------------------------------------------------------------------------------
template <int D>
struct DkvCfg {
  static constexpr int kPanels = D / 64;
  static constexpr int kBKeys = 128, kBQ = 64, kStages = 2;
  static constexpr int kKBytes = kBKeys * D * 2;  // K tile = V tile
  static constexpr int kQBytes = kBQ * D * 2;     // Q tile = dO tile (per stage)
  static constexpr int kPBytes = kBKeys * kBQ * 2;
  static constexpr int kTmemCols = (128 + 2 * D <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = 2 * kKBytes + kStages * 2 * kQBytes + 2 * kPBytes + 2 * 2 * kBQ * 4 + 1024 + 256;
};

template <int D>
__global__ void __launch_bounds__(kBwThreads, 1)
    flash_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                          const BwdArgs a) {
  using Cfg = DkvCfg<D>;
  constexpr int kPanels = Cfg::kPanels, kStages = Cfg::kStages, kBQ = Cfg::kBQ, kBKeys = Cfg::kBKeys;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sK = smem;                                   // kPanels x [128 keys][128 B]
  uint8_t* sV = sK + Cfg::kKBytes;
  uint8_t* sQ = sV + Cfg::kKBytes;                      // stages x kPanels x [64 q][128 B]
  uint8_t* sDO = sQ + kStages * Cfg::kQBytes;
  uint8_t* sP = sDO + kStages * Cfg::kQBytes;           // [128 keys][64 q] bf16 (P^T)
  uint8_t* sDS = sP + Cfg::kPBytes;                     // [128 keys][64 q] bf16 (dS^T)
  float* sLD = reinterpret_cast<float*>(sDS + Cfg::kPBytes);  // [2 buffers][lse2 | delta][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLD + 2 * 2 * kBQ);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;
  uint64_t* qdo_empty = qdo_full + kStages;
  uint64_t* st_full = qdo_empty + kStages;
  uint64_t* st_free = st_full + 1;
  uint64_t* pds_full = st_free + 1;
  uint64_t* pds_free = pds_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pds_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kBKeys;
  const int h = blockIdx.y, b = blockIdx.z;
  const int key_begin = a.kv_start ? a.kv_start[b] : 0;
  const int key_end = a.kv_len ? min(a.kv_len[b], a.Lk) : a.Lk;
  const int causal_off = a.Lk - a.Lq;
  // query tiles that can see at least one key of this tile
  int q_lo = 0;
  if (a.causal) q_lo = max(0, k0 - causal_off) / kBQ;
  const bool any_key = (k0 < key_end) && (k0 + kBKeys > key_begin);
  const int q_tiles_all = (a.Lq + kBQ - 1) / kBQ;
  const int n_tiles = any_key ? max(0, q_tiles_all - q_lo) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(st_full, 1);
    mbar_init(st_free, 4);
    mbar_init(pds_full, 4);
    mbar_init(pds_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDV = tmem_base + 128, tDK = tmem_base + 128 + D;

  if (warp == 0) {
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(kv_full, 2 * Cfg::kKBytes);
      for (int p = 0; p < kPanels; ++p) {
        tc_load(sK + p * (kBKeys * 128), &tmK, kv_full, a.k_dim, p * 64, k0, h, b);
        tc_load(sV + p * (kBKeys * 128), &tmV, kv_full, a.v_dim, p * 64, k0, h, b);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_tiles; ++i) {
        const int q0 = (q_lo + i) * kBQ;
        mbar_wait(&qdo_empty[stage], phase ^ 1u);
        mbar_expect_tx(&qdo_full[stage], 2 * Cfg::kQBytes);
        for (int p = 0; p < kPanels; ++p) {
          tc_load(sQ + stage * Cfg::kQBytes + p * (kBQ * 128), &tmQ, &qdo_full[stage], a.q_dim, p * 64, q0, h, b);
          tc_load(sDO + stage * Cfg::kQBytes + p * (kBQ * 128), &tmDO, &qdo_full[stage], a.do_dim, p * 64, q0, h, b);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_st = umma_idesc_bf16(kBKeys, kBQ, 0, 0);  // S^T, dP^T: [128 keys x 64 q]
      constexpr uint32_t idesc_acc = umma_idesc_bf16(kBKeys, D, 0, 1);   // dV, dK: A = P^T / dS^T, B = dO / Q (MN-major)
      mbar_wait(kv_full, 0);
      int stage = 0, prev_stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i <= n_tiles; ++i) {
        if (i < n_tiles) {
          mbar_wait(&qdo_full[stage], phase);
          if (i > 0) mbar_wait(st_free, (i - 1) & 1u);
          tc_fence_after_sync();
#pragma unroll
          for (int p = 0; p < kPanels; ++p) {
            const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(sK + p * (kBKeys * 128)));
            const uint64_t dv = umma_desc_kmajor_sw128(smem_u32(sV + p * (kBKeys * 128)));
            const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ + stage * Cfg::kQBytes + p * (kBQ * 128)));
            const uint64_t dd = umma_desc_kmajor_sw128(smem_u32(sDO + stage * Cfg::kQBytes + p * (kBQ * 128)));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ss(tS, dk + 2 * k, dq + 2 * k, idesc_st, (p | k) != 0 ? 1u : 0u);
              umma_bf16_ss(tDP, dv + 2 * k, dd + 2 * k, idesc_st, (p | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(st_full);
        }
        if (i > 0) {
          const int j = i - 1;  // accumulate tile j
          mbar_wait(pds_full, j & 1u);
          tc_fence_after_sync();
          const uint64_t dp = umma_desc_kmajor_sw128(smem_u32(sP));
          const uint64_t ds = umma_desc_kmajor_sw128(smem_u32(sDS));
          const uint64_t ddo = umma_desc_mnmajor_sw128(smem_u32(sDO + prev_stage * Cfg::kQBytes), kBQ * 128);
          const uint64_t dqq = umma_desc_mnmajor_sw128(smem_u32(sQ + prev_stage * Cfg::kQBytes), kBQ * 128);
#pragma unroll
          for (int k = 0; k < kBQ / 16; ++k) {
            umma_bf16_ss(tDV, dp + 2 * k, ddo + 128 * k, idesc_acc, (j > 0 || k != 0) ? 1u : 0u);
            umma_bf16_ss(tDK, ds + 2 * k, dqq + 128 * k, idesc_acc, (j > 0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&qdo_empty[prev_stage]);
          umma_commit(pds_free);
        }
        if (i < n_tiles) {
          prev_stage = stage;
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // key row inside the tile = TMEM lane
    const int kj = k0 + r;
    const int t = threadIdx.x - 64;  // 0..127
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const bool key_ok = kj >= key_begin && kj < key_end;
    const float* lse_bh = a.lse + (static_cast<long long>(b) * a.H + h) * a.Lq;
    const float* del_bh = a.delta + (static_cast<long long>(b) * a.H + h) * a.Lq;
    for (int i = 0; i < n_tiles; ++i) {
      const int q0 = (q_lo + i) * kBQ;
      float* ld = sLD + (i & 1) * (2 * kBQ);
      if (t < kBQ) {
        const int qi = q0 + t;
        float l2 = INFINITY, dl = 0.f;  // +inf: P = exp2(s - inf) = 0 for rows that do not exist / saw no key
        if (qi < a.Lq) {
          const float l = lse_bh[qi];
          if (l > -INFINITY) l2 = l * kLog2e;
          dl = del_bh[qi];
        }
        ld[t] = l2;
        ld[kBQ + t] = dl;
      }
      bw_bar_sync();
      mbar_wait(st_full, i & 1u);
      tc_fence_after_sync();
      uint32_t sv[2][32], dv[2][32];
      tmem_ld_32x32(tS + lane_off, sv[0]);
      tmem_ld_32x32(tS + lane_off + 32, sv[1]);
      tmem_ld_32x32(tDP + lane_off, dv[0]);
      tmem_ld_32x32(tDP + lane_off + 32, dv[1]);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(st_free);
      uint32_t pk[32], dk[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float p[2], g[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = hf * 32 + c + e;
            const int qi = q0 + col;
            const bool vis = key_ok && (!a.causal || kj <= qi + causal_off);
            const float pe = vis ? fast_exp2(fmaf(__uint_as_float(sv[hf][c + e]), a.scale_log2, -ld[col])) : 0.f;
            p[e] = pe;
            g[e] = pe * (__uint_as_float(dv[hf][c + e]) - ld[kBQ + col]) * a.scale;
          }
          pk[hf * 16 + c / 2] = pack_bf16x2(p[0], p[1]);
          dk[hf * 16 + c / 2] = pack_bf16x2(g[0], g[1]);
        }
      if (i > 0) {
        mbar_wait(pds_free, (i - 1) & 1u);
        tc_fence_after_sync();
      }
      store_row_sw128(sP, r, pk);
      store_row_sw128(sDS, r, dk);
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    // ---- epilogue
    bf16* dkrow = a.dk + b * a.dk_bs + static_cast<long long>(kj) * a.dk_rs + h * a.dk_hs;
    bf16* dvrow = a.dv + b * a.dv_bs + static_cast<long long>(kj) * a.dv_rs + h * a.dv_hs;
    if (n_tiles > 0) {
      mbar_wait(pds_free, (n_tiles - 1) & 1u);
      tc_fence_after_sync();
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        bf16* orow = which ? dkrow : dvrow;
        const uint32_t tacc = which ? tDK : tDV;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32(tacc + lane_off + c * 32, ov);
          tmem_ld_wait();
          if (kj < a.Lk) {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(ov[8 * i4 + 0]), __uint_as_float(ov[8 * i4 + 1]));
              o.y = pack_bf16x2(__uint_as_float(ov[8 * i4 + 2]), __uint_as_float(ov[8 * i4 + 3]));
              o.z = pack_bf16x2(__uint_as_float(ov[8 * i4 + 4]), __uint_as_float(ov[8 * i4 + 5]));
              o.w = pack_bf16x2(__uint_as_float(ov[8 * i4 + 6]), __uint_as_float(ov[8 * i4 + 7]));
              *reinterpret_cast<uint4*>(orow + c * 32 + 8 * i4) = o;
            }
          }
        }
      }
    } else if (kj < a.Lk) {
      for (int c = 0; c < D / 8; ++c) {
        *reinterpret_cast<uint4*>(dkrow + c * 8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dvrow + c * 8) = make_uint4(0, 0, 0, 0);
      }
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
// dQ
// ---------------------------------------------------------------------------------------------------
template <int D>
struct DqCfg {
  static constexpr int kPanels = D / 64;
  static constexpr int kBQ = 128, kBKeys = 64, kStages = 2;
  static constexpr int kQBytes = kBQ * D * 2;     // Q tile = dO tile
  static constexpr int kKBytes = kBKeys * D * 2;  // K tile = V tile (per stage)
  static constexpr int kDSBytes = kBQ * kBKeys * 2;
  static constexpr int kTmemCols = (128 + D <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = 2 * kQBytes + kStages * 2 * kKBytes + kDSBytes + 1024 + 256;
};

template <int D>
__global__ void __launch_bounds__(kBwThreads, 1)
    flash_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const BwdArgs a) {
  using Cfg = DqCfg<D>;
  constexpr int kPanels = Cfg::kPanels, kStages = Cfg::kStages, kBQ = Cfg::kBQ, kBKeys = Cfg::kBKeys;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                   // kPanels x [128 q][128 B]
  uint8_t* sDO = sQ + Cfg::kQBytes;
  uint8_t* sK = sDO + Cfg::kQBytes;                     // stages x kPanels x [64 keys][128 B]
  uint8_t* sV = sK + kStages * Cfg::kKBytes;
  uint8_t* sDS = sV + kStages * Cfg::kKBytes;           // [128 q][64 keys] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + Cfg::kDSBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kStages;
  uint64_t* st_full = kv_empty + kStages;
  uint64_t* st_free = st_full + 1;
  uint64_t* ds_full = st_free + 1;
  uint64_t* ds_free = ds_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ds_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = a.causal ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);
  const int q0 = qt * kBQ;
  const int h = blockIdx.y, b = blockIdx.z;
  int key_begin = a.kv_start ? a.kv_start[b] : 0;
  int key_end = a.kv_len ? min(a.kv_len[b], a.Lk) : a.Lk;
  const int causal_off = a.Lk - a.Lq;
  if (a.causal) key_end = min(key_end, q0 + kBQ + causal_off);
  const int tile_begin = key_begin / kBKeys;
  const int tile_end = key_end > key_begin ? (key_end + kBKeys - 1) / kBKeys : tile_begin;
  const int n_tiles = tile_end - tile_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(st_full, 1);
    mbar_init(st_free, 4);
    mbar_init(ds_full, 4);
    mbar_init(ds_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDQ = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(q_full, 2 * Cfg::kQBytes);
      for (int p = 0; p < kPanels; ++p) {
        tc_load(sQ + p * (kBQ * 128), &tmQ, q_full, a.q_dim, p * 64, q0, h, b);
        tc_load(sDO + p * (kBQ * 128), &tmDO, q_full, a.do_dim, p * 64, q0, h, b);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int t = tile_begin; t < tile_end; ++t) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        mbar_expect_tx(&kv_full[stage], 2 * Cfg::kKBytes);
        for (int p = 0; p < kPanels; ++p) {
          tc_load(sK + stage * Cfg::kKBytes + p * (kBKeys * 128), &tmK, &kv_full[stage], a.k_dim, p * 64, t * kBKeys, h, b);
          tc_load(sV + stage * Cfg::kKBytes + p * (kBKeys * 128), &tmV, &kv_full[stage], a.v_dim, p * 64, t * kBKeys, h, b);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kBQ, kBKeys, 0, 0);
      constexpr uint32_t idesc_dq = umma_idesc_bf16(kBQ, D, 0, 1);  // B = K tile read MN-major
      mbar_wait(q_full, 0);
      int stage = 0, prev_stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j <= n_tiles; ++j) {
        if (j < n_tiles) {
          mbar_wait(&kv_full[stage], phase);
          if (j > 0) mbar_wait(st_free, (j - 1) & 1u);
          tc_fence_after_sync();
#pragma unroll
          for (int p = 0; p < kPanels; ++p) {
            const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ + p * (kBQ * 128)));
            const uint64_t dd = umma_desc_kmajor_sw128(smem_u32(sDO + p * (kBQ * 128)));
            const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(sK + stage * Cfg::kKBytes + p * (kBKeys * 128)));
            const uint64_t dv = umma_desc_kmajor_sw128(smem_u32(sV + stage * Cfg::kKBytes + p * (kBKeys * 128)));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ss(tS, dq + 2 * k, dk + 2 * k, idesc_s, (p | k) != 0 ? 1u : 0u);
              umma_bf16_ss(tDP, dd + 2 * k, dv + 2 * k, idesc_s, (p | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(st_full);
        }
        if (j > 0) {
          const int jj = j - 1;
          mbar_wait(ds_full, jj & 1u);
          tc_fence_after_sync();
          const uint64_t ds = umma_desc_kmajor_sw128(smem_u32(sDS));
          const uint64_t dkm = umma_desc_mnmajor_sw128(smem_u32(sK + prev_stage * Cfg::kKBytes), kBKeys * 128);
#pragma unroll
          for (int k = 0; k < kBKeys / 16; ++k)
            umma_bf16_ss(tDQ, ds + 2 * k, dkm + 128 * k, idesc_dq, (jj > 0 || k != 0) ? 1u : 0u);
          umma_commit(&kv_empty[prev_stage]);
          umma_commit(ds_free);
        }
        if (j < n_tiles) {
          prev_stage = stage;
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    float l2 = INFINITY, dl = 0.f;
    if (qi < a.Lq) {
      const long long idx = (static_cast<long long>(b) * a.H + h) * a.Lq + qi;
      const float l = a.lse[idx];
      if (l > -INFINITY) l2 = l * kLog2e;
      dl = a.delta[idx];
    }
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(st_full, j & 1u);
      tc_fence_after_sync();
      uint32_t sv[2][32], dv[2][32];
      tmem_ld_32x32(tS + lane_off, sv[0]);
      tmem_ld_32x32(tS + lane_off + 32, sv[1]);
      tmem_ld_32x32(tDP + lane_off, dv[0]);
      tmem_ld_32x32(tDP + lane_off + 32, dv[1]);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(st_free);
      const int kbase = (tile_begin + j) * kBKeys;
      uint32_t dk[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float g[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int kj = kbase + hf * 32 + c + e;
            const bool vis = kj >= key_begin && kj < key_end && (!a.causal || kj <= qi + causal_off);
            const float pe = vis ? fast_exp2(fmaf(__uint_as_float(sv[hf][c + e]), a.scale_log2, -l2)) : 0.f;
            g[e] = pe * (__uint_as_float(dv[hf][c + e]) - dl) * a.scale;
          }
          dk[hf * 16 + c / 2] = pack_bf16x2(g[0], g[1]);
        }
      if (j > 0) {
        mbar_wait(ds_free, (j - 1) & 1u);
        tc_fence_after_sync();
      }
      store_row_sw128(sDS, r, dk);
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    bf16* orow = a.dq + b * a.dq_bs + static_cast<long long>(qi) * a.dq_rs + h * a.dq_hs;
    if (n_tiles > 0) {
      mbar_wait(ds_free, (n_tiles - 1) & 1u);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32(tDQ + lane_off + c * 32, ov);
        tmem_ld_wait();
        if (qi < a.Lq) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(ov[8 * i4 + 0]), __uint_as_float(ov[8 * i4 + 1]));
            o.y = pack_bf16x2(__uint_as_float(ov[8 * i4 + 2]), __uint_as_float(ov[8 * i4 + 3]));
            o.z = pack_bf16x2(__uint_as_float(ov[8 * i4 + 4]), __uint_as_float(ov[8 * i4 + 5]));
            o.w = pack_bf16x2(__uint_as_float(ov[8 * i4 + 6]), __uint_as_float(ov[8 * i4 + 7]));
            *reinterpret_cast<uint4*>(orow + c * 32 + 8 * i4) = o;
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

// Delta[b, h, i] = sum_d dO[b, i, h, d] * O[b, i, h, d]; one warp per row
__global__ void __launch_bounds__(128) attn_delta_kernel(const bf16* __restrict__ o, long long o_bs, long long o_rs,
                                                         long long o_hs, const bf16* __restrict__ d_o, long long d_bs,
                                                         long long d_rs, long long d_hs, int B, int H, int Lq, int D,
                                                         float* __restrict__ delta) {
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * H * Lq) return;
  const int lane = threadIdx.x & 31;
  const int i = static_cast<int>(row % Lq);
  const int h = static_cast<int>((row / Lq) % H);
  const int b = static_cast<int>(row / (static_cast<long long>(Lq) * H));
  const bf16* op = o + b * o_bs + i * o_rs + h * o_hs;
  const bf16* dp = d_o + b * d_bs + i * d_rs + h * d_hs;
  float acc = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    const uint32_t x = *reinterpret_cast<const uint32_t*>(op + d), y = *reinterpret_cast<const uint32_t*>(dp + d);
    acc += bf16lo(x) * bf16lo(y) + bf16hi(x) * bf16hi(y);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) delta[row] = acc;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
size_t flash_attn_bwd_workspace_bytes(int B, int H, int Lq) { return static_cast<size_t>(B) * H * Lq * sizeof(float); }

template <int D>
static int launch_bwd(const AttnArgs& f, const AttnBwdArgs& g, float* delta, cudaStream_t stream) {
  static bool cfg = false;
  if (!cfg) {
    B200_CUDA_OK(cudaFuncSetAttribute(flash_bwd_dkdv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      DkvCfg<D>::kSmemBytes));
    B200_CUDA_OK(cudaFuncSetAttribute(flash_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      DqCfg<D>::kSmemBytes));
    cfg = true;
  }
  BwdArgs a;
  a.dq = g.dq; a.dk = g.dk; a.dv = g.dv;
  a.dq_bs = g.dq_bs; a.dq_rs = g.dq_rs; a.dq_hs = g.dq_hs;
  a.dk_bs = g.dk_bs; a.dk_rs = g.dk_rs; a.dk_hs = g.dk_hs;
  a.dv_bs = g.dv_bs; a.dv_rs = g.dv_rs; a.dv_hs = g.dv_hs;
  a.lse = g.lse;
  a.delta = delta;
  a.B = f.B; a.H = f.H; a.Lq = f.Lq; a.Lk = f.Lk;
  a.kv_start = f.kv_start;
  a.kv_len = f.kv_len;
  a.causal = f.causal;
  a.scale_log2 = f.scale_log2;
  a.scale = f.scale_log2 / kLog2e;
  // the two kernels tile queries / keys differently: separate tensor maps (box rows 64 vs 128)
  CUtensorMap q64, k128, v128, do64, q128, k64, v64, do128;
  int dim[3];
  B200_TRY(make_tmap_attn(&q64, f.q, D, f.Lq, f.H, f.B, f.q_rs, f.q_hs, f.q_bs, 64, a.q_dim));
  B200_TRY(make_tmap_attn(&q128, f.q, D, f.Lq, f.H, f.B, f.q_rs, f.q_hs, f.q_bs, 128, dim));
  B200_TRY(make_tmap_attn(&k128, f.k, D, f.Lk, f.H, f.B, f.k_rs, f.k_hs, f.k_bs, 128, a.k_dim));
  B200_TRY(make_tmap_attn(&k64, f.k, D, f.Lk, f.H, f.B, f.k_rs, f.k_hs, f.k_bs, 64, dim));
  B200_TRY(make_tmap_attn(&v128, f.v, D, f.Lk, f.H, f.B, f.v_rs, f.v_hs, f.v_bs, 128, a.v_dim));
  B200_TRY(make_tmap_attn(&v64, f.v, D, f.Lk, f.H, f.B, f.v_rs, f.v_hs, f.v_bs, 64, dim));
  B200_TRY(make_tmap_attn(&do64, g.d_o, D, f.Lq, f.H, f.B, g.do_rs, g.do_hs, g.do_bs, 64, a.do_dim));
  B200_TRY(make_tmap_attn(&do128, g.d_o, D, f.Lq, f.H, f.B, g.do_rs, g.do_hs, g.do_bs, 128, dim));
  const double pairs = static_cast<double>(f.B) * f.H * f.Lq * f.Lk * (f.causal ? 0.5 : 1.0);
  LaunchScope scope(kFamTrain, stream, 2.0 * f.B * f.H * D * (4.0 * f.Lq + 4.0 * f.Lk), 14.0 * pairs * D, 3);
  const long long rows = static_cast<long long>(f.B) * f.H * f.Lq;
  attn_delta_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, stream>>>(
      f.o, f.o_bs, f.o_rs, f.o_hs, g.d_o, g.do_bs, g.do_rs, g.do_hs, f.B, f.H, f.Lq, D, delta);
  flash_bwd_dkdv_kernel<D><<<dim3((f.Lk + 127) / 128, f.H, f.B), kBwThreads, DkvCfg<D>::kSmemBytes, stream>>>(
      q64, k128, v128, do64, a);
  flash_bwd_dq_kernel<D><<<dim3((f.Lq + 127) / 128, f.H, f.B), kBwThreads, DqCfg<D>::kSmemBytes, stream>>>(
      q128, k64, v64, do128, a);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// f: the forward call's arguments (q, k, v, o, masks, scale); g: dO, lse and the gradient outputs.
int flash_attn_bwd(const AttnArgs& f, const AttnBwdArgs& g, int head_dim, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
  if (f.B <= 0 || f.H <= 0 || f.Lq <= 0 || f.Lk <= 0) return 0;
  if (g.lse == nullptr) return fail(-2, "flash_attn_bwd: the forward log-sum-exp is required");
  if (workspace == nullptr || workspace_bytes < flash_attn_bwd_workspace_bytes(f.B, f.H, f.Lq))
    return fail(-2, "flash_attn_bwd: workspace too small");
  const long long st[] = {g.dq_bs, g.dq_rs, g.dq_hs, g.dk_bs, g.dk_rs, g.dk_hs, g.dv_bs, g.dv_rs, g.dv_hs};
  for (long long v : st)
    if (v % 8 != 0) return fail(-2, "flash_attn_bwd: gradient strides must be multiples of 8 elements");
  float* delta = static_cast<float*>(workspace);
  if (head_dim == 64) return launch_bwd<64>(f, g, delta, stream);
  if (head_dim == 128) return launch_bwd<128>(f, g, delta, stream);
  return fail(-2, "flash_attn_bwd: head_dim %d not supported (64 or 128)", head_dim);
}

}  // namespace b200
