// tcgen05 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . B[N,K]^T), bf16 operands, fp32 accumulation in TMEM.
//
// This one kernel carries ~95 % of the MM2SG hot path's FLOPs: every nn.Linear of the CLIP ViT-L tower
// (reference call site clip_encoder.py:48), the BERT image pooler (multimodal_projector/builder.py:173),
// the mlp2x_gelu projector (llava_arch.py:182) and the Llama-7B decoder + lm_head (llava_llama.py:93).
// In the reference each of them is a cuBLAS call followed by separate bias / activation / residual kernels.
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of a 128 x 64 A tile and a BN x 64 B tile per
//               pipeline stage into 128B-swizzled shared memory, completion on an mbarrier (expect_tx).
//   warp 1      MMA issuer (+ TMEM allocator): one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//               (M = 128, N = BN, K = 16) four times per stage; tcgen05.commit releases the stage back to the
//               producer and, after the last K block, hands the accumulator to the epilogue.
//   warps 2..5  epilogue: tcgen05.ld the 128 x BN fp32 accumulator (one output row per thread), apply
//               bias / activation / SwiGLU pairing / residual / row scatter, convert to bf16 and store 16 B
//               vectors. TMEM holds two accumulators so the epilogue of tile i overlaps the main loop of i+1.
#include <cstdarg>
#include <cstdlib>
#include <mutex>

#include "common.h"
#include "ptx.cuh"

namespace b200 {

static constexpr int kBM = 128;
static constexpr int kBK = 64;  // 64 bf16 = 128 B = one swizzle row
static constexpr int kGemmThreads = 192;
static constexpr int kRasterGroupM = 16;  // legacy raster (B200_RASTER_MB=0): groups of 16 M tiles

struct GemmArgs {
  void* C;
  int ldc;
  int M, N, K;
  int m_tiles, n_tiles;
  // rasterisation (tile_coords): raster_n = 0 -> groups of raster_group consecutive M tiles walk over all N tiles (their
  // A panel stays L2-resident, B streams once per group); raster_n = 1 -> groups of N tiles walk over all M tiles
  int raster_n, raster_group;
  // NF4 variant (template flag NF4): the B operand is the 4-bit NormalFloat storage of an [N, K] weight (nf4.cu: blocks
  // of 64 consecutive K values, two codes per byte, one fp32 absmax per block); four extra warps expand it to bf16
  // straight into the swizzled B stage -- the dequantisation of bitsandbytes' Linear4bit (reference call site
  // train/train.py:1098-1114) fused into the GEMM's producer side, so only the packed base ever lives in HBM
  const uint8_t* nf4_codes;
  const float* nf4_absmax;
  // operand majors: 0 = K-major ([rows, K], the nn.Linear forward layout), 1 = MN-major (the matrix is stored
  // [K, rows]: the transposed operands of the backward GEMMs dX = dY W and dW = dY^T X, read in place)
  int a_mn, b_mn;
  GemmEpilogue epi;
};

template <int BN, int PER_SM = 1>
struct GemmCfg {
  static constexpr int kStageBytesA = kBM * kBK * 2;
  static constexpr int kStageBytesB = BN * kBK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  // BN in {32, 64, 128, 256} are the general-purpose tiles; 96 / 160 / 224 exist for the decode step's wide
  // projections, where the tile width decides whether the weight tiles fill the 148 SMs in one wave
  // (stages.cu::decode_bn): 12288 / 96 = 128 tiles, 22016 / 160 = 138, 32000 / 224 = 143.
  // PER_SM = 2 (decode step only): half the ring so that two CTAs fit one SM -- 3 x 32 KB (BN 128), 3 x 28 KB (BN 96),
  // 4 x 24 KB (BN 64) -- and their 2 x 2 accumulators the 512 TMEM columns.
  static constexpr int kStages = PER_SM == 2 ? (BN == 64 ? 4 : 3)
                                 : (BN == 256) ? 4 : (BN == 224) ? 5 : (BN >= 128) ? 6 : (BN == 96) ? 7 : 8;
  // two accumulators of BN columns; tcgen05.alloc takes a power of two >= 32
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 :
                                   (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN % 32 == 0 && BN <= 256, "the epilogue drains 32 accumulator columns at a time");
  static_assert(kStageBytesB % 1024 == 0, "SWIZZLE_128B tiles start on 1024-byte boundaries");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory per CTA");
  static_assert(PER_SM == 1 || (BN <= 128 && 2 * kTmemCols <= 512 && 2 * (kSmemBytes + 1024) <= 228 * 1024),
                "two CTAs per SM: TMEM columns and shared memory of both must fit");
};

__device__ __forceinline__ void tile_coords(int tile, const GemmArgs& g, int& mt, int& nt) {
  // grouped rasterisation: the tiles of a group share one operand panel while it is L2-hot (sizes chosen on the host,
  // gemm_bf16_ex: the panel that is kept fits the L2 budget, the other operand is streamed once per group)
  const int G = g.raster_group;
  if (g.raster_n == 0) {
    const int per_group = G * g.n_tiles;
    const int group = tile / per_group;
    const int first_m = group * G;
    const int gsize = min(G, g.m_tiles - first_m);
    const int r = tile - group * per_group;
    mt = first_m + r % gsize;
    nt = r / gsize;
  } else {
    const int per_group = G * g.m_tiles;
    const int group = tile / per_group;
    const int first_n = group * G;
    const int gsize = min(G, g.n_tiles - first_n);
    const int r = tile - group * per_group;
    nt = first_n + r % gsize;
    mt = r / gsize;
  }
}

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == kActQuickGelu) return x / (1.0f + __expf(-1.702f * x));
  if (act == kActGeluErf) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  return x;
}

// the 16 NormalFloat levels (QLoRA, appendix E; the same literals as nf4.cu::level)
static __device__ const float kNf4Levels[16] = {
    -1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f, -0.28444138169288635f,
    -0.18477343022823334f, -0.09105003625154495f, 0.0f, 0.07958029955625534f, 0.16093020141124725f,
    0.24611230194568634f, 0.33791524171829224f, 0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};
static constexpr int kNf4Threads = 128;  // dequantisation warps 6..9 of the NF4 variant

template <int BN, int PER_SM = 1, bool NF4 = false>
__global__ void __launch_bounds__(NF4 ? kGemmThreads + kNf4Threads : kGemmThreads, PER_SM)
    gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmArgs g) {
  using Cfg = GemmCfg<BN, PER_SM>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // 1024 B aligned for SWIZZLE_128B
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kStageBytesA;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* nf4_lut = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 192);  // 16 floats, NF4 only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = g.m_tiles * g.n_tiles;
  const int k_blocks = (g.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], NF4 ? 2 : 1);  // NF4: the TMA of the A tile + the dequantisation warps' arrival
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  if (NF4 && threadIdx.x >= kGemmThreads && threadIdx.x < kGemmThreads + 16)
    nf4_lut[threadIdx.x - kGemmThreads] = kNf4Levels[threadIdx.x - kGemmThreads];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // Programmatic dependent launch: weight tiles of the first k-blocks are requested before the previous kernel
      // has finished (they do not depend on it); the activation tiles of the same stages follow after the wait.
      int pre = 0;
      if (!NF4 && g.epi.b_const && !g.b_mn && static_cast<int>(blockIdx.x) < num_tiles) {
        int mt, nt;
        tile_coords(blockIdx.x, g, mt, nt);
        pre = k_blocks < kStages ? k_blocks : kStages;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_expect_tx(&full_bar[kb], Cfg::kStageBytes);
          tma_load_2d(smem_b + kb * Cfg::kStageBytesB, &tmB, &full_bar[kb], kb * kBK, nt * BN);
        }
      }
      griddep_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mt, nt;
        tile_coords(tile, g, mt, nt);
        for (int kb = 0; kb < k_blocks; ++kb) {
          const bool b_done = pre > 0;  // this stage's weight tile (and its expect_tx) was issued above
          if (b_done) --pre;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (!b_done) mbar_expect_tx(&full_bar[stage], NF4 ? Cfg::kStageBytesA : Cfg::kStageBytes);
          if (!g.a_mn) {
            tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmA, &full_bar[stage], kb * kBK, mt * kBM);
          } else {  // [64 k rows][64 m] boxes, one 8 KB panel per 64 rows of the tile
            for (int p = 0; p < kBM / 64; ++p)
              tma_load_2d(smem_a + stage * Cfg::kStageBytesA + p * 8192, &tmA, &full_bar[stage], mt * kBM + p * 64,
                          kb * kBK);
          }
          if (b_done || NF4) {
          } else if (!g.b_mn) {
            tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tmB, &full_bar[stage], kb * kBK, nt * BN);
          } else {
            for (int p = 0; p < BN / 64; ++p)
              tma_load_2d(smem_b + stage * Cfg::kStageBytesB + p * 8192, &tmB, &full_bar[stage], nt * BN + p * 64,
                          kb * kBK);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kBM, BN, g.a_mn, g.b_mn);
      // K-major: 16 k = 32 B inside the 128 B swizzle row (+2 in the addr>>4 field); MN-major: 16 k rows of 128 B
      const uint32_t a_step = g.a_mn ? 128u : 2u, b_step = g.b_mn ? 128u : 2u;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_addr = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::kStageBytesA);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::kStageBytesB);
          const uint64_t da = g.a_mn ? umma_desc_mnmajor_sw128(sa, 8192) : umma_desc_kmajor_sw128(sa);
          const uint64_t db = g.b_mn ? umma_desc_mnmajor_sw128(sb, 8192) : umma_desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16_ss(d_addr, da + a_step * k, db + b_step * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);  // stage reusable once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full[buf]);  // accumulator complete
      }
      // late programmatic-launch trigger: after the last MMA of this CTA's last tile (see gemm_skinny.cu)
      griddep_launch();
    }
  } else if (NF4 && warp >= 6) {
    // ===================== NF4 dequantisation (warps 6..9): codes -> bf16 B stage =====================
    // Thread t owns rows t, t + 128, ... of the BN x 64 tile: 32 bytes of codes + one absmax per row and k-block
    // (the k-block IS the quantisation block: 64 consecutive K values of a weight row). Value = level[code] * absmax
    // rounded to bf16 -- the arithmetic of nf4.cu::dequantize_kernel --, written in the K-major SWIZZLE_128B layout
    // the TMA would have produced: row r at r * 128 B, 16-byte chunk c at position c ^ (r & 7).
    constexpr int kRows = BN >= 128 ? BN / 128 : 1;  // tile rows per dequantisation thread (BN is 128 or 256 here)
    constexpr int kAhead = 4;  // k-blocks whose codes are in flight in registers (HBM latency ~ 1 us >> one MMA stage)
    const int t = threadIdx.x - kGemmThreads;
    const long long kblocks_total = g.K / kBK;
    const int my_tiles = (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const long long total = static_cast<long long>(my_tiles) * k_blocks;  // (tile, k-block) iterations of this CTA
    uint4 c0[kAhead][kRows], c1[kAhead][kRows];
    float am[kAhead][kRows];
    // codes + absmax of iteration `it` -> register slot `slot` (compile-time index: the loops below are unrolled)
    auto fetch = [&](long long it, int slot) {
      const int tile = static_cast<int>(blockIdx.x) + static_cast<int>(it / k_blocks) * static_cast<int>(gridDim.x);
      const int kb = static_cast<int>(it % k_blocks);
      int mt, nt;
      tile_coords(tile, g, mt, nt);
#pragma unroll
      for (int j = 0; j < kRows; ++j) {
        const long long n = static_cast<long long>(nt) * BN + t + 128 * j;
        c0[slot][j] = make_uint4(0u, 0u, 0u, 0u);
        c1[slot][j] = c0[slot][j];
        am[slot][j] = 0.f;
        if (n < g.N) {
          const uint4* src = reinterpret_cast<const uint4*>(g.nf4_codes + (n * g.K + static_cast<long long>(kb) * kBK) / 2);
          c0[slot][j] = __ldg(src);
          c1[slot][j] = __ldg(src + 1);
          am[slot][j] = __ldg(g.nf4_absmax + n * kblocks_total + kb);
        }
      }
    };
#pragma unroll
    for (int s0 = 0; s0 < kAhead; ++s0)
      if (s0 < total) fetch(s0, s0);
    int stage = 0;
    uint32_t phase = 0;
    for (long long it0 = 0; it0 < total; it0 += kAhead) {
#pragma unroll
      for (int s0 = 0; s0 < kAhead; ++s0) {
        const long long it = it0 + s0;
        if (it < total) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sb = smem_b + stage * Cfg::kStageBytesB;
#pragma unroll
          for (int j = 0; j < kRows; ++j) {
            const int r = t + 128 * j;
            const uint32_t words[8] = {c0[s0][j].x, c0[s0][j].y, c0[s0][j].z, c0[s0][j].w,
                                       c1[s0][j].x, c1[s0][j].y, c1[s0][j].z, c1[s0][j].w};
            const float m = am[s0][j];
#pragma unroll
            for (int c = 0; c < 8; ++c) {  // word c = bytes 4c..4c+3 = elements 8c..8c+7 (even element: high nibble)
              const uint32_t wd = words[c];
              uint4 o;
              o.x = pack_bf16x2(nf4_lut[(wd >> 4) & 15u] * m, nf4_lut[wd & 15u] * m);
              o.y = pack_bf16x2(nf4_lut[(wd >> 12) & 15u] * m, nf4_lut[(wd >> 8) & 15u] * m);
              o.z = pack_bf16x2(nf4_lut[(wd >> 20) & 15u] * m, nf4_lut[(wd >> 16) & 15u] * m);
              o.w = pack_bf16x2(nf4_lut[(wd >> 28) & 15u] * m, nf4_lut[(wd >> 24) & 15u] * m);
              *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = o;
            }
          }
          if (it + kAhead < total) fetch(it + kAhead, s0);         // refill this register slot
          fence_proxy_async_smem();                                  // generic-proxy writes -> visible to tcgen05.mma
          asm volatile("bar.sync 1, %0;" ::"n"(kNf4Threads) : "memory");  // all rows of the stage are in place
          if (t == 0) mbar_arrive(&full_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const GemmEpilogue& e = g.epi;
    griddep_wait();  // residual reads and C writes below must follow the previous kernel
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int mt, nt;
      tile_coords(tile, g, mt, nt);
      const int buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1u);
      tc_fence_after_sync();
      const int row = mt * kBM + q * 32 + lane;
      int out_row = row;
      bool ok = row < g.M;
      if (ok && e.row_map != nullptr) {
        out_row = e.row_map[row];
        ok = out_row >= 0;
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        const int n0 = nt * BN + c * 32;
        if (!ok || n0 >= g.N) continue;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) * e.scale;
        if (e.bias != nullptr) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (n0 + j8 * 8 < g.N) {
              const uint4 b = __ldg(reinterpret_cast<const uint4*>(e.bias + n0 + j8 * 8));
              x[j8 * 8 + 0] += bf16lo(b.x);
              x[j8 * 8 + 1] += bf16hi(b.x);
              x[j8 * 8 + 2] += bf16lo(b.y);
              x[j8 * 8 + 3] += bf16hi(b.y);
              x[j8 * 8 + 4] += bf16lo(b.z);
              x[j8 * 8 + 5] += bf16hi(b.z);
              x[j8 * 8 + 6] += bf16lo(b.w);
              x[j8 * 8 + 7] += bf16hi(b.w);
            }
          }
        }
        if (e.act == kActSwiGLU) {
          // 32 accumulator columns -> 16 outputs
          const int o0 = n0 >> 1;
          bf16* crow = reinterpret_cast<bf16*>(g.C) + static_cast<size_t>(out_row) * g.ldc + o0;
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            if (n0 + j8 * 16 < g.N) {
              float y[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float gt = x[j8 * 16 + 2 * j], up = x[j8 * 16 + 2 * j + 1];
                y[j] = gt / (1.0f + __expf(-gt)) * up;
              }
              uint4 o;
              o.x = pack_bf16x2(y[0], y[1]);
              o.y = pack_bf16x2(y[2], y[3]);
              o.z = pack_bf16x2(y[4], y[5]);
              o.w = pack_bf16x2(y[6], y[7]);
              *reinterpret_cast<uint4*>(crow + j8 * 8) = o;
            }
          }
          continue;
        }
        if (e.act != kActNone) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = act_apply(x[j], e.act);
        }
        if (e.residual != nullptr) {
          const size_t rr = e.res_group > 0
                                ? static_cast<size_t>(row / e.res_group) * e.res_group_stride + row % e.res_group
                                : static_cast<size_t>(row);
          const bf16* rrow = e.residual + rr * e.ldr + n0;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (n0 + j8 * 8 < g.N) {
              const uint4 r = *reinterpret_cast<const uint4*>(rrow + j8 * 8);
              x[j8 * 8 + 0] += bf16lo(r.x);
              x[j8 * 8 + 1] += bf16hi(r.x);
              x[j8 * 8 + 2] += bf16lo(r.y);
              x[j8 * 8 + 3] += bf16hi(r.y);
              x[j8 * 8 + 4] += bf16lo(r.z);
              x[j8 * 8 + 5] += bf16hi(r.z);
              x[j8 * 8 + 6] += bf16lo(r.w);
              x[j8 * 8 + 7] += bf16hi(r.w);
            }
          }
        }
        if (e.out_fp32) {
          float* crow = reinterpret_cast<float*>(g.C) + static_cast<size_t>(out_row) * g.ldc + n0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            if (n0 + j4 * 4 < g.N) {
              float4 o = make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
              if (e.accumulate) {  // C += result (gradient accumulation over micro-batches)
                const float4 old = *reinterpret_cast<const float4*>(crow + j4 * 4);
                o.x += old.x;
                o.y += old.y;
                o.z += old.z;
                o.w += old.w;
              }
              *reinterpret_cast<float4*>(crow + j4 * 4) = o;
            }
          }
        } else {
          const size_t coff = static_cast<size_t>(out_row) * g.ldc + n0;
          bf16* crow = reinterpret_cast<bf16*>(g.C) + coff;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (n0 + j8 * 8 < g.N) {
              uint4 o;
              o.x = pack_bf16x2(x[j8 * 8 + 0], x[j8 * 8 + 1]);
              o.y = pack_bf16x2(x[j8 * 8 + 2], x[j8 * 8 + 3]);
              o.z = pack_bf16x2(x[j8 * 8 + 4], x[j8 * 8 + 5]);
              o.w = pack_bf16x2(x[j8 * 8 + 6], x[j8 * 8 + 7]);
              if (e.n_peers == 0) {
                *reinterpret_cast<uint4*>(crow + j8 * 8) = o;
              } else {
                // all-gather fused into the epilogue: the tile goes straight into every rank's copy over NVLink
                for (int p = 0; p < e.n_peers; ++p)
                  *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.peer_c[p]) + coff + j8 * 8) = o;
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with leading dimension ld (elements); box = box_rows x 64
// elements, 128-byte swizzle, out-of-bounds elements read as zero.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 int l2_promotion = 2) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(-6, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(-2, "TMA operand must be 16-byte aligned (ptr %p, ld %llu)", base, (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  l2_promotion == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                  : l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                      : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int BN, int PER_SM = 1, bool NF4 = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PER_SM>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, PER_SM, NF4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
    // experiment knob (tools/corun_bench.py): the full 228 KB shared carve-out instead of the smallest one that holds
    // this kernel's ring, so that small-footprint kernels of another stream find room next to the persistent CTA
    const char* e = getenv("B200_GEMM_CARVEOUT");
    if (e != nullptr && e[0] == '1')
      B200_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, PER_SM, NF4>,
                                        cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  const int tiles = g.m_tiles * g.n_tiles;
  const int slots = num_sms() * PER_SM;
  const int grid = tiles < slots ? tiles : slots;
  const double mnk = static_cast<double>(g.M) * g.N * g.K;
  const int n_out = g.epi.act == kActSwiGLU ? g.N / 2 : g.N;
  const double bytes = 2.0 * (static_cast<double>(g.M) * g.K + static_cast<double>(g.N) * g.K) +
                       static_cast<double>(g.M) * n_out * (g.epi.out_fp32 ? 4 : 2) +
                       (g.epi.residual ? 2.0 * g.M * g.N : 0.0);
  LaunchScope scope(g.M <= 128 ? kFamGemmSkinny : kFamGemm, stream, bytes, 2.0 * mnk);
  B200_CUDA_OK(launch_ex(gemm_bf16_tn_kernel<BN, PER_SM, NF4>, dim3(grid), dim3(NF4 ? kGemmThreads + kNf4Threads : kGemmThreads),
                         Cfg::kSmemBytes, stream, 0, g.epi.b_const != 0, tmA, tmB, g));
  return 0;
}

// C[M, N] = epilogue(A[M, K] . dequant(W)[N, K]^T): W given as NF4 codes (N * K / 2 bytes, row-major [N, K]) and one fp32
// absmax per block of 64 K values (N * K / 64 floats). Same tiles, MMA order and epilogue as gemm_bf16_tn on the
// dequantised weight => bit-identical results.
int gemm_nf4_tn(const bf16* A, int lda, const uint8_t* codes, const float* absmax, void* C, int ldc, int M, int N, int K,
                const GemmEpilogue& epi, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0 || K % kBK != 0) return fail(-2, "gemm_nf4: K (%d) must be a positive multiple of 64 (the NF4 block)", K);
  if (A == nullptr || codes == nullptr || absmax == nullptr || C == nullptr) return fail(-2, "gemm_nf4: null operand");
  if ((reinterpret_cast<uintptr_t>(codes) & 15) != 0) return fail(-2, "gemm_nf4: codes must be 16-byte aligned");
  if (epi.act == kActSwiGLU || epi.n_peers != 0 || epi.ctas_per_sm == 2)
    return fail(-2, "gemm_nf4: SwiGLU pairing / peer stores / two CTAs per SM are not built for the NF4 operand");
  GemmArgs g;
  g.C = C;
  g.ldc = ldc;
  g.M = M;
  g.N = N;
  g.K = K;
  g.m_tiles = (M + kBM - 1) / kBM;
  const int bn = (g.m_tiles * ((N + 255) / 256) >= num_sms()) ? 256 : 128;
  g.n_tiles = (N + bn - 1) / bn;
  g.a_mn = g.b_mn = 0;
  g.epi = epi;
  g.epi.b_const = 0;
  g.raster_n = 0;
  g.raster_group = g.m_tiles < 32 ? g.m_tiles : 32;
  g.nf4_codes = codes;
  g.nf4_absmax = absmax;
  CUtensorMap tmA;
  B200_TRY(make_tmap_2d(&tmA, A, M, K, lda, kBM));
  if (bn == 256) return launch_gemm<256, 1, true>(tmA, tmA, g, stream);
  return launch_gemm<128, 1, true>(tmA, tmA, g, stream);
}

// A: [M, K] (a_mn = 0) or stored transposed [K, M] (a_mn = 1); B: [N, K] (b_mn = 0) or [K, N] (b_mn = 1).
int gemm_bf16_ex(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                 int K, const GemmEpilogue& epi, int bn_hint, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (N % 8 != 0) return fail(-2, "gemm: N (%d) must be a multiple of 8", N);
  if (!(a_mn && b_mn) && K % 8 != 0) return fail(-2, "gemm: K (%d) must be a multiple of 8 for a K-major operand", K);
  // (a transposed operand only needs a 16-byte aligned leading dimension; make_tmap_2d checks it)
  if (epi.act == kActSwiGLU && (N % 16 != 0 || epi.out_fp32 || epi.residual))
    return fail(-2, "gemm: SwiGLU epilogue needs N %% 16 == 0, bf16 output and no residual");
  if (epi.n_peers < 0 || epi.n_peers > 8 || (epi.n_peers > 0 && (epi.out_fp32 || epi.act == kActSwiGLU)))
    return fail(-2, "gemm: peer stores need 1..8 peers, bf16 output and no SwiGLU pairing");
  if (epi.accumulate && !epi.out_fp32) return fail(-2, "gemm: accumulation needs an fp32 output");
  const int m_tiles = (M + kBM - 1) / kBM;
  int bn = bn_hint;
  if (bn == 0) {
    // largest tile that still gives every SM work; small-M (decode) problems get narrow tiles so that all
    // 148 SMs stream weights
    bn = 256;
    const int sms = num_sms();
    while (bn > 32 && m_tiles * ((N + bn - 1) / bn) < sms) bn >>= 1;
  }
  if (b_mn && bn < 64) bn = 64;  // a transposed B tile is made of 64-column panels
  if (bn != 32 && bn != 64 && bn != 96 && bn != 128 && bn != 160 && bn != 224 && bn != 256)
    return fail(-2, "gemm: unsupported BN %d", bn);
  if (b_mn && bn % 64 != 0) return fail(-2, "gemm: a transposed B operand needs BN %% 64 == 0 (got %d)", bn);
  GemmArgs g;
  g.C = C;
  g.ldc = ldc;
  g.M = M;
  g.N = N;
  g.K = K;
  g.m_tiles = m_tiles;
  g.n_tiles = (N + bn - 1) / bn;
  g.a_mn = a_mn ? 1 : 0;
  g.b_mn = b_mn ? 1 : 0;
  g.epi = epi;
  g.nf4_codes = nullptr;
  g.nf4_absmax = nullptr;
  {
    // Raster choice. DRAM traffic of a grouped raster = (kept operand, once) + (streamed operand) x (number of groups);
    // the kept panel must stay in the 126 MB L2 next to the stream, so it gets a budget (default 40 MB; the round-1
    // raster -- 16 M tiles per group whatever K -- re-read the 180 MB gate_up weights 6.5 times per prefill chunk:
    // 1.37 GB of DRAM reads against 0.29 GB algorithmic, profiles/r2_ncu_gemm256llm.txt). B200_RASTER_MB=0 restores it.
    static const int budget_mb = [] {
      const char* e = getenv("B200_RASTER_MB");
      return e != nullptr ? atoi(e) : 40;
    }();
    g.raster_n = 0;
    g.raster_group = kRasterGroupM;
    if (budget_mb > 0) {
      const double budget = budget_mb * 1048576.0;
      const double a_tile = 2.0 * kBM * K, b_tile = 2.0 * bn * K;
      const double a_all = a_tile * g.m_tiles, b_all = b_tile * g.n_tiles;
      auto clampi = [](double v, int hi) { return v < 1.0 ? 1 : (v > hi ? hi : static_cast<int>(v)); };
      const int gm = clampi(budget / a_tile, g.m_tiles), gn = clampi(budget / b_tile, g.n_tiles);
      const int groups_m = (g.m_tiles + gm - 1) / gm, groups_n = (g.n_tiles + gn - 1) / gn;
      const double cost_m = a_all + b_all * groups_m, cost_n = b_all + a_all * groups_n;
      if (cost_n < cost_m) {
        g.raster_n = 1;
        g.raster_group = (g.n_tiles + groups_n - 1) / groups_n;   // equal-sized groups
      } else {
        g.raster_group = (g.m_tiles + groups_m - 1) / groups_m;
      }
    }
  }
  CUtensorMap tmA, tmB;
  if (a_mn)
    B200_TRY(make_tmap_2d(&tmA, A, K, M, lda, 64));
  else
    B200_TRY(make_tmap_2d(&tmA, A, M, K, lda, kBM));
  if (b_mn)
    B200_TRY(make_tmap_2d(&tmB, B, K, N, ldb, 64));
  else
    B200_TRY(make_tmap_2d(&tmB, B, N, K, ldb, bn));
  if (epi.ctas_per_sm == 2) {
    if (a_mn || b_mn) return fail(-2, "gemm: two CTAs per SM is the decode step's K-major shape only");
    switch (bn) {
      case 128: return launch_gemm<128, 2>(tmA, tmB, g, stream);
      case 96: return launch_gemm<96, 2>(tmA, tmB, g, stream);
      case 64: return launch_gemm<64, 2>(tmA, tmB, g, stream);
      default: return fail(-2, "gemm: two CTAs per SM needs BN 64 / 96 / 128 (got %d)", bn);
    }
  }
  switch (bn) {
    case 256: return launch_gemm<256>(tmA, tmB, g, stream);
    case 224: return launch_gemm<224>(tmA, tmB, g, stream);
    case 160: return launch_gemm<160>(tmA, tmB, g, stream);
    case 128: return launch_gemm<128>(tmA, tmB, g, stream);
    case 96: return launch_gemm<96>(tmA, tmB, g, stream);
    case 64: return launch_gemm<64>(tmA, tmB, g, stream);
    default: return launch_gemm<32>(tmA, tmB, g, stream);
  }
}

int gemm_bf16_tn(const bf16* A, int lda, const bf16* B, int ldb, void* C, int ldc, int M, int N, int K,
                 const GemmEpilogue& epi, int bn_hint, cudaStream_t stream) {
  return gemm_bf16_ex(A, lda, 0, B, ldb, 0, C, ldc, M, N, K, epi, bn_hint, stream);
}

}  // namespace b200
