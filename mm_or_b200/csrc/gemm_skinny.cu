// Weight-streaming tcgen05 GEMM for the decode step:  C[M,N] = epilogue(A[M,K] . W[N,K]^T) with M <= 256 rows
// (one row per sequence being decoded).
//
// Replaces, for the single-token step of llava_arch.py:192-201 -> HF LlamaDecoderLayer, the q/k/v/o/gate/up/down_proj
// and lm_head nn.Linear calls (llava_llama.py:93). At M = batch the work is HBM bound: 13.2 GB of bf16 weights are
// streamed once per step and every byte is used for only 2*M FLOPs. The prefill kernel (gemm_sm100.cu) is the wrong
// shape for that: with a 128-row activation tile and narrow weight tiles each CTA re-reads the whole activation
// matrix from L2 (4x the weight bytes at BN = 32) and N = 4096 outputs give only 32..128 CTAs.
//
// Design ("swap AB" + split K, persistent, warp specialised like the prefill kernel):
//   * the WEIGHT tile is the MMA's M operand: 128 weight rows x 64 k (16 KB, TMA, 128B swizzle); the activations are
//     the N operand: MT (= batch rounded up to 32/64/128/256) rows x 64 k. D[128 x MT] fp32 lives in TMEM (two
//     accumulators so that the epilogue of one work unit overlaps the main loop of the next).
//   * a work unit is (weight tile, K split). Splitting K gives every SM a stream of weights even when N / 128 < 148
//     (o_proj / down_proj: 32 tiles). Units are dealt round-robin to one persistent CTA per SM; 8 pipeline stages of
//     24 KB keep ~190 KB of loads in flight per SM, which is what saturating HBM3e needs (Little's law: ~4.5 MB
//     chip-wide at 6.4 TB/s x ~700 ns).
//   * split-K reduction without a second kernel: every unit stores its fp32 partial tile to the workspace
//     (column-major, coalesced across TMEM lanes), fences, and bumps a per-tile counter; the CTA that arrives last
//     sums the partials in split order (fixed order => deterministic bits), applies bias / activation / SwiGLU /
//     residual and writes bf16. Counters are left at zero for the next launch (CUDA-graph replay safe).
//   * epilogue threads own one weight row (= output column n) each, so for a fixed sequence m the 32 lanes of a
//     warp write 32 consecutive bf16 of C: coalesced 64 B segments. SwiGLU pairs (gate, up) sit in adjacent lanes
//     and are combined with one shuffle.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

static constexpr int kSkBW = 128;  // weight rows per tile (UMMA M)
static constexpr int kSkBK = 64;
static constexpr int kSkThreads = 192;

struct SkinnyArgs {
  void* C;
  int ldc;
  int M, N, K;
  int n_tiles, splits, kb_per_split, k_blocks;
  float* partial;  // [n_tiles * splits][MT][128] fp32
  int* counters;   // [n_tiles], zero on entry, zero on exit
  GemmEpilogue epi;
};

template <int MT>
struct SkinnyCfg {
  static constexpr int kStageBytesW = kSkBW * kSkBK * 2;
  static constexpr int kStageBytesA = MT * kSkBK * 2;
  static constexpr int kStageBytes = kStageBytesW + kStageBytesA;
  static constexpr int kStages = (MT <= 64) ? 8 : (MT == 128) ? 6 : 4;
  static constexpr int kTmemCols = (2 * MT < 32) ? 32 : 2 * MT;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

__device__ __forceinline__ float sk_act(float x, int act) {
  if (act == kActQuickGelu) return x / (1.0f + __expf(-1.702f * x));
  if (act == kActGeluErf) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  return x;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int MT>
__global__ void __launch_bounds__(kSkThreads, 1)
    gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                       const SkinnyArgs g) {
  using Cfg = SkinnyCfg<MT>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + kStages * Cfg::kStageBytesW;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_units = g.n_tiles * g.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int nt = u / g.splits, sp = u - nt * g.splits;
        const int kb0 = sp * g.kb_per_split;
        const int kb1 = min(g.k_blocks, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_w + stage * Cfg::kStageBytesW, &tmW, &full_bar[stage], kb * kSkBK, nt * kSkBW);
          tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmA, &full_bar[stage], kb * kSkBK, 0);
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
      constexpr uint32_t idesc = umma_idesc_bf16(kSkBW, MT);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
        const int nt = u / g.splits, sp = u - nt * g.splits;
        const int kb0 = sp * g.kb_per_split;
        const int kb1 = min(g.k_blocks, kb0 + g.kb_per_split);
        const int buf = it & 1;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_addr = tmem_base + static_cast<uint32_t>(buf * MT);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint64_t dw = umma_desc_kmajor_sw128(smem_u32(smem_w + stage * Cfg::kStageBytesW));
          const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem_a + stage * Cfg::kStageBytesA));
#pragma unroll
          for (int k = 0; k < kSkBK / 16; ++k)
            umma_bf16_ss(d_addr, dw + 2 * k, da + 2 * k, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // weight row inside the tile = TMEM lane
    const GemmEpilogue& e = g.epi;
    const bool swiglu = e.act == kActSwiGLU;
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int nt = u / g.splits, sp = u - nt * g.splits;
      const int buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1u);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * MT);
      const int n = nt * kSkBW + r;
      bool reduce_here = true;
      if (g.splits > 1) {
        // ---- park the partial tile, then find out whether this CTA is the last of its tile
        float* part = g.partial + static_cast<size_t>(u) * MT * kSkBW;
#pragma unroll 1
        for (int c = 0; c < MT / 32; ++c) {
          if (c * 32 >= g.M) break;
          uint32_t v[32];
          tmem_ld_32x32(t_row + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < g.M) part[(c * 32 + j) * kSkBW + r] = __uint_as_float(v[j]);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);  // accumulator drained: the MMA warp may reuse it
        __threadfence();
        epi_bar_sync();
        if (threadIdx.x == 64) {
          const int old = atomicAdd(&g.counters[nt], 1);
          const int last = old == g.splits - 1;
          if (last) g.counters[nt] = 0;  // every other split of this tile has already arrived
          *last_flag = last;
        }
        epi_bar_sync();
        reduce_here = *last_flag != 0;
        epi_bar_sync();  // last_flag may be rewritten by the next unit only after everyone has read it
        if (reduce_here) __threadfence();
      }
      if (reduce_here) {
        const float bias = (e.bias != nullptr && n < g.N) ? __bfloat162float(e.bias[n]) : 0.f;
#pragma unroll 1
        for (int c = 0; c < MT / 32; ++c) {
          if (c * 32 >= g.M) break;
          float x[32];
          if (g.splits > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = 0.f;
            for (int s = 0; s < g.splits; ++s) {
              const float* part = g.partial + static_cast<size_t>(nt * g.splits + s) * MT * kSkBW;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c * 32 + j < g.M) x[j] += __ldcg(part + (c * 32 + j) * kSkBW + r);
            }
          } else {
            uint32_t v[32];
            tmem_ld_32x32(t_row + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = c * 32 + j;
            float val = x[j] + bias;
            if (swiglu) {
              const float other = __shfl_xor_sync(0xffffffffu, val, 1);  // lane 2i: gate, lane 2i+1: up
              if (m < g.M && n < g.N && (lane & 1) == 0) {
                const float y = val / (1.0f + __expf(-val)) * other;
                reinterpret_cast<bf16*>(g.C)[static_cast<size_t>(m) * g.ldc + (n >> 1)] = __float2bfloat16(y);
              }
              continue;
            }
            if (m >= g.M || n >= g.N) continue;
            val = sk_act(val, e.act);
            if (e.residual != nullptr) val += __bfloat162float(e.residual[static_cast<size_t>(m) * e.ldr + n]);
            if (e.out_fp32)
              reinterpret_cast<float*>(g.C)[static_cast<size_t>(m) * g.ldc + n] = val;
            else
              reinterpret_cast<bf16*>(g.C)[static_cast<size_t>(m) * g.ldc + n] = __float2bfloat16(val);
          }
        }
      }
      if (g.splits == 1) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
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
// host side
// ---------------------------------------------------------------------------------------------------
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

static int skinny_mt(int M) { return M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256; }

static void skinny_plan(int M, int N, int K, int splits_hint, int& n_tiles, int& splits, int& kb_per) {
  const int k_blocks = (K + kSkBK - 1) / kSkBK;
  n_tiles = (N + kSkBW - 1) / kSkBW;
  int s = splits_hint;
  if (s <= 0) {
    // enough units for every SM to stream weights, but at least 8 k-blocks (128 KB of weights) per unit
    s = 1;
    const int sms = num_sms();
    while (n_tiles * s < sms && k_blocks / (s * 2) >= 8) s *= 2;
  }
  if (s > k_blocks) s = k_blocks;
  if (s > 32) s = 32;
  kb_per = (k_blocks + s - 1) / s;
  splits = (k_blocks + kb_per - 1) / kb_per;  // no empty split
  (void)M;
}

size_t gemm_skinny_workspace_bytes(int M, int N, int K) {
  // worst case over the automatic plan and any splits hint up to 32
  const int n_tiles = (N + kSkBW - 1) / kSkBW;
  const size_t counters = gemm_skinny_counter_bytes();
  int nt, s, kb;
  skinny_plan(M, N, K, 0, nt, s, kb);
  if (s < 8) s = 8;
  return counters + static_cast<size_t>(n_tiles) * s * skinny_mt(M) * kSkBW * sizeof(float);
}

template <int MT>
static int launch_skinny(const CUtensorMap& tmW, const CUtensorMap& tmA, const SkinnyArgs& g, cudaStream_t stream) {
  using Cfg = SkinnyCfg<MT>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
    configured = true;
  }
  const int units = g.n_tiles * g.splits;
  const int grid = units < num_sms() ? units : num_sms();
  const int n_out = g.epi.act == kActSwiGLU ? g.N / 2 : g.N;
  const double bytes = 2.0 * (static_cast<double>(g.M) * g.K + static_cast<double>(g.N) * g.K) +
                       static_cast<double>(g.M) * n_out * (g.epi.out_fp32 ? 4 : 2) +
                       (g.epi.residual ? 2.0 * g.M * g.N : 0.0);
  LaunchScope scope(kFamGemmSkinny, stream, bytes, 2.0 * g.M * g.N * g.K);
  gemm_skinny_kernel<MT><<<grid, kSkThreads, Cfg::kSmemBytes, stream>>>(tmW, tmA, g);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// counters region: the first gemm_skinny_counter_bytes() of the workspace; must be zero on entry (the decode step
// clears it once per step with a memset node; the kernel leaves it zero).
int gemm_skinny(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int M, int N, int K,
                const GemmEpilogue& epi, int splits_hint, void* workspace, size_t workspace_bytes,
                cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (M > 256) return fail(-2, "gemm_skinny: M %d > 256 (use gemm_bf16_tn)", M);
  if (K % 8 != 0) return fail(-2, "gemm_skinny: K (%d) must be a multiple of 8", K);
  if (epi.row_map != nullptr || epi.res_group != 0) return fail(-2, "gemm_skinny: row maps are not supported");
  if (epi.act == kActSwiGLU && (N % 2 != 0 || epi.out_fp32 || epi.residual))
    return fail(-2, "gemm_skinny: SwiGLU epilogue needs even N, bf16 output and no residual");
  SkinnyArgs g;
  g.C = C;
  g.ldc = ldc;
  g.M = M;
  g.N = N;
  g.K = K;
  g.k_blocks = (K + kSkBK - 1) / kSkBK;
  skinny_plan(M, N, K, splits_hint, g.n_tiles, g.splits, g.kb_per_split);
  g.epi = epi;
  const int mt = skinny_mt(M);
  const size_t counters = gemm_skinny_counter_bytes();
  if (static_cast<size_t>(g.n_tiles) * sizeof(int) > counters)
    return fail(-2, "gemm_skinny: N %d needs more than %zu tile counters", N, counters / sizeof(int));
  g.counters = nullptr;
  g.partial = nullptr;
  if (g.splits > 1) {
    const size_t need = counters + static_cast<size_t>(g.n_tiles) * g.splits * mt * kSkBW * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need)
      return fail(-2, "gemm_skinny: workspace too small (%zu < %zu)", workspace_bytes, need);
    g.counters = static_cast<int*>(workspace);
    g.partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + counters);
  }
  CUtensorMap tmW, tmA;
  B200_TRY(make_tmap_2d(&tmW, W, N, K, ldw, kSkBW));
  B200_TRY(make_tmap_2d(&tmA, A, M, K, lda, mt));
  switch (mt) {
    case 32: return launch_skinny<32>(tmW, tmA, g, stream);
    case 64: return launch_skinny<64>(tmW, tmA, g, stream);
    case 128: return launch_skinny<128>(tmW, tmA, g, stream);
    default: return launch_skinny<256>(tmW, tmA, g, stream);
  }
}

}  // namespace b200
