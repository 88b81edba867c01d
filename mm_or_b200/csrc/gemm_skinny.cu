// Weight-streaming tcgen05 GEMM for the decode step:  C[M,N] = epilogue(A[M,K] . W[N,K]^T) with M <= 256 rows
// (one row per sequence being decoded).
//
// Replaces, for the single-token step of llava_arch.py:192-201 -> HF LlamaDecoderLayer, the q/k/v/o/gate/up/down_proj
// and lm_head nn.Linear calls (llava_llama.py:93). At M = batch the work is HBM bound: 13.2 GB of bf16 weights are
// streamed once per step and every byte is used for only 2*M FLOPs. The prefill kernel (gemm_sm100.cu) is the wrong
// shape for that: with a 128-row activation tile and narrow weight tiles each CTA re-reads the whole activation
// matrix from L2 (4x the weight bytes at BN = 32) and N = 4096 outputs give only 32..128 CTAs.
//
// Design ("swap AB" + split K across a thread-block cluster):
//   * the WEIGHT tile is the MMA's M operand: 128 weight rows x 64 k (16 KB, TMA, 128B swizzle); the activations are
//     the N operand: MT (= batch rounded up to 32/64/128/256) rows x 64 k. D[128 x MT] fp32 lives in TMEM.
//   * one CTA per (weight tile, K split), two CTAs resident per SM (4 x 24 KB stages each). Measured (profiles/
//     r1_skinny_gemm_notes.md): an SM streams cold weights at ~30 GB/s with one resident CTA and ~41 GB/s with two,
//     independent of pipeline depth, TMA box shape, L2 promotion or weight layout -- roughly its fair share of HBM --
//     so what matters is that every SM holds two streaming CTAs for the whole kernel. This kernel is used where the
//     tiled kernel cannot do that: few weight tiles (N = 4096: o_proj, down_proj); the caller keeps the tiled kernel
//     for the wide projections.
//   * the S (<= 8) K-splits of one weight tile form a thread-block CLUSTER. After its main loop every CTA parks its
//     fp32 partial tile in its own shared memory (the drained pipeline buffers); after a cluster barrier CTA r sums
//     rows [r*128/S, (r+1)*128/S) of all S partials through distributed shared memory in split order (fixed order
//     => deterministic bits), applies bias / activation / SwiGLU / residual and writes bf16. No global scratch, no
//     atomics, no second kernel.
//   * epilogue threads own one weight row (= output column n) each, so for a fixed sequence m the lanes of a warp
//     write consecutive bf16 of C (coalesced). SwiGLU (gate, up) pairs sit in adjacent lanes: one shuffle.
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
  GemmEpilogue epi;
};

template <int MT, int STAGES>
struct SkinnyCfg {
  static constexpr int kStageBytesW = kSkBW * kSkBK * 2;
  static constexpr int kStageBytesA = MT * kSkBK * 2;
  static constexpr int kStageBytes = kStageBytesW + kStageBytesA;
  static constexpr int kStages = STAGES;
  static constexpr int kTmemCols = MT < 32 ? 32 : MT;
  static constexpr int kPartialBytes = MT * kSkBW * 4;  // fp32 partial tile, aliases the pipeline buffers
  static_assert(kPartialBytes <= kStages * kStageBytes, "partial tile must fit in the drained pipeline buffers");
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

__device__ __forceinline__ float sk_act(float x, int act) {
  if (act == kActQuickGelu) return x / (1.0f + __expf(-1.702f * x));
  if (act == kActGeluErf) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  return x;
}

__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t cta) {
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

// one output element: bias -> activation / SwiGLU pairing -> residual -> store. Called by all 32 lanes (shuffle).
__device__ __forceinline__ void sk_store(const SkinnyArgs& g, float val, float bias, int m, int n, bool live,
                                         int lane) {
  const GemmEpilogue& e = g.epi;
  val = val * e.scale + bias;
  if (e.act == kActSwiGLU) {
    const float other = __shfl_xor_sync(0xffffffffu, val, 1);  // lane 2i: gate, lane 2i+1: up
    if (live && (lane & 1) == 0) {
      const float y = val / (1.0f + __expf(-val)) * other;
      reinterpret_cast<bf16*>(g.C)[static_cast<size_t>(m) * g.ldc + (n >> 1)] = __float2bfloat16(y);
    }
    return;
  }
  if (!live) return;
  val = sk_act(val, e.act);
  if (e.residual != nullptr) val += __bfloat162float(e.residual[static_cast<size_t>(m) * e.ldr + n]);
  if (e.out_fp32)
    reinterpret_cast<float*>(g.C)[static_cast<size_t>(m) * g.ldc + n] = val;
  else
    reinterpret_cast<bf16*>(g.C)[static_cast<size_t>(m) * g.ldc + n] = __float2bfloat16(val);
}

template <int MT, int STAGES>
__global__ void __launch_bounds__(kSkThreads, (SkinnyCfg<MT, STAGES>::kSmemBytes <= 112 * 1024) ? 2 : 1)
    gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                       const SkinnyArgs g) {
  using Cfg = SkinnyCfg<MT, STAGES>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + kStages * Cfg::kStageBytesW;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* partial = reinterpret_cast<float*>(smem);  // [MT][128] fp32 once the pipeline has drained

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = g.splits;                       // == cluster size
  const int nt = blockIdx.x / S;
  const int sp = S > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int kb0 = sp * g.kb_per_split;
  const int kb1 = min(g.k_blocks, kb0 + g.kb_per_split);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
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
      // Programmatic dependent launch: the weight tiles of the first stages are requested while the previous kernel
      // is still draining (weights do not depend on it); the activation tiles follow after the wait.
      const int pre = g.epi.b_const ? (nkb < kStages ? nkb : kStages) : 0;
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(&full_bar[i], Cfg::kStageBytes);
        tma_load_2d(smem_w + i * Cfg::kStageBytesW, &tmW, &full_bar[i], (kb0 + i) * kSkBK, nt * kSkBW);
      }
      griddep_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb0 + i;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (i >= pre) {
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_w + stage * Cfg::kStageBytesW, &tmW, &full_bar[stage], kb * kSkBK, nt * kSkBW);
        }
        tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmA, &full_bar[stage], kb * kSkBK, 0);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kSkBW, MT);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint64_t dw = umma_desc_kmajor_sw128(smem_u32(smem_w + stage * Cfg::kStageBytesW));
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem_a + stage * Cfg::kStageBytesA));
#pragma unroll
        for (int k = 0; k < kSkBK / 16; ++k)
          umma_bf16_ss(tmem_base, dw + 2 * k, da + 2 * k, idesc, (i > 0 || k != 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (kb1 > kb0) umma_commit(tmem_full);  // fires when every MMA above has completed (and has read its smem)
      // Programmatic dependent launch, LATE trigger: the next kernel may be scheduled once every CTA of this grid has
      // issued its last MMA -- its prologue then overlaps this kernel's drain + epilogue only. (Round 1 triggered at
      // kernel entry: the successors' CTAs, parked in griddepcontrol.wait, held shared memory and CTA slots that the
      // running kernel's later waves needed, and the step got 4 % slower.) A no-op without the launch attribute.
      griddep_launch();
    }
  } else {
    // ===================== epilogue warps 2..5: drain TMEM =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // weight row inside the tile = TMEM lane
    const bool empty_split = kb1 <= kb0;  // only possible when S does not divide the k-blocks: partial = 0
    griddep_wait();  // residual reads and C writes below must follow the previous kernel
    if (!empty_split) mbar_wait(tmem_full, 0);
    tc_fence_after_sync();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int n = nt * kSkBW + r;
    const float bias = (S == 1 && g.epi.bias != nullptr && n < g.N) ? __bfloat162float(g.epi.bias[n]) : 0.f;
#pragma unroll 1
    for (int c = 0; c < MT / 32; ++c) {
      if (c * 32 >= g.M) break;
      uint32_t v[32];
      if (!empty_split) {
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (S > 1) {
        // park the partial tile: [m][128 weight rows] fp32, a warp writes 128 contiguous bytes per m
#pragma unroll
        for (int j = 0; j < 32; ++j) partial[(c * 32 + j) * kSkBW + r] = __uint_as_float(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int m = c * 32 + j;
          sk_store(g, __uint_as_float(v[j]), bias, m, n, m < g.M && n < g.N, lane);
        }
      }
    }
    tc_fence_before_sync();
  }

  if (S > 1) {
    // ===================== cluster reduction through distributed shared memory =====================
    __syncwarp();
    cluster_sync_all();  // every CTA of the cluster has parked its partial tile
    if (warp >= 2) {
      const int t = threadIdx.x - 64;        // 0..127
      const int R = kSkBW / S;               // weight rows reduced by this CTA
      const int n_local = sp * R + (t % R);  // row inside the 128-row tile
      const int mg = t / R;                  // this thread handles m = mg, mg + S, ...
      const int n = nt * kSkBW + n_local;
      const float bias = (g.epi.bias != nullptr && n < g.N) ? __bfloat162float(g.epi.bias[n]) : 0.f;
      const uint32_t base = smem_u32(partial) + static_cast<uint32_t>(n_local) * 4u;
      const int m_end = (g.M + S - 1) / S * S;  // whole warps run the same trip count (shuffle inside sk_store)
#pragma unroll 2
      for (int m = mg; m < m_end; m += S) {
        float acc = 0.f;
        if (m < g.M) {
          const uint32_t addr = base + static_cast<uint32_t>(m) * (kSkBW * 4u);
          for (int s = 0; s < S; ++s) acc += ld_dsmem_f32(addr, static_cast<uint32_t>(s));
        }
        sk_store(g, acc, bias, m, n, m < g.M && n < g.N, lane);
      }
    }
    __syncwarp();
    cluster_sync_all();  // nobody leaves (and frees its shared memory) while a peer may still be reading it
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
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 int l2_promotion);

static int skinny_mt(int M) { return M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256; }

// K splits (= cluster size, 1/2/4/8): as many CTAs as fit on the chip at once (two per SM up to MT = 128), but at
// least 8 k-blocks (128 KB of weights) per CTA so that the pipeline fill is amortised.
static int skinny_splits(int M, int N, int K, int hint) {
  const int k_blocks = (K + kSkBK - 1) / kSkBK;
  const int n_tiles = (N + kSkBW - 1) / kSkBW;
  int s = 1;
  if (hint > 0) {
    while (s * 2 <= hint && s < 8) s *= 2;
  } else {
    const int slots = num_sms() * (skinny_mt(M) <= 128 ? 2 : 1);
    while (s < 8 && n_tiles * s * 2 <= slots && k_blocks / (s * 2) >= 8) s *= 2;
  }
  while (s > 1 && s > k_blocks) s >>= 1;
  return s;
}

template <int MT, int STAGES>
static int launch_skinny(const CUtensorMap& tmW, const CUtensorMap& tmA, const SkinnyArgs& g, cudaStream_t stream) {
  using Cfg = SkinnyCfg<MT, STAGES>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<MT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
    configured = true;
  }
  const int n_out = g.epi.act == kActSwiGLU ? g.N / 2 : g.N;
  const double bytes = 2.0 * (static_cast<double>(g.M) * g.K + static_cast<double>(g.N) * g.K) +
                       static_cast<double>(g.M) * n_out * (g.epi.out_fp32 ? 4 : 2) +
                       (g.epi.residual ? 2.0 * g.M * g.N : 0.0);
  LaunchScope scope(kFamGemmSkinny, stream, bytes, 2.0 * g.M * g.N * g.K);
  B200_CUDA_OK(launch_ex(gemm_skinny_kernel<MT, STAGES>, dim3(static_cast<unsigned>(g.n_tiles * g.splits)),
                         dim3(kSkThreads), Cfg::kSmemBytes, stream, static_cast<unsigned>(g.splits),
                         g.epi.b_const != 0, tmW, tmA, g));
  return 0;
}

int gemm_skinny(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int M, int N, int K,
                const GemmEpilogue& epi, int splits_hint, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (M > 256) return fail(-2, "gemm_skinny: M %d > 256 (use gemm_bf16_tn)", M);
  if (K % 8 != 0) return fail(-2, "gemm_skinny: K (%d) must be a multiple of 8", K);
  if (epi.row_map != nullptr || epi.res_group != 0 || epi.n_peers != 0)
    return fail(-2, "gemm_skinny: row maps / peer stores are not supported");
  if (epi.act == kActSwiGLU && (N % 2 != 0 || epi.out_fp32 || epi.residual))
    return fail(-2, "gemm_skinny: SwiGLU epilogue needs even N, bf16 output and no residual");
  SkinnyArgs g;
  g.C = C;
  g.ldc = ldc;
  g.M = M;
  g.N = N;
  g.K = K;
  g.k_blocks = (K + kSkBK - 1) / kSkBK;
  g.n_tiles = (N + kSkBW - 1) / kSkBW;
  g.splits = skinny_splits(M, N, K, splits_hint);
  g.kb_per_split = (g.k_blocks + g.splits - 1) / g.splits;
  g.epi = epi;
  const int mt = skinny_mt(M);
  CUtensorMap tmW, tmA;
  B200_TRY(make_tmap_2d(&tmW, W, N, K, ldw, kSkBW, 2));
  B200_TRY(make_tmap_2d(&tmA, A, M, K, lda, mt, 2));
  switch (mt) {
    case 32: return launch_skinny<32, 4>(tmW, tmA, g, stream);
    case 64: return launch_skinny<64, 4>(tmW, tmA, g, stream);
    case 128: return launch_skinny<128, 3>(tmW, tmA, g, stream);
    default: return launch_skinny<256, 4>(tmW, tmA, g, stream);
  }
}

}  // namespace b200
