// Host-side shared declarations for libb200mmor.so (internal; the public C ABI is include/b200_mmor.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace b200 {

typedef __nv_bfloat16 bf16;

// error plumbing: every entry point returns 0 on success or a negative code and leaves a message
// retrievable through b200_last_error().
void set_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define B200_CUDA_OK(expr)                                                                         \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return b200::fail(-5, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                             __FILE__, __LINE__);                                  \
  } while (0)

#define B200_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != 0) return _rc;     \
  } while (0)

int num_sms();

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch for the decode step (opt-in: environment B200_PDL=1 or b200_set_pdl(1); off by
// default until it has been timed on hardware). With it on, the kernels of a decode step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: kernel n+1 is scheduled while kernel n drains, runs its
// prologue (and the GEMMs their first weight tiles, which do not depend on kernel n) and blocks in griddep_wait()
// (ptx.cuh) before it touches an activation. Every kernel launched through launch_ex(..., pdl_ok = true) MUST call
// griddep_wait() before its first dependent global access: a grid that skipped it could finish before its
// predecessor and release the grid after it too early.
// ---------------------------------------------------------------------------------------------
bool pdl_enabled();
void set_pdl(bool on);
// Tile widths 96 / 160 / 224 for the decode step's wide projections (stages.cu::decode_bn). Default 1 since round 2
// (timed on hardware: -4 % per decode step); B200_DECODE_TILES=0 / b200_set_option("decode_tiles", 0) restores the
// 128-column tiles. PDL stays off: it measured 4 % SLOWER (profiles/r2_decode_bench.json).
int decode_tiles_mode();       // 0 off, 1 widths for one CTA per SM, 2 two CTAs per SM (widths 64 / 96 / 128)
void set_decode_tiles(int mode);
// RoPE + KV append fused into the decode attention kernel (DecodeArgs::rope_k) whenever the step does not split the
// context; default on, B200_FUSED_ROPE=0 / b200_set_option("fused_rope", 0) runs rope_kv_kernel separately.
bool fused_rope_enabled();
void set_fused_rope(bool on);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             unsigned cluster_x, bool pdl_ok, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 0) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_ok && pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// launch accounting + optional per-family CUDA-event timing (errors.cu). Every host launcher opens a
// LaunchScope around its <<<>>>; the counter feeds bench.py's "gpu_launches", the event timing its
// "roofline" object (events are recorded on the launching stream; skipped while the stream is capturing).
// ---------------------------------------------------------------------------------------------
enum KernelFamily : int {
  kFamGemm = 0,     // tcgen05 GEMM, M > 128 (tensor bound)
  kFamGemmSkinny,   // tcgen05 GEMM, M <= 128 (decode: weight streaming, HBM bound)
  kFamFlashAttn,    // prefill / ViT / pooler attention
  kFamDecodeAttn,   // single-token attention over the KV cache (HBM bound)
  kFamNorm,         // LayerNorm / RMSNorm
  kFamRope,         // RoPE + KV-cache append
  kFamEmbed,        // embedding gather / pack rows
  kFamArgmax,
  kFamPatchify,
  kFamSegmask,
  kFamMisc,
  kFamTrain,        // backward / loss / optimizer kernels
  kFamPointCloud,   // PointTransformerV3 operators (ptv3.cu)
  kFamCount
};
struct LaunchScope {
  LaunchScope(int family, cudaStream_t stream, double alg_bytes = 0.0, double alg_flops = 0.0, int kernels = 1);
  ~LaunchScope();
  LaunchScope(const LaunchScope&) = delete;
  LaunchScope& operator=(const LaunchScope&) = delete;

 private:
  cudaStream_t stream_;
  int slot_;
};

// ---------------------------------------------------------------------------------------------
// GEMM  C[M,N] = epilogue(A[M,K] . B[N,K]^T)   (both operands K-major, bf16, fp32 accumulate)
// ---------------------------------------------------------------------------------------------
enum GemmAct : int {
  kActNone = 0,
  kActQuickGelu = 1,  // x * sigmoid(1.702 x)           (CLIP MLP)
  kActGeluErf = 2,    // 0.5 x (1 + erf(x / sqrt 2))    (BERT pooler, mm_projector)
  kActSwiGLU = 3      // columns are (gate, up) pairs: out[:, j] = silu(acc[:, 2j]) * acc[:, 2j+1]
};

struct GemmEpilogue {
  const bf16* bias = nullptr;      // [N] (pre-activation), optional
  const bf16* residual = nullptr;  // [M, ldr], added after the activation, optional (may alias C)
  const int* row_map = nullptr;    // [M] logical row -> output row, negative = drop, optional
  int ldr = 0;
  int res_group = 0;         // 0: residual row = row; else (row / res_group) * res_group_stride + row % res_group
  int res_group_stride = 0;  // (lets a compact (B*576)-row GEMM read its residual from a (B, S, D) tensor)
  int act = kActNone;
  int out_fp32 = 0;  // 0: C is bf16, 1: C is fp32
  int accumulate = 0;  // fp32 output only: C += result
  float scale = 1.f;   // accumulator is multiplied by this before bias / activation / residual (LoRA alpha / r)
  // programmatic dependent launch: operand B (the weights) is not written by the kernel launched before this one, so
  // its first tiles may be fetched before the grid dependency resolves. Only the decode step sets it.
  int b_const = 0;
  // 2: the decode step's weight-streaming shape -- half the pipeline depth so that two CTAs are resident per SM (an SM
  // streams cold weights at ~41 GB/s with two resident CTAs vs ~30 GB/s with one, profiles/r1_skinny_gemm_notes.md)
  // and the grid has 2 x SMs slots; BN <= 128 only (two CTAs x two accumulators must fit the 512 TMEM columns).
  int ctas_per_sm = 1;
  // fused GEMM -> all-gather: when n_peers > 0 every bf16 output vector is stored to the same offset of each
  // peer_c[p] (peer-mapped device pointers over NVLink, this rank's own buffer included) instead of C.
  int n_peers = 0;
  void* peer_c[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// bn_hint: 0 = choose automatically, else one of 32/64/128/256
int gemm_bf16_tn(const bf16* A, int lda, const bf16* B, int ldb, void* C, int ldc, int M, int N, int K,
                 const GemmEpilogue& epi, int bn_hint, cudaStream_t stream);

// General form: a_mn / b_mn = 1 means the operand is stored transposed ([K, M] / [K, N]) and is read in place as an
// MN-major tensor-core operand (backward GEMMs: dX = dY W uses b_mn, dW = dY^T X uses both).
int gemm_nf4_tn(const bf16* A, int lda, const uint8_t* codes, const float* absmax, void* C, int ldc, int M, int N, int K,
                const GemmEpilogue& epi, cudaStream_t stream);
int gemm_bf16_ex(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                 int K, const GemmEpilogue& epi, int bn_hint, cudaStream_t stream);

// Decode-step GEMM (gemm_skinny.cu): M <= 256, weights streamed once, K split across a thread-block cluster and
// reduced through distributed shared memory in split order. splits_hint: 0 = auto, else 1/2/4/8.
int gemm_skinny(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int M, int N, int K,
                const GemmEpilogue& epi, int splits_hint, cudaStream_t stream);

// training-side kernels (train.cu)
size_t weighted_ce_workspace_bytes(int Bn, int L);
int weighted_ce(const void* logits, int is_fp32, long long ld, const long long* labels, const float* vocab_w, int Bn,
                int L, int V, float grad_scale, void* dlogits, long long ldd, float* loss_out, void* workspace,
                size_t workspace_bytes, cudaStream_t stream);
size_t grad_norm_workspace_bytes();
int grad_sq_norm(const void* grad, int grad_fp32, long long n, int accumulate, float max_norm, float* out2,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t colsum_workspace_bytes(int N);
int colsum(const bf16* dy, long long ld, int M, int N, int accumulate, float* out, void* workspace,
           size_t workspace_bytes, cudaStream_t stream);
int act_forward(const bf16* z, bf16* y, long long n, int act, cudaStream_t stream);
int group_sum(const bf16* x, int groups, long long slab, int accumulate, float* out, cudaStream_t stream);
int gather_add_rows(const bf16* src, long long ld_src, const int* row_map, const bf16* add, int period, int rows, int D,
                    bf16* out, cudaStream_t stream);
int swiglu_forward(const bf16* z, bf16* h, long long n_out, cudaStream_t stream);
int rope_kv_backward(bf16* dqkv, const int* kv_start, const float* cos_t, const float* sin_t, int max_pos,
                     const bf16* dk_cache, const bf16* dv_cache, int Bn, int H, int Lq, int cap, cudaStream_t stream);
int act_backward(const bf16* z, const bf16* dy, bf16* dz, long long n_out, int act, cudaStream_t stream);
size_t norm_backward_workspace_bytes(int M, int D);
int norm_backward(const bf16* x, const bf16* dy, const bf16* gamma, float eps, int M, int D, int rms, const bf16* add,
                  bf16* dx, float* dgamma, float* dbeta, int accumulate, void* workspace, size_t workspace_bytes,
                  cudaStream_t stream);
int adamw_step(float* master, bf16* param, const void* grad, int grad_fp32, float* m, float* v, long long n, float lr,
               float beta1, float beta2, float eps, float weight_decay, int step, const float* clip_coef,
               cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// norms (norm.cu)
// ---------------------------------------------------------------------------------------------
int layernorm(const bf16* x, long long ldx, const int* row_map, const bf16* add, int period, const bf16* gamma,
              const bf16* beta, float eps, bf16* out, long long ldo, int M, int D, int out_group,
              int out_group_stride, cudaStream_t stream);
int rmsnorm(const bf16* x, long long ldx, const bf16* w, float eps, bf16* out, long long ldo, int M, int D,
            cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// attention (attention.cu)
// ---------------------------------------------------------------------------------------------
struct AttnArgs {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* o;
  // element strides: batch, row (token), head
  long long q_bs, q_rs, q_hs;
  long long k_bs, k_rs, k_hs;
  long long v_bs, v_rs, v_hs;
  long long o_bs, o_rs, o_hs;
  int B, H, Lq, Lk;
  const int* kv_start;  // [B] first valid key (left padding) or null
  const int* kv_len;    // [B] number of valid keys counted from 0 (key padding) or null
  int causal;           // key j visible to query i iff j <= i + (Lk - Lq)
  float scale_log2;     // softmax scale * log2(e)
  float* lse;           // optional [B, H, Lq] fp32: natural-log sum-exp of the scaled scores per query row (saved for
                        // the backward pass; -inf for fully masked rows)
};
int flash_attn(const AttnArgs& a, int head_dim, cudaStream_t stream);

// backward (attention_bwd_sm100.cu): the forward call's AttnArgs (q, k, v, o, masks, scale) plus dO, the saved lse and
// the gradient outputs; strides as in AttnArgs
struct AttnBwdArgs {
  const bf16* d_o;
  long long do_bs, do_rs, do_hs;
  const float* lse;  // [B, H, Lq] from the forward
  bf16* dq;
  bf16* dk;
  bf16* dv;
  long long dq_bs, dq_rs, dq_hs;
  long long dk_bs, dk_rs, dk_hs;
  long long dv_bs, dv_rs, dv_hs;
};
size_t flash_attn_bwd_workspace_bytes(int B, int H, int Lq);
int flash_attn_bwd(const AttnArgs& f, const AttnBwdArgs& g, int head_dim, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream);

struct DecodeArgs {
  const bf16* q;  // [B, q_rs] with head h at h*128
  long long q_rs;
  const bf16* kc;  // [B][H][cap][128]
  const bf16* vc;
  bf16* o;  // [B, o_rs]
  long long o_rs;
  int B, H, cap;
  int ctx;              // number of cache slots in use (keys 0..ctx-1) ...
  const int* ctx_dev;   // ... or, when non-null, *ctx_dev + ctx_add (decode under a CUDA graph); ctx is then the
  int ctx_add;          //     upper bound used for launch sizing
  const int* kv_start;  // [B] first valid slot (left padding), or null
  const int* finished;  // [B] or null: rows with finished[b] != 0 (EOS already emitted, HF greedy pads them from then
                        // on) read no cache at all and get o = 0 -- their logits are never used
  float scale_log2;
  int splits;  // 0 = choose
  float* part_o;
  float* part_ml;
  // Fused RoPE + KV append (the decode step, splits == 1 only; all null / 0 otherwise): the kernel of (row b, head h)
  // rotates q and the step's new key itself (position = slot - kv_start[b], clamped like rope_kv_kernel), appends the
  // rotated key and the value to the cache at slot ctx - 1, and scores / accumulates that one key from shared memory
  // -- what rope_kv_kernel + a re-read would do, in one launch less per layer.
  const bf16* rope_k;   // [B, q_rs]: the new key of head h at h * 128 (un-rotated), or null = no fusion
  const bf16* rope_v;   // [B, q_rs]: the new value
  bf16* kc_w;           // the cache again, writable
  bf16* vc_w;
  const float* cos_t;   // [max_pos, 64]
  const float* sin_t;
  int max_pos;
};
size_t decode_attn_workspace_bytes(int B, int H, int splits);
int decode_attn_pick_splits(int B, int H, int ctx);
int decode_attn(DecodeArgs a, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// misc.cu
// ---------------------------------------------------------------------------------------------
int patchify_im2col(const bf16* pixels, bf16* cols, int N, int C, int S, int P, int Kpad, cudaStream_t stream);
int rope_kv_write(bf16* qkv, const int* kv_start, const float* cos_t, const float* sin_t, int max_pos, bf16* kc,
                  bf16* vc, int B, int H, int Lq, int slot0, const int* slot0_dev, int cap, cudaStream_t stream);
int bump_counters(int* state, cudaStream_t stream);
int embed_rows(const int* ids, const bf16* table, bf16* out, long long ldo, int rows, int D, int vocab,
               cudaStream_t stream);
int argmax_rows(const void* logits, int is_fp32, long long ld, int rows, int V, int* out_tok, int* finished, int eos_id,
                int pad_id, int* history, int hist_ld, int step, const int* step_dev, cudaStream_t stream);
size_t segmask_workspace_bytes(int n_maps);
int segmask_forward(const uint8_t* cls, int n_maps, const bf16* emb, const bf16* const* conv_w,
                    const bf16* const* conv_b, bf16* out, long long out_ld, const int* out_row_map, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream);

// image preprocessing (preprocess.cu)
size_t preprocess_workspace_bytes(int n, int canvas_h, int res_w);
int preprocess_images(const uint8_t* img, int n, int H, int W, int pad, const uint8_t* bg, const int* bx, const int* kx,
                      int ksize_x, const int* by, const int* ky, int ksize_y, int res_h, int res_w, int out,
                      const float* mean, const float* stdv, bf16* dst, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream);

}  // namespace b200
