/*
 * b200_mmor.h -- C ABI of libb200mmor.so, the B200 (sm_100a) implementation of MM2SG's multimodal hot path
 * (reference: egeozsoy/MM-OR, scene_graph_generation/LLaVA/llava; paths below are relative to that directory).
 *
 * The reference has no FFI: the path sits behind Python objects that call HF transformers / torch ops
 * (SURVEY.md 8b). These entry points are what the Python mirror of that interface (mm_or_b200/model/*.py) binds
 * with ctypes. Each entry point cites the reference call site whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless documented as host; bf16 tensors are row-major and dense unless a
 *     leading dimension is given; sizes are in elements; weight structs are plain host structs of device pointers.
 *   - every function returns 0 on success, a negative code on failure (-2 bad argument, -5 CUDA error,
 *     -6 driver entry point missing); b200_last_error() returns the message for the calling thread.
 *   - functions enqueue work on `stream` (a cudaStream_t passed as void*) and never synchronise or allocate;
 *     scratch memory is passed in as (workspace, workspace_bytes) and sized by the matching *_workspace_bytes.
 *   - re-entrant per stream; the only global state is lazily initialised kernel attributes.
 */
#ifndef B200_MMOR_H_
#define B200_MMOR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200_stream_t; /* cudaStream_t */

const char* b200_last_error(void);
int b200_abi_version(void);
/* sizeof() of the weight structs below, in declaration order (0 = b200_vit_layer ... 8 = b200_kv_cache); lets a
 * binding verify its struct mirrors. Returns 0 for an unknown index. */
size_t b200_sizeof_struct(int which);

/* Launch accounting (no reference counterpart: the reference launches its kernels from Python, op by op).
 * b200_launch_count: kernels this library has launched (or captured into a CUDA graph) in this process.
 * b200_prof_enable(1) starts bracketing every launch with CUDA events on the launching stream, grouped by kernel
 * family (b200_prof_family_name); b200_prof_collect waits for the recorded events (the one call here that blocks
 * the host) and returns, per family, the summed device milliseconds, the algorithmic bytes / FLOPs the launchers
 * were asked for, and the launch count. Arrays must hold b200_prof_family_count() entries. Launches made while the
 * stream is capturing a CUDA graph are counted but not timed. */
long long b200_launch_count(void);
int b200_prof_enable(int on);
int b200_prof_family_count(void);
const char* b200_prof_family_name(int family);
int b200_prof_collect(double* ms, double* alg_bytes, double* alg_flops, long long* launches);

/* Process-wide tuning switches of the decode step (no reference counterpart: the reference's decode step is ~1900
 * separate eager launches from Python). Set them here or through the environment variable named below BEFORE the
 * step is launched / captured into a CUDA graph; results are bit-identical in every mode (same arithmetic in the same
 * order per output element: "pdl" only changes scheduling, "decode_tiles" only which CTA owns an output column).
 * Defaults as timed on B200: "pdl" 0 (it measured 4 % slower), "decode_tiles" 1 (4 % faster).
 *   "pdl"          (B200_PDL=1): programmatic dependent launch -- the step's kernels are launched with programmatic
 *                  stream serialization, so each kernel's prologue, and the first weight tiles of the GEMMs (which do
 *                  not depend on the previous kernel), overlap the tail of the kernel before it.
 *   "decode_tiles" (B200_DECODE_TILES=0|1|2): the wide projections (qkv, gate_up, lm_head) choose their weight-tile width
 *                  so that the tiles cover the CTA slots in as few waves as possible instead of always 128 columns.
 *                  1: one CTA per SM, widths {96, 128, 160, 224, 256} (12288 / 96 = 128 tiles, 22016 / 160 = 138,
 *                  32000 / 224 = 143 on 148 SMs); 2: two CTAs per SM with half-depth rings, widths {64, 96, 128}
 *                  (192 / 230 / 250 tiles on 296 slots).
 *   "fused_rope"   (B200_FUSED_ROPE=0|1, default 1): RoPE of q / k and the KV-cache append of a decode step run inside
 *                  the decode attention kernel (one launch less per layer) whenever the step has enough (row, head)
 *                  pairs to fill the chip without splitting the context; 0 runs the separate RoPE kernel.
 * b200_set_option returns 0, or -2 for an unknown name / value; b200_get_option returns the value, or -2. */
int b200_set_option(const char* name, int value);
int b200_get_option(const char* name);
/* The weight-tile width "decode_tiles" picks for a projection with `n` output features at `rows` sequences on a
 * device with `sms` SMs and 1 or 2 CTAs per SM (host-side arithmetic only; exposed so that the policy can be pinned
 * without a GPU). */
int b200_decode_tile_width(int rows, int n, int sms, int ctas_per_sm);

/* ============================================================================================================
 * Operator level (used by the stage entry points below and by the parity tests)
 * ========================================================================================================== */

/* Dense linear:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)      bf16 in, fp32 accumulate (tcgen05 / TMEM)
 * Replaces every nn.Linear on the path: CLIP q/k/v/out_proj, fc1, fc2 (model/multimodal_encoder/clip_encoder.py:48 ->
 * HF CLIPEncoderLayer), BERT pooler dense layers (model/multimodal_projector/builder.py:173), mm_projector
 * (model/llava_arch.py:182), Llama q/k/v/o/gate/up/down_proj and lm_head (model/language_model/llava_llama.py:93).
 *   act: 0 none, 1 quick_gelu, 2 gelu(erf), 3 SwiGLU over interleaved (gate, up) column pairs (C has N/2 columns)
 *   bias [N] / residual [M, ldr] / row_map [M] (output row per logical row, <0 drops the row) may be NULL
 *   out_fp32: C is float instead of bf16.  bn_hint: 0 = auto tile width, or 32/64/96/128/160/224/256; -64/-96/-128 =
 *   that width on half-depth rings with two CTAs per SM (the decode step's "decode_tiles" = 2 shape). */
int b200_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                   const void* bias, const void* residual, int ldr, const int32_t* row_map, int act, int out_fp32,
                   int bn_hint, b200_stream_t stream);

/* Same contract for M <= 256 rows (the single-token decode step, model/llava_arch.py:192-201 -> HF LlamaDecoderLayer /
 * lm_head), HBM bound: weights are the tensor-core M operand and are streamed exactly once; K is split over a
 * thread-block cluster and reduced through distributed shared memory in split order (deterministic). No row_map.
 * splits: 0 = auto, else 1/2/4/8. */
int b200_gemm_bf16_skinny(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                          const void* bias, const void* residual, int ldr, int act, int out_fp32, int splits,
                          b200_stream_t stream);

/* y = LayerNorm(gather(x)[row] + add[row % period]) * gamma + beta  (HF CLIP LayerNorm eps 1e-5, BERT eps 1e-12).
 * row_map (source row per output row, <0 = zero row), add may be NULL. */
int b200_layernorm(const void* x, int64_t ldx, const int32_t* row_map, const void* add, int period, const void* gamma,
                   const void* beta, float eps, void* out, int64_t ldo, int M, int D, b200_stream_t stream);

/* y = w * x * rsqrt(mean(x^2) + eps)   (HF LlamaRMSNorm) */
int b200_rmsnorm(const void* x, int64_t ldx, const void* w, float eps, void* out, int64_t ldo, int M, int D,
                 b200_stream_t stream);

/* Fused attention softmax(Q K^T * scale + mask) V with online softmax (scores never reach HBM).
 * Strides are in elements: *_bs batch, *_rs row/token, *_hs head. head_dim 64 or 128.
 * kv_start[B] (first visible key: left padding), kv_len[B] (keys >= kv_len masked: key padding) may be NULL;
 * causal != 0: key j visible to query i iff j <= i + (Lk - Lq).
 * Replaces HF CLIPAttention / BertSelfAttention / LlamaAttention eager bmm-softmax-bmm and the training-time
 * flash_attn_varlen_qkvpacked_func (train/llama_flash_attn_monkey_patch.py:78-89). */
int b200_flash_attention(const void* q, int64_t q_bs, int64_t q_rs, int64_t q_hs, const void* k, int64_t k_bs,
                         int64_t k_rs, int64_t k_hs, const void* v, int64_t v_bs, int64_t v_rs, int64_t v_hs, void* o,
                         int64_t o_bs, int64_t o_rs, int64_t o_hs, int B, int H, int Lq, int Lk, int head_dim,
                         const int32_t* kv_start, const int32_t* kv_len, int causal, float scale,
                         b200_stream_t stream);

/* Training-time variant: additionally writes lse [B, H, Lq] fp32, the natural-log sum-exp of every query row's scaled
 * scores (-inf for fully masked rows), which the attention backward pass needs (FlashAttention-2 saves the same,
 * train/llama_flash_attn_monkey_patch.py:78-89 -> flash_attn_varlen_qkvpacked_func). */
int b200_flash_attention_lse(const void* q, int64_t q_bs, int64_t q_rs, int64_t q_hs, const void* k, int64_t k_bs,
                             int64_t k_rs, int64_t k_hs, const void* v, int64_t v_bs, int64_t v_rs, int64_t v_hs,
                             void* o, int64_t o_bs, int64_t o_rs, int64_t o_hs, int B, int H, int Lq, int Lk,
                             int head_dim, const int32_t* kv_start, const int32_t* kv_len, int causal, float scale,
                             float* lse, b200_stream_t stream);

/* Attention backward (autograd of the three attention sites; for the Llama layers the reference runs FlashAttention-2's
 * backward behind train/llama_flash_attn_monkey_patch.py:78-89). P is recomputed from lse; dO has O's layout and
 * dq / dk / dv have the layout (strides) of q / k / v. No atomics: bit-deterministic. workspace: *_workspace_bytes. */
size_t b200_flash_attention_bwd_workspace_bytes(int B, int H, int Lq);
int b200_flash_attention_bwd(const void* q, int64_t q_bs, int64_t q_rs, int64_t q_hs, const void* k, int64_t k_bs,
                             int64_t k_rs, int64_t k_hs, const void* v, int64_t v_bs, int64_t v_rs, int64_t v_hs,
                             const void* o, const void* d_o, int64_t o_bs, int64_t o_rs, int64_t o_hs, const float* lse,
                             void* dq, void* dk, void* dv, int B, int H, int Lq, int Lk, int head_dim,
                             const int32_t* kv_start, const int32_t* kv_len, int causal, float scale, void* workspace,
                             size_t workspace_bytes, b200_stream_t stream);

/* Single-token attention against the KV cache [B][H][cap][128] (decode step, model/llava_arch.py:192-201 + HF
 * LlamaAttention with past_key_values). ctx = slots in use; kv_start as above. splits: 0 = auto. */
size_t b200_decode_attention_workspace_bytes(int B, int H, int ctx);
int b200_decode_attention(const void* q, int64_t q_rs, const void* k_cache, const void* v_cache, void* o, int64_t o_rs,
                          int B, int H, int cap, int ctx, const int32_t* kv_start, float scale, int splits,
                          void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* RoPE (theta via the fp32 cos/sin tables [max_pos, 64]) on q (in place) and k, plus KV-cache append of k and v.
 * qkv: [B*Lq, 3*H*128]. Token (b, l) goes to slot slot0 + l, rotary position slot - kv_start[b]. */
int b200_rope_kv_write(void* qkv, const int32_t* kv_start, const float* cos_table, const float* sin_table, int max_pos,
                       void* k_cache, void* v_cache, int B, int H, int Lq, int slot0, int cap, b200_stream_t stream);

/* out[r] = table[ids[r]] (ids >= 0), zeros (ids == -1), untouched (ids <= -2): embed_tokens gather + zero pad rows
 * of the multimodal pack (model/llava_arch.py:262, :317-338). */
int b200_embed_rows(const int32_t* ids, const void* table, void* out, int64_t ldo, int rows, int D, int vocab,
                    b200_stream_t stream);

/* Greedy argmax per row, lowest index wins ties. finished/history may be NULL. */
int b200_argmax(const void* logits, int is_fp32, int64_t ld, int rows, int V, int32_t* out_tok, int32_t* finished,
                int eos_id, int pad_id, b200_stream_t stream);

/* im2col for the CLIP patch embedding conv (k = s = P): pixels [N,C,S,S] -> cols [N*(S/P)^2, Kpad]. */
int b200_patchify(const void* pixels, void* cols, int N, int C, int S, int P, int Kpad, b200_stream_t stream);

/* ============================================================================================================
 * Stage level
 * ========================================================================================================== */

/* ---- CLIP ViT tower: CLIPVisionTower.forward + feature_select (model/multimodal_encoder/clip_encoder.py:29-51) ---- */
typedef struct {
  const void *ln1_w, *ln1_b;
  const void *qkv_w, *qkv_b; /* [3*hidden, hidden] = cat(q*head_dim^-0.5, k, v), bias likewise */
  const void *out_w, *out_b;
  const void *ln2_w, *ln2_b;
  const void *fc1_w, *fc1_b;
  const void *fc2_w, *fc2_b;
} b200_vit_layer;

typedef struct {
  int hidden, heads, ffn, image_size, patch, kpad;
  int n_layers;            /* layers to execute: hidden_states[select_layer] (select_layer=-2 of 24 -> 23) */
  float ln_eps;
  const void* patch_w;     /* [hidden, kpad] patch_embedding.weight flattened (c, ky, kx), zero padded */
  const void* pos_cls;     /* [1 + (image/patch)^2, hidden] position_embedding, row 0 += class_embedding */
  const void *pre_ln_w, *pre_ln_b;
  const b200_vit_layer* layers; /* host array [n_layers] */
} b200_vit_weights;

size_t b200_vit_workspace_bytes(const b200_vit_weights* w, int n_img);
/* pixels [n_img, 3, S, S] bf16 -> hidden [n_img, 1 + P, hidden] bf16 (CLS row kept; callers drop it by row maps). */
int b200_vit_forward(const b200_vit_weights* w, const void* pixels, void* hidden_out, int n_img, void* workspace,
                     size_t workspace_bytes, b200_stream_t stream);

/* ---- image pooler: ImageEmbeddingPooler.forward BERT part (model/multimodal_projector/builder.py:169-175) ---- */
typedef struct {
  const void *qkv_w, *qkv_b; /* [3*hidden, hidden] = cat(query, key, value) */
  const void *ao_w, *ao_b, *ao_ln_w, *ao_ln_b;     /* attention.output.dense + LayerNorm */
  const void *fc1_w, *fc1_b;                       /* intermediate.dense */
  const void *fc2_w, *fc2_b, *out_ln_w, *out_ln_b; /* output.dense + LayerNorm */
} b200_bert_layer;

typedef struct {
  int hidden, heads, ffn, n_layers, max_pos;
  float ln_eps;
  const void* pos_type;  /* [max_pos, hidden] position_embeddings + token_type_embeddings[0] */
  const void *emb_ln_w, *emb_ln_b;
  const b200_bert_layer* layers; /* host array */
} b200_pooler_weights;

size_t b200_pooler_workspace_bytes(const b200_pooler_weights* w, int B, int S);
/* src [*, hidden] rows addressed through gather_map [B*S] (<0 = zero padding row, model/llava_arch.py:143-170);
 * kv_len[B] = valid tokens per sample; keeps the first `keep` (576) tokens of every sample and writes them to
 * out[b * out_tokens + t] (out is [B, out_tokens, hidden], extra-modality tokens follow at t >= keep). */
int b200_pooler_forward(const b200_pooler_weights* w, const void* src, int64_t src_ld, const int32_t* gather_map,
                        const int32_t* kv_len, int B, int S, int keep, void* out, int out_tokens, void* workspace,
                        size_t workspace_bytes, b200_stream_t stream);

/* ---- seg-mask tokens: SegmentationMapFeatureExtractor.forward
 *      (model/multimodal_projector/segmentation_map_feature_extractor.py:53-75) ---- */
typedef struct {
  const void* emb;       /* [30, 8] */
  const void* conv_w[5]; /* [Cout, Cin, 3, 3], channels 8-64-128-256-512-1024 */
  const void* conv_b[5];
} b200_segmask_weights;
size_t b200_segmask_workspace_bytes(int n_maps);
/* cls [n_maps, 32, 32] uint8 -> out[out_row_map[n]][0..1024) (row stride out_ld) */
int b200_segmask_forward(const b200_segmask_weights* w, const uint8_t* cls, int n_maps, void* out, int64_t out_ld,
                         const int32_t* out_row_map, void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* ---- mm_projector + multimodal token pack (model/llava_arch.py:182 and :235-338) ---- */
typedef struct {
  int in_dim, hidden;
  const void *w0, *b0; /* Linear(in_dim, hidden) */
  const void *w2, *b2; /* Linear(hidden, hidden) */
} b200_projector_weights;
size_t b200_projector_workspace_bytes(const b200_projector_weights* w, int n_tokens);
/* tokens [n_tokens, in_dim] -> GELU MLP -> rows scattered into embeds[row_map[i]] (row stride = hidden);
 * then text rows gathered from embed_tokens by text_ids [n_rows] (see b200_embed_rows). text_ids may be NULL. */
int b200_projector_pack(const b200_projector_weights* w, const void* tokens, int n_tokens, const int32_t* row_map,
                        const int32_t* text_ids, const void* embed_table, int vocab, void* embeds, int n_rows,
                        void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* Multi-GPU variant (BASELINE.json north_star / SURVEY.md 8e: "NCCL all-gather of visual tokens over NVLink before LLM
 * fusion"; the reference itself has no multi-GPU inference path). The second projector GEMM stores every output
 * tile to the same slot of all `n_peers` (<= 8) peer-mapped buffers -- peers[p] + slot_offset_bytes + row * hidden * 2 --
 * so the all-gather rides on the GEMM epilogue over NVLink instead of running as a separate collective. peers is a
 * HOST array of device pointers valid on this GPU (peer mappings, own buffer included). The caller orders readers
 * behind the writers (a barrier over the symmetric-memory signal pads after this call). */
int b200_projector_gather(const b200_projector_weights* w, const void* tokens, int n_tokens, void* const* peers,
                          int n_peers, size_t slot_offset_bytes, void* workspace, size_t workspace_bytes,
                          b200_stream_t stream);

/* ---- Llama decoder (HF LlamaForCausalLM.forward reached from model/language_model/llava_llama.py:93) ---- */
typedef struct {
  const void* attn_norm;
  const void* qkv_w;     /* [3*hidden, hidden] = cat(q_proj, k_proj, v_proj) */
  const void* o_w;
  const void* mlp_norm;
  const void* gate_up_w; /* [2*ffn, hidden], rows interleaved: 2j = gate_proj[j], 2j+1 = up_proj[j] */
  const void* down_w;    /* [hidden, ffn] */
} b200_llama_layer;

typedef struct {
  int hidden, heads, ffn, n_layers, vocab, max_pos;
  float rms_eps;
  const b200_llama_layer* layers; /* host array */
  const void* final_norm;
  const void* lm_head;      /* [vocab, hidden] */
  const void* embed_tokens; /* [vocab, hidden] */
  const float* rope_cos;    /* [max_pos, 64] fp32 */
  const float* rope_sin;
} b200_llama_weights;

typedef struct {
  void* k;              /* [n_layers][batch][heads][cap][128] bf16 */
  void* v;
  int64_t layer_stride; /* elements between layers (= batch_total * heads * cap * 128) */
  int cap;
} b200_kv_cache;

size_t b200_llama_prefill_workspace_bytes(const b200_llama_weights* w, int B, int L, int all_logits);
/* x [B*L, hidden]: packed inputs_embeds, overwritten with the final hidden states. Tokens of sample b occupy cache
 * slots [0, L); kv_start[b] = number of left-pad rows (NULL = none); kv_len[b] = valid rows for right padding
 * (NULL = L). cache.k/v must already point at this batch slice. logits: [B, vocab] for the last position
 * (all_logits = 0) or [B*L, vocab] (all_logits = 1); logits_fp32 selects float output. */
int b200_llama_prefill(const b200_llama_weights* w, void* x, const int32_t* kv_start, const int32_t* kv_len,
                       const b200_kv_cache* cache, int B, int L, void* logits, int all_logits, int logits_fp32,
                       void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* Prefill that continues a cache (prefix-KV reuse of the online mode, scene_graph_prediction_model.py:140-199: every
 * prompt starts with the same system text): slots [0, q0) of every row already hold keys / values -- the caller copied
 * the shared prefix there, row b's copy starting at kv_start[b] --; x [B*Lq, hidden] are the tokens of slots
 * [q0, q0 + Lq), which attend to all q0 + Lq keys. q0 = 0 is b200_llama_prefill. Workspace as for (B, Lq). */
int b200_llama_prefill_from(const b200_llama_weights* w, void* x, const int32_t* kv_start, const int32_t* kv_len,
                            const b200_kv_cache* cache, int B, int Lq, int q0, void* logits, int all_logits,
                            int logits_fp32, void* workspace, size_t workspace_bytes, b200_stream_t stream);

size_t b200_llama_decode_workspace_bytes(const b200_llama_weights* w, int B, int cap);
/* One greedy decode step for B sequences (model/llava_arch.py:192-201 + HF greedy_search):
 * tokens [B] int32 (in: token to feed, out: argmax token), state[2] int32 device = {slots in use, step index} --
 * read on the device and incremented at the end, so the identical launch sequence can be replayed from a CUDA graph.
 * finished [B] int32 (HF unfinished_sequences semantics: finished rows emit pad_id; rows finish on eos_id), may be
 * NULL; history [B, hist_ld] int32 receives the produced token at column state[1]; logits_out (bf16 [B, vocab]) may
 * be NULL to use workspace. ctx_bound = upper bound of slots in use for this call (launch sizing). */
int b200_llama_decode_step(const b200_llama_weights* w, int32_t* tokens, int32_t* state, const int32_t* kv_start,
                           const b200_kv_cache* cache, int B, int ctx_bound, int32_t* finished, int eos_id, int pad_id,
                           int32_t* history, int hist_ld, void* logits_out, void* workspace, size_t workspace_bytes,
                           b200_stream_t stream);

/* ============================================================================================================
 * Image preprocessing in front of the tower: mm_utils.py:14-40 expand2square + process_images ->
 * HF CLIPImageProcessor.preprocess (PIL BICUBIC resize, centre crop, rescale, normalise) -> bf16.
 * images: device [n, H, W, 3] uint8 RGB. pad_to_square: paste on a max(H, W) square of background_rgb (HOST uint8[3])
 * like expand2square. The canvas is resized to resized_h x resized_w with Pillow's two-pass antialiased bicubic --
 * bounds_* [out, 2] int32 (first source index, taps) and coeffs_* [out, ksize_*] int32 22-bit fixed point are the
 * DEVICE tables of Pillow's precompute_coeffs / normalize_coeffs_8bpc (NULL for an axis that keeps its size) -- then
 * centre-cropped to crop x crop, rescaled by 1/255, normalised with mean / std (HOST float[3]) in float32 and written
 * as bf16 [n, 3, crop, crop]. Bit-exact with the host path (integer resampling, IEEE float tail).
 * ========================================================================================================== */
size_t b200_preprocess_workspace_bytes(int n, int canvas_h, int resized_w);
int b200_preprocess_images(const uint8_t* images, int n, int H, int W, int pad_to_square, const uint8_t* background_rgb,
                           const int32_t* bounds_x, const int32_t* coeffs_x, int ksize_x, const int32_t* bounds_y,
                           const int32_t* coeffs_y, int ksize_y, int resized_h, int resized_w, int crop,
                           const float* mean, const float* std, void* out, void* workspace, size_t workspace_bytes,
                           b200_stream_t stream);

/* ============================================================================================================
 * Training-side operators (fine-tune step)
 * ========================================================================================================== */

/* Backward building blocks of every nn.Linear / activation / norm on the path (autograd of the modules cited above;
 * the reference gets them from torch autograd + cuBLAS).
 * b200_gemm_bf16_ex: the GEMM of b200_gemm_bf16 with optionally TRANSPOSED operands read in place as MN-major
 * tensor-core operands: a_transposed: A is stored [K, M]; w_transposed: W is stored [K, N]. With Y = X W^T:
 *     dX[M,K] = dY[M,N] W[N,K]      -> (A = dY, W = W with w_transposed = 1, "N" = K, "K" = N)
 *     dW[N,K] = dY^T[N,M] X[M,K]    -> (A = dY with a_transposed = 1, W = X with w_transposed = 1, "M" = N, "N" = K, "K" = M)
 * accumulate != 0 (fp32 output only): C += result (gradient accumulation). scale multiplies the accumulator before
 * bias / activation / residual (the LoRA alpha / r factor; 1 otherwise).
 * b200_colsum: out[n] (+)= sum_m dy[m, n] (bias gradients), deterministic.
 * b200_act_backward: dz = dy * act'(z) from the saved pre-activation z; act 1 quick_gelu, 2 gelu(erf), 3 SwiGLU
 * (z = interleaved (gate, up) [M, 2F], dy [M, F], dz [M, 2F]); n_out = elements of dy.
 * b200_norm_backward: LayerNorm (rms = 0) / RMSNorm (rms = 1) backward from the saved input x: dx bf16, dgamma / dbeta
 * fp32 [D] ((+)= with accumulate; dbeta ignored for RMSNorm). `add` (bf16 [M, D], may be NULL, may alias dx) is added
 * to dx: the gradient that reaches x through the residual connection around the norm. D in {512, 1024, 4096}. */
/* C[M, N] (bf16) = scale * A[M, K] . dequant(W)^T (+ residual): W [N, K] given in 4-bit NormalFloat storage -- `codes`
 * N * K / 2 bytes (row-major, two codes per byte, even element in the high nibble, b200_nf4_quantize's layout) and
 * `absmax` one fp32 per block of 64 consecutive K values (N * K / 64 floats). The dequantisation bitsandbytes'
 * Linear4bit performs in front of every matmul of the QLoRA recipe (reference: train/train.py:1098-1114) runs inside the
 * GEMM's producer side: the packed weight is all that lives in HBM. Bit-identical to b200_gemm_bf16 on
 * b200_nf4_dequantize(codes, absmax). K % 64 == 0, codes 16-byte aligned. */
int b200_gemm_nf4(const void* A, int lda, const uint8_t* codes, const float* absmax, void* C, int ldc, int M, int N, int K,
                  const void* residual, int ldr, float scale, b200_stream_t stream);
int b200_gemm_bf16_ex(const void* A, int lda, int a_transposed, const void* W, int ldw, int w_transposed, void* C, int ldc,
                      int M, int N, int K, const void* bias, const void* residual, int ldr, int act, int out_fp32,
                      int accumulate, float scale, int bn_hint, b200_stream_t stream);
size_t b200_colsum_workspace_bytes(int N);
int b200_colsum(const void* dy, int64_t ld, int M, int N, int accumulate, float* out, void* workspace,
                size_t workspace_bytes, b200_stream_t stream);
int b200_act_backward(const void* z, const void* dy, void* dz, int64_t n_out, int act, b200_stream_t stream);
/* b200_swiglu_forward: h[m, j] = silu(z[m, 2j]) * z[m, 2j+1] from the stored pre-activation (training forward; the
 * inference path fuses it into the GEMM epilogue). b200_rope_kv_backward: backward of b200_rope_kv_write -- dq (in
 * place in dqkv) and dk (from the head-major cache layout) are rotated by -theta, dk / dv are gathered back into the
 * token-major dqkv [B*Lq, 3*H*128]. */
int b200_swiglu_forward(const void* z, void* h, int64_t n_out, b200_stream_t stream);
/* y = act(z) elementwise from a stored pre-activation (act 1 quick_gelu, 2 gelu(erf)): training forward of the CLIP
 * MLP, the pooler FFN and the mm_projector. b200_group_sum: out[i] (+)= sum_g x[g * slab + i], the gradient of a table
 * that is broadcast over the batch (the pooler's position + token-type embeddings). */
int b200_act_forward(const void* z, void* y, int64_t n, int act, b200_stream_t stream);
/* out[r] = src[row_map[r]] (zeros when < 0) + add[r % period]: the pooler's pre-LayerNorm embedding sum
 * (model/llava_arch.py:143-170 pad_embeddings + HF BertEmbeddings), materialised for the backward pass. */
int b200_gather_add_rows(const void* src, int64_t ld_src, const int32_t* row_map, const void* add, int period, int rows,
                         int D, void* out, b200_stream_t stream);
int b200_group_sum(const void* x, int groups, int64_t slab, int accumulate, float* out, b200_stream_t stream);
int b200_rope_kv_backward(void* dqkv, const int32_t* kv_start, const float* cos_table, const float* sin_table,
                          int max_pos, const void* dk_cache, const void* dv_cache, int B, int H, int Lq, int cap,
                          b200_stream_t stream);
size_t b200_norm_backward_workspace_bytes(int M, int D);
int b200_norm_backward(const void* x, const void* dy, const void* gamma, float eps, int M, int D, int rms,
                       const void* add, void* dx, float* dgamma, float* dbeta, int accumulate, void* workspace, size_t workspace_bytes,
                       b200_stream_t stream);

/* LLaVATrainer.compute_loss (train/llava_trainer.py:136-174): shifted cross-entropy of `logits` [B*L, V] against the
 * UNshifted modified_labels [B, L] (row (b, l) is scored against label (b, l + 1); -100 = ignored) with per-class
 * weights vocab_weight [V] (nn.CrossEntropyLoss(weight=...) semantics, NULL = unweighted):
 *     loss_out[0] = sum_i w[y_i] nll_i / sum_i w[y_i],   loss_out[1] = sum_i w[y_i].
 * dlogits (same dtype / shape as logits, may alias them, may be NULL) receives grad_scale * dloss/dlogits; rows
 * without a label are zero-filled. Deterministic (fixed-order reductions). */
size_t b200_weighted_ce_workspace_bytes(int B, int L);
int b200_weighted_ce(const void* logits, int logits_fp32, int64_t ld, const int64_t* labels, const float* vocab_weight,
                     int B, int L, int V, float grad_scale, void* dlogits, int64_t ldd, float* loss_out,
                     void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* HF Trainer gradient clipping (max_grad_norm, README.md:151) + torch.optim.AdamW as configured by
 * LLaVATrainer.create_optimizer (train/llava_trainer.py:191-278), on flat buffers.
 * Gradients are bf16 (grad_fp32 = 0) or fp32 (grad_fp32 = 1, what the backward GEMMs accumulate into).
 * b200_grad_sq_norm: out2[0] (+)= sum(grad^2) over a gradient buffer (accumulate != 0 chains buffers),
 * out2[1] = min(1, max_norm / (sqrt(out2[0]) + 1e-6)) (1 when max_norm <= 0). Deterministic two-stage sum.
 * b200_adamw_step: one pass over fp32 master weights / m / v with bf16 gradients scaled by *clip_coef (device
 * pointer, NULL = 1): decoupled weight decay, bias-corrected Adam update (step counts from 1), bf16 copy of the new
 * weights written to `param` (may be NULL). */
size_t b200_grad_norm_workspace_bytes(void);
int b200_grad_sq_norm(const void* grad, int grad_fp32, int64_t n, int accumulate, float max_norm, float* out2,
                      void* workspace, size_t workspace_bytes, b200_stream_t stream);
int b200_adamw_step(float* master, void* param, const void* grad, int grad_fp32, float* m, float* v, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, const float* clip_coef,
                    b200_stream_t stream);

/* ============================================================================================================
 * Point-cloud branch of the image pooler: PointTransformerV3 in cls_mode, fp32  (ptv3.cu)
 * Replaces ImageEmbeddingPooler._encode_pc (model/multimodal_projector/builder.py:93-148) and what it calls:
 * Point.serialization / sparsify (model/multimodal_projector/pointtransformerv3.py:84-177), the serialization codes
 * (model/multimodal_projector/serialization/{default,z_order,hilbert}.py), spconv SubMConv3d (:549,:768),
 * SerializedAttention over flash-attn varlen (:443-494), SerializedPooling over torch_scatter.segment_csr (:643-713).
 * The host (mm_or_b200/model/point_transformer.py) keeps every level sorted by z-order code; all feature tensors are
 * fp32 row-major [points, channels]; index tensors are int32; codes are int64.
 * ========================================================================================================== */

/* grid = trunc((xyz - min over ALL points) / grid_size) in IEEE fp32 (pointtransformerv3.py:96-98).
 * pts [n, ld] (xyz first), min3 [3] out, grid [n, 3] out, max_out [1] out = largest grid coordinate (the host derives
 * the serialization depth from it, :101-103). */
int b200_pc_grid_coords(const float* pts, int ld, int n, float grid_size, float* min3, int32_t* grid,
                        int32_t* max_out, b200_stream_t stream);

/* code[i] = batch[i] << 3*depth | key(grid[i])  (serialization/default.py:9-25); order 0 "z", 1 "z-trans", 2 "hilbert",
 * 3 "hilbert-trans". Bit-exact with the reference's int64 codes. */
int b200_pc_encode(const int32_t* grid, const int32_t* batch, int n, int depth, int order, int64_t* code,
                   b200_stream_t stream);

/* order = argsort(code) (stable, unsigned, low `bits` bits), code_sorted = code[order]  (pointtransformerv3.py:118). */
size_t b200_pc_argsort_workspace_bytes(int n);
int b200_pc_argsort(const int64_t* code, int n, int bits, int64_t* code_sorted, int32_t* order, void* workspace,
                    size_t workspace_bytes, b200_stream_t stream);

/* dst[i, 0:width] = src[idx[i], 0:width] over 4-byte elements (the physical re-ordering into z-order). */
int b200_pc_gather_rows(const void* src, int ld_src, const int32_t* idx, int n, int width, void* dst, int ld_dst,
                        b200_stream_t stream);

/* Submanifold-convolution neighbour table (spconv indice pairs): nbr[i, t] = row of the active voxel at
 * grid[i] + tap(t) - ksize/2, taps (a, b, c) row-major over (x, y, z), -1 where absent. zcode: sorted "z" codes of
 * the rows. ksize 3 or 5. dup_flag[0] is set to 1 when two rows share a voxel (undefined for spconv too). */
int b200_pc_neighbors(const int64_t* zcode, const int32_t* grid, const int32_t* batch, int n, int depth, int ksize,
                      int32_t* nbr, int32_t* dup_flag, b200_stream_t stream);

/* SerializedPooling plan (pointtransformerv3.py:655-701): clusters are the runs of equal zcode >> 3*pooling_depth.
 * seg_start [n+1] out (first n_out+1 entries valid), n_out [1] out (device), grid_out [n, 3] / batch_out [n] out
 * (first n_out rows valid): grid >> pooling_depth and batch of each cluster. */
int b200_pc_pool_plan(const int64_t* zcode, const int32_t* grid, const int32_t* batch, int n, int pooling_depth,
                      int32_t* seg_start, int32_t* n_out, int32_t* grid_out, int32_t* batch_out,
                      b200_stream_t stream);

/* off[b] = first row of cloud b, b = 0..n_clouds (batch sorted ascending). */
int b200_pc_cloud_offsets(const int32_t* batch, int n, int n_clouds, int32_t* off, b200_stream_t stream);

/* Gather-GEMM: C[m, :] = epilogue(sum_t A[idx[m, t], 0:K] . W[t*K:(t+1)*K, 0:N]); idx NULL = plain Linear (taps 1).
 * W is [taps*K, N] row-major (the loader transposes nn.Linear / SubMConv3d weights once). Epilogue: + bias[N],
 * * scale[N] + shift[N] (BatchNorm1d in eval mode, folded), act 0 none / 2 GELU(erf), + residual[M, ldr]; C fp32, or
 * bf16 when out_bf16 (project_pc writes the pooler's token rows). When few rows meet many taps (the deep stages) the taps
 * are split over CTAs and the partial sums pass through `workspace` (b200_pc_gemm_workspace_bytes; NULL / too small =
 * single pass); the splits are added in order, so the result does not depend on scheduling. */
size_t b200_pc_gemm_workspace_bytes(int M, int N, int K, int taps);
int b200_pc_gemm_f32(const float* A, int lda, const int32_t* idx, int taps, const float* W, const float* bias,
                     const float* scale, const float* shift, int act, const float* residual, int ldr, void* C, int ldc,
                     int out_bf16, int M, int N, int K, void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* out = residual + LayerNorm(x) * gamma + beta (nn.LayerNorm, biased variance; residual may be NULL). */
int b200_pc_layernorm_f32(const float* x, int ld, const float* gamma, const float* beta, float eps,
                          const float* residual, int ldr, float* out, int ldo, int M, int C, b200_stream_t stream);

/* Serialized patch attention with flash-attn varlen semantics (q, k, v and the output rounded to fp16, fp32 softmax).
 * qkv [n, ld] = (q | k | v) of `channels` each, head h at h*16; order [n] = serialized order of the block; patches
 * [n_patches, 4] = (q_begin, q_len, k_begin, k_len) in positions of `order` (get_padding_and_inverse,
 * pointtransformerv3.py:385-441: a topped-up last patch has q = its own points, k = the cloud's last patch_size
 * points). out [n, ldo], row r written by the patch that owns r. head_dim must be 16. */
int b200_pc_patch_attention(const float* qkv, int ld, const int32_t* order, const int32_t* patches, int n_patches,
                            int max_q_len, int channels, int heads, float scale, float* out, int ldo,
                            b200_stream_t stream);

/* out[s, c] = act((max over rows seg_start[s] .. seg_start[s+1]-1 of x[., c]) * scale[c] + shift[c])
 * (torch_scatter.segment_csr(reduce="max") + BatchNorm1d + GELU, pointtransformerv3.py:686-713). */
int b200_pc_segment_max(const float* x, int ld, const int32_t* seg_start, int n_seg, int C, const float* scale,
                        const float* shift, int act, float* out, int ldo, b200_stream_t stream);

/* out[row_map[b], :] = mean of rows cloud_off[b] .. cloud_off[b+1]-1 (AdaptiveAvgPool1d(1), builder.py:139-144). */
int b200_pc_cloud_mean(const float* x, int ld, const int32_t* cloud_off, int n_clouds, int C, const int32_t* row_map,
                       float* out, int ldo, b200_stream_t stream);


/* ---- fine-tuning of the small modality encoders and of embed_tokens (train_extras.cu) ----
 * Seg-mask CNN (model/multimodal_projector/segmentation_map_feature_extractor.py:53-75) under training: the forward keeps
 * every activation (fp32, `acts`), the backward is what torch autograd computes for the reference
 * (`image_pooler.segmasks_encoder.*` is trainable, train/train.py:1257-1261): d_conv_w[l] [Cout, Cin, 3, 3], d_conv_b[l]
 * [Cout], d_emb [30, 8], all fp32, overwritten or (accumulate) added to. d_out: bf16 token-gradient rows, map n reads
 * row row_map[n] (or n), < 0 = no gradient. d_conv_w / d_conv_b are HOST arrays of 5 device pointers. */
size_t b200_segmask_train_acts_bytes(int n_maps);
int b200_segmask_forward_train(const b200_segmask_weights* w, const uint8_t* cls, int n_maps, void* acts,
                               size_t acts_bytes, void* out, int64_t out_ld, const int32_t* out_row_map,
                               b200_stream_t stream);
size_t b200_segmask_backward_workspace_bytes(int n_maps);
int b200_segmask_backward(const b200_segmask_weights* w, const uint8_t* cls, int n_maps, const void* acts,
                          const void* d_out, int64_t d_ld, const int32_t* row_map, float* const* d_conv_w,
                          float* const* d_conv_b, float* d_emb, int accumulate, void* workspace, size_t workspace_bytes,
                          b200_stream_t stream);
/* nn.Embedding backward for embed_tokens (language_model/llava_llama.py:93 -> HF LlamaModel.embed_tokens under autograd):
 * d_table[seg_token[s], :] (+)= sum of the bf16 rows d_rows[row_list[j]] for j in [seg_start[s], seg_start[s+1]); the
 * host groups the text rows by token id, so the sum order is fixed. */
int b200_embed_grad(const void* d_rows, int64_t ld, const int32_t* row_list, const int32_t* seg_start,
                    const int32_t* seg_token, int n_seg, int D, float* d_table, int accumulate, b200_stream_t stream);

/* Inverted dropout on bf16 (the reference's train-mode nn.Dropout, e.g. lora_dropout 0.05, train/train.py:111,1165):
 * y = mask * x / (1 - p), or y += ... with accumulate = 1 (the backward: dx += mask * g / (1 - p) with the SAME seed).
 * The mask is a pure function of (seed, element index) -- Philox4x32-10, element i dropped iff word i mod 4 of
 * Philox(counter i / 4, key seed) < floor(p * 2^32) -- so nothing is stored between forward and backward. torch's own
 * generator stream is not reproducible outside torch: the distribution is kept, not the individual masks. */
int b200_dropout(const void* x_bf16, void* y_bf16, int64_t n, float p, uint64_t seed, int accumulate,
                 b200_stream_t stream);

/* ============================================================================================================
 * NF4 storage of the frozen base weights (QLoRA recipe). Replaces bitsandbytes' Linear4bit storage as configured by
 * LLaVA/llava/train/train.py:1098-1114 (BitsAndBytesConfig(load_in_4bit, bnb_4bit_quant_type='nf4', compute dtype
 * bf16)); bitsandbytes==0.41.0 is not vendored, its published algorithm is restated (mm_or_b200/csrc/nf4.cu).
 * w_bf16 / out_bf16: n bf16 values (n a multiple of 64); packed: n / 2 bytes, two 4-bit codes per byte with the even
 * element in the high nibble; absmax: n / 64 fp32 block maxima. quantize followed by dequantize yields the weights
 * a Linear4bit layer multiplies with.
 * ============================================================================================================ */
int b200_nf4_quantize(const void* w_bf16, int64_t n, uint8_t* packed, float* absmax, b200_stream_t stream);
int b200_nf4_dequantize(const uint8_t* packed, const float* absmax, int64_t n, void* out_bf16, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_MMOR_H_ */
