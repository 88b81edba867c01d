// extern "C" surface of libb200mmor.so (declared in include/b200_mmor.h).
#include "../../include/b200_mmor.h"

#include "common.h"

namespace b200 {
const char* last_error_cstr();
}

using namespace b200;

extern "C" {

const char* b200_last_error(void) { return last_error_cstr(); }

int b200_abi_version(void) { return 1; }

size_t b200_sizeof_struct(int which) {
  switch (which) {
    case 0: return sizeof(b200_vit_layer);
    case 1: return sizeof(b200_vit_weights);
    case 2: return sizeof(b200_bert_layer);
    case 3: return sizeof(b200_pooler_weights);
    case 4: return sizeof(b200_segmask_weights);
    case 5: return sizeof(b200_projector_weights);
    case 6: return sizeof(b200_llama_layer);
    case 7: return sizeof(b200_llama_weights);
    case 8: return sizeof(b200_kv_cache);
    default: return 0;
  }
}

int b200_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                   const void* bias, const void* residual, int ldr, const int32_t* row_map, int act, int out_fp32,
                   int bn_hint, b200_stream_t stream) {
  GemmEpilogue e;
  e.bias = static_cast<const bf16*>(bias);
  e.residual = static_cast<const bf16*>(residual);
  e.ldr = ldr;
  e.row_map = row_map;
  e.act = act;
  e.out_fp32 = out_fp32;
  if (bn_hint < 0) {  // -64 / -96 / -128: the decode step's two-CTAs-per-SM shape of that width
    e.ctas_per_sm = 2;
    bn_hint = -bn_hint;
  }
  return gemm_bf16_tn(static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), ldw, C, ldc, M, N, K, e,
                      bn_hint, static_cast<cudaStream_t>(stream));
}

int b200_gemm_bf16_skinny(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                          const void* bias, const void* residual, int ldr, int act, int out_fp32, int splits,
                          b200_stream_t stream) {
  GemmEpilogue e;
  e.bias = static_cast<const bf16*>(bias);
  e.residual = static_cast<const bf16*>(residual);
  e.ldr = ldr;
  e.act = act;
  e.out_fp32 = out_fp32;
  return gemm_skinny(static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), ldw, C, ldc, M, N, K, e, splits,
                     static_cast<cudaStream_t>(stream));
}

int b200_gemm_nf4(const void* A, int lda, const uint8_t* codes, const float* absmax, void* C, int ldc, int M, int N, int K,
                  const void* residual, int ldr, float scale, b200_stream_t stream) {
  GemmEpilogue e;
  e.scale = scale;
  e.residual = static_cast<const bf16*>(residual);
  e.ldr = ldr;
  return gemm_nf4_tn(static_cast<const bf16*>(A), lda, codes, absmax, C, ldc, M, N, K, e,
                     static_cast<cudaStream_t>(stream));
}

int b200_gemm_bf16_ex(const void* A, int lda, int a_transposed, const void* W, int ldw, int w_transposed, void* C, int ldc,
                      int M, int N, int K, const void* bias, const void* residual, int ldr, int act, int out_fp32,
                      int accumulate, float scale, int bn_hint, b200_stream_t stream) {
  GemmEpilogue e;
  e.scale = scale;
  e.bias = static_cast<const bf16*>(bias);
  e.residual = static_cast<const bf16*>(residual);
  e.ldr = ldr;
  e.act = act;
  e.out_fp32 = out_fp32;
  e.accumulate = accumulate;
  return gemm_bf16_ex(static_cast<const bf16*>(A), lda, a_transposed, static_cast<const bf16*>(W), ldw, w_transposed, C,
                      ldc, M, N, K, e, bn_hint, static_cast<cudaStream_t>(stream));
}

size_t b200_colsum_workspace_bytes(int N) { return colsum_workspace_bytes(N); }
int b200_colsum(const void* dy, int64_t ld, int M, int N, int accumulate, float* out, void* workspace,
                size_t workspace_bytes, b200_stream_t stream) {
  return colsum(static_cast<const bf16*>(dy), ld, M, N, accumulate, out, workspace, workspace_bytes,
                static_cast<cudaStream_t>(stream));
}

int b200_act_backward(const void* z, const void* dy, void* dz, int64_t n_out, int act, b200_stream_t stream) {
  return act_backward(static_cast<const bf16*>(z), static_cast<const bf16*>(dy), static_cast<bf16*>(dz), n_out, act,
                      static_cast<cudaStream_t>(stream));
}

int b200_act_forward(const void* z, void* y, int64_t n, int act, b200_stream_t stream) {
  return act_forward(static_cast<const bf16*>(z), static_cast<bf16*>(y), n, act, static_cast<cudaStream_t>(stream));
}

int b200_group_sum(const void* x, int groups, int64_t slab, int accumulate, float* out, b200_stream_t stream) {
  return group_sum(static_cast<const bf16*>(x), groups, slab, accumulate, out, static_cast<cudaStream_t>(stream));
}

int b200_gather_add_rows(const void* src, int64_t ld_src, const int32_t* row_map, const void* add, int period, int rows,
                         int D, void* out, b200_stream_t stream) {
  return gather_add_rows(static_cast<const bf16*>(src), ld_src, row_map, static_cast<const bf16*>(add), period, rows, D,
                         static_cast<bf16*>(out), static_cast<cudaStream_t>(stream));
}

int b200_swiglu_forward(const void* z, void* h, int64_t n_out, b200_stream_t stream) {
  return swiglu_forward(static_cast<const bf16*>(z), static_cast<bf16*>(h), n_out, static_cast<cudaStream_t>(stream));
}

int b200_rope_kv_backward(void* dqkv, const int32_t* kv_start, const float* cos_table, const float* sin_table,
                          int max_pos, const void* dk_cache, const void* dv_cache, int B, int H, int Lq, int cap,
                          b200_stream_t stream) {
  return rope_kv_backward(static_cast<bf16*>(dqkv), kv_start, cos_table, sin_table, max_pos,
                          static_cast<const bf16*>(dk_cache), static_cast<const bf16*>(dv_cache), B, H, Lq, cap,
                          static_cast<cudaStream_t>(stream));
}

size_t b200_norm_backward_workspace_bytes(int M, int D) { return norm_backward_workspace_bytes(M, D); }
int b200_norm_backward(const void* x, const void* dy, const void* gamma, float eps, int M, int D, int rms,
                       const void* add, void* dx, float* dgamma, float* dbeta, int accumulate, void* workspace, size_t workspace_bytes,
                       b200_stream_t stream) {
  return norm_backward(static_cast<const bf16*>(x), static_cast<const bf16*>(dy), static_cast<const bf16*>(gamma), eps,
                       M, D, rms, static_cast<const bf16*>(add), static_cast<bf16*>(dx), dgamma, dbeta, accumulate, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

size_t b200_weighted_ce_workspace_bytes(int B, int L) { return weighted_ce_workspace_bytes(B, L); }

int b200_weighted_ce(const void* logits, int logits_fp32, int64_t ld, const int64_t* labels, const float* vocab_weight,
                     int B, int L, int V, float grad_scale, void* dlogits, int64_t ldd, float* loss_out,
                     void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  return weighted_ce(logits, logits_fp32, ld, reinterpret_cast<const long long*>(labels), vocab_weight, B, L, V,
                     grad_scale, dlogits, ldd, loss_out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t b200_grad_norm_workspace_bytes(void) { return grad_norm_workspace_bytes(); }

int b200_grad_sq_norm(const void* grad, int grad_fp32, int64_t n, int accumulate, float max_norm, float* out2,
                      void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  return grad_sq_norm(grad, grad_fp32, n, accumulate, max_norm, out2, workspace, workspace_bytes,
                      static_cast<cudaStream_t>(stream));
}

int b200_adamw_step(float* master, void* param, const void* grad, int grad_fp32, float* m, float* v, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, const float* clip_coef,
                    b200_stream_t stream) {
  return adamw_step(master, static_cast<bf16*>(param), grad, grad_fp32, m, v, n, lr, beta1, beta2, eps, weight_decay,
                    step, clip_coef, static_cast<cudaStream_t>(stream));
}

int b200_layernorm(const void* x, int64_t ldx, const int32_t* row_map, const void* add, int period, const void* gamma,
                   const void* beta, float eps, void* out, int64_t ldo, int M, int D, b200_stream_t stream) {
  return layernorm(static_cast<const bf16*>(x), ldx, row_map, static_cast<const bf16*>(add), period,
                   static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta), eps, static_cast<bf16*>(out), ldo, M,
                   D, 0, 0, static_cast<cudaStream_t>(stream));
}

int b200_rmsnorm(const void* x, int64_t ldx, const void* w, float eps, void* out, int64_t ldo, int M, int D,
                 b200_stream_t stream) {
  return rmsnorm(static_cast<const bf16*>(x), ldx, static_cast<const bf16*>(w), eps, static_cast<bf16*>(out), ldo, M, D,
                 static_cast<cudaStream_t>(stream));
}

int b200_flash_attention(const void* q, int64_t q_bs, int64_t q_rs, int64_t q_hs, const void* k, int64_t k_bs,
                         int64_t k_rs, int64_t k_hs, const void* v, int64_t v_bs, int64_t v_rs, int64_t v_hs, void* o,
                         int64_t o_bs, int64_t o_rs, int64_t o_hs, int B, int H, int Lq, int Lk, int head_dim,
                         const int32_t* kv_start, const int32_t* kv_len, int causal, float scale,
                         b200_stream_t stream) {
  AttnArgs a{};
  a.q = static_cast<const bf16*>(q);
  a.k = static_cast<const bf16*>(k);
  a.v = static_cast<const bf16*>(v);
  a.o = static_cast<bf16*>(o);
  a.q_bs = q_bs; a.q_rs = q_rs; a.q_hs = q_hs;
  a.k_bs = k_bs; a.k_rs = k_rs; a.k_hs = k_hs;
  a.v_bs = v_bs; a.v_rs = v_rs; a.v_hs = v_hs;
  a.o_bs = o_bs; a.o_rs = o_rs; a.o_hs = o_hs;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
  a.kv_start = kv_start;
  a.kv_len = kv_len;
  a.causal = causal;
  a.scale_log2 = scale * 1.4426950408889634f;
  return flash_attn(a, head_dim, static_cast<cudaStream_t>(stream));
}

int b200_flash_attention_lse(const void* q, int64_t q_bs, int64_t q_rs, int64_t q_hs, const void* k, int64_t k_bs,
                             int64_t k_rs, int64_t k_hs, const void* v, int64_t v_bs, int64_t v_rs, int64_t v_hs,
                             void* o, int64_t o_bs, int64_t o_rs, int64_t o_hs, int B, int H, int Lq, int Lk,
                             int head_dim, const int32_t* kv_start, const int32_t* kv_len, int causal, float scale,
                             float* lse, b200_stream_t stream) {
  AttnArgs a{};
  a.q = static_cast<const bf16*>(q);
  a.k = static_cast<const bf16*>(k);
  a.v = static_cast<const bf16*>(v);
  a.o = static_cast<bf16*>(o);
  a.q_bs = q_bs; a.q_rs = q_rs; a.q_hs = q_hs;
  a.k_bs = k_bs; a.k_rs = k_rs; a.k_hs = k_hs;
  a.v_bs = v_bs; a.v_rs = v_rs; a.v_hs = v_hs;
  a.o_bs = o_bs; a.o_rs = o_rs; a.o_hs = o_hs;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
  a.kv_start = kv_start;
  a.kv_len = kv_len;
  a.causal = causal;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.lse = lse;
  return flash_attn(a, head_dim, static_cast<cudaStream_t>(stream));
}

size_t b200_flash_attention_bwd_workspace_bytes(int B, int H, int Lq) { return flash_attn_bwd_workspace_bytes(B, H, Lq); }

int b200_flash_attention_bwd(const void* q, int64_t q_bs, int64_t q_rs, int64_t q_hs, const void* k, int64_t k_bs,
                             int64_t k_rs, int64_t k_hs, const void* v, int64_t v_bs, int64_t v_rs, int64_t v_hs,
                             const void* o, const void* d_o, int64_t o_bs, int64_t o_rs, int64_t o_hs, const float* lse,
                             void* dq, void* dk, void* dv, int B, int H, int Lq, int Lk, int head_dim,
                             const int32_t* kv_start, const int32_t* kv_len, int causal, float scale, void* workspace,
                             size_t workspace_bytes, b200_stream_t stream) {
  AttnArgs a{};
  a.q = static_cast<const bf16*>(q);
  a.k = static_cast<const bf16*>(k);
  a.v = static_cast<const bf16*>(v);
  a.o = const_cast<bf16*>(static_cast<const bf16*>(o));
  a.q_bs = q_bs; a.q_rs = q_rs; a.q_hs = q_hs;
  a.k_bs = k_bs; a.k_rs = k_rs; a.k_hs = k_hs;
  a.v_bs = v_bs; a.v_rs = v_rs; a.v_hs = v_hs;
  a.o_bs = o_bs; a.o_rs = o_rs; a.o_hs = o_hs;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
  a.kv_start = kv_start;
  a.kv_len = kv_len;
  a.causal = causal;
  a.scale_log2 = scale * 1.4426950408889634f;
  AttnBwdArgs g{};
  g.d_o = static_cast<const bf16*>(d_o);
  g.do_bs = o_bs; g.do_rs = o_rs; g.do_hs = o_hs;      // dO shares O's layout
  g.lse = lse;
  g.dq = static_cast<bf16*>(dq);
  g.dk = static_cast<bf16*>(dk);
  g.dv = static_cast<bf16*>(dv);
  g.dq_bs = q_bs; g.dq_rs = q_rs; g.dq_hs = q_hs;      // gradients share their tensor's layout
  g.dk_bs = k_bs; g.dk_rs = k_rs; g.dk_hs = k_hs;
  g.dv_bs = v_bs; g.dv_rs = v_rs; g.dv_hs = v_hs;
  return flash_attn_bwd(a, g, head_dim, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t b200_decode_attention_workspace_bytes(int B, int H, int ctx) {
  (void)ctx;
  return decode_attn_workspace_bytes(B, H, 16);
}

int b200_decode_attention(const void* q, int64_t q_rs, const void* k_cache, const void* v_cache, void* o, int64_t o_rs,
                          int B, int H, int cap, int ctx, const int32_t* kv_start, float scale, int splits,
                          void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  DecodeArgs a{};
  a.q = static_cast<const bf16*>(q);
  a.q_rs = q_rs;
  a.kc = static_cast<const bf16*>(k_cache);
  a.vc = static_cast<const bf16*>(v_cache);
  a.o = static_cast<bf16*>(o);
  a.o_rs = o_rs;
  a.B = B; a.H = H; a.cap = cap; a.ctx = ctx;
  a.kv_start = kv_start;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.splits = splits;
  return decode_attn(a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int b200_rope_kv_write(void* qkv, const int32_t* kv_start, const float* cos_table, const float* sin_table, int max_pos,
                       void* k_cache, void* v_cache, int B, int H, int Lq, int slot0, int cap, b200_stream_t stream) {
  return rope_kv_write(static_cast<bf16*>(qkv), kv_start, cos_table, sin_table, max_pos, static_cast<bf16*>(k_cache),
                       static_cast<bf16*>(v_cache), B, H, Lq, slot0, nullptr, cap, static_cast<cudaStream_t>(stream));
}

int b200_embed_rows(const int32_t* ids, const void* table, void* out, int64_t ldo, int rows, int D, int vocab,
                    b200_stream_t stream) {
  return embed_rows(ids, static_cast<const bf16*>(table), static_cast<bf16*>(out), ldo, rows, D, vocab,
                    static_cast<cudaStream_t>(stream));
}

int b200_argmax(const void* logits, int is_fp32, int64_t ld, int rows, int V, int32_t* out_tok, int32_t* finished,
                int eos_id, int pad_id, b200_stream_t stream) {
  return argmax_rows(logits, is_fp32, ld, rows, V, out_tok, finished, eos_id, pad_id, nullptr, 0, 0, nullptr,
                     static_cast<cudaStream_t>(stream));
}

int b200_patchify(const void* pixels, void* cols, int N, int C, int S, int P, int Kpad, b200_stream_t stream) {
  return patchify_im2col(static_cast<const bf16*>(pixels), static_cast<bf16*>(cols), N, C, S, P, Kpad,
                         static_cast<cudaStream_t>(stream));
}

size_t b200_preprocess_workspace_bytes(int n, int canvas_h, int resized_w) {
  return preprocess_workspace_bytes(n, canvas_h, resized_w);
}

int b200_preprocess_images(const uint8_t* images, int n, int H, int W, int pad_to_square, const uint8_t* background_rgb,
                           const int32_t* bounds_x, const int32_t* coeffs_x, int ksize_x, const int32_t* bounds_y,
                           const int32_t* coeffs_y, int ksize_y, int resized_h, int resized_w, int crop,
                           const float* mean, const float* std, void* out, void* workspace, size_t workspace_bytes,
                           b200_stream_t stream) {
  return preprocess_images(images, n, H, W, pad_to_square, background_rgb, bounds_x, coeffs_x, ksize_x, bounds_y,
                           coeffs_y, ksize_y, resized_h, resized_w, crop, mean, std, static_cast<bf16*>(out), workspace,
                           workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
