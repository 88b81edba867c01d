// Stage-level entry points: the launch sequences that replace the reference's Python module forwards
// (CLIPVisionTower / ImageEmbeddingPooler / mm_projector + pack / LlamaForCausalLM prefill and decode step).
// Host code only: every arithmetic step is one of the kernels in gemm_sm100.cu / attention.cu / norm.cu / misc.cu.
#include <cmath>

#include "../../include/b200_mmor.h"
#include "common.h"

namespace b200 {

static constexpr float kLog2e = 1.4426950408889634f;

// bump allocator over the caller-provided workspace (256-byte aligned slices)
struct Arena {
  uint8_t* base;
  size_t cap, used;
  Arena(void* p, size_t n) : base(static_cast<uint8_t*>(p)), cap(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) & ~static_cast<size_t>(255);
    T* r = reinterpret_cast<T*>(base + used);
    used += bytes;
    return r;
  }
  bool ok() const { return used <= cap && (base != nullptr || used == 0); }
};
static size_t al(size_t bytes) { return (bytes + 255) & ~static_cast<size_t>(255); }

static const bf16* B(const void* p) { return static_cast<const bf16*>(p); }

// -----------------------------------------------------------------------------------------------
// ViT
// -----------------------------------------------------------------------------------------------
static size_t vit_ws(const b200_vit_weights* w, int n) {
  const size_t G = w->image_size / w->patch, P = G * G, T = P + 1;
  const size_t rows = static_cast<size_t>(n) * T;
  const size_t colbuf = static_cast<size_t>(n) * P * w->kpad;       // im2col
  const size_t fc1buf = rows * w->ffn;                               // fc1 output (aliases im2col: used later)
  size_t s = 0;
  s += al((colbuf > fc1buf ? colbuf : fc1buf) * 2);
  s += al(static_cast<size_t>(n) * P * w->hidden * 2);  // patch embeddings
  s += al(rows * w->hidden * 2);                        // LN output / attention context
  s += al(rows * 3 * w->hidden * 2);                    // qkv
  s += al(rows * sizeof(int));                          // gather map
  return s;
}

__global__ void vit_row_map_kernel(int* map, int n_img, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * T) return;
  const int t = i % T, n = i / T;
  map[i] = t == 0 ? -1 : n * (T - 1) + t - 1;  // CLS row has no patch source; its value lives in pos_cls[0]
}

static int vit_forward(const b200_vit_weights* w, const bf16* pixels, bf16* x, int n, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  if (n <= 0) return 0;
  const int G = w->image_size / w->patch, P = G * G, T = P + 1, D = w->hidden;
  if (D % w->heads != 0 || D / w->heads != 64) return fail(-2, "vit: head_dim must be 64");
  const int rows = n * T;
  if (ws_bytes < vit_ws(w, n)) return fail(-2, "vit_forward: workspace too small (%zu < %zu)", ws_bytes, vit_ws(w, n));
  Arena a(ws, ws_bytes);
  const size_t colbuf = static_cast<size_t>(n) * P * w->kpad, fc1buf = static_cast<size_t>(rows) * w->ffn;
  bf16* big = a.take<bf16>(colbuf > fc1buf ? colbuf : fc1buf);
  bf16* patches = a.take<bf16>(static_cast<size_t>(n) * P * D);
  bf16* lnbuf = a.take<bf16>(static_cast<size_t>(rows) * D);
  bf16* qkv = a.take<bf16>(static_cast<size_t>(rows) * 3 * D);
  int* map = a.take<int>(rows);

  // patch embedding: im2col + GEMM (conv k = s = patch, no bias)
  B200_TRY(patchify_im2col(pixels, big, n, 3, w->image_size, w->patch, w->kpad, st));
  GemmEpilogue e0;
  B200_TRY(gemm_bf16_tn(big, w->kpad, B(w->patch_w), w->kpad, patches, D, n * P, D, w->kpad, e0, 0, st));
  // x = pre_layrnorm([CLS ; patches] + position_embedding)
  {
    LaunchScope scope(kFamMisc, st);
    vit_row_map_kernel<<<(rows + 255) / 256, 256, 0, st>>>(map, n, T);
  }
  B200_CUDA_OK(cudaGetLastError());
  B200_TRY(layernorm(patches, D, map, B(w->pos_cls), T, B(w->pre_ln_w), B(w->pre_ln_b), w->ln_eps, x, D, rows, D, 0, 0,
                     st));
  for (int l = 0; l < w->n_layers; ++l) {
    const b200_vit_layer& L = w->layers[l];
    B200_TRY(layernorm(x, D, nullptr, nullptr, 1, B(L.ln1_w), B(L.ln1_b), w->ln_eps, lnbuf, D, rows, D, 0, 0, st));
    GemmEpilogue eq;
    eq.bias = B(L.qkv_b);
    B200_TRY(gemm_bf16_tn(lnbuf, D, B(L.qkv_w), D, qkv, 3 * D, rows, 3 * D, D, eq, 0, st));
    AttnArgs at{};
    at.q = qkv;
    at.k = qkv + D;
    at.v = qkv + 2 * D;
    at.o = lnbuf;
    at.q_bs = at.k_bs = at.v_bs = static_cast<long long>(T) * 3 * D;
    at.q_rs = at.k_rs = at.v_rs = 3 * D;
    at.q_hs = at.k_hs = at.v_hs = 64;
    at.o_bs = static_cast<long long>(T) * D;
    at.o_rs = D;
    at.o_hs = 64;
    at.B = n;
    at.H = w->heads;
    at.Lq = at.Lk = T;
    at.scale_log2 = kLog2e;  // the 64^-0.5 query scale is folded into qkv_w / qkv_b (exact: power of two)
    B200_TRY(flash_attn(at, 64, st));
    GemmEpilogue eo;
    eo.bias = B(L.out_b);
    eo.residual = x;
    eo.ldr = D;
    B200_TRY(gemm_bf16_tn(lnbuf, D, B(L.out_w), D, x, D, rows, D, D, eo, 0, st));
    B200_TRY(layernorm(x, D, nullptr, nullptr, 1, B(L.ln2_w), B(L.ln2_b), w->ln_eps, lnbuf, D, rows, D, 0, 0, st));
    GemmEpilogue e1;
    e1.bias = B(L.fc1_b);
    e1.act = kActQuickGelu;
    B200_TRY(gemm_bf16_tn(lnbuf, D, B(L.fc1_w), D, big, w->ffn, rows, w->ffn, D, e1, 0, st));
    GemmEpilogue e2;
    e2.bias = B(L.fc2_b);
    e2.residual = x;
    e2.ldr = D;
    B200_TRY(gemm_bf16_tn(big, w->ffn, B(L.fc2_w), w->ffn, x, D, rows, D, w->ffn, e2, 0, st));
  }
  return 0;
}

// -----------------------------------------------------------------------------------------------
// pooler (BERT encoder over all views of a sample)
// -----------------------------------------------------------------------------------------------
static size_t pooler_ws(const b200_pooler_weights* w, int Bn, int S) {
  const size_t rows = static_cast<size_t>(Bn) * S;
  size_t s = 0;
  s += al(rows * w->hidden * 2) * 3;   // e, ctx/tmp, e_next
  s += al(rows * 3 * w->hidden * 2);   // qkv
  s += al(rows * w->ffn * 2);          // ffn hidden
  return s;
}

static int pooler_forward(const b200_pooler_weights* w, const bf16* src, long long src_ld, const int* gmap,
                          const int* kv_len, int Bn, int S, int keep, bf16* out, int out_tokens, void* ws,
                          size_t ws_bytes, cudaStream_t st) {
  if (Bn <= 0) return 0;
  const int D = w->hidden;
  if (D / w->heads != 128 || D % w->heads) return fail(-2, "pooler: head_dim must be 128");
  if (S > w->max_pos) return fail(-2, "pooler: %d tokens exceed max_position_embeddings %d", S, w->max_pos);
  if (keep > S || keep > out_tokens) return fail(-2, "pooler: keep %d > S %d or out_tokens %d", keep, S, out_tokens);
  if (ws_bytes < pooler_ws(w, Bn, S)) return fail(-2, "pooler_forward: workspace too small");
  Arena a(ws, ws_bytes);
  const size_t rows_all = static_cast<size_t>(Bn) * S;
  bf16* e = a.take<bf16>(rows_all * D);
  bf16* tmp = a.take<bf16>(rows_all * D);
  bf16* e2 = a.take<bf16>(rows_all * D);
  bf16* qkv = a.take<bf16>(rows_all * 3 * D);
  bf16* ffn = a.take<bf16>(rows_all * w->ffn);
  // embeddings: LN(x + position + token_type[0])
  B200_TRY(layernorm(src, src_ld, gmap, B(w->pos_type), S, B(w->emb_ln_w), B(w->emb_ln_b), w->ln_eps, e, D,
                     static_cast<int>(rows_all), D, 0, 0, st));
  for (int l = 0; l < w->n_layers; ++l) {
    const b200_bert_layer& L = w->layers[l];
    const bool last = l == w->n_layers - 1;
    const int Lq = last ? keep : S;  // only the first `keep` rows of the last layer are consumed (builder.py:175)
    const int rows = Bn * Lq;
    GemmEpilogue eq;
    eq.bias = B(L.qkv_b);
    B200_TRY(gemm_bf16_tn(e, D, B(L.qkv_w), D, qkv, 3 * D, static_cast<int>(rows_all), 3 * D, D, eq, 0, st));
    AttnArgs at{};
    at.q = qkv;
    at.k = qkv + D;
    at.v = qkv + 2 * D;
    at.o = tmp;
    at.q_bs = at.k_bs = at.v_bs = static_cast<long long>(S) * 3 * D;
    at.q_rs = at.k_rs = at.v_rs = 3 * D;
    at.q_hs = at.k_hs = at.v_hs = 128;
    at.o_bs = static_cast<long long>(Lq) * D;
    at.o_rs = D;
    at.o_hs = 128;
    at.B = Bn;
    at.H = w->heads;
    at.Lq = Lq;
    at.Lk = S;
    at.kv_len = kv_len;
    at.scale_log2 = kLog2e / sqrtf(128.f);
    B200_TRY(flash_attn(at, 128, st));
    // attention.output: LN(dense(ctx) + e)
    GemmEpilogue eo;
    eo.bias = B(L.ao_b);
    eo.residual = e;
    eo.ldr = D;
    if (last && Lq != S) {
      eo.res_group = Lq;
      eo.res_group_stride = S;
    }
    B200_TRY(gemm_bf16_tn(tmp, D, B(L.ao_w), D, e2, D, rows, D, D, eo, 0, st));
    B200_TRY(layernorm(e2, D, nullptr, nullptr, 1, B(L.ao_ln_w), B(L.ao_ln_b), w->ln_eps, tmp, D, rows, D, 0, 0, st));
    // intermediate + output: LN(fc2(gelu(fc1(h))) + h)
    GemmEpilogue e1;
    e1.bias = B(L.fc1_b);
    e1.act = kActGeluErf;
    B200_TRY(gemm_bf16_tn(tmp, D, B(L.fc1_w), D, ffn, w->ffn, rows, w->ffn, D, e1, 0, st));
    GemmEpilogue ef;
    ef.bias = B(L.fc2_b);
    ef.residual = tmp;
    ef.ldr = D;
    B200_TRY(gemm_bf16_tn(ffn, w->ffn, B(L.fc2_w), w->ffn, e2, D, rows, D, w->ffn, ef, 0, st));
    if (last) {
      B200_TRY(layernorm(e2, D, nullptr, nullptr, 1, B(L.out_ln_w), B(L.out_ln_b), w->ln_eps, out, D, rows, D, keep,
                         out_tokens, st));
    } else {
      B200_TRY(layernorm(e2, D, nullptr, nullptr, 1, B(L.out_ln_w), B(L.out_ln_b), w->ln_eps, e, D, rows, D, 0, 0, st));
    }
  }
  return 0;
}

// -----------------------------------------------------------------------------------------------
// Llama
// -----------------------------------------------------------------------------------------------
static size_t llama_prefill_ws(const b200_llama_weights* w, int Bn, int L, int all_logits) {
  const size_t T = static_cast<size_t>(Bn) * L;
  size_t s = 0;
  s += al(T * w->hidden * 2);      // norm output / attention context
  s += al(T * 3 * w->hidden * 2);  // qkv
  s += al(T * w->ffn * 2);         // swiglu output
  s += al(static_cast<size_t>(all_logits ? T : Bn) * w->hidden * 2);  // final norm rows
  return s;
}

// nn.Linear dispatch for the decode step (T <= 256 rows, HBM bound). Measured on B200 with cold weights
// (profiles/r1_skinny_gemm_notes.md): projections with few 128-row weight tiles (N = 4096: o_proj, down_proj) are
// 1.3-1.9x faster on the split-K cluster kernel, which keeps two streaming CTAs on every SM; the wide ones (qkv,
// gate_up, lm_head) already fill the chip with 128-column tiles of the tiled kernel.
// Weight-tile width of a wide decode-step projection (opt-in, common.h: decode_tiles_mode): the kernel is persistent
// and HBM bound, so its time is (waves of tiles over the CTA slots) x (bytes of one tile); pick the width that minimises
// waves x width. Mode 1, one CTA per SM (148 slots): qkv 12288 -> 96 (128 tiles, one wave), gate_up 22016 -> 160 (138),
// lm_head 32000 -> 224 (143); with 128 columns they take 96 / 172 / 250 tiles. Mode 2, two CTAs per SM (296 slots,
// half-depth rings, widths 64 / 96 / 128): qkv -> 64 (192 tiles), gate_up -> 96 (230), lm_head -> 128 (250), one wave
// each, and every SM that holds two CTAs streams faster (profiles/r1_skinny_gemm_notes.md).
static int decode_bn(int T, int N, int sms, int per_sm) {
  static const int wide[] = {256, 224, 160, 128, 96};
  static const int half[] = {128, 96, 64};
  const int* widths = per_sm == 2 ? half : wide;
  const int n_widths = per_sm == 2 ? 3 : 5;
  const int m_tiles = (T + 127) / 128;
  if (sms <= 0) return 128;
  const long long slots = static_cast<long long>(sms) * per_sm;
  int best = 128;
  long long best_cost = -1;
  for (int w = 0; w < n_widths; ++w) {
    const int bn = widths[w];
    const long long tiles = static_cast<long long>(m_tiles) * ((N + bn - 1) / bn);
    const long long cost = ((tiles + slots - 1) / slots) * bn;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

static int linear(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int T, int N, int K,
                  const GemmEpilogue& e_in, bool decode, cudaStream_t st) {
  GemmEpilogue e = e_in;
  // decode step with programmatic dependent launch: W is a weight matrix no kernel writes, so its first tiles may be
  // requested before the previous kernel has finished (common.h). Off => the launch and load order of every GEMM is
  // exactly the one validated on hardware.
  e.b_const = (decode && pdl_enabled()) ? 1 : 0;
  if (decode && T <= 256) {
    const int n_tiles = (N + 127) / 128;
    if (n_tiles * 2 <= num_sms()) return gemm_skinny(A, lda, W, ldw, C, ldc, T, N, K, e, 0, st);
    const int mode = decode_tiles_mode();
    if (mode == 2) e.ctas_per_sm = 2;
    if (mode != 0) return gemm_bf16_tn(A, lda, W, ldw, C, ldc, T, N, K, e, decode_bn(T, N, num_sms(), mode), st);
    if (T <= 128) return gemm_bf16_tn(A, lda, W, ldw, C, ldc, T, N, K, e, 128, st);
  }
  return gemm_bf16_tn(A, lda, W, ldw, C, ldc, T, N, K, e, 0, st);
}

static int llama_layer(const b200_llama_weights* w, const b200_llama_layer& Ly, bf16* x, bf16* nbuf, bf16* qkv,
                       bf16* act, const int* kv_start, const int* kv_len, bf16* kc, bf16* vc, int cap, int Bn, int L,
                       int decode, const int* state, int ctx_bound, void* dec_ws, size_t dec_ws_bytes,
                       const int* finished, cudaStream_t st, int q0 = 0) {
  // q0 (prefill only): cache slots [0, q0) of every row already hold keys / values (a shared prompt prefix copied in by
  // the host); the L rows of x are the tokens of slots [q0, q0 + L) and attend to all q0 + L keys (bottom-right causal)
  const bool sk = decode != 0;
  const int D = w->hidden, H = w->heads, T = Bn * L;
  B200_TRY(rmsnorm(x, D, B(Ly.attn_norm), w->rms_eps, nbuf, D, T, D, st));
  GemmEpilogue e;
  B200_TRY(linear(nbuf, D, B(Ly.qkv_w), D, qkv, 3 * D, T, 3 * D, D, e, sk, st));
  // decode step with enough (row, head) pairs to fill the chip without splitting the context: the attention kernel
  // rotates q / k and appends k / v itself (DecodeArgs::rope_k) -- one launch less per layer
  const bool fuse_rope = decode && fused_rope_enabled() && decode_attn_pick_splits(Bn, H, ctx_bound + 1) == 1;
  if (!fuse_rope)
    B200_TRY(rope_kv_write(qkv, kv_start, w->rope_cos, w->rope_sin, w->max_pos, kc, vc, Bn, H, L, decode ? 0 : q0,
                           decode ? state : nullptr, cap, st));
  if (!decode) {
    AttnArgs at{};
    at.q = qkv;
    at.k = kc;
    at.v = vc;
    at.o = nbuf;
    at.q_bs = static_cast<long long>(L) * 3 * D;
    at.q_rs = 3 * D;
    at.q_hs = 128;
    at.k_bs = at.v_bs = static_cast<long long>(H) * cap * 128;
    at.k_rs = at.v_rs = 128;
    at.k_hs = at.v_hs = static_cast<long long>(cap) * 128;
    at.o_bs = static_cast<long long>(L) * D;
    at.o_rs = D;
    at.o_hs = 128;
    at.B = Bn;
    at.H = H;
    at.Lq = L;
    at.Lk = q0 + L;
    at.kv_start = kv_start;
    at.kv_len = kv_len;
    at.causal = 1;
    at.scale_log2 = kLog2e / sqrtf(128.f);
    B200_TRY(flash_attn(at, 128, st));
  } else {
    DecodeArgs da{};
    da.q = qkv;
    da.q_rs = 3 * D;
    da.kc = kc;
    da.vc = vc;
    da.o = nbuf;
    da.o_rs = D;
    da.B = Bn;
    da.H = H;
    da.cap = cap;
    da.ctx = ctx_bound + 1;
    da.ctx_dev = state;
    da.ctx_add = 1;  // the token written by rope_kv_write above is visible to itself
    da.kv_start = kv_start;
    da.finished = finished;
    da.scale_log2 = kLog2e / sqrtf(128.f);
    da.splits = 0;
    if (fuse_rope) {
      da.splits = 1;
      da.rope_k = qkv + D;
      da.rope_v = qkv + 2 * D;
      da.kc_w = kc;
      da.vc_w = vc;
      da.cos_t = w->rope_cos;
      da.sin_t = w->rope_sin;
      da.max_pos = w->max_pos;
    }
    B200_TRY(decode_attn(da, dec_ws, dec_ws_bytes, st));
  }
  GemmEpilogue eo;
  eo.residual = x;
  eo.ldr = D;
  B200_TRY(linear(nbuf, D, B(Ly.o_w), D, x, D, T, D, D, eo, sk, st));
  B200_TRY(rmsnorm(x, D, B(Ly.mlp_norm), w->rms_eps, nbuf, D, T, D, st));
  GemmEpilogue eg;
  eg.act = kActSwiGLU;
  B200_TRY(linear(nbuf, D, B(Ly.gate_up_w), D, act, w->ffn, T, 2 * w->ffn, D, eg, sk, st));
  GemmEpilogue ed;
  ed.residual = x;
  ed.ldr = D;
  B200_TRY(linear(act, w->ffn, B(Ly.down_w), w->ffn, x, D, T, D, w->ffn, ed, sk, st));
  return 0;
}

static int llama_prefill(const b200_llama_weights* w, bf16* x, const int* kv_start, const int* kv_len,
                         const b200_kv_cache* c, int Bn, int L, void* logits, int all_logits, int logits_fp32, void* ws,
                         size_t ws_bytes, cudaStream_t st, int q0 = 0) {
  if (Bn <= 0 || L <= 0) return 0;
  const int D = w->hidden;
  if (D / w->heads != 128 || D % w->heads) return fail(-2, "llama: head_dim must be 128");
  if (q0 < 0) return fail(-2, "llama_prefill: negative first slot %d", q0);
  if (q0 + L > c->cap) return fail(-2, "llama_prefill: %d + %d slots exceed KV capacity %d", q0, L, c->cap);
  if (ws_bytes < llama_prefill_ws(w, Bn, L, all_logits)) return fail(-2, "llama_prefill: workspace too small");
  Arena a(ws, ws_bytes);
  const size_t T = static_cast<size_t>(Bn) * L;
  bf16* nbuf = a.take<bf16>(T * D);
  bf16* qkv = a.take<bf16>(T * 3 * D);
  bf16* act = a.take<bf16>(T * w->ffn);
  bf16* fin = a.take<bf16>(static_cast<size_t>(all_logits ? T : Bn) * D);
  for (int l = 0; l < w->n_layers; ++l) {
    bf16* kc = static_cast<bf16*>(c->k) + l * c->layer_stride;
    bf16* vc = static_cast<bf16*>(c->v) + l * c->layer_stride;
    B200_TRY(llama_layer(w, w->layers[l], x, nbuf, qkv, act, kv_start, kv_len, kc, vc, c->cap, Bn, L, 0, nullptr, 0,
                         nullptr, 0, nullptr, st, q0));
  }
  if (logits != nullptr) {
    GemmEpilogue e;
    e.out_fp32 = logits_fp32;
    if (all_logits) {
      B200_TRY(rmsnorm(x, D, B(w->final_norm), w->rms_eps, fin, D, static_cast<int>(T), D, st));
      B200_TRY(gemm_bf16_tn(fin, D, B(w->lm_head), D, logits, w->vocab, static_cast<int>(T), w->vocab, D, e, 0, st));
    } else {
      // HF generate consumes logits[:, -1] only: normalise and project the last row of every sample
      B200_TRY(rmsnorm(x + static_cast<size_t>(L - 1) * D, static_cast<long long>(L) * D, B(w->final_norm), w->rms_eps,
                       fin, D, Bn, D, st));
      B200_TRY(gemm_bf16_tn(fin, D, B(w->lm_head), D, logits, w->vocab, Bn, w->vocab, D, e, 0, st));
    }
  }
  return 0;
}

static size_t llama_decode_ws(const b200_llama_weights* w, int Bn, int cap) {
  size_t s = 0;
  s += al(static_cast<size_t>(Bn) * w->hidden * 2) * 2;  // x, norm/context
  s += al(static_cast<size_t>(Bn) * 3 * w->hidden * 2);  // qkv
  s += al(static_cast<size_t>(Bn) * w->ffn * 2);         // swiglu
  s += al(static_cast<size_t>(Bn) * w->vocab * 2);       // logits
  s += al(decode_attn_workspace_bytes(Bn, w->heads, 16));
  (void)cap;
  return s;
}

static int llama_decode_step(const b200_llama_weights* w, int* tokens, int* state, const int* kv_start,
                             const b200_kv_cache* c, int Bn, int ctx_bound, int* finished, int eos_id, int pad_id,
                             int* history, int hist_ld, void* logits_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (Bn <= 0) return 0;
  const int D = w->hidden;
  if (ctx_bound + 1 > c->cap) return fail(-2, "llama_decode_step: KV cache full (%d slots)", c->cap);
  if (ws_bytes < llama_decode_ws(w, Bn, c->cap)) return fail(-2, "llama_decode_step: workspace too small");
  Arena a(ws, ws_bytes);
  bf16* x = a.take<bf16>(static_cast<size_t>(Bn) * D);
  bf16* nbuf = a.take<bf16>(static_cast<size_t>(Bn) * D);
  bf16* qkv = a.take<bf16>(static_cast<size_t>(Bn) * 3 * D);
  bf16* act = a.take<bf16>(static_cast<size_t>(Bn) * w->ffn);
  bf16* logits = a.take<bf16>(static_cast<size_t>(Bn) * w->vocab);
  const size_t dws_bytes = decode_attn_workspace_bytes(Bn, w->heads, 16);
  void* dws = a.take<uint8_t>(dws_bytes);
  if (logits_out != nullptr) logits = static_cast<bf16*>(logits_out);
  B200_TRY(embed_rows(tokens, B(w->embed_tokens), x, D, Bn, D, w->vocab, st));
  for (int l = 0; l < w->n_layers; ++l) {
    bf16* kc = static_cast<bf16*>(c->k) + l * c->layer_stride;
    bf16* vc = static_cast<bf16*>(c->v) + l * c->layer_stride;
    B200_TRY(llama_layer(w, w->layers[l], x, nbuf, qkv, act, kv_start, nullptr, kc, vc, c->cap, Bn, 1, 1, state,
                         ctx_bound, dws, dws_bytes, logits_out == nullptr ? finished : nullptr, st));
  }
  B200_TRY(rmsnorm(x, D, B(w->final_norm), w->rms_eps, nbuf, D, Bn, D, st));
  GemmEpilogue e;
  B200_TRY(linear(nbuf, D, B(w->lm_head), D, logits, w->vocab, Bn, w->vocab, D, e, true, st));
  B200_TRY(argmax_rows(logits, 0, w->vocab, Bn, w->vocab, tokens, finished, eos_id, pad_id, history, hist_ld, 0,
                       state + 1, st));
  B200_TRY(bump_counters(state, st));
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_vit_workspace_bytes(const b200_vit_weights* w, int n_img) { return vit_ws(w, n_img); }
int b200_vit_forward(const b200_vit_weights* w, const void* pixels, void* hidden_out, int n_img, void* workspace,
                     size_t workspace_bytes, b200_stream_t stream) {
  return vit_forward(w, static_cast<const bf16*>(pixels), static_cast<bf16*>(hidden_out), n_img, workspace,
                     workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t b200_pooler_workspace_bytes(const b200_pooler_weights* w, int Bn, int S) { return pooler_ws(w, Bn, S); }
int b200_pooler_forward(const b200_pooler_weights* w, const void* src, int64_t src_ld, const int32_t* gather_map,
                        const int32_t* kv_len, int Bn, int S, int keep, void* out, int out_tokens, void* workspace,
                        size_t workspace_bytes, b200_stream_t stream) {
  return pooler_forward(w, static_cast<const bf16*>(src), src_ld, gather_map, kv_len, Bn, S, keep,
                        static_cast<bf16*>(out), out_tokens, workspace, workspace_bytes,
                        static_cast<cudaStream_t>(stream));
}

size_t b200_segmask_workspace_bytes(int n_maps) { return segmask_workspace_bytes(n_maps); }
int b200_segmask_forward(const b200_segmask_weights* w, const uint8_t* cls, int n_maps, void* out, int64_t out_ld,
                         const int32_t* out_row_map, void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  const bf16* cw[5];
  const bf16* cb[5];
  for (int i = 0; i < 5; ++i) {
    cw[i] = static_cast<const bf16*>(w->conv_w[i]);
    cb[i] = static_cast<const bf16*>(w->conv_b[i]);
  }
  return segmask_forward(cls, n_maps, static_cast<const bf16*>(w->emb), cw, cb, static_cast<bf16*>(out), out_ld,
                         out_row_map, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t b200_projector_workspace_bytes(const b200_projector_weights* w, int n_tokens) {
  return al(static_cast<size_t>(n_tokens) * w->hidden * 2);
}
int b200_projector_pack(const b200_projector_weights* w, const void* tokens, int n_tokens, const int32_t* row_map,
                        const int32_t* text_ids, const void* embed_table, int vocab, void* embeds, int n_rows,
                        void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_tokens > 0) {
    if (workspace_bytes < b200_projector_workspace_bytes(w, n_tokens))
      return fail(-2, "projector_pack: workspace too small");
    bf16* h = static_cast<bf16*>(workspace);
    GemmEpilogue e0;
    e0.bias = B(w->b0);
    e0.act = kActGeluErf;
    B200_TRY(gemm_bf16_tn(B(tokens), w->in_dim, B(w->w0), w->in_dim, h, w->hidden, n_tokens, w->hidden, w->in_dim, e0,
                          0, st));
    GemmEpilogue e2;
    e2.bias = B(w->b2);
    e2.row_map = row_map;
    B200_TRY(gemm_bf16_tn(h, w->hidden, B(w->w2), w->hidden, embeds, w->hidden, n_tokens, w->hidden, w->hidden, e2, 0,
                          st));
  }
  if (text_ids != nullptr)
    B200_TRY(embed_rows(text_ids, B(embed_table), static_cast<bf16*>(embeds), w->hidden, n_rows, w->hidden, vocab, st));
  return 0;
}

int b200_projector_gather(const b200_projector_weights* w, const void* tokens, int n_tokens, void* const* peers,
                          int n_peers, size_t slot_offset_bytes, void* workspace, size_t workspace_bytes,
                          b200_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_tokens <= 0) return 0;
  if (n_peers < 1 || n_peers > 8 || peers == nullptr) return fail(-2, "projector_gather: 1..8 peer buffers expected");
  if (slot_offset_bytes % 16 != 0) return fail(-2, "projector_gather: slot offset must be 16-byte aligned");
  if (workspace_bytes < b200_projector_workspace_bytes(w, n_tokens))
    return fail(-2, "projector_gather: workspace too small");
  bf16* h = static_cast<bf16*>(workspace);
  GemmEpilogue e0;
  e0.bias = B(w->b0);
  e0.act = kActGeluErf;
  B200_TRY(gemm_bf16_tn(B(tokens), w->in_dim, B(w->w0), w->in_dim, h, w->hidden, n_tokens, w->hidden, w->in_dim, e0, 0,
                        st));
  GemmEpilogue e2;
  e2.bias = B(w->b2);
  e2.n_peers = n_peers;
  for (int p = 0; p < n_peers; ++p) {
    if (peers[p] == nullptr) return fail(-2, "projector_gather: peer %d is NULL", p);
    e2.peer_c[p] = static_cast<uint8_t*>(peers[p]) + slot_offset_bytes;
  }
  return gemm_bf16_tn(h, w->hidden, B(w->w2), w->hidden, e2.peer_c[0], w->hidden, n_tokens, w->hidden, w->hidden, e2, 0,
                      st);
}

size_t b200_llama_prefill_workspace_bytes(const b200_llama_weights* w, int Bn, int L, int all_logits) {
  return llama_prefill_ws(w, Bn, L, all_logits);
}
int b200_llama_prefill(const b200_llama_weights* w, void* x, const int32_t* kv_start, const int32_t* kv_len,
                       const b200_kv_cache* cache, int Bn, int L, void* logits, int all_logits, int logits_fp32,
                       void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  return llama_prefill(w, static_cast<bf16*>(x), kv_start, kv_len, cache, Bn, L, logits, all_logits, logits_fp32,
                       workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int b200_llama_prefill_from(const b200_llama_weights* w, void* x, const int32_t* kv_start, const int32_t* kv_len,
                            const b200_kv_cache* cache, int Bn, int Lq, int q0, void* logits, int all_logits,
                            int logits_fp32, void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  return llama_prefill(w, static_cast<bf16*>(x), kv_start, kv_len, cache, Bn, Lq, logits, all_logits, logits_fp32,
                       workspace, workspace_bytes, static_cast<cudaStream_t>(stream), q0);
}

int b200_decode_tile_width(int rows, int n, int sms, int ctas_per_sm) {
  return decode_bn(rows, n, sms, ctas_per_sm == 2 ? 2 : 1);
}

size_t b200_llama_decode_workspace_bytes(const b200_llama_weights* w, int Bn, int cap) {
  return llama_decode_ws(w, Bn, cap);
}
int b200_llama_decode_step(const b200_llama_weights* w, int32_t* tokens, int32_t* state, const int32_t* kv_start,
                           const b200_kv_cache* cache, int Bn, int ctx_bound, int32_t* finished, int eos_id, int pad_id,
                           int32_t* history, int hist_ld, void* logits_out, void* workspace, size_t workspace_bytes,
                           b200_stream_t stream) {
  return llama_decode_step(w, tokens, state, kv_start, cache, Bn, ctx_bound, finished, eos_id, pad_id, history, hist_ld,
                           logits_out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
