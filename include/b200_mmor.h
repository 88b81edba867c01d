/*
 * b200_mmor.h -- C ABI of libb200mmor.so, the B200 (sm_100a) implementation of MM2SG's multimodal hot path
 * (reference: egeozsoy/MM-OR, scene_graph_generation/LLaVA/llava).
 *
 * The reference has no FFI: the path sits behind Python objects that call HF transformers / torch ops
 * (SURVEY.md 8b). These entry points are what the Python mirror of that interface (mm_or_b200/model/*.py) binds
 * with ctypes. Each entry point cites the reference call site whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; bf16 tensors are row-major and dense
 *     unless a leading dimension is given; sizes are in elements.
 *   - every function returns 0 on success, a negative code on failure (-2 bad argument, -5 CUDA error,
 *     -6 driver entry point missing); b200_last_error() returns the message for the calling thread.
 *   - functions enqueue work on `stream` (a cudaStream_t passed as void*) and never synchronise or allocate;
 *     scratch memory is passed in as (workspace, workspace_bytes) and sized by the matching *_workspace_bytes.
 *   - re-entrant per stream; no global mutable state besides lazily initialised function attributes.
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

/* ------------------------------------------------------------------------------------------------------------
 * Dense linear:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)      bf16 in, fp32 accumulate (tcgen05 / TMEM)
 * Replaces every nn.Linear on the path: CLIP q/k/v/out_proj, fc1, fc2 (clip_encoder.py:48 -> HF CLIPEncoderLayer),
 * BERT pooler dense layers (multimodal_projector/builder.py:173), mm_projector (llava_arch.py:182),
 * Llama q/k/v/o/gate/up/down_proj and lm_head (llava_llama.py:93).
 *   act: 0 none, 1 quick_gelu, 2 gelu(erf), 3 SwiGLU over interleaved (gate, up) column pairs (C has N/2 columns)
 *   bias [N] / residual [M, ldr] / row_map [M] (output row per logical row, <0 drops the row) may be NULL
 *   out_fp32: C is float instead of bf16.  bn_hint: 0 = auto tile width, or 32/64/128/256.
 * ---------------------------------------------------------------------------------------------------------- */
int b200_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                   const void* bias, const void* residual, int ldr, const int32_t* row_map, int act, int out_fp32,
                   int bn_hint, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_MMOR_H_ */
