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
  int act = kActNone;
  int out_fp32 = 0;  // 0: C is bf16, 1: C is fp32
};

// bn_hint: 0 = choose automatically, else one of 32/64/128/256
int gemm_bf16_tn(const bf16* A, int lda, const bf16* B, int ldb, void* C, int ldc, int M, int N, int K,
                 const GemmEpilogue& epi, int bn_hint, cudaStream_t stream);

}  // namespace b200
