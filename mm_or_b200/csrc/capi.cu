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
  return gemm_bf16_tn(static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), ldw, C, ldc, M, N, K, e,
                      bn_hint, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
