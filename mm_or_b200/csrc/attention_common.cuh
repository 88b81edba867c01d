// Shared pieces of the tcgen05 attention kernels (forward: attention_sm100.cu, backward: attention_bwd_sm100.cu).
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace b200 {

// 2^x on the SFU (ex2.approx.ftz): -inf -> 0, no denormal fix-up code around it
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// issue one box load with (token, head, sample) routed to the tensor map's dimension order
__device__ __forceinline__ void tc_load(void* dst, const CUtensorMap* tm, uint64_t* bar, const int (&dim)[3], int d0,
                                        int token, int head, int sample) {
  int c[4] = {d0, 0, 0, 0};
  c[dim[0]] = token;
  c[dim[1]] = head;
  c[dim[2]] = sample;
  tma_load_4d(dst, tm, bar, c[0], c[1], c[2], c[3]);
}

typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 4-D bf16 map over (d, then token / head / sample sorted by ascending stride -- cuTensorMapEncodeTiled wants every
// stride to be a multiple of the previous one, which holds for all layouts on this path once sorted); box = 64 d x
// box_rows tokens. dim_of[0..2] receive the tensor-map dimension (1..3) of token / head / sample.
static inline int make_tmap_attn(CUtensorMap* out, const bf16* base, int D, int L, int H, int B, long long rs, long long hs,
                          long long bs, int box_rows, int (&dim_of)[3]) {
  static EncodeTiledFn4 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return fail(-6, "cuTensorMapEncodeTiled entry point not available");
    fn = reinterpret_cast<EncodeTiledFn4>(p);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (rs * 2) % 16 != 0 || (hs * 2) % 16 != 0 || (bs * 2) % 16 != 0)
    return fail(-2, "flash_attn: q/k/v pointers and strides must be 16-byte aligned");
  struct Dim {
    long long extent, stride_b;
    int role;  // 0 token, 1 head, 2 sample
  } d[3] = {{L, rs * 2, 0}, {H, hs * 2, 1}, {B, bs * 2, 2}};
  // an extent-1 dimension may come with any stride (even 0): park it last with a legal one
  long long span = D * 2;
  for (int i = 0; i < 3; ++i)
    if (d[i].extent > 1 && d[i].stride_b * d[i].extent > span) span = d[i].stride_b * d[i].extent;
  for (int i = 0; i < 3; ++i)
    if (d[i].extent <= 1) d[i].stride_b = span;
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (d[j].stride_b < d[i].stride_b || (d[j].stride_b == d[i].stride_b && d[j].extent > d[i].extent)) {
        const Dim t = d[i];
        d[i] = d[j];
        d[j] = t;
      }
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(D), 0, 0, 0};
  cuuint64_t strides[3];
  cuuint32_t box[4] = {64, 1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i) {
    dims[i + 1] = static_cast<cuuint64_t>(d[i].extent);
    strides[i] = static_cast<cuuint64_t>(d[i].stride_b);
    dim_of[d[i].role] = i + 1;
    if (d[i].role == 0) box[i + 1] = static_cast<cuuint32_t>(box_rows);
  }
  CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    return fail(-6, "cuTensorMapEncodeTiled (attention) failed with CUresult %d (strides %lld/%lld/%lld B)", (int)rc,
                d[0].stride_b, d[1].stride_b, d[2].stride_b);
  return 0;
}


}  // namespace b200
