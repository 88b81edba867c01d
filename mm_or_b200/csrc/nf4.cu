// nf4.cu -- 4-bit NormalFloat storage of the frozen base weights for the QLoRA recipe.
//
// Reference call site: LLaVA/llava/train/train.py:1098-1114 loads the decoder with
// BitsAndBytesConfig(load_in_4bit, bnb_4bit_quant_type='nf4', bnb_4bit_use_double_quant=True, compute dtype bf16), i.e.
// every nn.Linear outside mm_projector / image_pooler is replaced by bitsandbytes' Linear4bit: the weight is stored as
// 4-bit codes with one fp32 absmax per block of 64 values and is DEQUANTISED to bf16 in front of every matmul. The
// arithmetic lives in bitsandbytes==0.41.0 (SGG/pyproject.toml:18), which is not vendored and not installed here; it
// is restated from its published algorithm (csrc/kernels.cu: kQuantizeBlockwise<.., NF4>, dQuantizeNF4,
// dDequantizeNF4, kDequantizeBlockwise; QLoRA, Dettmers et al. 2023, appendix E for the 16 code values):
//   absmax_b = max |w| over the block (fp32);   code = nearest NF4 level of w * (1 / absmax_b)  (decision tree over the
//   midpoints between levels; a NaN -- the all-zero block, 0 * inf -- falls through to code 0);   two codes per byte,
//   even element in the high nibble;   dequantised value = level[code] * absmax_b, rounded to bf16.
// The second-level ("double") quantisation of the absmax vector is load-time host work (mm_or_b200/train/nf4.py).
//
// Byte / integer work, HBM bound: 2 B read + 0.5625 B written per weight when quantising (once, at load), 0.5625 B
// read + 2 B written when dequantising. One warp owns one block of 64 weights: 128 B in, 32 B of codes out (or the
// reverse), consecutive warps take consecutive blocks, so every global access of a warp is one contiguous segment.
// Like ptv3.cu / train_extras.cu the file also compiles with g++ -DB200_EMU against tests/emu/cuda_emu.h: the CPU
// test-suite runs these kernels against the numpy restatement (oracle/nf4_oracle.py); the bar is bit-exact.
#ifdef B200_EMU
#include "cuda_emu.h"
#include "emu_common.h"
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define NF4_CHECK_LAUNCH(what)
#else
#include "../../include/b200_mmor.h"
#include "common.h"
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define NF4_CHECK_LAUNCH(what)                                                                         \
  do {                                                                                                 \
    cudaError_t _e = cudaGetLastError();                                                               \
    if (_e != cudaSuccess) return fail(-5, "%s: launch failed: %s", what, cudaGetErrorString(_e));     \
  } while (0)
#endif

namespace b200 {
namespace nf4 {

constexpr int kBlock = 64;  // weights per quantisation block (bitsandbytes' blocksize for 4-bit types)

__device__ __forceinline__ float bf16_to_f32(uint16_t u) {
  uint32_t w = (uint32_t)u << 16;
  float f;
  memcpy(&f, &w, 4);
  return f;
}
__device__ __forceinline__ uint16_t f32_to_bf16(float x) {  // round to nearest even
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// the 16 NormalFloat levels (quantiles of N(0, 1) normalised to [-1, 1], zero exactly representable)
__device__ __forceinline__ float level(int code) {
  switch (code & 15) {
    case 0: return -1.0f;
    case 1: return -0.6961928009986877f;
    case 2: return -0.5250730514526367f;
    case 3: return -0.39491748809814453f;
    case 4: return -0.28444138169288635f;
    case 5: return -0.18477343022823334f;
    case 6: return -0.09105003625154495f;
    case 7: return 0.0f;
    case 8: return 0.07958029955625534f;
    case 9: return 0.16093020141124725f;
    case 10: return 0.24611230194568634f;
    case 11: return 0.33791524171829224f;
    case 12: return 0.44070982933044434f;
    case 13: return 0.5626170039176941f;
    case 14: return 0.7229568362236023f;
    default: return 1.0f;
  }
}

// nearest level by the decision tree over the midpoints (strict >, so a value on a midpoint takes the lower level and
// NaN takes code 0)
__device__ __forceinline__ int nearest_code(float x) {
  if (x > 0.03979014977812767f) {
    if (x > 0.3893125355243683f) {
      if (x > 0.6427869200706482f) return x > 0.8614784181118011f ? 15 : 14;
      return x > 0.5016634166240692f ? 13 : 12;
    }
    if (x > 0.2035212516784668f) return x > 0.2920137718319893f ? 11 : 10;
    return x > 0.1202552504837513f ? 9 : 8;
  }
  if (x > -0.33967943489551544f) {
    if (x > -0.13791173323988914f) return x > -0.045525018125772476f ? 7 : 6;
    return x > -0.23460740596055984f ? 5 : 4;
  }
  if (x > -0.6106329262256622f) return x > -0.4599952697753906f ? 3 : 2;
  return x > -0.8480964004993439f ? 1 : 0;
}

// one warp per block of 64 weights; lane l holds weights 2l and 2l+1 (one 4-byte load) and writes byte l of the block
__global__ void __launch_bounds__(128) quantize_kernel(const uint16_t* __restrict__ w, long long n_blocks,
                                                       uint8_t* __restrict__ packed, float* __restrict__ absmax) {
  const long long blk = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const bool live = blk < n_blocks;  // (whole warps stay in the shuffles)
  float a = 0.f, b = 0.f;
  if (live) {
    const uint32_t two = *reinterpret_cast<const uint32_t*>(w + blk * kBlock + 2 * lane);
    a = bf16_to_f32((uint16_t)(two & 0xffffu));
    b = bf16_to_f32((uint16_t)(two >> 16));
  }
  float m = fmaxf(fabsf(a), fabsf(b));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if (!live) return;
  const float inv = 1.0f / m;
  packed[blk * (kBlock / 2) + lane] = (uint8_t)((nearest_code(a * inv) << 4) | nearest_code(b * inv));
  if (lane == 0) absmax[blk] = m;
}

__global__ void __launch_bounds__(128) dequantize_kernel(const uint8_t* __restrict__ packed,
                                                         const float* __restrict__ absmax, long long n_blocks,
                                                         uint16_t* __restrict__ out) {
  const long long blk = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (blk >= n_blocks) return;
  const float m = absmax[blk];
  const uint8_t q = packed[blk * (kBlock / 2) + lane];
  const uint32_t lo = f32_to_bf16(level(q >> 4) * m), hi = f32_to_bf16(level(q & 15) * m);
  *reinterpret_cast<uint32_t*>(out + blk * kBlock + 2 * lane) = lo | (hi << 16);
}

}  // namespace nf4
}  // namespace b200

using namespace b200;

extern "C" {

int b200_nf4_quantize(const void* w_bf16, int64_t n, uint8_t* packed, float* absmax, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) return 0;
  if (!w_bf16 || !packed || !absmax || n < 0 || n % nf4::kBlock != 0)
    return fail(-2, "b200_nf4_quantize: bad argument (n=%lld must be a positive multiple of 64)", (long long)n);
  const long long n_blocks = n / nf4::kBlock;
  if ((n_blocks + 3) / 4 > 0x7fffffffLL) return fail(-2, "b200_nf4_quantize: tensor too large (n=%lld)", (long long)n);
  LaunchScope ls(kFamTrain, stream, 2.5625 * (double)n, 0.0);
  B200_LAUNCH(nf4::quantize_kernel, dim3((unsigned)((n_blocks + 3) / 4)), dim3(128), 0, stream,
              reinterpret_cast<const uint16_t*>(w_bf16), n_blocks, packed, absmax);
  NF4_CHECK_LAUNCH("b200_nf4_quantize");
  return 0;
}

int b200_nf4_dequantize(const uint8_t* packed, const float* absmax, int64_t n, void* out_bf16, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) return 0;
  if (!packed || !absmax || !out_bf16 || n < 0 || n % nf4::kBlock != 0)
    return fail(-2, "b200_nf4_dequantize: bad argument (n=%lld must be a positive multiple of 64)", (long long)n);
  const long long n_blocks = n / nf4::kBlock;
  if ((n_blocks + 3) / 4 > 0x7fffffffLL) return fail(-2, "b200_nf4_dequantize: tensor too large (n=%lld)", (long long)n);
  LaunchScope ls(kFamTrain, stream, 2.5625 * (double)n, 0.0);
  B200_LAUNCH(nf4::dequantize_kernel, dim3((unsigned)((n_blocks + 3) / 4)), dim3(128), 0, stream, packed, absmax,
              n_blocks, reinterpret_cast<uint16_t*>(out_bf16));
  NF4_CHECK_LAUNCH("b200_nf4_dequantize");
  return 0;
}

}  // extern "C"
