// ptv3.cu -- point-cloud branch of the image pooler: PointTransformerV3 (cls_mode) operators, fp32.
//
// Reference: ImageEmbeddingPooler._encode_pc (model/multimodal_projector/builder.py:93-148) ->
// PointTransformerV3.forward (model/multimodal_projector/pointtransformerv3.py:982-1002), which the reference runs in
// fp32 (builder.py:95 `.float()`, autocast disabled builder.py:176) on spconv / torch_scatter / flash-attn kernels.
//
// B200 design (DESIGN.md "Point clouds"): points are kept PHYSICALLY sorted by their z-order code at every level, so
//   * the "z" serialization is the identity, pooling clusters (code >> 3) are contiguous runs and the pooled level is
//     born sorted -- no unique / scatter, a head-flag scan is the whole pooling plan;
//   * the submanifold-convolution neighbour of a voxel is found by binary search of its z-code in the sorted code
//     array (L2-resident) -- no hash table to build or probe;
//   * convolutions and linears are one gather-GEMM kernel (K flattened over taps x channels) with bias / folded
//     BatchNorm / GELU / residual epilogues; gathers hit neighbouring rows because z-order keeps space-neighbours close.
// Arithmetic is fp32 on the CUDA cores: channel widths are 32..512 and the reference runs true fp32 (TF32 / bf16 tensor
// cores would break parity); attention rounds q, k, v and its output to fp16 exactly where flash-attn does.
//
// The same source is compiled with g++ -DB200_EMU (tests/emu/) so that the CPU test-suite executes these kernels
// against the oracle; only SIMT features that the emulator provides are used here.
#ifdef B200_EMU
#include "cuda_emu.h"
#include "emu_common.h"
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define B200_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::st().dyn_smem.data())
#else
#include <cuda_fp16.h>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/b200_mmor.h"
#include "common.h"
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define B200_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type* name = reinterpret_cast<type*>(name##_raw)
#endif

namespace b200 {
namespace pc {

// ---------------------------------------------------------------------------------------------------------------
// rounding helpers (identical results on device and in the emulator)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float round_f16(float x) {
#ifdef B200_EMU
  // round-to-nearest-even to IEEE binary16 and back (normal / subnormal / overflow to inf)
  uint32_t u;
  memcpy(&u, &x, 4);
  uint32_t sign = u & 0x80000000u;
  uint32_t a = u & 0x7fffffffu;
  float r;
  if (a >= 0x7f800000u) {
    r = x;
  } else if (a >= 0x477ff000u) {  // >= 65520 rounds to inf
    uint32_t inf = sign | 0x7f800000u;
    memcpy(&r, &inf, 4);
    return r;
  } else if (a < 0x38800000u) {  // below 2^-14: subnormal half, spacing 2^-24
    float ax = fabsf(x);
    float q = nearbyintf(ax * 16777216.0f) / 16777216.0f;
    r = sign ? -q : q;
  } else {
    uint32_t lsb = (a >> 13) & 1u;
    a += 0xfffu + lsb;
    a &= ~0x1fffu;
    a |= sign;
    memcpy(&r, &a, 4);
  }
  return r;
#else
  return __half2float(__float2half_rn(x));
#endif
}

__device__ __forceinline__ uint16_t bf16_bits(float x) {
#ifdef B200_EMU
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
#else
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
#endif
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---------------------------------------------------------------------------------------------------------------
// geometry: grid coordinates (pointtransformerv3.py:96-98)
// ---------------------------------------------------------------------------------------------------------------
// single block: min over all points of x / y / z, fixed reduction order
__global__ void coord_min_kernel(const float* __restrict__ pts, int ld, int n, float* __restrict__ out3) {
  __shared__ float red[3][1024];
  float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* p = pts + (size_t)i * ld;
    m0 = fminf(m0, p[0]);
    m1 = fminf(m1, p[1]);
    m2 = fminf(m2, p[2]);
  }
  red[0][threadIdx.x] = m0;
  red[1][threadIdx.x] = m1;
  red[2][threadIdx.x] = m2;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      for (int d = 0; d < 3; ++d) red[d][threadIdx.x] = fminf(red[d][threadIdx.x], red[d][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) out3[threadIdx.x] = red[threadIdx.x][0];
}

// grid = trunc((coord - min) / grid_size) with IEEE fp32 subtraction and division, as torch.div(rounding_mode="trunc")
__global__ void grid_coord_kernel(const float* __restrict__ pts, int ld, int n, const float* __restrict__ min3,
                                  float grid_size, int* __restrict__ grid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pts + (size_t)i * ld;
  for (int d = 0; d < 3; ++d) grid[(size_t)i * 3 + d] = (int)truncf(__fdiv_rn(__fsub_rn(p[d], min3[d]), grid_size));
}

// single block: max over all 3n grid coordinates
__global__ void grid_max_kernel(const int* __restrict__ grid, int n3, int* __restrict__ out) {
  __shared__ int red[1024];
  int m = 0;
  for (int i = threadIdx.x; i < n3; i += blockDim.x) m = grid[i] > m ? grid[i] : m;
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] = red[threadIdx.x] > red[threadIdx.x + s] ? red[threadIdx.x]
                                                                                        : red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// ---------------------------------------------------------------------------------------------------------------
// serialization codes (serialization/default.py:9-25, z_order.py:74-103, hilbert.py:91-193)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long spread3(unsigned v) {  // bit i -> bit 3i (16 bits in)
  unsigned long long x = v & 0xffffu;
  x = (x | (x << 32)) & 0x001f00000000ffffull;
  x = (x | (x << 16)) & 0x001f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ unsigned long long z_key(unsigned x, unsigned y, unsigned z) {
  return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}
__device__ __forceinline__ unsigned long long hilbert_key(unsigned x0, unsigned x1, unsigned x2, int depth) {
  unsigned X[3] = {x0, x1, x2};
  for (int bit = depth - 1; bit >= 0; --bit) {  // Skilling's transform, top bit first
    unsigned q = 1u << bit, p = q - 1u;
    for (int d = 0; d < 3; ++d) {
      if (X[d] & q) {
        X[0] ^= p;
      } else {
        unsigned t = (X[0] ^ X[d]) & p;
        X[0] ^= t;
        X[d] ^= t;
      }
    }
  }
  unsigned long long g = z_key(X[0], X[1], X[2]);
  for (int s = 1; s < 3 * depth; s <<= 1) g ^= g >> s;  // Gray -> binary over the interleaved string
  return g;
}

// order: 0 z, 1 z-trans, 2 hilbert, 3 hilbert-trans ("trans" swaps x and y)
__global__ void encode_kernel(const int* __restrict__ grid, const int* __restrict__ batch, int n, int depth, int order,
                              long long* __restrict__ code) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (unsigned)grid[(size_t)i * 3], y = (unsigned)grid[(size_t)i * 3 + 1], z = (unsigned)grid[(size_t)i * 3 + 2];
  if (order & 1) {
    unsigned t = x;
    x = y;
    y = t;
  }
  unsigned mask = (1u << depth) - 1u;  // both encoders keep the low `depth` bits (z_order.py:89-93, hilbert.py:137-143)
  x &= mask;
  y &= mask;
  z &= mask;
  unsigned long long c = (order & 2) ? hilbert_key(x, y, z, depth) : z_key(x, y, z);
  code[i] = (long long)(((unsigned long long)batch[i] << (3 * depth)) | c);
}

__global__ void iota_kernel(int* __restrict__ v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// dst[i, 0:width] = src[idx[i], 0:width] over 4-byte elements
__global__ void gather_rows_kernel(const uint32_t* __restrict__ src, int ld_src, const int* __restrict__ idx, int n,
                                   int width, uint32_t* __restrict__ dst, int ld_dst) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * width) return;
  int i = (int)(t / width), c = (int)(t % width);
  dst[(size_t)i * ld_dst + c] = src[(size_t)idx[i] * ld_src + c];
}

// ---------------------------------------------------------------------------------------------------------------
// submanifold-convolution neighbour table: nbr[i, t] = index of the active voxel at grid[i] + tap(t) - centre, or -1.
// Points are sorted by z-code, so the lookup is a binary search. dup[0] is raised when two points share a voxel.
// ---------------------------------------------------------------------------------------------------------------
__global__ void neighbor_kernel(const long long* __restrict__ zcode, const int* __restrict__ grid,
                                const int* __restrict__ batch, int n, int depth, int ksize, int* __restrict__ nbr,
                                int* __restrict__ dup) {
  int taps = ksize * ksize * ksize;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * taps) return;
  int i = (int)(t / taps), tap = (int)(t % taps);
  int c = ksize / 2;
  int dx = tap / (ksize * ksize) - c, dy = (tap / ksize) % ksize - c, dz = tap % ksize - c;
  int x = grid[(size_t)i * 3] + dx, y = grid[(size_t)i * 3 + 1] + dy, z = grid[(size_t)i * 3 + 2] + dz;
  int lim = 1 << depth;
  int found = -1;
  if (x >= 0 && y >= 0 && z >= 0 && x < lim && y < lim && z < lim) {
    long long key = (long long)(((unsigned long long)batch[i] << (3 * depth)) | z_key((unsigned)x, (unsigned)y, (unsigned)z));
    int lo = 0, hi = n;  // first position with zcode >= key
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (zcode[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo < n && zcode[lo] == key) found = lo;
  }
  nbr[t] = found;
  if (tap == taps / 2) {
    if (found != i) dup[0] = 1;  // the centre tap must find the point itself: otherwise codes are unsorted / duplicated
    if (i + 1 < n && zcode[i + 1] == zcode[i]) dup[0] = 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pooling plan (SerializedPooling, pointtransformerv3.py:655-672): clusters = runs of equal (zcode >> shift).
// single block: seg_start[c] = first member of cluster c, seg_start[n_out] = n, n_out_dev[0] = n_out.
// ---------------------------------------------------------------------------------------------------------------
__global__ void pool_plan_kernel(const long long* __restrict__ zcode, int n, int shift, int* __restrict__ seg_start,
                                 int* __restrict__ n_out_dev) {
  __shared__ int part[1024];
  int T = blockDim.x, t = threadIdx.x;
  int chunk = (n + T - 1) / T;
  int b = t * chunk, e = b + chunk < n ? b + chunk : n;
  int cnt = 0;
  for (int i = b; i < e; ++i) cnt += (i == 0 || (zcode[i] >> shift) != (zcode[i - 1] >> shift)) ? 1 : 0;
  part[t] = cnt;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int j = 0; j < T; ++j) {
      int c = part[j];
      part[j] = run;
      run += c;
    }
    n_out_dev[0] = run;
    seg_start[run] = n;
  }
  __syncthreads();
  int pos = part[t];
  for (int i = b; i < e; ++i)
    if (i == 0 || (zcode[i] >> shift) != (zcode[i - 1] >> shift)) seg_start[pos++] = i;
}

// next level's coordinates: grid >> pd and batch of each cluster's first member (pointtransformerv3.py:695-701).
// The cluster count lives on the device: launched for the upper bound and bounded by *n_out_dev.
__global__ void pool_coords_kernel(const int* __restrict__ seg_start, const int* __restrict__ n_out_dev,
                                   const int* __restrict__ grid, const int* __restrict__ batch, int pd,
                                   int* __restrict__ grid_o, int* __restrict__ batch_o) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_out_dev[0]) return;
  int h = seg_start[c];
  for (int d = 0; d < 3; ++d) grid_o[(size_t)c * 3 + d] = grid[(size_t)h * 3 + d] >> pd;
  batch_o[c] = batch[h];
}

// off[b] = first index with batch >= b, b = 0..n_clouds (batch is sorted)
__global__ void cloud_offsets_kernel(const int* __restrict__ batch, int n, int n_clouds, int* __restrict__ off) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_clouds) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (batch[mid] < b) lo = mid + 1; else hi = mid;
  }
  off[b] = lo;
}

// ---------------------------------------------------------------------------------------------------------------
// gather-GEMM, fp32:  C[m, :] = epilogue( sum_{t < taps} A[idx[m, t], 0:K] . W[t*K : (t+1)*K, 0:N] )
//   idx == nullptr: taps = 1, row m itself; idx < 0: the tap is absent. W is [taps*K, N] row-major (N contiguous).
//   epilogue: (+ bias) -> (* scale + shift: folded BatchNorm) -> (GELU) -> (+ residual); fp32 or bf16 store.
// Generic kernel (any shape): 64 x 64 tile, 16-deep K slices over the flattened (tap, channel) axis, 256 threads,
// 4 x 4 outputs per thread. The network's own shapes take gather_gemm_tiled_kernel below.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GM = 64, GN = 64, GK = 16;

__global__ void __launch_bounds__(256) gather_gemm_kernel(
    const float* __restrict__ A, int lda, const int* __restrict__ idx, int taps, const float* __restrict__ W,
    const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift, int act,
    const float* __restrict__ residual, int ldr, void* __restrict__ Cout, int ldc, int out_bf16, int M, int N, int K) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int tx = tid % 16, ty = tid / 16;  // thread computes rows ty*4..+3, cols tx*4..+3
  const int KK = taps * K;
  float acc[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < KK; k0 += GK) {
    // A slice: 64 rows x 16 flattened-k; thread loads 4 elements: row = e / 16, kk = e % 16
    for (int e = tid; e < GM * GK; e += 256) {
      int r = e / GK, kk = e % GK;
      int m = m0 + r, kf = k0 + kk;
      float v = 0.f;
      if (m < M && kf < KK) {
        int t = kf / K, k = kf - t * K;
        int row = idx ? idx[(size_t)m * taps + t] : m;
        if (row >= 0) v = A[(size_t)row * lda + k];
      }
      As[kk][r] = v;
    }
    for (int e = tid; e < GK * GN; e += 256) {
      int kk = e / GN, c = e % GN;
      int kf = k0 + kk, nn = n0 + c;
      Bs[kk][c] = (kf < KK && nn < N) ? W[(size_t)kf * N + nn] : 0.f;
    }
    __syncthreads();
    for (int kk = 0; kk < GK; ++kk) {
      float a[4], b[4];
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    for (int j = 0; j < 4; ++j) {
      int nn = n0 + tx * 4 + j;
      if (nn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[nn];
      if (scale) v = v * scale[nn] + shift[nn];
      if (act == 2) v = gelu_erf(v);
      if (residual) v += residual[(size_t)m * ldr + nn];
      if (out_bf16) reinterpret_cast<uint16_t*>(Cout)[(size_t)m * ldc + nn] = bf16_bits(v);
      else reinterpret_cast<float*>(Cout)[(size_t)m * ldc + nn] = v;
    }
  }
}

// Fast path of the same contract for the shapes the network produces (K, lda, N multiples of 4, 16-byte aligned A / W):
// TM x BN tile (TM = 128, or 64 when 128-row tiles would leave SMs idle; BN = 32 / 64 / 128 by output width), one
// (tap, 16-channel) slice per step so the gathered row index is read once per slice and never divided, 128-bit global
// and shared loads, (TM / 16) x (BN / 16) outputs per thread laid out so that the shared-memory reads of a quarter-warp
// are contiguous (no bank conflicts). Deep stages have few rows and 27 taps: the taps are then split over blockIdx.z,
// every split writes its raw partial sums to a workspace, and splitk_epilogue_kernel adds them in split order
// (deterministic) and applies the epilogue.
constexpr int TK = 16;

__device__ __forceinline__ void gemm_epilogue_store(float v, int m, int nn, const float* bias, const float* scale,
                                                    const float* shift, int act, const float* residual, int ldr,
                                                    void* Cout, int ldc, int out_bf16) {
  if (bias) v += bias[nn];
  if (scale) v = v * scale[nn] + shift[nn];
  if (act == 2) v = gelu_erf(v);
  if (residual) v += residual[(size_t)m * ldr + nn];
  if (out_bf16) reinterpret_cast<uint16_t*>(Cout)[(size_t)m * ldc + nn] = bf16_bits(v);
  else reinterpret_cast<float*>(Cout)[(size_t)m * ldc + nn] = v;
}

template <int TMv, int BN>
__global__ void __launch_bounds__(256) gather_gemm_tiled_kernel(
    const float* __restrict__ A, int lda, const int* __restrict__ idx, int taps, const float* __restrict__ W,
    const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift, int act,
    const float* __restrict__ residual, int ldr, void* __restrict__ Cout, int ldc, int out_bf16, int M, int N, int K,
    int taps_per_split, float* __restrict__ part) {
  constexpr int NH = BN == 128 ? 2 : 1;  // column halves per thread
  constexpr int CW = BN / 16 / NH;       // contiguous columns per half: 4, 4, 2
  constexpr int RH = TMv / 64;           // row halves per thread: rows h*64 + ty*4 .. +3
  constexpr int LPR = 256 / TMv;         // loader threads per tile row: 2 or 4
  constexpr int KPT = TK / LPR;          // channels of a slice per loader thread: 8 or 4
  __shared__ __align__(16) float As[TK][TMv + 4];
  __shared__ __align__(16) float Bs[TK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TMv, n0 = blockIdx.x * BN;
  const int tx = tid % 16, ty = tid / 16;
  const int lr = tid / LPR, lk = (tid % LPR) * KPT;
  const int lm = m0 + lr;
  float acc[RH * 4][NH * CW];
#pragma unroll
  for (int i = 0; i < RH * 4; ++i) {
#pragma unroll
    for (int j = 0; j < NH * CW; ++j) acc[i][j] = 0.f;
  }
  const int t_begin = blockIdx.z * taps_per_split;
  const int t_end = t_begin + taps_per_split < taps ? t_begin + taps_per_split : taps;

  for (int t = t_begin; t < t_end; ++t) {
    int row = -1;
    if (lm < M) row = idx ? idx[(size_t)lm * taps + t] : lm;
    const float* arow = row >= 0 ? A + (size_t)row * lda : nullptr;
    const float* wt = W + (size_t)t * K * N;
    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
      for (int q = 0; q < KPT / 4; ++q) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        int ka = k0 + lk + 4 * q;
        if (arow && ka < K) a = *reinterpret_cast<const float4*>(arow + ka);
        As[lk + 4 * q + 0][lr] = a.x;
        As[lk + 4 * q + 1][lr] = a.y;
        As[lk + 4 * q + 2][lr] = a.z;
        As[lk + 4 * q + 3][lr] = a.w;
      }
      for (int v = tid; v < TK * BN / 4; v += 256) {
        int kk = v / (BN / 4), c4 = v % (BN / 4);
        int nn = n0 + c4 * 4;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + kk < K && nn < N) b = *reinterpret_cast<const float4*>(wt + (size_t)(k0 + kk) * N + nn);
        *reinterpret_cast<float4*>(&Bs[kk][c4 * 4]) = b;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float a[RH * 4], b[NH * CW];
#pragma unroll
        for (int h = 0; h < RH; ++h) {
          float4 x = *reinterpret_cast<const float4*>(&As[kk][h * 64 + ty * 4]);
          a[h * 4 + 0] = x.x; a[h * 4 + 1] = x.y; a[h * 4 + 2] = x.z; a[h * 4 + 3] = x.w;
        }
        if constexpr (CW == 4) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            float4 y = *reinterpret_cast<const float4*>(&Bs[kk][h * 64 + tx * 4]);
            b[h * 4 + 0] = y.x; b[h * 4 + 1] = y.y; b[h * 4 + 2] = y.z; b[h * 4 + 3] = y.w;
          }
        } else {
          float2 y = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
          b[0] = y.x; b[1] = y.y;
        }
#pragma unroll
        for (int i = 0; i < RH * 4; ++i) {
#pragma unroll
          for (int j = 0; j < NH * CW; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < RH * 4; ++i) {
    int m = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
    if (m >= M) continue;
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        int nn = n0 + h * 64 + tx * CW + j;
        if (nn >= N) continue;
        float v = acc[i][h * CW + j];
        if (part) part[((size_t)blockIdx.z * M + m) * N + nn] = v;
        else gemm_epilogue_store(v, m, nn, bias, scale, shift, act, residual, ldr, Cout, ldc, out_bf16);
      }
  }
}

// C[m, n] = epilogue(part[0][m][n] + part[1][m][n] + ...), splits added in order
__global__ void splitk_epilogue_kernel(const float* __restrict__ part, int splits, const float* __restrict__ bias,
                                       const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                       const float* __restrict__ residual, int ldr, void* __restrict__ Cout, int ldc,
                                       int out_bf16, int M, int N) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)M * N) return;
  int m = (int)(t / N), nn = (int)(t % N);
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += part[(size_t)s * M * N + t];
  gemm_epilogue_store(v, m, nn, bias, scale, shift, act, residual, ldr, Cout, ldc, out_bf16);
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm, fp32, one warp per row: out = (residual) + (x - mean) / sqrt(var + eps) * gamma + beta  (biased variance)
// ---------------------------------------------------------------------------------------------------------------
__global__ void layernorm_f32_kernel(const float* __restrict__ x, int ld, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, const float* __restrict__ residual,
                                     int ldr, float* __restrict__ out, int ldo, int M, int C) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  bool active = warp < M;
  int row = active ? warp : M - 1;  // keep every lane in the shuffles
  const float* xr = x + (size_t)row * ld;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    float d = xr[c] - mean;
    v += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  float rstd = 1.0f / sqrtf(v / (float)C + eps);
  if (!active) return;
  for (int c = lane; c < C; c += 32) {
    float y = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    if (residual) y += residual[(size_t)row * ldr + c];
    out[(size_t)row * ldo + c] = y;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// serialized patch attention (SerializedAttention.forward, pointtransformerv3.py:443-494, flash-attn varlen semantics):
// patch p = {q_begin, q_len, k_begin, k_len} in positions of the serialized order; queries and keys are the rows
// order[pos]; q, k, v are rounded to fp16, scores / softmax / PV accumulate in fp32, the output is rounded to fp16.
// grid (q tiles of 256, heads, patches), 128 threads, two queries per thread (each K / V row read from shared memory
// feeds 2 x 16 FMAs), 128-key tiles of K and V in shared memory, online softmax over chunks of 8 keys.
// head_dim is fixed at 16 (every PTv3 stage: channels / heads = 16).
// ---------------------------------------------------------------------------------------------------------------
constexpr int AD = 16, AT = 128, AQ = 2;  // head_dim, threads (= keys per shared tile), queries per thread

__global__ void __launch_bounds__(AT) patch_attention_kernel(const float* __restrict__ qkv, int ld,
                                                            const int* __restrict__ order,
                                                            const int* __restrict__ patches, int C, float scale,
                                                            float* __restrict__ out, int ldo) {
  __shared__ __align__(16) float Ks[AT][AD];
  __shared__ __align__(16) float Vs[AT][AD];
  const int* pd = patches + (size_t)blockIdx.z * 4;
  const int q_begin = pd[0], q_len = pd[1], k_begin = pd[2], k_len = pd[3];
  const int h = blockIdx.y;
  const int q0 = blockIdx.x * (AT * AQ);
  if (q0 >= q_len) return;  // whole block out of range (uniform)
  bool active[AQ];
  int qrow[AQ];
  float q[AQ][AD], o[AQ][AD], m[AQ], l[AQ];
#pragma unroll
  for (int u = 0; u < AQ; ++u) {
    int qi = q0 + u * AT + threadIdx.x;
    active[u] = qi < q_len;
    qrow[u] = active[u] ? order[q_begin + qi] : 0;
    m[u] = -INFINITY;
    l[u] = 0.f;
#pragma unroll
    for (int d = 0; d < AD; ++d) {
      q[u][d] = active[u] ? round_f16(qkv[(size_t)qrow[u] * ld + h * AD + d]) : 0.f;
      o[u][d] = 0.f;
    }
  }
  for (int k0 = 0; k0 < k_len; k0 += AT) {
    int kj = k0 + threadIdx.x;
    if (kj < k_len) {
      int krow = order[k_begin + kj];
      const float* kp = qkv + (size_t)krow * ld + C + h * AD;
      const float* vp = qkv + (size_t)krow * ld + 2 * C + h * AD;
#pragma unroll
      for (int d = 0; d < AD; ++d) {
        Ks[threadIdx.x][d] = round_f16(kp[d]);
        Vs[threadIdx.x][d] = round_f16(vp[d]);
      }
    }
    __syncthreads();
    int nk = k_len - k0 < AT ? k_len - k0 : AT;
    for (int c0 = 0; c0 < nk; c0 += 8) {
      float s[AQ][8];
      float cm[AQ];
#pragma unroll
      for (int u = 0; u < AQ; ++u) cm[u] = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bool ok = c0 + j < nk;
        float kd[AD];
#pragma unroll
        for (int d = 0; d < AD; ++d) kd[d] = Ks[ok ? c0 + j : c0][d];
#pragma unroll
        for (int u = 0; u < AQ; ++u) {
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < AD; ++d) a = fmaf(q[u][d], kd[d], a);
          a = ok ? a * scale : -INFINITY;
          s[u][j] = a;
          cm[u] = fmaxf(cm[u], a);
        }
      }
#pragma unroll
      for (int u = 0; u < AQ; ++u) {
        float mn = fmaxf(m[u], cm[u]);
        float alpha = expf(m[u] - mn);  // m = -inf on the first chunk: exp(-inf) = 0
        l[u] *= alpha;
#pragma unroll
        for (int d = 0; d < AD; ++d) o[u][d] *= alpha;
        m[u] = mn;
#pragma unroll
        for (int j = 0; j < 8; ++j) s[u][j] = expf(s[u][j] - mn);  // exp(-inf) = 0 for the keys past the end
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float vd[AD];
        int kr = c0 + j < nk ? c0 + j : c0;
#pragma unroll
        for (int d = 0; d < AD; ++d) vd[d] = Vs[kr][d];
#pragma unroll
        for (int u = 0; u < AQ; ++u) {
          l[u] += s[u][j];
#pragma unroll
          for (int d = 0; d < AD; ++d) o[u][d] = fmaf(s[u][j], vd[d], o[u][d]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < AQ; ++u) {
    if (!active[u]) continue;
    float inv = 1.0f / l[u];
#pragma unroll
    for (int d = 0; d < AD; ++d) out[(size_t)qrow[u] * ldo + h * AD + d] = round_f16(o[u][d] * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// segment max over contiguous clusters + folded BatchNorm + GELU (SerializedPooling, pointtransformerv3.py:686-713)
// ---------------------------------------------------------------------------------------------------------------
__global__ void segment_max_kernel(const float* __restrict__ x, int ld, const int* __restrict__ seg_start, int n_seg,
                                   int C, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                   float* __restrict__ out, int ldo) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_seg * C) return;
  int s = (int)(t / C), c = (int)(t % C);
  float m = -INFINITY;
  for (int i = seg_start[s]; i < seg_start[s + 1]; ++i) m = fmaxf(m, x[(size_t)i * ld + c]);
  if (scale) m = m * scale[c] + shift[c];
  if (act == 2) m = gelu_erf(m);
  out[(size_t)s * ldo + c] = m;
}

// per-cloud mean of the last level's features into row row_map[b] of out (builder.py:139-144); sequential sum per
// (cloud, channel): the last level holds a handful of points per cloud
__global__ void cloud_mean_kernel(const float* __restrict__ x, int ld, const int* __restrict__ cloud_off, int n_clouds,
                                  int C, const int* __restrict__ row_map, float* __restrict__ out, int ldo) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_clouds * C) return;
  int b = t / C, c = t % C;
  int i0 = cloud_off[b], i1 = cloud_off[b + 1];
  float s = 0.f;
  for (int i = i0; i < i1; ++i) s += x[(size_t)i * ld + c];
  out[(size_t)row_map[b] * ldo + c] = s / (float)(i1 - i0);
}

}  // namespace pc
}  // namespace b200

using namespace b200;
using namespace b200::pc;

#define PC_CHECK_LAUNCH(name)                                                                  \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) return fail(-5, "%s launch failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

static inline unsigned blocks_for(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }

extern "C" {

int b200_pc_grid_coords(const float* pts, int ld, int n, float grid_size, float* min3, int32_t* grid, int32_t* max_out,
                        b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!pts || !min3 || !grid || !max_out || n <= 0 || ld < 3 || !(grid_size > 0.f))
    return fail(-2, "b200_pc_grid_coords: bad argument (n=%d ld=%d grid_size=%g)", n, ld, (double)grid_size);
  LaunchScope ls(kFamPointCloud, stream, (double)n * 24.0, 0.0, 3);
  B200_LAUNCH(coord_min_kernel, dim3(1), dim3(1024), 0, stream, pts, ld, n, min3);
  B200_LAUNCH(grid_coord_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, stream, pts, ld, n, min3, grid_size, grid);
  B200_LAUNCH(grid_max_kernel, dim3(1), dim3(1024), 0, stream, grid, 3 * n, max_out);
  PC_CHECK_LAUNCH("b200_pc_grid_coords");
  return 0;
}

int b200_pc_encode(const int32_t* grid, const int32_t* batch, int n, int depth, int order, int64_t* code,
                   b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!grid || !batch || !code || n <= 0 || depth < 0 || depth > 16 || order < 0 || order > 3)
    return fail(-2, "b200_pc_encode: bad argument (n=%d depth=%d order=%d)", n, depth, order);
  LaunchScope ls(kFamPointCloud, stream, (double)n * 24.0);
  B200_LAUNCH(encode_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, stream, grid, batch, n, depth, order,
              reinterpret_cast<long long*>(code));
  PC_CHECK_LAUNCH("b200_pc_encode");
  return 0;
}

size_t b200_pc_argsort_workspace_bytes(int n) {
  if (n <= 0) return 0;
  size_t iota = ((size_t)n * sizeof(int) + 255) / 256 * 256;
#ifdef B200_EMU
  return iota;
#else
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, n);
  return iota + (tmp + 255) / 256 * 256;
#endif
}

int b200_pc_argsort(const int64_t* code, int n, int bits, int64_t* code_sorted, int32_t* order, void* workspace,
                    size_t workspace_bytes, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!code || !code_sorted || !order || !workspace || n <= 0 || bits <= 0 || bits > 64)
    return fail(-2, "b200_pc_argsort: bad argument (n=%d bits=%d)", n, bits);
  if (workspace_bytes < b200_pc_argsort_workspace_bytes(n)) return fail(-2, "b200_pc_argsort: workspace too small");
  int* iota = reinterpret_cast<int*>(workspace);
  LaunchScope ls(kFamPointCloud, stream, (double)n * 40.0, 0.0, 2);
  B200_LAUNCH(iota_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, stream, iota, n);
#ifdef B200_EMU
  std::vector<int> idx(iota, iota + n);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (uint64_t)code[a] < (uint64_t)code[b]; });
  for (int i = 0; i < n; ++i) {
    order[i] = idx[i];
    code_sorted[i] = code[idx[i]];
  }
#else
  size_t off = ((size_t)n * sizeof(int) + 255) / 256 * 256;
  size_t tmp = workspace_bytes - off;
  B200_CUDA_OK(cub::DeviceRadixSort::SortPairs(reinterpret_cast<char*>(workspace) + off, tmp,
                                               reinterpret_cast<const unsigned long long*>(code),
                                               reinterpret_cast<unsigned long long*>(code_sorted), iota, order, n, 0,
                                               bits, stream));
#endif
  PC_CHECK_LAUNCH("b200_pc_argsort");
  return 0;
}

int b200_pc_gather_rows(const void* src, int ld_src, const int32_t* idx, int n, int width, void* dst, int ld_dst,
                        b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!src || !idx || !dst || n <= 0 || width <= 0 || ld_src < width || ld_dst < width)
    return fail(-2, "b200_pc_gather_rows: bad argument (n=%d width=%d)", n, width);
  LaunchScope ls(kFamPointCloud, stream, (double)n * width * 8.0);
  B200_LAUNCH(gather_rows_kernel, dim3(blocks_for((long long)n * width, 256)), dim3(256), 0, stream,
              reinterpret_cast<const uint32_t*>(src), ld_src, idx, n, width, reinterpret_cast<uint32_t*>(dst), ld_dst);
  PC_CHECK_LAUNCH("b200_pc_gather_rows");
  return 0;
}

int b200_pc_neighbors(const int64_t* zcode, const int32_t* grid, const int32_t* batch, int n, int depth, int ksize,
                      int32_t* nbr, int32_t* dup_flag, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!zcode || !grid || !batch || !nbr || !dup_flag || n <= 0 || depth < 0 || depth > 16 || (ksize != 3 && ksize != 5))
    return fail(-2, "b200_pc_neighbors: bad argument (n=%d depth=%d ksize=%d)", n, depth, ksize);
  long long work = (long long)n * ksize * ksize * ksize;
  LaunchScope ls(kFamPointCloud, stream, (double)work * 4.0);
  B200_LAUNCH(neighbor_kernel, dim3(blocks_for(work, 256)), dim3(256), 0, stream,
              reinterpret_cast<const long long*>(zcode), grid, batch, n, depth, ksize, nbr, dup_flag);
  PC_CHECK_LAUNCH("b200_pc_neighbors");
  return 0;
}

int b200_pc_pool_plan(const int64_t* zcode, const int32_t* grid, const int32_t* batch, int n, int pooling_depth,
                      int32_t* seg_start, int32_t* n_out, int32_t* grid_out, int32_t* batch_out,
                      b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!zcode || !grid || !batch || !seg_start || !n_out || !grid_out || !batch_out || n <= 0 || pooling_depth < 0 ||
      pooling_depth > 4)
    return fail(-2, "b200_pc_pool_plan: bad argument (n=%d pooling_depth=%d)", n, pooling_depth);
  LaunchScope ls(kFamPointCloud, stream, (double)n * 32.0, 0.0, 2);
  B200_LAUNCH(pool_plan_kernel, dim3(1), dim3(1024), 0, stream, reinterpret_cast<const long long*>(zcode), n,
              3 * pooling_depth, seg_start, n_out);
  // the cluster count is only known on the device: size the launch for the upper bound n and let threads read it
  B200_LAUNCH(pool_coords_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, stream, seg_start, n_out, grid, batch,
              pooling_depth, grid_out, batch_out);
  PC_CHECK_LAUNCH("b200_pc_pool_plan");
  return 0;
}

int b200_pc_cloud_offsets(const int32_t* batch, int n, int n_clouds, int32_t* off, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!batch || !off || n <= 0 || n_clouds <= 0) return fail(-2, "b200_pc_cloud_offsets: bad argument");
  LaunchScope ls(kFamPointCloud, stream, (double)n_clouds * 64.0);
  B200_LAUNCH(cloud_offsets_kernel, dim3(blocks_for(n_clouds + 1, 128)), dim3(128), 0, stream, batch, n, n_clouds, off);
  PC_CHECK_LAUNCH("b200_pc_cloud_offsets");
  return 0;
}

// tile / split selection shared by the launcher and the workspace query
struct GemmPlan {
  bool fast;
  int bn, tm, splits, taps_per_split;
};
static GemmPlan plan_gemm(int M, int N, int K, int taps, int lda, const void* A, const void* W) {
  GemmPlan p{};
  p.fast = K % 4 == 0 && lda % 4 == 0 && N % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  p.bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  p.tm = 128;
  p.splits = 1;
  p.taps_per_split = taps;
  if (!p.fast) return p;
  const long long sms = num_sms();
  long long cols = (N + p.bn - 1) / p.bn;
  if (cols * ((M + 127) / 128) < 2 * sms) p.tm = 64;  // 128-row tiles would leave SMs idle
  long long ctas = cols * ((M + p.tm - 1) / p.tm);
  if (taps > 1 && ctas < 2 * sms) {  // still starved: split the taps (deterministic two-pass reduction)
    long long want = (2 * sms + ctas - 1) / ctas;
    int s = (int)(want < taps ? want : taps);
    if (s > 9) s = 9;
    p.taps_per_split = (taps + s - 1) / s;
    p.splits = (taps + p.taps_per_split - 1) / p.taps_per_split;
  }
  return p;
}

size_t b200_pc_gemm_workspace_bytes(int M, int N, int K, int taps) {
  if (M <= 0 || N <= 0 || K <= 0 || taps <= 0) return 0;
  GemmPlan p = plan_gemm(M, N, K, taps, K % 4 == 0 ? 4 : 1, nullptr, nullptr);  // aligned operands assumed
  return p.splits > 1 ? (size_t)p.splits * M * N * sizeof(float) : 0;
}

int b200_pc_gemm_f32(const float* A, int lda, const int32_t* idx, int taps, const float* W, const float* bias,
                     const float* scale, const float* shift, int act, const float* residual, int ldr, void* C, int ldc,
                     int out_bf16, int M, int N, int K, void* workspace, size_t workspace_bytes,
                     b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || taps <= 0 || lda < K || ldc < N || (idx == nullptr && taps != 1) ||
      (act != 0 && act != 2) || ((scale == nullptr) != (shift == nullptr)) || (residual && ldr < N))
    return fail(-2, "b200_pc_gemm_f32: bad argument (M=%d N=%d K=%d taps=%d act=%d)", M, N, K, taps, act);
  if ((M + 63) / 64 > 65535) return fail(-2, "b200_pc_gemm_f32: M=%d exceeds the grid (max %d rows)", M, 65535 * 64);
  double flops = 2.0 * M * N * (double)K * taps;
  GemmPlan p = plan_gemm(M, N, K, taps, lda, A, W);
  if (p.splits > 1 && (!workspace || workspace_bytes < (size_t)p.splits * M * N * sizeof(float))) {
    p.splits = 1;  // no room for the partial sums: one pass over all taps
    p.taps_per_split = taps;
  }
  LaunchScope ls(kFamPointCloud, stream, 4.0 * ((double)M * K * taps + (double)K * taps * N + (double)M * N), flops,
                 p.splits > 1 ? 2 : 1);
  if (p.fast) {
    dim3 grid((N + p.bn - 1) / p.bn, (M + p.tm - 1) / p.tm, p.splits);
    float* part = p.splits > 1 ? reinterpret_cast<float*>(workspace) : nullptr;
#define PC_GEMM_LAUNCH(TMV, BNV)                                                                                   \
  B200_LAUNCH((gather_gemm_tiled_kernel<TMV, BNV>), grid, dim3(256), 0, stream, A, lda, idx, taps, W, bias, scale, \
              shift, act, residual, ldr, C, ldc, out_bf16, M, N, K, p.taps_per_split, part)
    if (p.tm == 128) {
      if (p.bn == 32) PC_GEMM_LAUNCH(128, 32);
      else if (p.bn == 64) PC_GEMM_LAUNCH(128, 64);
      else PC_GEMM_LAUNCH(128, 128);
    } else {
      if (p.bn == 32) PC_GEMM_LAUNCH(64, 32);
      else if (p.bn == 64) PC_GEMM_LAUNCH(64, 64);
      else PC_GEMM_LAUNCH(64, 128);
    }
#undef PC_GEMM_LAUNCH
    if (p.splits > 1) {
      B200_LAUNCH(splitk_epilogue_kernel, dim3(blocks_for((long long)M * N, 256)), dim3(256), 0, stream, part, p.splits,
                  bias, scale, shift, act, residual, ldr, C, ldc, out_bf16, M, N);
    }
  } else {  // odd shapes (the stem: K = 6) take the generic kernel
    dim3 grid((N + GN - 1) / GN, (M + GM - 1) / GM);
    B200_LAUNCH(gather_gemm_kernel, grid, dim3(256), 0, stream, A, lda, idx, taps, W, bias, scale, shift, act,
                residual, ldr, C, ldc, out_bf16, M, N, K);
  }
  PC_CHECK_LAUNCH("b200_pc_gemm_f32");
  return 0;
}

int b200_pc_layernorm_f32(const float* x, int ld, const float* gamma, const float* beta, float eps,
                          const float* residual, int ldr, float* out, int ldo, int M, int C, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !gamma || !beta || !out || M <= 0 || C <= 0 || ld < C || ldo < C || (residual && ldr < C))
    return fail(-2, "b200_pc_layernorm_f32: bad argument (M=%d C=%d)", M, C);
  LaunchScope ls(kFamPointCloud, stream, 8.0 * M * C);
  B200_LAUNCH(layernorm_f32_kernel, dim3(blocks_for((long long)M * 32, 256)), dim3(256), 0, stream, x, ld, gamma, beta,
              eps, residual, ldr, out, ldo, M, C);
  PC_CHECK_LAUNCH("b200_pc_layernorm_f32");
  return 0;
}

int b200_pc_patch_attention(const float* qkv, int ld, const int32_t* order, const int32_t* patches, int n_patches,
                            int max_q_len, int channels, int heads, float scale, float* out, int ldo,
                            b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!qkv || !order || !patches || !out || n_patches <= 0 || max_q_len <= 0 || heads <= 0 || channels != heads * AD ||
      ld < 3 * channels || ldo < channels)
    return fail(-2, "b200_pc_patch_attention: bad argument (patches=%d channels=%d heads=%d; head_dim must be %d)",
                n_patches, channels, heads, AD);
  LaunchScope ls(kFamPointCloud, stream, 0.0, 4.0 * n_patches * (double)max_q_len * max_q_len * channels);
  dim3 grid((max_q_len + AT * AQ - 1) / (AT * AQ), heads, n_patches);
  B200_LAUNCH(patch_attention_kernel, grid, dim3(AT), 0, stream, qkv, ld, order, patches, channels, scale, out, ldo);
  PC_CHECK_LAUNCH("b200_pc_patch_attention");
  return 0;
}

int b200_pc_segment_max(const float* x, int ld, const int32_t* seg_start, int n_seg, int C, const float* scale,
                        const float* shift, int act, float* out, int ldo, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !seg_start || !out || n_seg <= 0 || C <= 0 || ld < C || ldo < C || ((scale == nullptr) != (shift == nullptr)) ||
      (act != 0 && act != 2))
    return fail(-2, "b200_pc_segment_max: bad argument (n_seg=%d C=%d)", n_seg, C);
  LaunchScope ls(kFamPointCloud, stream, 8.0 * n_seg * C);
  B200_LAUNCH(segment_max_kernel, dim3(blocks_for((long long)n_seg * C, 256)), dim3(256), 0, stream, x, ld, seg_start,
              n_seg, C, scale, shift, act, out, ldo);
  PC_CHECK_LAUNCH("b200_pc_segment_max");
  return 0;
}

int b200_pc_cloud_mean(const float* x, int ld, const int32_t* cloud_off, int n_clouds, int C, const int32_t* row_map,
                       float* out, int ldo, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !cloud_off || !row_map || !out || n_clouds <= 0 || C <= 0 || ld < C || ldo < C)
    return fail(-2, "b200_pc_cloud_mean: bad argument (n_clouds=%d C=%d)", n_clouds, C);
  LaunchScope ls(kFamPointCloud, stream, 4.0 * n_clouds * C);
  B200_LAUNCH(cloud_mean_kernel, dim3(blocks_for((long long)n_clouds * C, 128)), dim3(128), 0, stream, x, ld, cloud_off,
              n_clouds, C, row_map, out, ldo);
  PC_CHECK_LAUNCH("b200_pc_cloud_mean");
  return 0;
}

}  // extern "C"
