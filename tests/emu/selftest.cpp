// selftest.cpp -- TEST INFRASTRUCTURE ONLY: kernels that exercise the emulator itself (cuda_emu.h): block barriers with
// shared memory, full-warp shuffles, multi-dimensional grids / blocks, early exits before a barrier-free tail. Built
// into the emulator library by tests/emu_lib.py and driven by tests/test_emu_selftest.py.
#include "cuda_emu.h"

namespace {

// block-wide inclusive scan through shared memory: needs a correct __syncthreads on every step
__global__ void scan_kernel(const int* in, int* out, int n) {
  __shared__ int buf[2][256];
  int t = threadIdx.x;
  int i = blockIdx.x * blockDim.x + t;
  buf[0][t] = i < n ? in[i] : 0;
  __syncthreads();
  int cur = 0;
  for (int d = 1; d < (int)blockDim.x; d <<= 1) {
    buf[1 - cur][t] = buf[cur][t] + (t >= d ? buf[cur][t - d] : 0);
    cur = 1 - cur;
    __syncthreads();
  }
  if (i < n) out[i] = buf[cur][t];
}

// per-warp butterfly sum + broadcast of lane 3 + shift-down
__global__ void shuffle_kernel(const float* in, float* sum, float* lane3, float* down) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = in[i];
  float s = v;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  sum[i] = s;
  lane3[i] = __shfl_sync(0xffffffffu, v, 3);
  down[i] = __shfl_down_sync(0xffffffffu, v, 5);
}

// 3-D grid and 2-D block indexing; threads of odd rows leave early
__global__ void index_kernel(int* out, int nx, int ny) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  int z = blockIdx.z;
  if (x >= nx || y >= ny) return;
  if (y & 1) return;
  out[(z * ny + y) * nx + x] = 1000000 * z + 1000 * y + x + gridDim.x * 0 + blockDim.y * 0;
}

}  // namespace

extern "C" int b200_emu_selftest_scan(const int* in, int* out, int n) {
  emu::launch(dim3((n + 255) / 256), dim3(256), 0, [&]() { scan_kernel(in, out, n); });
  return 0;
}
extern "C" int b200_emu_selftest_shuffle(const float* in, float* sum, float* lane3, float* down, int n) {
  if (n % 64) return -2;
  emu::launch(dim3(n / 64), dim3(64), 0, [&]() { shuffle_kernel(in, sum, lane3, down); });
  return 0;
}
extern "C" int b200_emu_selftest_index(int* out, int nx, int ny, int nz) {
  emu::launch(dim3((nx + 7) / 8, (ny + 3) / 4, nz), dim3(8, 4), 0, [&]() { index_kernel(out, nx, ny); });
  return 0;
}
