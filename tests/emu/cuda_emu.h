// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: runs plain (SIMT, shared-memory, warp-shuffle) CUDA kernels on the host.
//
// The container that builds this repo has no GPU. Kernels that use nothing beyond threadIdx/blockIdx, __shared__,
// __syncthreads and full-warp shuffles (mm_or_b200/csrc/ptv3.cu) are compiled a second time with g++ and -DB200_EMU
// against this header into tests/emu/_build/libb200emu.so, and the `-m "not gpu"` tests run the SAME kernel source on
// the CPU against the oracle. Every thread of a block is a ucontext fiber on one OS thread; blocks run one after the
// other; __syncthreads / warp shuffles are generation barriers that yield to the next fiber. The product never loads
// this library (mm_or_b200/_lib.py only opens libb200mmor.so and rejects host pointers); tcgen05 / TMA / cluster
// kernels cannot be emulated and are validated on the GPU only.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu {
  unsigned x, y, z;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
struct alignas(8) float2 {
  float x, y;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct alignas(8) uint2 {
  unsigned x, y;
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
#define __align__(n) __attribute__((aligned(n)))
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

namespace emu {

struct Fiber {
  ucontext_t ctx;
  bool done;
};
struct Bar {
  int count = 0;
  unsigned gen = 0;
};
struct State {
  uint3_emu tid, bid;
  dim3 bdim, gdim;
  int cur = 0, nthreads = 0;
  ucontext_t sched;
  std::vector<Fiber> fibers;
  std::vector<char> stacks;
  Bar block_bar;
  std::vector<Bar> warp_bar;
  std::vector<uint64_t> warp_slot;  // [warp][32]
  std::vector<char> dyn_smem;
  const std::function<void()>* body = nullptr;
};
inline State& st() {
  static State s;
  return s;
}
static const size_t kStack = 96 * 1024;

inline void set_tid(int t) {
  State& s = st();
  s.cur = t;
  s.tid.x = t % s.bdim.x;
  s.tid.y = (t / s.bdim.x) % s.bdim.y;
  s.tid.z = t / (s.bdim.x * s.bdim.y);
}
inline void yield() {
  State& s = st();
  int me = s.cur;
  swapcontext(&s.fibers[me].ctx, &s.sched);
}
inline void barrier(Bar& b, int n) {
  unsigned gen = b.gen;
  if (++b.count == n) {
    b.count = 0;
    b.gen++;
  } else {
    while (b.gen == gen) yield();
  }
}
inline void sync_block() { barrier(st().block_bar, st().nthreads); }
inline void sync_warp() {
  State& s = st();
  int w = s.cur / 32;
  int lanes = std::min(32, s.nthreads - w * 32);
  barrier(s.warp_bar[w], lanes);
}
inline void trampoline() {
  State& s = st();
  (*s.body)();
  s.fibers[s.cur].done = true;
  swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  State& s = st();
  s.bdim = block;
  s.gdim = grid;
  s.nthreads = block.x * block.y * block.z;
  s.body = &body;
  s.dyn_smem.assign(smem_bytes + 16, 0);
  int nw = (s.nthreads + 31) / 32;
  if (s.stacks.size() < kStack * (size_t)s.nthreads) s.stacks.resize(kStack * (size_t)s.nthreads);
  s.fibers.resize(s.nthreads);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bid = {bx, by, bz};
        s.block_bar = Bar();
        s.warp_bar.assign(nw, Bar());
        s.warp_slot.assign((size_t)nw * 32, 0);
        for (int t = 0; t < s.nthreads; ++t) {
          Fiber& f = s.fibers[t];
          f.done = false;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = s.stacks.data() + kStack * (size_t)t;
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = &s.sched;
          makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        int live = s.nthreads;
        while (live > 0) {
          for (int t = 0; t < s.nthreads; ++t) {
            if (s.fibers[t].done) continue;
            set_tid(t);
            swapcontext(&s.sched, &s.fibers[t].ctx);
            if (s.fibers[t].done) --live;
          }
        }
      }
  s.body = nullptr;
}

template <typename T>
inline T shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of <= 8-byte values only");
  State& s = st();
  int w = s.cur / 32, lane = s.cur % 32;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  s.warp_slot[(size_t)w * 32 + lane] = raw;
  sync_warp();
  uint64_t got = s.warp_slot[(size_t)w * 32 + (src_lane & 31)];
  sync_warp();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}

}  // namespace emu

#define threadIdx (emu::st().tid)
#define blockIdx (emu::st().bid)
#define blockDim (emu::st().bdim)
#define gridDim (emu::st().gdim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __syncthreads() emu::sync_block()
#define __syncwarp(...) emu::sync_warp()

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return emu::shfl(v, (emu::st().cur % 32) ^ lane_mask);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int delta) {
  int lane = emu::st().cur % 32;
  return emu::shfl(v, lane + delta < 32 ? lane + delta : lane);
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
  return emu::shfl(v, src);
}

inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __expf(float x) { return expf(x); }
inline int __float2int_rz(float x) { return (int)x; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
