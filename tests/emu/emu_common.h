// emu_common.h -- TEST INFRASTRUCTURE ONLY: the few declarations of mm_or_b200/csrc/common.h and include/b200_mmor.h
// that ptv3.cu / train_extras.cu use, for the host (g++ -DB200_EMU) build of the kernels. See cuda_emu.h.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/b200_mmor.h"

namespace b200 {
int fail(int code, const char* fmt, ...);
inline int num_sms() { return 148; }  // B200
enum { kFamTrain = 11, kFamPointCloud = 12 };
struct LaunchScope {
  LaunchScope(int, cudaStream_t, double = 0.0, double = 0.0, int = 1) {}
};
}  // namespace b200

#define B200_CUDA_OK(expr) \
  do {                     \
    (void)(expr);          \
  } while (0)
