// emu_support.cpp -- TEST INFRASTRUCTURE ONLY: error plumbing for the host build of ptv3.cu (see cuda_emu.h).
#include <cstdarg>
#include <cstdio>
#include <string>

namespace b200 {
static thread_local std::string g_last_error;
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
}  // namespace b200

extern "C" const char* b200_last_error(void) { return b200::g_last_error.c_str(); }
extern "C" int b200_emu_marker(void) { return 1; }
