// Error state and version of libsobolev_b200 (the C ABI declared in include/sobolev_b200.h).
#include <cstdarg>
#include <cstdio>

#include "../../include/sobolev_b200.h"
#include "common.cuh"

namespace sab {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace sab

extern "C" const char* sa_last_error(void) { return sab::g_error; }
extern "C" int sa_version(void) { return 100; }
