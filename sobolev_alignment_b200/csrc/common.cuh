// Error plumbing shared by every translation unit of libsobolev_b200.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace sab {

void set_error(const char* fmt, ...);  // thread-local message behind sa_last_error()

// Tuning / diagnostic options: sa_set_option() value, else the SAB_* environment
// variable read once at first use.
bool option(const char* name, char (&out)[64]);
int option_int(const char* name, int dflt);

#define SAB_CHECK_CUDA(expr)                                                                 \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::sab::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                       __LINE__);                                                            \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define SAB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::sab::set_error(__VA_ARGS__);    \
      return 2;                         \
    }                                   \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace sab
