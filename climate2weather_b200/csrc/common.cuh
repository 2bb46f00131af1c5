// Shared host-side helpers for the C-ABI library: error slot, checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace c2w {

enum : int {
  C2W_OK = 0,
  C2W_ERR_INVALID = -1,   // bad argument / unsupported shape
  C2W_ERR_CUDA = -2,      // CUDA runtime / driver failure
  C2W_ERR_STATE = -3,     // call order violated (weights not finalised, workspace not bound ...)
  C2W_ERR_MISSING = -4,   // a weight the architecture needs was never loaded
};

inline char* error_slot() {
  static thread_local char buf[1024] = {0};
  return buf;
}

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_slot(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

#define C2W_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::c2w::fail(::c2w::C2W_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                         cudaGetErrorString(_e));                                                   \
  } while (0)

#define C2W_REQUIRE(cond, ...)                                          \
  do {                                                                  \
    if (!(cond)) return ::c2w::fail(::c2w::C2W_ERR_INVALID, __VA_ARGS__); \
  } while (0)

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace c2w
