// Shared helpers of the tor10_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tor10_b200.h"

namespace t10 {

void set_error(const char* fmt, ...);

#define T10_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      t10::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return T10_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define T10_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      t10::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

#define T10_LAUNCH_CHECK() T10_CUDA(cudaGetLastError())

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();

// ordinal of the current device, clamped into [0, T10_MAX_DEVICES): index of per-device set-up caches (function
// attributes and occupancy are properties of a (kernel, device) pair, not of the process -- ADVICE r1)
constexpr int T10_MAX_DEVICES = 64;
static inline int cur_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= T10_MAX_DEVICES) d = 0;
  return d;
}

// stream-ordered scratch that frees itself (used only while building maps)
struct Scratch {
  cudaStream_t st;
  void* ptrs[32];
  int n = 0;
  explicit Scratch(cudaStream_t s) : st(s) {}
  template <typename T>
  cudaError_t alloc(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, (count ? count : 1) * sizeof(T), st);
    if (e == cudaSuccess) ptrs[n++] = q;
    *p = (T*)q;
    return e;
  }
  ~Scratch() {
    for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st);
  }
};

}  // namespace t10
