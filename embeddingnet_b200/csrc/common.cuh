// Shared host-side plumbing for the C ABI: thread-local error text, launch accounting, argument checks.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/embeddingnet_b200.h"

namespace en {

char* last_error_buf();          // thread-local, 512 bytes
int64_t& launch_counter();       // thread-local

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int cuda_fail(cudaError_t e, const char* what) {
  snprintf(last_error_buf(), 512, "%s: %s", what, cudaGetErrorString(e));
  return static_cast<int>(e);
}

#define EN_REQUIRE(cond, ...) \
  do {                        \
    if (!(cond)) return ::en::fail(EN_ERR_ARG, __VA_ARGS__); \
  } while (0)

// Check the launch that was just issued and count it.
#define EN_LAUNCHED(what)                                   \
  do {                                                      \
    cudaError_t e_ = cudaGetLastError();                    \
    if (e_ != cudaSuccess) return ::en::cuda_fail(e_, what); \
    ++::en::launch_counter();                               \
  } while (0)

#define EN_CUDA(call)                                        \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) return ::en::cuda_fail(e_, #call); \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Workspace {
  uint8_t* base;
  size_t size;
  size_t off = 0;
  Workspace(void* p, size_t n) : base(static_cast<uint8_t*>(p)), size(n) {}
  template <class T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= size && (reinterpret_cast<uintptr_t>(base) & 255) == 0; }
};

// Optional timing of the dominant kernel of a call: when enabled (en_prof_enable) the entry points bracket their
// distance-GEMM / streaming-scan launch with CUDA events recorded on the launching stream.
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st);

int device_sm_count();  // SM count of the current device (cached per device), <0 on error
int check_sm100();      // 0 when the current device is compute capability 10.x

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace en
