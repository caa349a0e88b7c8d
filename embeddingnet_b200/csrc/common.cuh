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
void prof_mark(cudaStream_t st, int slot);  // stage boundaries of a multi-kernel entry point (en_prof_marks_ms)

int device_sm_count();  // SM count of the current device (cached per device), <0 on error
int check_sm100();      // 0 when the current device is compute capability 10.x

// Rigorous bound c with |dot~ - dot| <= c |q| |b| for the scan arithmetic (x1.5 safety):
//   operand split: the dropped products lo*lo, r*b, q*r with |x - hi| <= u |x|, |x - hi - lo| <= u^2 |x|
//                  (u = 2^-11 TF32 round-to-nearest, 2^-8 BF16)                      -> 3 u^2 (1 + u)
//   accumulation : the tensor core truncates when it adds into the fp32 accumulator, <= 2^-22 of the partial sum
//                  per MMA (measured bias on B200: ~2^-24 per link), d / K_mma links     -> 2^-22 (links + 1)
//   fp32 stream  : d/32 sequential FMAs per lane + 5 shuffle adds, 2^-24 each.
inline double cert_bound(int precision, int d) {
  const double p22 = 1.0 / 4194304.0;
  double c;
  if (precision == EN_PREC_BF16X3) c = 3.0 / 65536.0 * (1.0 + 1.0 / 256.0) + p22 * ((d + 15) / 16 + 1);
  else if (precision == EN_PREC_TF32X3) c = 3.0 * p22 * (1.0 + 1.0 / 2048.0) + p22 * ((d + 7) / 8 + 1);
  else c = p22 / 4.0 * ((d + 31) / 32 + 8);
  return 1.5 * c;
}

// Per-(row, column range, half) partial of a pair loss: the sum of its terms and, for batch-all, how many were > 0.
struct PairPartial {
  double sum;
  unsigned long long npos;
};

// What pair_tc_launch() leaves in its workspace for pair_tc_finish() (csrc/pair_tc.cu)
struct PairTcFinish {
  const float* mu;      // [d] column means of the embeddings
  const float* rowsum;  // [B] row sums of the pair coefficients
};

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
