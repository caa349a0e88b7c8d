// Library-wide plumbing (version, thread-local error text, launch accounting, device queries) and the
// deterministic synthetic-data generator used by bench.py and the tests.
#include "common.cuh"

namespace en {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int64_t& launch_counter() {
  static thread_local int64_t c = 0;
  return c;
}

namespace {
struct ProfState {
  bool on = false;
  bool have = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEvent_t marks[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n_marks = 0;
};
ProfState& prof() {
  static thread_local ProfState p;
  return p;
}
}  // namespace

void prof_begin(cudaStream_t st) {
  ProfState& p = prof();
  if (!p.on) return;
  if (!p.e0) {
    cudaEventCreate(&p.e0);
    cudaEventCreate(&p.e1);
  }
  cudaEventRecord(p.e0, st);
}
void prof_end(cudaStream_t st) {
  ProfState& p = prof();
  if (!p.on || !p.e0) return;
  cudaEventRecord(p.e1, st);
  p.have = true;
}

// stage boundaries of a multi-kernel entry point (en_prof_marks_ms); slot 0 = entry
void prof_mark(cudaStream_t st, int slot) {
  ProfState& p = prof();
  if (!p.on || slot < 0 || slot >= 8) return;
  if (!p.marks[slot]) cudaEventCreate(&p.marks[slot]);
  cudaEventRecord(p.marks[slot], st);
  if (slot == 0) p.n_marks = 1;
  else if (slot + 1 > p.n_marks) p.n_marks = slot + 1;
}

int device_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cache[dev] = n;
  }
  return cache[dev];
}

int check_sm100() {
  static int cache[64] = {0};  // 0 unknown, 1 ok, 2 bad
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return EN_ERR_ARCH;
  if (cache[dev] == 0) {
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cache[dev] = major == 10 ? 1 : 2;
  }
  if (cache[dev] != 1) return fail(EN_ERR_ARCH, "embeddingnet_b200 requires an sm_100 (B200) device");
  return 0;
}

namespace {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// u in [-1, 1), a multiple of 2^-23: exactly representable in fp32, no transcendental functions.
__device__ inline float hash_u(uint64_t r, uint64_t c, uint64_t seed) {
  const uint64_t h = splitmix64(seed ^ (r * 2654435761ull + c));
  return static_cast<float>(static_cast<int64_t>(h >> 40)) * (1.0f / 8388608.0f) - 1.0f;
}

__global__ void synth_fill_kernel(float* __restrict__ x, int64_t rows, int d, int64_t row_offset, uint64_t seed_c,
                                  uint64_t seed_n, int64_t n_classes, int64_t rows_per_class, float noise, int relu,
                                  int32_t* __restrict__ labels_out) {
  const int64_t total = rows * d;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / d;
    const int c = static_cast<int>(i - r * d);
    const uint64_t R = static_cast<uint64_t>(r + row_offset);
    float v;
    if (n_classes > 0) {
      const uint64_t label = rows_per_class > 0 ? (R / static_cast<uint64_t>(rows_per_class)) % n_classes : R % n_classes;
      v = __fadd_rn(hash_u(label, c, seed_c), __fmul_rn(noise, hash_u(R, c, seed_n)));
      if (c == 0 && labels_out) labels_out[r] = static_cast<int32_t>(label);
    } else {
      v = hash_u(R, c, seed_n);
    }
    if (relu) v = fmaxf(v, 0.f);
    x[i] = v;
  }
}

}  // namespace
}  // namespace en

using namespace en;

extern "C" {

const char* en_version(void) { return "embeddingnet_b200 0.1.0 (sm_100a)"; }
const char* en_last_error(void) { return last_error_buf(); }
int64_t en_launch_count(void) { return launch_counter(); }
void en_launch_count_reset(void) { launch_counter() = 0; }

int en_prof_enable(int on) {
  prof().on = on != 0;
  prof().have = false;
  return EN_OK;
}

int en_prof_last_ms(float* ms_host) {
  EN_REQUIRE(ms_host != nullptr, "en_prof_last_ms: null output");
  ProfState& p = prof();
  if (!p.have) return fail(EN_ERR_ARG, "en_prof_last_ms: no timed launch recorded (call en_prof_enable(1) first)");
  EN_CUDA(cudaEventSynchronize(p.e1));
  EN_CUDA(cudaEventElapsedTime(ms_host, p.e0, p.e1));
  return EN_OK;
}

int en_prof_marks_ms(float* ms_host, int capacity, int* n_out) {
  EN_REQUIRE(ms_host != nullptr && n_out != nullptr && capacity > 0, "en_prof_marks_ms: bad arguments");
  ProfState& p = prof();
  const int n = p.n_marks - 1 < capacity ? p.n_marks - 1 : capacity;
  *n_out = n < 0 ? 0 : n;
  for (int i = 0; i < n; ++i) {
    EN_CUDA(cudaEventSynchronize(p.marks[i + 1]));
    EN_CUDA(cudaEventElapsedTime(&ms_host[i], p.marks[i], p.marks[i + 1]));
  }
  return EN_OK;
}

int en_synth_fill(float* x, int64_t rows, int d, int64_t row_offset, uint64_t seed_centre, uint64_t seed_noise,
                  int64_t n_classes, int64_t rows_per_class, float noise, int relu, int32_t* labels_out,
                  void* stream) {
  EN_REQUIRE(x && rows >= 0 && d > 0, "en_synth_fill: bad arguments");
  if (rows == 0) return EN_OK;
  const int64_t total = rows * d;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  synth_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(
      x, rows, d, row_offset, seed_centre, seed_noise, n_classes, rows_per_class, noise, relu, labels_out);
  EN_LAUNCHED("synth_fill_kernel");
  return EN_OK;
}

}  // extern "C"
