// Standalone bring-up test for the tcgen05/TMA distance GEMM (embeddingnet_b200/csrc/tc_engine.cuh).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I embeddingnet_b200/csrc \
//              tools/tc_selftest.cu -o build/tc_selftest
// Run on a B200 (gpurun):  ./build/tc_selftest
// Compares C = A.B^T from the tensor-core engine against an fp64 CUDA-core reference on several shapes
// (aligned, ragged M/N, ragged d), for the 1-pass (plain TF32) and 3-pass (fp32-faithful) modes, then times
// the 4096x4096x512 case with a cheap epilogue.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_engine_pair.cuh"

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                           \
    }                                                                                    \
  } while (0)

using namespace en;

struct EpStore {
  struct Params {
    float* out;
    int64_t ld;
    int64_t N;
  };
  struct Row {};
  static constexpr int kSmemBytes = 0;
  static __device__ void item_begin(const Params&, Row&, const tc::Ctx&, int64_t, bool, int, int) {}
  static __device__ void chunk(const Params& p, Row&, const tc::Ctx&, int64_t row, bool valid, int64_t col0, const float (&dot)[32]) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < p.N) p.out[row * p.ld + col0 + j] = dot[j];
  }
  static __device__ void tile_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int) {}
  static __device__ void item_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int, int) {}
};

struct EpRowMax {
  struct Params {
    float* out;  // [M][n_splits]
    int n_splits;
  };
  struct Row {
    float m;
  };
  static constexpr int kSmemBytes = 0;
  static __device__ void item_begin(const Params&, Row& r, const tc::Ctx&, int64_t, bool, int, int) { r.m = -3.4e38f; }
  static __device__ void chunk(const Params&, Row& r, const tc::Ctx&, int64_t, bool, int64_t, const float (&dot)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r.m = fmaxf(r.m, dot[j]);
  }
  static __device__ void tile_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int) {}
  static __device__ void item_end(const Params& p, Row& r, const tc::Ctx&, int64_t row, bool valid, int, int split) {
    if (valid) p.out[(row * p.n_splits + split) * tc::EPI_H + 0] = r.m;  // both halves write (timing only)
  }
};

__global__ void ref_dot_kernel(const float* a, const float* b, int64_t M, int64_t N, int d, double* out) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t i = blockIdx.y;
  if (j >= N || i >= M) return;
  double acc = 0;
  for (int k = 0; k < d; ++k) acc += (double)a[i * d + k] * (double)b[j * d + k];
  out[i * N + j] = acc;
}

static uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

static int run_case(int64_t M, int64_t N, int d, int passes, bool same, int num_sms, bool pair = false,
                    int bf16 = 0) {
  int dpad = tc::dpad_for(d, bf16);
  std::vector<float> ha(M * d), hb(N * d);
  for (int64_t i = 0; i < M * d; ++i) ha[i] = (float)((double)(splitmix(i + 17) >> 40) / 8388608.0 - 1.0);
  for (int64_t i = 0; i < N * d; ++i) hb[i] = (float)((double)(splitmix(i + 9999991) >> 40) / 8388608.0 - 1.0);
  if (same) hb.assign(ha.begin(), ha.begin() + std::min(M, N) * d), hb.resize(N * d, 0.5f);
  float *a, *b, *ahi, *alo, *bhi, *blo, *na, *nb, *out;
  double* ref;
  CK(cudaMalloc(&a, M * d * 4));
  CK(cudaMalloc(&b, N * d * 4));
  CK(cudaMalloc(&ahi, M * dpad * 4));
  CK(cudaMalloc(&alo, M * dpad * 4));
  CK(cudaMalloc(&bhi, N * dpad * 4));
  CK(cudaMalloc(&blo, N * dpad * 4));
  CK(cudaMalloc(&na, M * 4));
  CK(cudaMalloc(&nb, N * 4));
  CK(cudaMalloc(&out, M * N * 4));
  CK(cudaMalloc(&ref, M * N * 8));
  CK(cudaMemcpy(a, ha.data(), M * d * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b, hb.data(), N * d * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(out, 0xFF, M * N * 4));
  CUtensorMap tah, tal, tbh, tbl;
  const int bbox = pair ? tc::BN / 2 : tc::BN;
  if (bf16) {  // planes allocated at fp32 size: more than enough for bf16
    CK(tc::launch_split_bf16(a, M, d, d, dpad, ahi, alo, na, 0));
    CK(tc::launch_split_bf16(b, N, d, d, dpad, bhi, blo, nb, 0));
    if (tc::make_plane_tmap_bf16(&tah, ahi, M, dpad) || tc::make_plane_tmap_bf16(&tal, alo, M, dpad) ||
        tc::make_plane_tmap_bf16(&tbh, bhi, N, dpad) || tc::make_plane_tmap_bf16(&tbl, blo, N, dpad)) {
      printf("tensor map encode failed\n");
      return 1;
    }
  } else {
    CK(tc::launch_split(a, M, d, d, dpad, ahi, alo, na, 0));
    CK(tc::launch_split(b, N, d, d, dpad, bhi, blo, nb, 0));
    if (tc::make_plane_tmap(&tah, ahi, M, dpad) || tc::make_plane_tmap(&tal, alo, M, dpad) ||
        tc::make_plane_tmap(&tbh, bhi, N, dpad, bbox) || tc::make_plane_tmap(&tbl, blo, N, dpad, bbox)) {
      printf("tensor map encode failed\n");
      return 1;
    }
  }
  tc::Shape sh = tc::make_shape(M, N, d, 1 << 30, passes, bf16);
  EpStore::Params ep{out, N, N};
  if (pair) CK(tc::pair::launch<EpStore>(tah, tal, tbh, tbl, sh, ep, num_sms, 0));
  else CK(tc::launch<EpStore>(tah, tal, tbh, tbl, sh, ep, num_sms, 0));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("case M=%lld N=%lld d=%d passes=%d: kernel failed: %s\n", (long long)M, (long long)N, d, passes,
           cudaGetErrorString(e));
    return 1;
  }
  dim3 g((unsigned)((N + 127) / 128), (unsigned)M);
  ref_dot_kernel<<<g, 128>>>(a, b, M, N, d, ref);
  CK(cudaDeviceSynchronize());
  std::vector<float> ho(M * N);
  std::vector<double> hr(M * N);
  CK(cudaMemcpy(ho.data(), out, M * N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hr.data(), ref, M * N * 8, cudaMemcpyDeviceToHost));
  double max_abs = 0, max_ref = 0;
  int64_t bad = 0, nan = 0, worst = 0;
  for (int64_t i = 0; i < M * N; ++i) {
    if (!(ho[i] == ho[i])) { ++nan; continue; }
    double err = fabs((double)ho[i] - hr[i]);
    if (err > max_abs) { max_abs = err; worst = i; }
    max_ref = fmax(max_ref, fabs(hr[i]));
  }
  // scale: |a||b| ~ d/3 for uniform(-1,1)
  double scale = d / 3.0;
  double tol = bf16 ? (passes == 1 ? 2e-2 : 6e-5) : (passes == 1 ? 2e-3 : 8e-6);
  if (max_abs / scale > tol) bad = 1;
  printf("%s%s M=%-5lld N=%-5lld d=%-4d passes=%d same=%d: max_abs_err=%.3e (rel to |a||b| %.3e) nan=%lld worst@(%lld,%lld) got=%.7f ref=%.7f  %s\n",
         pair ? "pair" : "case", bf16 ? "-bf16" : "", (long long)M, (long long)N, d, passes, (int)same, max_abs, max_abs / scale, (long long)nan,
         (long long)(worst / N), (long long)(worst % N), ho[worst], hr[worst], (bad || nan) ? "FAIL" : "ok");
  cudaFree(a); cudaFree(b); cudaFree(ahi); cudaFree(alo); cudaFree(bhi); cudaFree(blo);
  cudaFree(na); cudaFree(nb); cudaFree(out); cudaFree(ref);
  return (bad || nan) ? 1 : 0;
}

static void time_case(int64_t M, int64_t N, int d, int passes, int n_splits, int num_sms, bool pair = false,
                      int bf16 = 0) {
  int dpad = tc::dpad_for(d, bf16);
  float *a, *ahi, *alo, *na, *out;
  CK(cudaMalloc(&a, M * d * 4));
  CK(cudaMalloc(&ahi, M * dpad * 4));
  CK(cudaMalloc(&alo, M * dpad * 4));
  CK(cudaMalloc(&na, M * 4));
  std::vector<float> ha(M * d);
  for (int64_t i = 0; i < M * d; ++i) ha[i] = (float)((double)(splitmix(i + 17) >> 40) / 8388608.0 - 1.0);
  CK(cudaMemcpy(a, ha.data(), M * d * 4, cudaMemcpyHostToDevice));
  CUtensorMap tah, tal, tbh, tbl;
  if (bf16) {
    CK(tc::launch_split_bf16(a, M, d, d, dpad, ahi, alo, na, 0));
    tc::make_plane_tmap_bf16(&tah, ahi, M, dpad);
    tc::make_plane_tmap_bf16(&tal, alo, M, dpad);
    tc::make_plane_tmap_bf16(&tbh, ahi, M, dpad);
    tc::make_plane_tmap_bf16(&tbl, alo, M, dpad);
  } else {
    CK(tc::launch_split(a, M, d, d, dpad, ahi, alo, na, 0));
    tc::make_plane_tmap(&tah, ahi, M, dpad);
    tc::make_plane_tmap(&tal, alo, M, dpad);
    tc::make_plane_tmap(&tbh, ahi, M, dpad, pair ? tc::BN / 2 : tc::BN);
    tc::make_plane_tmap(&tbl, alo, M, dpad, pair ? tc::BN / 2 : tc::BN);
  }
  tc::Shape sh = tc::make_shape(M, N, d, n_splits, passes, bf16);
  CK(cudaMalloc(&out, M * sh.n_splits * tc::EPI_H * 4));
  EpRowMax::Params ep{out, sh.n_splits};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto go = [&]() {
    return pair ? tc::pair::launch<EpRowMax>(tah, tal, tbh, tbl, sh, ep, num_sms, 0)
                : tc::launch<EpRowMax>(tah, tal, tbh, tbl, sh, ep, num_sms, 0);
  };
  for (int i = 0; i < 3; ++i) CK(go());
  CK(cudaDeviceSynchronize());
  const int iters = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) CK(go());
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double us = ms * 1000.0 / iters;
  double flops = 2.0 * M * N * d;
  printf("%s%s M=%lld N=%lld d=%d passes=%d n_splits=%d items=%d: %.2f us/launch  -> %.1f TFLOP/s algorithmic, %.1f TFLOP/s issued\n",
         pair ? "time-pair" : "time", bf16 ? "-bf16" : "", (long long)M, (long long)N, d, passes, sh.n_splits, sh.tiles_m * sh.n_splits, us, flops / us * 1e-6,
         flops * passes / us * 1e-6);
  cudaFree(a); cudaFree(ahi); cudaFree(alo); cudaFree(na); cudaFree(out);
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int sms = prop.multiProcessorCount;
  int fails = 0;
  fails += run_case(128, 128, 32, 1, false, sms);
  fails += run_case(128, 128, 32, 3, false, sms);
  fails += run_case(128, 128, 128, 3, false, sms);
  fails += run_case(256, 384, 64, 3, false, sms);
  fails += run_case(300, 200, 100, 3, false, sms);
  fails += run_case(1000, 1000, 512, 1, true, sms);
  fails += run_case(1000, 1000, 512, 3, true, sms);
  fails += run_case(2048, 4096, 256, 3, false, sms);
  // BF16 planes (kind::f16), 1 and 3 passes
  fails += run_case(128, 128, 64, 1, false, sms, false, 1);
  fails += run_case(128, 128, 64, 3, false, sms, false, 1);
  fails += run_case(256, 384, 128, 3, false, sms, false, 1);
  fails += run_case(300, 200, 100, 3, false, sms, false, 1);
  fails += run_case(1000, 1000, 512, 3, true, sms, false, 1);
  fails += run_case(2048, 4096, 256, 3, false, sms, false, 1);
  if (argc > 1) {
    fails += run_case(256, 128, 32, 1, false, sms, true);
    fails += run_case(256, 128, 32, 3, false, sms, true);
    fails += run_case(256, 384, 64, 3, false, sms, true);
    fails += run_case(300, 200, 100, 3, false, sms, true);
    fails += run_case(1000, 1000, 512, 3, true, sms, true);
    fails += run_case(2048, 4096, 256, 3, false, sms, true);
  }
  printf("selftest: %d failing cases\n", fails);
  if (true) {
    time_case(4096, 4096, 512, 3, 32, sms);
    time_case(4096, 4096, 512, 3, 4, sms);
    time_case(4096, 4096, 512, 1, 32, sms);
    time_case(16384, 16384, 512, 3, 16, sms);
    time_case(16384, 16384, 512, 1, 16, sms);
    time_case(4096, 4096, 512, 3, 32, sms, false, 1);
    time_case(16384, 16384, 512, 3, 16, sms, false, 1);
    time_case(16384, 16384, 512, 1, 16, sms, false, 1);
    if (argc > 1) {
      time_case(4096, 4096, 512, 3, 32, sms, true);
      time_case(16384, 16384, 512, 3, 16, sms, true);
      time_case(16384, 16384, 512, 1, 16, sms, true);
    }
  }
  return fails ? 1 : 0;
}
