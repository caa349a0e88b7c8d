// Bring-up test: tcgen05.mma kind::f16 with the A operand in TENSOR MEMORY (BF16 packed two per 32-bit column).
// Checks the packing convention the pair-backward kernel relies on (csrc/ptx_sm100.cuh: mma_bf16_ts) against a host
// reference, for N = 128 and N = 256, K = 64 (one 128-byte swizzle row of B = 4 instructions).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I embeddingnet_b200/csrc tools/tmem_a_bf16_test.cu -o build/tmem_a_bf16_test
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ptx_sm100.cuh"

using namespace en;

static inline uint16_t f2bf(float x) {  // exact for the small integers used here
  uint32_t u;
  memcpy(&u, &x, 4);
  return static_cast<uint16_t>(u >> 16);
}

// swap = 0: element 2c in the low half of column c;  swap = 1: in the high half
template <int N>
__global__ void __launch_bounds__(128, 1) test_kernel(const uint16_t* __restrict__ a /*128 x 64*/,
                                                      const uint16_t* __restrict__ b /*N x 64*/, float* __restrict__ out,
                                                      int swap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* bt = smem;                                   // N rows x 128 bytes, 128B swizzle
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + N * 128);
  uint32_t* tbase = reinterpret_cast<uint32_t*>(smem + N * 128 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // B tile: row n, 16-byte chunk c (8 bf16) lives at n * 128 + ((c ^ (n & 7)) * 16)
  for (int i = threadIdx.x; i < N * 8; i += 128) {
    const int n = i >> 3, c = i & 7;
    uint4 v = *reinterpret_cast<const uint4*>(b + n * 64 + c * 8);
    *reinterpret_cast<uint4*>(bt + n * 128 + ((c ^ (n & 7)) * 16)) = v;
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  if (warp == 0) ptx::tmem_alloc<512>(tbase);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tbase;
  // A: row r = threadIdx.x (TMEM lane), 64 bf16 -> 32 packed columns at TMEM columns [256, 288)
  {
    float v[32];
    const int r = threadIdx.x;
    for (int c = 0; c < 32; ++c) {
      const uint32_t e0 = a[r * 64 + 2 * c], e1 = a[r * 64 + 2 * c + 1];
      v[c] = __uint_as_float(swap ? (e1 | (e0 << 16)) : (e0 | (e1 << 16)));
    }
    ptx::tmem_st_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 256, v);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const uint64_t bd = ptx::make_kmajor_sw128_desc(ptx::smem_u32(bt));
    for (int k = 0; k < 4; ++k) ptx::mma_bf16_ts(tmem, tmem + 256 + k * 8, bd + k * 2, idesc, k != 0);
    ptx::mma_commit(bar);
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  for (int c = 0; c < N / 32; ++c) {
    float v[32];
    ptx::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[threadIdx.x * N + c * 32 + j] = v[j];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

template <int N>
static int run(int swap) {
  std::vector<uint16_t> a(128 * 64), b(N * 64);
  std::vector<float> af(128 * 64), bf(N * 64);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 64; ++k) {
      af[r * 64 + k] = static_cast<float>((r * 7 + k * 3) % 17 - 8);
      a[r * 64 + k] = f2bf(af[r * 64 + k]);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 64; ++k) {
      bf[n * 64 + k] = static_cast<float>((n * 5 + k * 11) % 13 - 6);
      b[n * 64 + k] = f2bf(bf[n * 64 + k]);
    }
  uint16_t *da, *db;
  float* dout;
  cudaMalloc(&da, a.size() * 2);
  cudaMalloc(&db, b.size() * 2);
  cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  const int smem_bytes = N * 128 + 128;
  cudaFuncSetAttribute(test_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  test_kernel<N><<<1, 128, smem_bytes>>>(da, db, dout, swap);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d swap=%d: CUDA error %s\n", N, swap, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> out(128 * N);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  double worst = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += static_cast<double>(af[r * 64 + k]) * bf[n * 64 + k];
      const double err = fabs(ref - out[r * N + n]);
      if (err > worst) worst = err;
      if (err > 1e-3) ++bad;
    }
  printf("N=%d swap=%d: mismatches %d / %d, worst |err| %.3g  -> %s\n", N, swap, bad, 128 * N, worst,
         bad == 0 ? "PACKING OK" : "no");
  cudaFree(da);
  cudaFree(db);
  cudaFree(dout);
  return bad == 0 ? 0 : 1;
}

int main() {
  int ok = 0;
  ok |= run<128>(0) == 0 ? 1 : 0;
  run<128>(1);
  ok |= run<256>(0) == 0 ? 2 : 0;
  run<256>(1);
  printf("result: %s\n", ok == 3 ? "low-half-even-k packing confirmed for N=128 and N=256" : "NOT confirmed");
  return ok == 3 ? 0 : 1;
}
