// Wide-tile variant of the distance engine for the SYMMETRIC split-BF16 case (batch-hard, A == B): a CTA owns one row
// tile and TWO adjacent column tiles (128 x 256), so that the A tiles of a k-block are fetched once for both.
//
// Why (tools/trace_bh.py on B200, B = 4096, d = 512): per 128 x 128 tile the operand stream is 8 k-blocks x 64 KiB =
// 512 KiB; the MMA thread needed 5.5 us per tile where the 96 MMAs take 3.2 us -- 148 SMs x 512 KiB / 5.5 us =
// 13.8 TB/s is the L2 -> SM limit of the chip, not the tensor pipe (the epilogue, 1.7-3.8 us per tile, hides behind
// either).  With A shared by two column tiles the stream is 96 KiB per k-block and tile PAIR: a quarter less L2
// traffic per tile, and the three products per k-step become three N = 256 instructions (A read once from shared
// memory for 256 columns).
//
// Differences from dist_gemm_kernel (csrc/tc_engine.cuh):
//   * stage = A_hi | A_lo (128 rows each) | B_hi | B_lo (256 rows each: one TMA box) = 96 KiB, two stages;
//   * ONE accumulator of 256 columns per tile pair, double buffered (2 x 256 = all of TMEM): hi*hi, hi*lo and lo*hi
//     accumulate together, so the accumulator-truncation term of the error bound counts 3 d/16 links instead of
//     d/16 (the caller widens its band accordingly: batch-hard only SELECTS on these values);
//   * work item = (row tile I, column-tile pair P) with 2P+1 >= I; the epilogue skips the sub-tile left of the
//     diagonal (odd I, P = (I-1)/2) and a phantom sub-tile past the last column tile (odd tile count).
// The epilogue interface (Ep::item_begin / chunk / tile_end / item_end, Ctx) is the one of tc_engine.cuh.
#pragma once
#include "tc_engine.cuh"

namespace en {
namespace tc {
namespace wide {

constexpr int STAGES = 2;
constexpr int BNW = 2 * BN;                        // columns per work item
constexpr int STAGE_BYTES = 6 * TILE_BYTES;        // A_hi, A_lo: 16 KiB each; B_hi, B_lo: 32 KiB each
constexpr int NUM_ACC = 2;
constexpr int ACC_COLS = BNW;                      // one merged accumulator per tile pair
constexpr int SMEM_BASE_BYTES = STAGES * STAGE_BYTES + 256 + EPI_WARPS * WARP_SCRATCH_BYTES;
constexpr int SMEM_EP_MAX = 232448 - SMEM_BASE_BYTES;
static_assert(NUM_ACC * ACC_COLS == 512, "the two accumulators fill tensor memory");

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[NUM_ACC];
  uint64_t tmem_empty[NUM_ACC];
  uint32_t tmem_base;
};

inline int pairs_of(int tiles_n) { return (tiles_n + 1) / 2; }
inline int num_items(int tiles_n) {
  const int pairs = pairs_of(tiles_n);
  int n = 0;
  for (int I = 0; I < tiles_n; ++I) n += pairs - (I >> 1);
  return n;
}

struct Item {
  int tile_m, pair;
};
// row-major over the row tiles; row I holds the pairs I/2 .. pairs-1
__device__ __forceinline__ Item decode_item(int pairs, int item) {
  int I = 0, rem = item, len = pairs;
  while (rem >= len) {
    rem -= len;
    ++I;
    len = pairs - (I >> 1);
  }
  return Item{I, (I >> 1) + rem};
}

// tm_a_*: box (64 columns, 128 rows); tm_b_*: box (64 columns, 256 rows) over the SAME planes.
template <class Ep>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dist_gemm_wide_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                      const Shape shape, const int n_items, const typename Ep::Params ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES);
  uint8_t* warp_scratch = smem + STAGES * STAGE_BYTES + 256;
  uint8_t* ep_smem = warp_scratch + EPI_WARPS * WARP_SCRATCH_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pairs = (shape.tiles_n + 1) >> 1;
  if (threadIdx.x == 0) trace_stamp(shape, 0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a_hi);
    ptx::prefetch_tmap(&tm_a_lo);
    ptx::prefetch_tmap(&tm_b_hi);
    ptx::prefetch_tmap(&tm_b_lo);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      ptx::mbar_init(&bars->tmem_full[a], 1);
      ptx::mbar_init(&bars->tmem_empty[a], EPI_WARPS);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) ptx::tmem_alloc<512>(&bars->tmem_base);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (threadIdx.x == 0) trace_stamp(shape, 1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(pairs, item);
        for (int kb = 0; kb < shape.kblocks; ++kb) {
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&bars->full[stage], STAGE_BYTES);
          const int kc = kb * shape.bk;
          ptx::tma_load_2d(&tm_a_hi, &bars->full[stage], st + 0 * TILE_BYTES, kc, it.tile_m * BM);
          ptx::tma_load_2d(&tm_a_lo, &bars->full[stage], st + 1 * TILE_BYTES, kc, it.tile_m * BM);
          ptx::tma_load_2d(&tm_b_hi, &bars->full[stage], st + 2 * TILE_BYTES, kc, it.pair * BNW);
          ptx::tma_load_2d(&tm_b_lo, &bars->full[stage], st + 4 * TILE_BYTES, kc, it.pair * BNW);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BNW);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
        const uint32_t acc = acc_it % NUM_ACC;
        const uint32_t acc_phase = (acc_it / NUM_ACC) & 1;
        ptx::mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        trace_stamp(shape, 8 + 4 * static_cast<int>(acc_it));
        const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < shape.kblocks; ++kb) {
          ptx::mbar_wait(&bars->full[stage], phase);
          ptx::tc_fence_after();
          if (kb == 0) trace_stamp(shape, 9 + 4 * static_cast<int>(acc_it));
          const uint32_t st = ptx::smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_hi = ptx::make_kmajor_sw128_desc(st + 0 * TILE_BYTES);
          const uint64_t a_lo = ptx::make_kmajor_sw128_desc(st + 1 * TILE_BYTES);
          const uint64_t b_hi = ptx::make_kmajor_sw128_desc(st + 2 * TILE_BYTES);  // 256 rows: 32 groups of 8
          const uint64_t b_lo = ptx::make_kmajor_sw128_desc(st + 4 * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK16 / 16; ++k) {
            const uint64_t koff = static_cast<uint64_t>(k * 16 * 2 / 16);  // 32 bytes along K per instruction
            ptx::mma_bf16_ss(tmem_d, a_lo + koff, b_hi + koff, idesc, (kb | k) != 0);
            ptx::mma_bf16_ss(tmem_d, a_hi + koff, b_lo + koff, idesc, 1);
            ptx::mma_bf16_ss(tmem_d, a_hi + koff, b_hi + koff, idesc, 1);
          }
          ptx::mma_commit(&bars->empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&bars->tmem_full[acc]);
        trace_stamp(shape, 10 + 4 * static_cast<int>(acc_it));
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (TMEM -> registers)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t acc_it = 0;
    typename Ep::Row rs;
    uint8_t* ws = warp_scratch + (warp - 2) * WARP_SCRATCH_BYTES;
    const Ctx ctx{ep_smem, reinterpret_cast<float*>(ws), reinterpret_cast<int32_t*>(ws + 128), quarter * 32 + lane,
                  half, lane, quarter};
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
      const Item it = decode_item(pairs, item);
      const int tile_m = it.tile_m;
      const int64_t row = static_cast<int64_t>(tile_m) * BM + quarter * 32 + lane;
      const bool row_valid = row < shape.M;
      Ep::item_begin(ep, rs, ctx, row, row_valid, tile_m, it.pair);
      const uint32_t acc = acc_it % NUM_ACC;
      const uint32_t acc_phase = (acc_it / NUM_ACC) & 1;
      ptx::mbar_wait(&bars->tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      if (warp == 2 && lane == 0) trace_stamp(shape, 40 + 2 * static_cast<int>(acc_it));
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * ACC_COLS;
      const int t0 = 2 * it.pair, t1 = 2 * it.pair + 1;
      const bool do0 = t0 >= tile_m;                // left of the diagonal otherwise (odd row tile, first pair)
      const bool do1 = t1 < shape.tiles_n;          // phantom sub-tile otherwise (odd number of column tiles)
      if (do0) {
#pragma unroll 1
        for (int c = half * (COLS_PER_EPI_WARP / 32); c < (half + 1) * (COLS_PER_EPI_WARP / 32); ++c) {
          float dot[32];
          ptx::tmem_ld_32x32(taddr + c * 32, dot);
          ptx::tmem_ld_wait();
          Ep::chunk(ep, rs, ctx, row, row_valid, static_cast<int64_t>(t0) * BN + c * 32, dot);
        }
        Ep::tile_end(ep, rs, ctx, row, row_valid, t0);
      }
      if (do1) {
#pragma unroll 1
        for (int c = half * (COLS_PER_EPI_WARP / 32); c < (half + 1) * (COLS_PER_EPI_WARP / 32); ++c) {
          float dot[32];
          ptx::tmem_ld_32x32(taddr + BN + c * 32, dot);
          ptx::tmem_ld_wait();
          Ep::chunk(ep, rs, ctx, row, row_valid, static_cast<int64_t>(t1) * BN + c * 32, dot);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
      if (do1) Ep::tile_end(ep, rs, ctx, row, row_valid, t1);
      if (warp == 2 && lane == 0) trace_stamp(shape, 41 + 2 * static_cast<int>(acc_it));
      Ep::item_end(ep, rs, ctx, row, row_valid, tile_m, it.pair);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// shape: make_shape_symmetric(N, d, 3, /*bf16*/ 1); a_*: 128-row boxes, b_*: 256-row boxes over the same planes
template <class Ep>
inline cudaError_t launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                          const CUtensorMap& b_lo, const Shape& shape, const typename Ep::Params& ep, int num_sms,
                          cudaStream_t stream) {
  static_assert(Ep::kSmemBytes <= SMEM_EP_MAX, "epilogue scratch does not fit beside the wide operand pipeline");
  constexpr int SMEM_BYTES = SMEM_BASE_BYTES + Ep::kSmemBytes;
  cudaError_t e =
      cudaFuncSetAttribute(dist_gemm_wide_kernel<Ep>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const int items = num_items(shape.tiles_n);
  const int grid = items < num_sms ? items : num_sms;
  dist_gemm_wide_kernel<Ep><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(a_hi, a_lo, b_hi, b_lo, shape, items, ep);
  return cudaGetLastError();
}

}  // namespace wide
}  // namespace tc
}  // namespace en
