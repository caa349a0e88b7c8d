// EXPERIMENT (not linked into libembeddingnet_b200.so): CTA-pair variant of the distance engine.
//
// Result on B200 (tools/tc_selftest pair, round 1): numerically identical to the single-CTA kernel on every test
// shape, but 15 % SLOWER (631 vs 743 TFLOP/s TF32 issued at 16384 x 16384 x 512, 3 passes).  The single-CTA kernel
// already runs at ~90 % of the TF32 equivalent of the measured cuBLAS bf16 burst peak under the 1 kW power cap
// (SM clock ~1.5 GHz), so relieving shared-memory traffic buys nothing, while the cross-CTA barrier hops lengthen
// the producer -> MMA -> producer loop.  Kept because a 256 x 256 pair tile (N = 256 MMAs) is the natural next
// experiment; only tools/tc_selftest.cu includes this header.
#pragma once
#include "tc_engine.cuh"

namespace en {
namespace tc {

// ================================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs owns a 256-row x 128-column tile.  Each CTA keeps its own
// 128 rows of A and only HALF of the B tile (64 rows) in shared memory; the leader CTA issues M = 256 MMAs that
// read A and B from both CTAs and write each CTA's 128 x 128 accumulator into its own TMEM.  Per CTA and k-block
// that is 48 KiB of TMA writes instead of 64 KiB and 6 KiB instead of 8 KiB of operand reads per MMA, which lifts
// the shared-memory-port limit of the single-CTA kernel and halves the L2 -> SM traffic for B.
// ================================================================================================
namespace pair {

constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;          // 16 KiB (hi or lo)
constexpr int B_BYTES = (BN / 2) * BK * 4;    // 8 KiB: this CTA's half of the B tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 48 KiB
constexpr int SMEM_BASE_BYTES = STAGES * STAGE_BYTES + 256 + EPI_WARPS * WARP_SCRATCH_BYTES;
constexpr int SMEM_EP_MAX = 232448 - SMEM_BASE_BYTES;

struct Barriers {
  uint64_t full[STAGES];        // used in the leader CTA only: both CTAs' TMA bytes land here
  uint64_t empty[STAGES];       // leader's commit multicasts to both CTAs
  uint64_t tmem_full[NUM_ACC];  // leader's commit multicasts to both CTAs
  uint64_t tmem_empty[NUM_ACC]; // leader's copy collects the arrivals of both CTAs' epilogue warps
  uint32_t tmem_base;
};

// Work item -> (row-tile PAIR, column-tile range)
__device__ __forceinline__ Item decode_pair_item(const Shape& sh, int item) {
  Item it;
  const int pairs_m = (sh.tiles_m + 1) / 2;
  if (sh.symmetric) {
    // pair P covers row tiles 2P and 2P+1; it needs the column tiles J >= 2P
    int P = 0, rem = item, len = sh.tiles_n;
    while (rem >= len) {
      rem -= len;
      len -= 2;
      ++P;
    }
    it.tile_m = P;
    it.split = 2 * P + rem;
    it.nt0 = it.split;
    it.nt1 = it.nt0 + 1;
  } else {
    it.tile_m = item % pairs_m;
    it.split = item / pairs_m;
    it.nt0 = it.split * sh.tiles_per_split;
    it.nt1 = min(it.nt0 + sh.tiles_per_split, sh.tiles_n);
  }
  return it;
}

inline int num_pair_items(const Shape& sh) {
  const int pairs_m = (sh.tiles_m + 1) / 2;
  if (!sh.symmetric) return pairs_m * sh.n_splits;
  int n = 0;
  for (int P = 0; P < pairs_m; ++P) n += sh.tiles_n - 2 * P > 0 ? sh.tiles_n - 2 * P : 0;
  return n;
}

template <class Ep>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
dist_gemm_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
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
  const uint32_t rank = ptx::cluster_ctarank();   // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

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
      ptx::mbar_init(&bars->tmem_empty[a], 2 * EPI_WARPS);  // epilogue warps of BOTH CTAs
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) ptx::tmem_alloc_pair<TMEM_COLS>(&bars->tmem_base);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // the peer's barriers exist before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one per CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const bool lo = shape.passes > 1;
      const uint32_t stage_tx = 2u * (lo ? STAGE_BYTES : (A_BYTES + B_BYTES));  // both CTAs
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const Item it = decode_pair_item(shape, item);
        const int row0 = (2 * it.tile_m + static_cast<int>(rank)) * BM;
        for (int nt = it.nt0; nt < it.nt1; ++nt) {
          const int col0 = nt * BN + static_cast<int>(rank) * (BN / 2);
          for (int kb = 0; kb < shape.kblocks; ++kb) {
            ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* st = smem + stage * STAGE_BYTES;
            if (rank == 0) ptx::mbar_arrive_expect_tx(&bars->full[stage], stage_tx);
            ptx::tma_load_2d_pair(&tm_a_hi, &bars->full[stage], st, kb * BK, row0);
            ptx::tma_load_2d_pair(&tm_b_hi, &bars->full[stage], st + 2 * A_BYTES, kb * BK, col0);
            if (lo) {
              ptx::tma_load_2d_pair(&tm_a_lo, &bars->full[stage], st + A_BYTES, kb * BK, row0);
              ptx::tma_load_2d_pair(&tm_b_lo, &bars->full[stage], st + 2 * A_BYTES + B_BYTES, kb * BK, col0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: leader CTA only
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const Item it = decode_pair_item(shape, item);
        for (int nt = it.nt0; nt < it.nt1; ++nt, ++acc_it) {
          const uint32_t acc = acc_it % NUM_ACC;
          const uint32_t acc_phase = (acc_it / NUM_ACC) & 1;
          ptx::mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
          const uint32_t tmem_x = tmem_d + BN;
          for (int kb = 0; kb < shape.kblocks; ++kb) {
            ptx::mbar_wait(&bars->full[stage], phase);
            ptx::tc_fence_after();
            const uint32_t st = ptx::smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t a_hi = ptx::make_kmajor_sw128_desc(st);
            const uint64_t a_lo = ptx::make_kmajor_sw128_desc(st + A_BYTES);
            const uint64_t b_hi = ptx::make_kmajor_sw128_desc(st + 2 * A_BYTES);
            const uint64_t b_lo = ptx::make_kmajor_sw128_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = static_cast<uint64_t>(k * UMMA_K * 4 / 16);
              if (shape.passes > 1) {
                ptx::mma_tf32_ss_pair(tmem_x, a_lo + koff, b_hi + koff, idesc, (kb | k) != 0);
                ptx::mma_tf32_ss_pair(tmem_x, a_hi + koff, b_lo + koff, idesc, 1);
                ptx::mma_tf32_ss_pair(tmem_d, a_hi + koff, b_hi + koff, idesc, (kb | k) != 0);
              } else {
                ptx::mma_tf32_ss_pair(tmem_d, a_hi + koff, b_hi + koff, idesc, (kb | k) != 0);
              }
            }
            ptx::mma_commit_pair(&bars->empty[stage]);   // frees the slot in BOTH CTAs
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          ptx::mma_commit_pair(&bars->tmem_full[acc]);   // accumulators ready in BOTH CTAs
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (each CTA: its own 128 rows)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t acc_it = 0;
    typename Ep::Row rs;
    uint8_t* ws = warp_scratch + (warp - 2) * WARP_SCRATCH_BYTES;
    const Ctx ctx{ep_smem, reinterpret_cast<float*>(ws), reinterpret_cast<int32_t*>(ws + 128), quarter * 32 + lane,
                  half, lane, quarter};
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const Item it = decode_pair_item(shape, item);
      const int tile_m = 2 * it.tile_m + static_cast<int>(rank);
      const int64_t row = static_cast<int64_t>(tile_m) * BM + quarter * 32 + lane;
      const bool row_valid = row < shape.M;
      Ep::item_begin(ep, rs, ctx, row, row_valid, tile_m, it.split);
      for (int nt = it.nt0; nt < it.nt1; ++nt, ++acc_it) {
        const uint32_t acc = acc_it % NUM_ACC;
        const uint32_t acc_phase = (acc_it / NUM_ACC) & 1;
        ptx::mbar_wait(&bars->tmem_full[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * ACC_COLS;
#pragma unroll 1
        for (int c = half * (COLS_PER_EPI_WARP / 32); c < (half + 1) * (COLS_PER_EPI_WARP / 32); ++c) {
          float dot[32];
          ptx::tmem_ld_32x32(taddr + c * 32, dot);
          if (shape.passes > 1) {
            float cross[32];
            ptx::tmem_ld_32x32(taddr + BN + c * 32, cross);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) dot[j] += cross[j];
          } else {
            ptx::tmem_ld_wait();
          }
          Ep::chunk(ep, rs, ctx, row, row_valid, static_cast<int64_t>(nt) * BN + c * 32, dot);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_remote(&bars->tmem_empty[acc], 0);  // the leader's barrier
        Ep::tile_end(ep, rs, ctx, row, row_valid, nt);
      }
      Ep::item_end(ep, rs, ctx, row, row_valid, tile_m, it.split);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // neither CTA may retire while the other can still touch its smem / barriers / TMEM
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair<TMEM_COLS>(tmem_base);
  }
}

template <class Ep>
inline cudaError_t launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                          const CUtensorMap& b_lo, const Shape& shape, const typename Ep::Params& ep, int num_sms,
                          cudaStream_t stream) {
  static_assert(Ep::kSmemBytes <= SMEM_EP_MAX, "epilogue scratch does not fit beside the operand pipeline");
  constexpr int SMEM_BYTES = SMEM_BASE_BYTES + Ep::kSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(dist_gemm_pair_kernel<Ep>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const int items = num_pair_items(shape);
  int clusters = num_sms / 2;
  if (items < clusters) clusters = items;
  if (clusters < 1) clusters = 1;
  dist_gemm_pair_kernel<Ep><<<2 * clusters, NUM_THREADS, SMEM_BYTES, stream>>>(a_hi, a_lo, b_hi, b_lo, shape, items,
                                                                               ep);
  return cudaGetLastError();
}

}  // namespace pair

}  // namespace tc
}  // namespace en
