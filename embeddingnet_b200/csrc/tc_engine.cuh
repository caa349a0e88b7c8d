// Tensor-core distance engine for sm_100a: C = A . B^T as a tcgen05/TMEM GEMM fed by TMA, with the FP32
// inputs split into two planes (hi, lo) so that hi*hi + hi*lo + lo*hi reproduces the FP32 product: TF32 planes
// ("3xTF32", ~2^-21, kind::tf32) where the value itself is used, BF16 planes ("split-BF16", ~2^-16 worst case,
// kind::f16 at twice the rate) where the GEMM only selects candidates that are re-evaluated exactly.  The 128x128
// accumulator tile never leaves the SM: a pluggable epilogue functor consumes it straight out of TMEM (label masks,
// per-anchor reductions, top-k, candidate counts, ...).
//
// Replaces, for the large shapes, the arithmetic the reference delegates to
//   sklearn.metrics.pairwise_distances      (embedding_net/datagenerators.py:219)
//   sklearn.neighbors.KNeighborsClassifier  (embedding_net/models.py:136-138)
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue.
// A warp can only read the TMEM lane quarter (warp_id % 4), so each quarter (32 tile rows) is served by EPI_H = 2
// warps that split the 128 accumulator columns between them: two epilogue warps per SM sub-partition, which is what
// hides the ALU / shared-memory latency of the per-element epilogue math (ncu, round 1: with one warp per
// sub-partition the epilogue ran at IPC 0.2 and took twice as long as the MMAs it was supposed to hide behind).
// Three mbarrier pipelines: smem full/empty, TMEM full/empty.
#pragma once
#include "ptx_sm100.cuh"

namespace en {
namespace tc {

constexpr int BM = 128;            // rows of A per tile (UMMA M)
constexpr int BN = 128;            // rows of B per tile (UMMA N)
constexpr int BK = 32;             // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;          // tf32: 32 bytes per MMA
constexpr int STAGES = 3;
constexpr int NUM_ACC = 2;         // TMEM accumulator double buffer
constexpr int TILE_BYTES = BM * BK * 4;            // 16 KiB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;        // Ahi, Alo, Bhi, Blo
// Each accumulator buffer holds TWO 128-column tiles: the dominant hi*hi sum and, separately, the small cross
// terms hi*lo + lo*hi.  The tensor core truncates (does not round) when it adds into the FP32 accumulator, so
// the bias grows with the number of MMAs chained into one cell; keeping the cross terms apart cuts that chain
// from 3*d/8 to d/8 links for the large term (measured on B200: 6e-6 -> ~2e-6 relative on all-positive sums).
constexpr int ACC_COLS = 2 * BN;                   // main | cross
constexpr int TMEM_COLS = NUM_ACC * ACC_COLS;      // 512 = all of TMEM (1 CTA/SM anyway, by smem)
constexpr int EPI_H = 2;                           // epilogue warps per TMEM lane quarter (column halves)
constexpr int EPI_WARPS = 4 * EPI_H;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;   // 320
constexpr int COLS_PER_EPI_WARP = BN / EPI_H;      // 64
constexpr int WARP_SCRATCH_BYTES = 256;            // per-epilogue-warp column cache: 32 floats + 32 ints
// No alignment slack: the kernel has no static shared memory, so the dynamic window starts at the CTA's shared
// base, which is 1024-byte aligned; the kernel traps if that ever stops being true.
constexpr int SMEM_BASE_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + EPI_WARPS * WARP_SCRATCH_BYTES;
constexpr int SMEM_EP_MAX = 232448 - SMEM_BASE_BYTES;  // what is left of the 227 KB for an epilogue's scratch
static_assert(BN == BM, "A and B tiles share TILE_BYTES");

struct Shape {
  int64_t M;        // rows of A (anchors / queries)
  int64_t N;        // rows of B (candidates / bank rows)
  int kblocks;      // ceil(d / BK); planes are zero padded to kblocks*BK columns
  int tiles_m;
  int tiles_n;
  int n_splits;     // column-tile ranges per row tile (work item = (row tile, range))
  int tiles_per_split;
  int passes;       // 3 = hi*hi+hi*lo+lo*hi (fp32 faithful), 1 = hi*hi only (plain TF32; diagnostics)
  int symmetric;    // A == B: only tiles with column tile >= row tile are computed (one tile per work item);
                    // the epilogue must reduce each tile both along rows and along columns
  int num_items;
  int nt_base;      // first column tile of this launch (bank scans walk the bank in L2-sized chunks, one launch each)
  int bf16;         // operand planes are BF16 (hi = bf16(x), lo = bf16(x - hi)), kind::f16 MMAs at twice the TF32
                    // rate; ~2^-16 relative instead of ~2^-22: for paths that only SELECT candidates which are then
                    // re-evaluated exactly (bank scan, batch-hard)
  int bk;           // elements per k-block (one 128-byte swizzle row): 32 fp32/TF32 or 64 BF16
  unsigned long long* trace;  // developer aid (null in production): per CTA 64 globaltimer stamps, see trace_stamp()
};

// Developer aid: CTA `blockIdx.x` writes stamp `slot` (ns) when the launch carries a trace buffer.
// Slots: 0 kernel entry, 1 set-up done; MMA thread: 8+4t tile t accumulator free, 9+4t first operands landed,
// 10+4t all MMAs issued; epilogue warp 2 lane 0: 40+2t tile t accumulator full, 41+2t tile t drained.
__device__ __forceinline__ void trace_stamp(const Shape& sh, int slot) {
  if (sh.trace != nullptr && slot < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    sh.trace[blockIdx.x * 64 + slot] = t;
  }
}

// Work item -> (row tile, column-tile range).  Same arithmetic in all three warp roles.
struct Item {
  int tile_m, split, nt0, nt1;
};
__device__ __forceinline__ Item decode_item(const Shape& sh, int item) {
  Item it;
  if (sh.symmetric) {
    // upper-triangular enumeration, row-major: row I holds tiles_n - I items
    int I = 0, rem = item, len = sh.tiles_n;
    while (rem >= len) {
      rem -= len;
      --len;
      ++I;
    }
    it.tile_m = I;
    it.split = I + rem;
    it.nt0 = I + rem;
    it.nt1 = it.nt0 + 1;
  } else {
    // row tile fastest: CTAs that run concurrently share the column (bank) range and differ in the row (query)
    // tile, so they stream the same B tiles at the same time and L2 serves all but the first reader (ncu, round 1:
    // with the split index fastest a 1.6 GB bank cost 48 GB of DRAM reads per scan)
    it.tile_m = item % sh.tiles_m;
    it.split = item / sh.tiles_m;
    it.nt0 = it.split * sh.tiles_per_split;
    it.nt1 = min(it.nt0 + sh.tiles_per_split, sh.tiles_n);
  }
  return it;
}

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[NUM_ACC];
  uint64_t tmem_empty[NUM_ACC];
  uint32_t tmem_base;
};

// Epilogue concept (EPI_H threads share one row of the 128-row tile, each owning BN / EPI_H of its columns, for
// the whole work item; results are published per (row, ..., ctx.half)):
//   struct Ep { struct Params; struct Row; static constexpr int kSmemBytes;   // scratch, <= SMEM_EP_MAX
//     static __device__ void item_begin(const Params&, Row&, const Ctx&, int64_t row, bool row_valid, int tile_m, int split);
//     static __device__ void chunk(const Params&, Row&, const Ctx&, int64_t row, bool row_valid, int64_t col0, const float (&dot)[32]);
//     static __device__ void tile_end(const Params&, Row&, const Ctx&, int64_t row, bool row_valid, int tile_n);
//     static __device__ void item_end(const Params&, Row&, const Ctx&, int64_t row, bool row_valid, int tile_m, int split); };
// `chunk` receives dot[j] = <A[row], B[col0 + j]> for 32 consecutive candidate rows (columns past N hold 0).
struct Ctx {
  uint8_t* smem;   // Ep::kSmemBytes of shared scratch (16-byte aligned), shared by all epilogue threads
  float* wf;       // per-warp scratch: 32 floats (column norms of the current 32-column chunk)
  int32_t* wi;     // per-warp scratch: 32 ints   (column labels of the current chunk)
  int erow;        // this thread's row inside the tile, 0..127
  int half;        // which column share of the row this thread owns, 0..EPI_H-1
  int lane;
  int quarter;     // TMEM lane quarter = 32-row group of the tile this warp serves
};

// Stage the per-column side data of a 32-column chunk once per warp (one coalesced load) instead of one broadcast
// global load per element per thread; read back with 128-bit shared loads (4 columns per instruction).
__device__ __forceinline__ void stage_columns(const Ctx& ctx, const float* __restrict__ norms,
                                              const int32_t* __restrict__ labels, int64_t col0, int64_t n_cols) {
  const int64_t c = col0 + ctx.lane;
  __syncwarp();
  ctx.wf[ctx.lane] = (norms != nullptr && c < n_cols) ? __ldg(&norms[c]) : 0.f;
  if (labels != nullptr) ctx.wi[ctx.lane] = c < n_cols ? __ldg(&labels[c]) : 0;
  __syncwarp();
}

template <class Ep>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dist_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                 const Shape shape, const typename Ep::Params ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tile bases.
  uint8_t* smem = smem_raw;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES);
  uint8_t* warp_scratch = smem + STAGES * STAGE_BYTES + 256;
  uint8_t* ep_smem = warp_scratch + EPI_WARPS * WARP_SCRATCH_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = shape.num_items;
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
      ptx::mbar_init(&bars->tmem_empty[a], EPI_WARPS);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(&bars->tmem_base);
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
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(shape, item);
        const int tile_m = it.tile_m;
        for (int nt = it.nt0; nt < it.nt1; ++nt) {
          for (int kb = 0; kb < shape.kblocks; ++kb) {
            ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* st = smem + stage * STAGE_BYTES;
            const bool lo = shape.passes > 1;
            ptx::mbar_arrive_expect_tx(&bars->full[stage], lo ? STAGE_BYTES : 2 * TILE_BYTES);
            const int kc = kb * shape.bk;
            ptx::tma_load_2d(&tm_a_hi, &bars->full[stage], st + 0 * TILE_BYTES, kc, tile_m * BM);
            ptx::tma_load_2d(&tm_b_hi, &bars->full[stage], st + 2 * TILE_BYTES, kc, (shape.nt_base + nt) * BN);
            if (lo) {
              ptx::tma_load_2d(&tm_a_lo, &bars->full[stage], st + 1 * TILE_BYTES, kc, tile_m * BM);
              ptx::tma_load_2d(&tm_b_lo, &bars->full[stage], st + 3 * TILE_BYTES, kc, (shape.nt_base + nt) * BN);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(BM, BN);
      constexpr uint32_t idesc16 = ptx::make_idesc_bf16(BM, BN);
      constexpr uint32_t idesc_w = ptx::make_idesc_tf32(BM, 2 * BN);      // B = [B_hi ; B_lo], D = main | cross
      constexpr uint32_t idesc16_w = ptx::make_idesc_bf16(BM, 2 * BN);
      const bool bf16 = shape.bf16 != 0;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(shape, item);
        for (int nt = it.nt0; nt < it.nt1; ++nt, ++acc_it) {
          const uint32_t acc = acc_it % NUM_ACC;
          const uint32_t acc_phase = (acc_it / NUM_ACC) & 1;
          ptx::mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
          ptx::tc_fence_after();
          trace_stamp(shape, 8 + 4 * static_cast<int>(acc_it));
          const uint32_t tmem_d = tmem_base + acc * ACC_COLS;   // hi*hi
          const uint32_t tmem_x = tmem_d + BN;                   // hi*lo + lo*hi
          for (int kb = 0; kb < shape.kblocks; ++kb) {
            ptx::mbar_wait(&bars->full[stage], phase);
            ptx::tc_fence_after();
            if (kb == 0) trace_stamp(shape, 9 + 4 * static_cast<int>(acc_it));
            const uint32_t st = ptx::smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t a_hi = ptx::make_kmajor_sw128_desc(st + 0 * TILE_BYTES);
            const uint64_t a_lo = ptx::make_kmajor_sw128_desc(st + 1 * TILE_BYTES);
            const uint64_t b_hi = ptx::make_kmajor_sw128_desc(st + 2 * TILE_BYTES);
            const uint64_t b_lo = ptx::make_kmajor_sw128_desc(st + 3 * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 32 bytes along K inside the 128B swizzle row: +2 in the (addr>>4) field
              // (8 TF32 or 16 BF16 elements per instruction: the same 32 bytes either way)
              const uint64_t koff = static_cast<uint64_t>(k * UMMA_K * 4 / 16);
              // Three products from TWO instructions: the B_hi and B_lo tiles are adjacent in the stage, and so
              // are the main and cross accumulators, so A_hi x [B_hi ; B_lo] (N = 256) leaves hi*hi in the main
              // columns and hi*lo in the cross columns while reading A_hi once; A_lo x B_hi (N = 128) follows.
              // Operand reads from shared memory drop from 24 to 20 KiB per k-step (the MMAs were paced by shared-
              // memory bandwidth: 8 KiB per 64-cycle instruction is the SM's 128 B/clk, with TMA writing beside).
              if (bf16) {
                if (shape.passes > 1) {
                  ptx::mma_bf16_ss(tmem_d, a_hi + koff, b_hi + koff, idesc16_w, (kb | k) != 0);
                  ptx::mma_bf16_ss(tmem_x, a_lo + koff, b_hi + koff, idesc16, 1);
                } else {
                  ptx::mma_bf16_ss(tmem_d, a_hi + koff, b_hi + koff, idesc16, (kb | k) != 0);
                }
              } else if (shape.passes > 1) {
                ptx::mma_tf32_ss(tmem_d, a_hi + koff, b_hi + koff, idesc_w, (kb | k) != 0);
                ptx::mma_tf32_ss(tmem_x, a_lo + koff, b_hi + koff, idesc, 1);
              } else {
                ptx::mma_tf32_ss(tmem_d, a_hi + koff, b_hi + koff, idesc, (kb | k) != 0);
              }
            }
            ptx::mma_commit(&bars->empty[stage]);  // frees the smem slot once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          ptx::mma_commit(&bars->tmem_full[acc]);  // accumulator complete -> epilogue
          trace_stamp(shape, 10 + 4 * static_cast<int>(acc_it));
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (TMEM -> registers)
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are reachable from this warp
    const int half = (warp - 2) >> 2;
    uint32_t acc_it = 0;
    typename Ep::Row rs;
    uint8_t* ws = warp_scratch + (warp - 2) * WARP_SCRATCH_BYTES;
    const Ctx ctx{ep_smem, reinterpret_cast<float*>(ws), reinterpret_cast<int32_t*>(ws + 128), quarter * 32 + lane,
                  half, lane, quarter};
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item it = decode_item(shape, item);
      const int tile_m = it.tile_m, split = it.split, nt0 = it.nt0, nt1 = it.nt1;
      const int64_t row = static_cast<int64_t>(tile_m) * BM + quarter * 32 + lane;
      const bool row_valid = row < shape.M;
      Ep::item_begin(ep, rs, ctx, row, row_valid, tile_m, split);
      for (int nt = nt0; nt < nt1; ++nt, ++acc_it) {
        const uint32_t acc = acc_it % NUM_ACC;
        const uint32_t acc_phase = (acc_it / NUM_ACC) & 1;
        ptx::mbar_wait(&bars->tmem_full[acc], acc_phase);
        ptx::tc_fence_after();
        if (warp == 2 && lane == 0) trace_stamp(shape, 40 + 2 * static_cast<int>(acc_it));
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
          Ep::chunk(ep, rs, ctx, row, row_valid, static_cast<int64_t>(shape.nt_base + nt) * BN + c * 32, dot);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
        Ep::tile_end(ep, rs, ctx, row, row_valid, shape.nt_base + nt);
        if (warp == 2 && lane == 0) trace_stamp(shape, 41 + 2 * static_cast<int>(acc_it));
      }
      Ep::item_end(ep, rs, ctx, row, row_valid, tile_m, split);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Tensor map over a split plane: `rows` x `cols_padded` fp32, row-major, box = (BK cols, 128 rows), 128B swizzle.
// Rows past `rows` are zero-filled by TMA, so ragged M/N need no host padding.
inline int make_plane_tmap(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols_padded,
                           int box_rows = BM) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols_padded), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols_padded) * 4};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

// Same over a BF16 plane: `rows` x `cols_padded` bf16 (cols_padded a multiple of BK16 = 64), box = (64 cols, 128
// rows) = the same 128-byte swizzle rows and 16 KiB tiles as the TF32 planes.
constexpr int BK16 = 64;
inline int make_plane_tmap_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols_padded,
                                int box_rows = BM) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols_padded), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols_padded) * 2};
  cuuint32_t box[2] = {BK16, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}
inline int dpad_for(int d, int bf16) { return bf16 ? (d + BK16 - 1) / BK16 * BK16 : (d + BK - 1) / BK * BK; }

inline Shape make_shape(int64_t M, int64_t N, int d, int n_splits, int passes, int bf16 = 0) {
  Shape s;
  s.M = M;
  s.N = N;
  s.bf16 = bf16;
  s.bk = bf16 ? BK16 : BK;
  s.kblocks = (d + s.bk - 1) / s.bk;
  s.tiles_m = static_cast<int>((M + BM - 1) / BM);
  s.tiles_n = static_cast<int>((N + BN - 1) / BN);
  if (n_splits < 1) n_splits = 1;
  if (n_splits > s.tiles_n) n_splits = s.tiles_n;
  s.tiles_per_split = (s.tiles_n + n_splits - 1) / n_splits;
  s.n_splits = (s.tiles_n + s.tiles_per_split - 1) / s.tiles_per_split;
  s.passes = passes;
  s.symmetric = 0;
  s.num_items = s.tiles_m * s.n_splits;
  s.nt_base = 0;
  s.trace = nullptr;
  return s;
}

// A == B (M == N): upper-triangular tile schedule, one tile per item.
inline Shape make_shape_symmetric(int64_t N, int d, int passes, int bf16 = 0) {
  Shape s = make_shape(N, N, d, 1 << 30, passes, bf16);
  s.symmetric = 1;
  s.num_items = s.tiles_n * (s.tiles_n + 1) / 2;
  return s;
}

template <class Ep>
inline cudaError_t launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                          const CUtensorMap& b_lo, const Shape& shape, const typename Ep::Params& ep, int num_sms,
                          cudaStream_t stream) {
  static_assert(Ep::kSmemBytes <= SMEM_EP_MAX, "epilogue scratch does not fit beside the operand pipeline");
  constexpr int SMEM_BYTES = SMEM_BASE_BYTES + Ep::kSmemBytes;
  // idempotent, cheap; set on every launch so it also holds after a device switch (one attribute per device)
  cudaError_t e =
      cudaFuncSetAttribute(dist_gemm_kernel<Ep>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const int items = shape.num_items;
  const int grid = items < num_sms ? items : num_sms;
  dist_gemm_kernel<Ep><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(a_hi, a_lo, b_hi, b_lo, shape, ep);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- operand prep
// x (rows x d, leading dimension ldx) -> hi/lo TF32 planes (rows x dpad, zero padded) + squared row norms.
// hi = rna_tf32(x), lo = rna_tf32(x - hi): both exactly representable in TF32, so the tensor core sees them
// unmodified.  One warp per row; float4 loads when the row is 16B aligned.
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// `mu` (optional, d floats): subtracted from every row first.  Distances are translation invariant, and centred
// rows give dot products of mixed sign, which removes the tensor core's one-sided accumulation (truncation) bias.
static __global__ void split_planes_kernel(const float* __restrict__ x, int64_t rows, int d, int64_t ldx, int dpad,
                                    float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ norms,
                                    const float* __restrict__ mu) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float* hr = hi + row * dpad;
  float* lr = lo + row * dpad;
  double acc = 0.0;
  for (int c = lane; c < dpad; c += 32) {
    float v = c < d ? xr[c] : 0.0f;
    if (mu != nullptr && c < d) v -= __ldg(&mu[c]);
    float h = to_tf32(v);
    float l = to_tf32(v - h);
    hr[c] = h;
    lr[c] = l;
    acc += static_cast<double>(v) * static_cast<double>(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && norms) norms[row] = static_cast<float>(acc);
}

// BF16 variant: hi = bf16_rn(x), lo = bf16_rn(x - hi); x - hi - lo is below 2^-16 |x|.
__device__ __forceinline__ uint16_t to_bf16_bits(float x) {
  uint16_t r;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}
// `zero_rows` (optional, rows x d floats) is cleared and `zero_word[0..1]` (optional) reset on the way: the fused
// loss + gradient step needs a zeroed gradient buffer and counter, and a store here is cheaper than memset nodes.
static __global__ void split_planes_bf16_kernel(const float* __restrict__ x, int64_t rows, int d, int64_t ldx, int dpad,
                                                uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                float* __restrict__ norms, float* __restrict__ zero_rows,
                                                unsigned* __restrict__ zero_word, const float* __restrict__ mu) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  if (zero_word != nullptr && row == 0 && lane < 2) zero_word[lane] = 0u;  // two adjacent counters
  if (zero_rows != nullptr) {
    float* z = zero_rows + row * d;
    if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(zero_rows) & 15) == 0)
      for (int c = 4 * lane; c < d; c += 128) *reinterpret_cast<float4*>(z + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    else
      for (int c = lane; c < d; c += 32) z[c] = 0.f;
  }
  const float* xr = x + row * ldx;
  uint32_t* hr = reinterpret_cast<uint32_t*>(hi + row * dpad);  // dpad is even (multiple of 64): 2 bf16 per store
  uint32_t* lr = reinterpret_cast<uint32_t*>(lo + row * dpad);
  double acc = 0.0;
  if ((d & 3) == 0 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // four elements per lane and trip: one 128-bit load, two 64-bit stores (the 2-element form below was latency
    // bound: 7.4 us for 4096 x 512 where the DRAM read takes 1.3 us)
    uint2* h2 = reinterpret_cast<uint2*>(hr);
    uint2* l2 = reinterpret_cast<uint2*>(lr);
#pragma unroll 4
    for (int c = 4 * lane; c < dpad; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        v = *reinterpret_cast<const float4*>(xr + c);
        if (mu != nullptr) {
          const float4 m = __ldg(reinterpret_cast<const float4*>(mu + c));
          v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
        }
      }
      const uint32_t h0 = to_bf16_bits(v.x), h1 = to_bf16_bits(v.y), h2b = to_bf16_bits(v.z), h3 = to_bf16_bits(v.w);
      const uint32_t r0 = to_bf16_bits(v.x - __uint_as_float(h0 << 16)), r1 = to_bf16_bits(v.y - __uint_as_float(h1 << 16)),
                     r2 = to_bf16_bits(v.z - __uint_as_float(h2b << 16)), r3 = to_bf16_bits(v.w - __uint_as_float(h3 << 16));
      h2[c >> 2] = make_uint2(h0 | (h1 << 16), h2b | (h3 << 16));
      l2[c >> 2] = make_uint2(r0 | (r1 << 16), r2 | (r3 << 16));
      acc += (static_cast<double>(v.x) * static_cast<double>(v.x) + static_cast<double>(v.y) * static_cast<double>(v.y)) +
             (static_cast<double>(v.z) * static_cast<double>(v.z) + static_cast<double>(v.w) * static_cast<double>(v.w));
    }
  } else
  for (int c = 2 * lane; c < dpad; c += 64) {
    float v0 = c < d ? xr[c] : 0.0f;
    float v1 = c + 1 < d ? xr[c + 1] : 0.0f;
    if (mu != nullptr) {  // centred rows (see split_planes_kernel)
      if (c < d) v0 -= __ldg(&mu[c]);
      if (c + 1 < d) v1 -= __ldg(&mu[c + 1]);
    }
    const uint16_t h0 = to_bf16_bits(v0), h1 = to_bf16_bits(v1);
    const float r0 = v0 - __uint_as_float(static_cast<uint32_t>(h0) << 16);
    const float r1 = v1 - __uint_as_float(static_cast<uint32_t>(h1) << 16);
    hr[c >> 1] = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
    lr[c >> 1] = static_cast<uint32_t>(to_bf16_bits(r0)) | (static_cast<uint32_t>(to_bf16_bits(r1)) << 16);
    acc += static_cast<double>(v0) * static_cast<double>(v0) + static_cast<double>(v1) * static_cast<double>(v1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && norms) norms[row] = static_cast<float>(acc);
}

// Column means of x (B x d), float64 accumulation in a fixed order (deterministic).  Used to centre the operands
// (see split_planes_kernel).
constexpr int kMeanCols = 8;     // columns per block: 8 floats = one 32-byte sector per row
constexpr int kMeanRows = 128;   // row groups per block (blockDim.y)
static __global__ void column_mean_kernel(const float* __restrict__ e, int64_t B, int d, float* __restrict__ mu) {
  // one block of (kMeanCols, kMeanRows) threads per 8 columns (64 blocks at d = 512: the 32-column version kept only
  // 16 SMs busy and took 23 us at 4096 x 512); each thread keeps four independent partial sums over its rows so that
  // the loads overlap, then the row groups are reduced through shared memory in a fixed order (deterministic)
  __shared__ double part[kMeanRows][kMeanCols + 1];
  const int c = blockIdx.x * kMeanCols + threadIdx.x;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (c < d) {
    int64_t r = threadIdx.y;
    for (; r + 3 * kMeanRows < B; r += 4 * kMeanRows) {
      const float x0 = e[r * d + c], x1 = e[(r + kMeanRows) * d + c], x2 = e[(r + 2 * kMeanRows) * d + c],
                  x3 = e[(r + 3 * kMeanRows) * d + c];
      a0 += static_cast<double>(x0);
      a1 += static_cast<double>(x1);
      a2 += static_cast<double>(x2);
      a3 += static_cast<double>(x3);
    }
    for (; r < B; r += kMeanRows) a0 += static_cast<double>(e[r * d + c]);
  }
  part[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  // fixed-shape tree over the row groups
  for (int h = kMeanRows / 2; h > 0; h >>= 1) {
    if (static_cast<int>(threadIdx.y) < h) part[threadIdx.y][threadIdx.x] += part[threadIdx.y + h][threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.y == 0 && c < d) mu[c] = static_cast<float>(part[0][threadIdx.x] / static_cast<double>(B));
}
inline void launch_column_mean(const float* e, int64_t B, int d, float* mu, cudaStream_t st) {
  column_mean_kernel<<<static_cast<unsigned>((d + kMeanCols - 1) / kMeanCols), dim3(kMeanCols, kMeanRows), 0, st>>>(e, B, d, mu);
}

inline cudaError_t launch_split(const float* x, int64_t rows, int d, int64_t ldx, int dpad, float* hi, float* lo,
                                float* norms, cudaStream_t stream, const float* mu = nullptr) {
  if (rows == 0) return cudaSuccess;
  const int threads = 256;
  const int64_t blocks = (rows * 32 + threads - 1) / threads;
  split_planes_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(x, rows, d, ldx, dpad, hi, lo, norms, mu);
  return cudaGetLastError();
}

inline cudaError_t launch_split_bf16(const float* x, int64_t rows, int d, int64_t ldx, int dpad, void* hi, void* lo,
                                     float* norms, cudaStream_t stream, float* zero_rows = nullptr,
                                     unsigned* zero_word = nullptr, const float* mu = nullptr) {
  if (rows == 0) return cudaSuccess;
  const int threads = 256;
  const int64_t blocks = (rows * 32 + threads - 1) / threads;
  split_planes_bf16_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
      x, rows, d, ldx, dpad, static_cast<uint16_t*>(hi), static_cast<uint16_t*>(lo), norms, zero_rows, zero_word,
      mu);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace en
