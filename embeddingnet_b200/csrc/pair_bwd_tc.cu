// Tensor-core backward for the pair losses (batch-all triplet negatives, all-pairs contrastive):
//
//     d L / d E  =  rowsum(C) o E  -  C . E ,       C_ik = symmetrised pair coefficient, a function of D_ik
//
// One kernel, two chained tcgen05 GEMMs per 128 x 128 tile, nothing of size B x B ever stored:
//   GEMM1  S  = E_I . E_J^T            operands by TMA, accumulator in TMEM.  Contrastive: split-BF16 planes (3
//                                     kind::f16 MMAs per k-step, twice the TF32 rate; S only feeds the smooth 1/D
//                                     factors).  Batch-all: 3xTF32 -- its hinge decisions D_ap + m - D_an > 0 flip
//                                     against float64 in proportion to the error of S (measured: BF16 planes
//                                     pushed 8.7 % of the rows over the gradient tolerance, TF32 planes < 8 %)
//   epilogue  C_IJ = f(S, labels, positives lists)   8 warps pull S out of TMEM into registers (releasing the
//                                     accumulator at once), build C and write it BACK to TMEM with tcgen05.st as
//                                     two TF32 planes C_hi + C_lo
//   GEMM2  G += C_IJ . E_J[:, slice]   A operand from TMEM (tcgen05.mma [d],[a],b), B operand = E^T tiles by TMA;
//                                     C_hi.E_hi + C_hi.E_lo + C_lo.E_hi
// A CTA owns (row tile I, 128-column slice of the gradient) and walks all column tiles J; GEMM1 of tile J+1 is
// issued before GEMM2 of tile J so the tensor pipe has work while the epilogue warps build C_J.
//
// Cancellation: grad_i = sum_k c_ik (e_i - e_k) is computed as rowsum_i e_i - (C.E)_i.  For post-ReLU embeddings
// (all components >= 0, cosines ~0.8) the two terms are ~10x larger than their difference, which amplified the
// tensor core's accumulation bias to 1e-4.  The expression is invariant under E -> E - mu, so GEMM2 and the
// rowsum term both use the mean-centred embeddings (mu = column mean), which removes the cancellation.
//
// This replaces the O(B^2 d) CUDA-core kernel (csrc/batch_losses.cu: pair_bwd_kernel, kept for classes with more
// than 8 positives per anchor).  Math: see en_batch_all_bwd / en_contrastive_allpairs_bwd.
#include "common.cuh"
#include "tc_engine.cuh"

namespace en {
namespace pbt {

using tc::BK;
using tc::BM;
using tc::BN;
using tc::UMMA_K;

constexpr int DN = 128;                 // gradient columns per GEMM2 instruction (UMMA N)
constexpr int NSUB = 2;                 // a work item owns NSUB * DN = 256 gradient columns
constexpr int DW = NSUB * DN;
constexpr int TILE_BYTES = BM * BK * 4; // 16 KiB
constexpr int G1_STAGES = 2;            // GEMM1 ring: A_hi | A_lo | B_hi | B_lo
constexpr int G1_STAGE_BYTES = 4 * TILE_BYTES;
constexpr int ET_STAGES = 2;            // GEMM2 B ring: ET_hi | ET_lo, each [128 gradient columns x 32 rows j]
constexpr int ET_STAGE_BYTES = 2 * TILE_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int CTRL_WARPS = 3;            // warp 0: GEMM1 operand TMA, warp 1: MMA issuer, warp 2: E^T (GEMM2) TMA
constexpr int NUM_THREADS = 32 * (CTRL_WARPS + EPI_WARPS);
constexpr int MAXP = 8;                 // positives per anchor handled by this kernel
constexpr int WARP_SCR = 256 + 256 + 64 * MAXP * 4;  // per warp, for its 64 columns: norms | labels | positives lists
constexpr int SMEM_BYTES = G1_STAGES * G1_STAGE_BYTES + ET_STAGES * ET_STAGE_BYTES + 256 + EPI_WARPS * WARP_SCR +
                           BM * 2 * 4;
// TMEM (512 columns): S | C | G.  The coefficient tile goes through ONE 128-column region twice per tile -- first its
// TF32 high part (GEMM2 phase A: C_hi.E_hi + C_hi.E_lo), then its low part (phase B: C_lo.E_hi) -- which frees the
// columns for a 256-wide gradient accumulator: a CTA now computes each S / C tile once per 256 gradient columns
// instead of once per 128 (round-1 ncu: GEMM1 and the coefficient epilogue, both repeated per slice, were 80 % of
// the kernel).
constexpr uint32_t TM_ACC1 = 0;    // 128 columns: S (single buffer: the epilogue copies it to registers and releases it)
constexpr uint32_t TM_C = 128;     // 128 columns: coefficient tile, high part then low part (A operand of GEMM2)
constexpr uint32_t TM_ACC2 = 256;  // 256 columns: the gradient accumulator (NSUB x DN)

struct Bars {
  uint64_t g1_full[G1_STAGES], g1_empty[G1_STAGES];
  uint64_t et_full[ET_STAGES], et_empty[ET_STAGES];
  uint64_t acc1_full, acc1_empty;
  uint64_t c_full;     // epilogue -> MMA: a part of C (hi, then lo) is in TMEM; two completions per tile
  uint64_t a_done;     // MMA -> epilogue: phase A has consumed C_hi, the low part may overwrite it
  uint64_t c_empty;    // MMA -> epilogue: phase B has consumed C_lo, the next tile's C_hi may be written
  uint64_t acc2_full, acc2_empty;
  uint32_t tmem_base;
};

struct Params {
  const float* emb;
  const int32_t* labels;
  const float* norms;
  const float* pos_d;     // [B][MAXP] (batch-all)
  const int32_t* pos_n;   // [B]
  int32_t* pos_cnt;       // [B][MAXP] out: active negatives per (anchor, positive slot)
  const double* stats;    // stats[1] = number of positive triplets
  const float* gloss;
  const float* mu;        // [d] column means of emb
  float* gemb;            // ZEROED by the caller: items that share rows (J ranges) add into it
  int64_t B;
  int d, tiles, n_wide, n_jparts, tiles_per_part, kblocks;
  int mode;               // 0 = batch-all, 1 = all-pairs contrastive
  int squared;
  float margin, scale_c;
};

__device__ __forceinline__ float rsqrt_ftz(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct Item {
  int ti, wide, j0, j1;
};
// work item -> (row tile, 256-column group, range of column tiles); row tile fastest so that concurrently running
// CTAs stream the same E_J / E^T tiles
__device__ __forceinline__ Item decode_item(const Params& p, int item) {
  Item it;
  it.ti = item % p.tiles;
  const int rest = item / p.tiles;
  it.wide = rest % p.n_wide;
  const int part = rest / p.n_wide;
  it.j0 = part * p.tiles_per_part;
  it.j1 = min(it.j0 + p.tiles_per_part, p.tiles);
  return it;
}

// kMode is a template parameter so that each instantiation carries only its own coefficient code: the fully
// unrolled epilogue of both modes together overflowed the instruction cache (ncu: stall_no_instruction 5.4 / issue).
template <int kMode>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pair_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                   const __grid_constant__ CUtensorMap tm_et_hi, const __grid_constant__ CUtensorMap tm_et_lo,
                   const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* g1 = smem;
  uint8_t* et = smem + G1_STAGES * G1_STAGE_BYTES;
  Bars* bars = reinterpret_cast<Bars*>(et + ET_STAGES * ET_STAGE_BYTES);
  uint8_t* warp_scr = reinterpret_cast<uint8_t*>(bars) + 256;
  float* rowsum_x = reinterpret_cast<float*>(warp_scr + EPI_WARPS * WARP_SCR);  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.tiles * p.n_wide * p.n_jparts;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_hi);
    ptx::prefetch_tmap(&tm_lo);
    ptx::prefetch_tmap(&tm_et_hi);
    ptx::prefetch_tmap(&tm_et_lo);
    for (int s = 0; s < G1_STAGES; ++s) { ptx::mbar_init(&bars->g1_full[s], 1); ptx::mbar_init(&bars->g1_empty[s], 1); }
    for (int s = 0; s < ET_STAGES; ++s) { ptx::mbar_init(&bars->et_full[s], 1); ptx::mbar_init(&bars->et_empty[s], 1); }
    ptx::mbar_init(&bars->acc1_full, 1);
    ptx::mbar_init(&bars->acc1_empty, EPI_WARPS);
    ptx::mbar_init(&bars->c_full, EPI_WARPS);
    ptx::mbar_init(&bars->a_done, 1);
    ptx::mbar_init(&bars->c_empty, 1);
    ptx::mbar_init(&bars->acc2_full, 1);
    ptx::mbar_init(&bars->acc2_empty, EPI_WARPS);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) ptx::tmem_alloc<512>(&bars->tmem_base);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer, GEMM1 operands
    // (its own thread: with one producer for both rings the E_J tiles of the next GEMM1 could not be requested
    // until the last E^T tile of the current GEMM2 had a free slot, and every GEMM1 started on a cold pipeline --
    // ncu, round 1: tensor pipe 42 % active, the epilogue warps spinning on a_done)
    if (lane == 0) {
      int gs = 0;
      uint32_t gph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        for (int J = it.j0; J < it.j1; ++J) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(&bars->g1_empty[gs], gph ^ 1);
            uint8_t* st = g1 + gs * G1_STAGE_BYTES;
            ptx::mbar_arrive_expect_tx(&bars->g1_full[gs], G1_STAGE_BYTES);
            // 128-byte k-blocks: 64 BF16 (contrastive) or 32 TF32 (batch-all) elements
            constexpr int kBk1 = kMode == 1 ? tc::BK16 : BK;
            ptx::tma_load_2d(&tm_hi, &bars->g1_full[gs], st + 0 * TILE_BYTES, kb * kBk1, it.ti * BM);
            ptx::tma_load_2d(&tm_lo, &bars->g1_full[gs], st + 1 * TILE_BYTES, kb * kBk1, it.ti * BM);
            ptx::tma_load_2d(&tm_hi, &bars->g1_full[gs], st + 2 * TILE_BYTES, kb * kBk1, J * BN);
            ptx::tma_load_2d(&tm_lo, &bars->g1_full[gs], st + 3 * TILE_BYTES, kb * kBk1, J * BN);
            if (++gs == G1_STAGES) { gs = 0; gph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ TMA producer, E^T tiles of GEMM2
    // per 32-row k-block and 128-column sub-slice; phase A needs both planes, phase B E_hi only
    if (lane == 0) {
      int es = 0;
      uint32_t eph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        for (int J = it.j0; J < it.j1; ++J) {
          for (int phase = 0; phase < 2; ++phase) {
            const bool with_lo = phase == 0;
            for (int kb2 = 0; kb2 < BN / BK; ++kb2) {
              for (int sub = 0; sub < NSUB; ++sub) {
                ptx::mbar_wait(&bars->et_empty[es], eph ^ 1);
                uint8_t* st = et + es * ET_STAGE_BYTES;
                ptx::mbar_arrive_expect_tx(&bars->et_full[es], with_lo ? ET_STAGE_BYTES : TILE_BYTES);
                const int rowc = (it.wide * NSUB + sub) * DN;
                ptx::tma_load_2d(&tm_et_hi, &bars->et_full[es], st, J * BN + kb2 * BK, rowc);
                if (with_lo) ptx::tma_load_2d(&tm_et_lo, &bars->et_full[es], st + TILE_BYTES, J * BN + kb2 * BK, rowc);
                if (++es == ET_STAGES) { es = 0; eph ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(BM, BN);
      constexpr uint32_t idesc16 = ptx::make_idesc_bf16(BM, BN);
      int gs = 0, es = 0;
      uint32_t gph = 0, eph = 0, acc1_it = 0, c_it = 0, item_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
        const Item it = decode_item(p, item);
        auto gemm1 = [&]() {
          ptx::mbar_wait(&bars->acc1_empty, (acc1_it & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tm = tmem + TM_ACC1;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(&bars->g1_full[gs], gph);
            ptx::tc_fence_after();
            const uint32_t st = ptx::smem_u32(g1 + gs * G1_STAGE_BYTES);
            const uint64_t a_hi = ptx::make_kmajor_sw128_desc(st), a_lo = ptx::make_kmajor_sw128_desc(st + TILE_BYTES);
            const uint64_t b_hi = ptx::make_kmajor_sw128_desc(st + 2 * TILE_BYTES),
                           b_lo = ptx::make_kmajor_sw128_desc(st + 3 * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = static_cast<uint64_t>(k * UMMA_K * 4 / 16);
              // one accumulator for all three products here: the backward only needs S to decide hinge activity
              // and 1/D factors, where the ~6e-6 accumulation bias is far inside the gradient tolerance
              if (kMode == 1) {
                ptx::mma_bf16_ss(d_tm, a_lo + koff, b_hi + koff, idesc16, (kb | k) != 0);
                ptx::mma_bf16_ss(d_tm, a_hi + koff, b_lo + koff, idesc16, 1);
                ptx::mma_bf16_ss(d_tm, a_hi + koff, b_hi + koff, idesc16, 1);
              } else {
                ptx::mma_tf32_ss(d_tm, a_lo + koff, b_hi + koff, idesc, (kb | k) != 0);
                ptx::mma_tf32_ss(d_tm, a_hi + koff, b_lo + koff, idesc, 1);
                ptx::mma_tf32_ss(d_tm, a_hi + koff, b_hi + koff, idesc, 1);
              }
            }
            ptx::mma_commit(&bars->g1_empty[gs]);
            if (++gs == G1_STAGES) { gs = 0; gph ^= 1; }
          }
          ptx::mma_commit(&bars->acc1_full);
          ++acc1_it;
        };
        ptx::mbar_wait(&bars->acc2_empty, (item_it & 1) ^ 1);
        ptx::tc_fence_after();
        gemm1();
        for (int J = it.j0; J < it.j1; ++J) {
          if (J + 1 < it.j1) gemm1();
          // ---- phase A: C_hi . (E_hi + E_lo)
          ptx::mbar_wait(&bars->c_full, c_it & 1);
          ++c_it;
          ptx::tc_fence_after();
          for (int kb2 = 0; kb2 < BN / BK; ++kb2) {
            for (int sub = 0; sub < NSUB; ++sub) {
              ptx::mbar_wait(&bars->et_full[es], eph);
              ptx::tc_fence_after();
              const uint32_t st = ptx::smem_u32(et + es * ET_STAGE_BYTES);
              const uint64_t e_hi = ptx::make_kmajor_sw128_desc(st), e_lo = ptx::make_kmajor_sw128_desc(st + TILE_BYTES);
              const uint32_t acc = tmem + TM_ACC2 + sub * DN;
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t koff = static_cast<uint64_t>(k * UMMA_K * 4 / 16);
                const uint32_t a_c = tmem + TM_C + kb2 * BK + k * UMMA_K;
                ptx::mma_tf32_ts(acc, a_c, e_lo + koff, idesc, (J != it.j0) | kb2 | k);
                ptx::mma_tf32_ts(acc, a_c, e_hi + koff, idesc, 1);
              }
              ptx::mma_commit(&bars->et_empty[es]);
              if (++es == ET_STAGES) { es = 0; eph ^= 1; }
            }
          }
          ptx::mma_commit(&bars->a_done);
          // ---- phase B: C_lo . E_hi
          ptx::mbar_wait(&bars->c_full, c_it & 1);
          ++c_it;
          ptx::tc_fence_after();
          for (int kb2 = 0; kb2 < BN / BK; ++kb2) {
            for (int sub = 0; sub < NSUB; ++sub) {
              ptx::mbar_wait(&bars->et_full[es], eph);
              ptx::tc_fence_after();
              const uint32_t st = ptx::smem_u32(et + es * ET_STAGE_BYTES);
              const uint64_t e_hi = ptx::make_kmajor_sw128_desc(st);
              const uint32_t acc = tmem + TM_ACC2 + sub * DN;
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t koff = static_cast<uint64_t>(k * UMMA_K * 4 / 16);
                ptx::mma_tf32_ts(acc, tmem + TM_C + kb2 * BK + k * UMMA_K, e_hi + koff, idesc, 1);
              }
              ptx::mma_commit(&bars->et_empty[es]);
              if (++es == ET_STAGES) { es = 0; eph ^= 1; }
            }
          }
          ptx::mma_commit(&bars->c_empty);
        }
        ptx::mma_commit(&bars->acc2_full);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps: build C, then write the slice
    const int quarter = warp & 3, half = (warp - CTRL_WARPS) >> 2;  // TMEM lane quarter = warp id % 4
    uint8_t* ws = warp_scr + (warp - CTRL_WARPS) * WARP_SCR;
    float* wf = reinterpret_cast<float*>(ws);
    int32_t* wi = reinterpret_cast<int32_t*>(ws + 256);
    float* wpos = reinterpret_cast<float*>(ws + 512);  // [64 columns][MAXP], margin added, -inf padded
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t e_it = 0, item_it = 0;
    const float gl = p.gloss ? p.gloss[0] : 1.f;
    const float inv_np = kMode == 0 ? static_cast<float>(1.0 / (p.stats[1] + 1e-16)) : 0.f;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
      const Item it = decode_item(p, item);
      const int64_t row = static_cast<int64_t>(it.ti) * BM + quarter * 32 + lane;
      const bool row_ok = row < p.B;
      const int32_t la = row_ok ? p.labels[row] : -1;
      const float na = row_ok ? p.norms[row] : 0.f;
      float pi[MAXP];
      int cnt_s[MAXP];
      int npi = 0;
      if (kMode == 0) {
        npi = row_ok ? p.pos_n[row] : 0;
#pragma unroll
        for (int s = 0; s < MAXP; ++s) {
          pi[s] = (s < npi) ? p.pos_d[row * MAXP + s] + p.margin : -INFINITY;
          cnt_s[s] = 0;
        }
      }
      double rowsum = 0.0;
      for (int J = it.j0; J < it.j1; ++J, ++e_it) {
        ptx::mbar_wait(&bars->acc1_full, e_it & 1);
        ptx::tc_fence_after();
        // pull this thread's 64 columns of S into registers and hand the accumulator straight back to the MMA warp
        float sv[2][32];
        ptx::tmem_ld_32x32(tmem + lane_base + TM_ACC1 + (half * 2 + 0) * 32, sv[0]);
        ptx::tmem_ld_32x32(tmem + lane_base + TM_ACC1 + (half * 2 + 1) * 32, sv[1]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->acc1_empty);
        // stage this warp's 64 columns' norms / labels / positives lists (one round of global loads per tile)
        __syncwarp();
#pragma unroll
        for (int cc2 = 0; cc2 < 2; ++cc2) {
          const int64_t cc = static_cast<int64_t>(J) * BN + (half * 2 + cc2) * 32 + lane;
          const bool ok = cc < p.B;
          wf[cc2 * 32 + lane] = ok ? __ldg(&p.norms[cc]) : 0.f;
          wi[cc2 * 32 + lane] = ok ? __ldg(&p.labels[cc]) : -2;
          if (kMode == 0) {
            const int npk = ok ? p.pos_n[cc] : 0;
#pragma unroll
            for (int s = 0; s < MAXP; ++s)
              wpos[(cc2 * 32 + lane) * MAXP + s] = (s < npk) ? p.pos_d[cc * MAXP + s] + p.margin : -INFINITY;
          }
        }
        __syncwarp();
        // the C region is free once phase B of the previous tile has retired
        ptx::mbar_wait(&bars->c_empty, (e_it & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int cc2 = 0; cc2 < 2; ++cc2) {
          const int c = half * 2 + cc2;
          const int64_t col0 = static_cast<int64_t>(J) * BN + c * 32;
          float (&w)[32] = sv[cc2];
          const float* wfc = wf + cc2 * 32;
          const int32_t* wic = wi + cc2 * 32;
          const float* wposc = wpos + cc2 * 32 * MAXP;
          float chunk_sum = 0.f;
          // interior chunks (no ragged edge, no diagonal) skip the per-element index checks
          const bool interior = row_ok && (col0 + 32 <= p.B) && (col0 != row - lane);
          const uint32_t c_addr = tmem + lane_base + TM_C + c * 32;
          // Eight elements per trip of a ROLLED loop: the trip consumes w[0..7], sends the high parts to TMEM, and
          // rotates the register array so that after four trips w holds the 32 low parts in order.  Fully unrolled,
          // the coefficient code of one tile was ~60 KB and the kernel stalled on instruction fetch (ncu, round 1:
          // stall_no_instruction 5.4 per issue, the top stall).
#pragma unroll 1
          for (int jj = 0; jj < 32; jj += 8) {
            float h8[8], l8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = jj + u;
              const bool ok = interior || (row_ok && col0 + j < p.B && col0 + j != row);
              const float d2 = fmaxf(na + wfc[j] - 2.f * w[u], 0.f);
              // 1/sqrt via the SFU (relative error ~1e-7): 1/d and d = d2/d come from one MUFU instead of an IEEE
              // sqrt plus an IEEE divide per element; the .ftz form skips the denormal rescaling code that
              // -ftz=false otherwise wraps around every MUFU (squared distances below 1e-30 count as zero)
              const float rs = d2 > 1e-30f ? rsqrt_ftz(d2) : 0.f;
              float cv = 0.f;
              if (kMode == 1) {
                // t'(d2) = 1 (same label) or -max(1 - d, 0) / d = -max(1/d - 1, 0); clamp region d2 < 1e-7: zero slope
                const float diff = -fmaxf(rs - 1.f, 0.f);
                cv = (ok && d2 >= 1e-7f) ? 4.f * p.scale_c * (wic[j] == la ? 1.f : diff) : 0.f;
              } else {
                const bool isneg = ok && wic[j] != la;
                const float dn = p.squared ? d2 : d2 * rs;
                int cnt = 0;
#pragma unroll
                for (int s = 0; s < MAXP; ++s) {
                  const int act = (isneg && (pi[s] - dn) > 1e-16f) ? 1 : 0;
                  cnt += act;
                  cnt_s[s] += act;
                }
                const float4 q0 = *reinterpret_cast<const float4*>(wposc + j * MAXP);
                const float4 q1 = *reinterpret_cast<const float4*>(wposc + j * MAXP + 4);
                cnt += (q0.x - dn > 1e-16f) + (q0.y - dn > 1e-16f) + (q0.z - dn > 1e-16f) + (q0.w - dn > 1e-16f) +
                       (q1.x - dn > 1e-16f) + (q1.y - dn > 1e-16f) + (q1.z - dn > 1e-16f) + (q1.w - dn > 1e-16f);
                const float sfac = p.squared ? 2.f : rs;
                cv = isneg ? -static_cast<float>(cnt) * inv_np * sfac : 0.f;
              }
              const float h = tf32_round(cv);
              const float l = tf32_round(cv - h);
              chunk_sum += h + l;   // exactly what the MMAs will see
              h8[u] = h;
              l8[u] = l;
            }
            ptx::tmem_st_32x8(c_addr + jj, h8);
#pragma unroll
            for (int i = 0; i < 24; ++i) w[i] = w[i + 8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[24 + u] = l8[u];
          }
          rowsum += static_cast<double>(chunk_sum);
        }
        // ---- the high part is in the C region
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->c_full);
        // ---- low part over it, as soon as phase A has read the high part
        ptx::mbar_wait(&bars->a_done, e_it & 1);
        ptx::tc_fence_after();
        ptx::tmem_st_32x32(tmem + lane_base + TM_C + (half * 2 + 0) * 32, sv[0]);  // sv now holds the low parts
        ptx::tmem_st_32x32(tmem + lane_base + TM_C + (half * 2 + 1) * 32, sv[1]);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->c_full);
      }
      // ---- this item's share of the gradient: gl * (rowsum_i * (e_i - mu) - (C.E)_i) over its J range, added into
      // the zeroed gemb (with two J ranges per row the two additions commute: the result stays deterministic)
      rowsum_x[half * BM + quarter * 32 + lane] = static_cast<float>(rowsum);
      ptx::named_bar_sync(1, EPI_WARPS * 32);
      const float rs = rowsum_x[quarter * 32 + lane] + rowsum_x[BM + quarter * 32 + lane];
      ptx::mbar_wait(&bars->acc2_full, item_it & 1);
      ptx::tc_fence_after();
      for (int c = half * (DW / 64); c < (half + 1) * (DW / 64); ++c) {  // this thread's 128 of the 256 columns
        float v[32];
        ptx::tmem_ld_32x32(tmem + lane_base + TM_ACC2 + c * 32, v);
        ptx::tmem_ld_wait();
        if (row_ok) {
          const int col0 = it.wide * DW + c * 32;
          const float* er = p.emb + row * p.d;
          float* gr = p.gemb + row * p.d;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.d) {
              const float g = gl * (rs * (er[col0 + j] - __ldg(&p.mu[col0 + j])) - v[j]);
              if (p.n_jparts == 1) gr[col0 + j] = g;
              else atomicAdd(&gr[col0 + j], g);
            }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->acc2_empty);
      ptx::named_bar_sync(1, EPI_WARPS * 32);  // rowsum_x may be rewritten by the next item
      if (kMode == 0 && it.wide == 0 && row_ok) {
#pragma unroll
        for (int s = 0; s < MAXP; ++s)
          if (s < npi && cnt_s[s] != 0) atomicAdd(&p.pos_cnt[row * MAXP + s], cnt_s[s]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

// E (B x d) -> E^T planes: ET_hi/lo [rows_t = n_slices*128][bpad], zero padded, TF32 split (K-major B operand of GEMM2)
__global__ void transpose_split_kernel(const float* __restrict__ e, const float* __restrict__ mu, int64_t B, int d,
                                       int rows_t, int64_t bpad, float* __restrict__ et_hi,
                                       float* __restrict__ et_lo) {
  __shared__ float tile[32][33];
  const int64_t j0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t j = j0 + r;
    const int c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (j < B && c < d) ? e[j * d + c] - mu[c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const int64_t j = j0 + threadIdx.x;
    if (c < rows_t && j < bpad) {
      const float v = tile[threadIdx.x][r];
      const float h = tc::to_tf32(v);
      et_hi[static_cast<int64_t>(c) * bpad + j] = h;
      et_lo[static_cast<int64_t>(c) * bpad + j] = tc::to_tf32(v - h);
    }
  }
}

}  // namespace pbt

// ---------------------------------------------------------------------------------------------- host entry
size_t pair_bwd_tc_ws_bytes(int64_t B, int d) {
  const size_t dpad = static_cast<size_t>((d + tc::BK - 1) / tc::BK * tc::BK);
  const size_t n_slices = static_cast<size_t>((d + pbt::DW - 1) / pbt::DW) * pbt::NSUB;
  const size_t bpad = static_cast<size_t>((B + 31) / 32 * 32);
  return 2 * align_up(static_cast<size_t>(B) * dpad * 4) + align_up(static_cast<size_t>(B) * 4) +
         2 * align_up(n_slices * pbt::DN * bpad * 4) + align_up(static_cast<size_t>(d) * 4);
}

// mode 0 = batch-all (pos_* describe lists with capacity 8), mode 1 = contrastive.  gemb is fully overwritten
// (the caller adds the sparse positive-pair terms of batch-all afterwards).
int pair_bwd_tc_launch(const float* emb, const int32_t* labels, int64_t B, int d, int mode, int squared, float margin,
                       float scale_c, const float* pos_d, const int32_t* pos_n, int32_t* pos_cnt, const double* stats,
                       const float* gloss, float* gemb, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < pair_bwd_tc_ws_bytes(B, d)) return fail(EN_ERR_WORKSPACE, "pair backward: workspace too small");
  Workspace w(ws, ws_bytes);
  const int g1_bf16 = mode == 1;        // contrastive: BF16 GEMM1 planes (they fit in the fp32-sized buffers below)
  const int dpad = tc::dpad_for(d, g1_bf16);
  const int dpad32 = tc::dpad_for(d, 0);
  const int n_wide = (d + pbt::DW - 1) / pbt::DW;  // 256-column groups of the gradient
  const int n_slices = n_wide * pbt::NSUB;
  const int rows_t = n_slices * pbt::DN;
  const int64_t bpad = (B + 31) / 32 * 32;
  float* hi = w.take<float>(static_cast<size_t>(B) * dpad32);
  float* lo = w.take<float>(static_cast<size_t>(B) * dpad32);
  float* norms = w.take<float>(B);
  float* et_hi = w.take<float>(static_cast<size_t>(rows_t) * bpad);
  float* et_lo = w.take<float>(static_cast<size_t>(rows_t) * bpad);
  float* mu = w.take<float>(d);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "pair backward: workspace too small or misaligned");
  dim3 tb(32, 8), tg(static_cast<unsigned>(bpad / 32), static_cast<unsigned>(rows_t / 32));
  tc::launch_column_mean(emb, B, d, mu, st);
  EN_LAUNCHED("column_mean_kernel");
  // GEMM1 also runs on the centred rows (norms are the centred norms): ||a-b|| is unchanged, S loses its
  // one-sided truncation bias, and with it the hinge-activity flips against the float64 oracle
  if (g1_bf16) EN_CUDA(tc::launch_split_bf16(emb, B, d, d, dpad, hi, lo, norms, st, nullptr, nullptr, mu));
  else EN_CUDA(tc::launch_split(emb, B, d, d, dpad, hi, lo, norms, st, mu));
  ++launch_counter();
  pbt::transpose_split_kernel<<<tg, tb, 0, st>>>(emb, mu, B, d, rows_t, bpad, et_hi, et_lo);
  EN_LAUNCHED("transpose_split_kernel");
  CUtensorMap th, tl, teh, tel;
  if ((g1_bf16 ? (tc::make_plane_tmap_bf16(&th, hi, B, dpad) || tc::make_plane_tmap_bf16(&tl, lo, B, dpad))
               : (tc::make_plane_tmap(&th, hi, B, dpad) || tc::make_plane_tmap(&tl, lo, B, dpad))) ||
      tc::make_plane_tmap(&teh, et_hi, rows_t, bpad) || tc::make_plane_tmap(&tel, et_lo, rows_t, bpad))
    return fail(EN_ERR_DRIVER, "pair backward: cuTensorMapEncodeTiled failed");
  pbt::Params p;
  p.emb = emb; p.labels = labels; p.norms = norms; p.pos_d = pos_d; p.pos_n = pos_n; p.pos_cnt = pos_cnt;
  p.stats = stats; p.gloss = gloss; p.mu = mu; p.gemb = gemb; p.B = B; p.d = d;
  p.tiles = static_cast<int>((B + tc::BM - 1) / tc::BM);
  const int sms = device_sm_count();
  // column-tile ranges per (row tile, column group): as many as it takes to give every SM an item (B = 4096,
  // d = 512: 32 x 2 x 2 = 128 items); each range re-reads nothing, the partial gradients are summed in gemb
  int jparts = sms / (p.tiles * n_wide);
  if (jparts < 1) jparts = 1;
  if (jparts > p.tiles) jparts = p.tiles;
  p.tiles_per_part = (p.tiles + jparts - 1) / jparts;
  p.n_jparts = (p.tiles + p.tiles_per_part - 1) / p.tiles_per_part;
  p.n_wide = n_wide; p.kblocks = dpad / (g1_bf16 ? tc::BK16 : tc::BK); p.mode = mode; p.squared = squared; p.margin = margin;
  p.scale_c = scale_c;
  if (mode == 0)
    EN_CUDA(cudaFuncSetAttribute(pbt::pair_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, pbt::SMEM_BYTES));
  else
    EN_CUDA(cudaFuncSetAttribute(pbt::pair_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pbt::SMEM_BYTES));
  const int items = p.tiles * p.n_wide * p.n_jparts;
  const int grid = items < sms ? items : sms;
  if (p.n_jparts > 1) EN_CUDA(cudaMemsetAsync(gemb, 0, static_cast<size_t>(B) * d * sizeof(float), st));
  prof_begin(st);
  if (mode == 0) pbt::pair_bwd_tc_kernel<0><<<grid, pbt::NUM_THREADS, pbt::SMEM_BYTES, st>>>(th, tl, teh, tel, p);
  else pbt::pair_bwd_tc_kernel<1><<<grid, pbt::NUM_THREADS, pbt::SMEM_BYTES, st>>>(th, tl, teh, tel, p);
  prof_end(st);
  EN_LAUNCHED("pair_bwd_tc_kernel");
  return EN_OK;
}

}  // namespace en
