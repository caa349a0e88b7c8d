// Fused tensor-core kernel of the pair losses (batch-all triplet, all-pairs contrastive): loss AND gradient from one
// pass over the B x B distance tiles, nothing of size B x B ever stored.
//
//     d L / d E  =  rowsum(C) o E  -  C . E ,       C_ik = symmetrised pair coefficient, a function of D_ik
//
// Two chained tcgen05 GEMMs per 128 x 128 tile:
//   GEMM1  S = E_I . E_J^T             on mean-centred planes, operands by TMA, accumulator in TMEM.  Batch-all: 3xTF32
//                                     (kind::tf32) -- S decides the hinges D_ap + m - D_an > 0, and BF16 planes
//                                     pushed 8.7 % of the rows over the gradient tolerance through flipped
//                                     decisions (TF32 planes < 8 %).  Contrastive: split-BF16 planes at twice the
//                                     rate -- S only feeds smooth terms there.
//   epilogue  S -> registers (accumulator released at once) -> per element: distance, loss term (kLoss), coefficient
//             C_IJ = f(S, labels, positives lists) -> split into BF16 hi + lo, packed two per 32-bit column and written
//             BACK to tensor memory with tcgen05.st: 64 columns per plane, both planes resident at once
//   GEMM2  G += C_IJ . E_J[:, 256-column slice]   kind::f16, A operand from TMEM (tcgen05.mma [d],[a],b), B operand =
//             BF16 planes of E^T by TMA; C_hi.E_hi + C_hi.E_lo + C_lo.E_hi (split-BF16: relative error ~4e-6 of
//             sum |c||e|, far inside the 1e-4 gradient tolerance; the rowsum term uses the same rounded values)
// A CTA owns (row tile I, 256 gradient columns, a range of column tiles J); GEMM1 of tile J+1 is issued before GEMM2 of
// tile J so the tensor pipe works while the epilogue warps build C_J.
//
// Round-2 redesign (ncu r1: pair_bwd_tc_kernel tensor pipe 38-42 % active, the MMA thread spinning
// 298 times per c_full wait): the TF32 coefficient tile needed 128 columns per plane and went through ONE region
// twice per tile (hi -> phase A -> lo -> phase B, four hand-offs).  Packed BF16 planes halve GEMM2's tensor time,
// remove two of the four hand-offs and leave room for S | C_hi C_lo | G(256) in the 512 columns.  The forward loss is
// accumulated by the same epilogue (each ordered pair (i, k) is visited exactly once), so a training step computes the
// distance tiles once instead of twice, and the positives lists / centred planes are built once.
//
// Cancellation: grad_i = sum_k c_ik (e_i - e_k) is computed as rowsum_i e_i - (C.E)_i.  For post-ReLU embeddings the
// two terms are ~10x larger than their difference; the expression is invariant under E -> E - mu, so GEMM2 and the
// rowsum term both use the mean-centred embeddings (mu = column mean), which removes the cancellation.
//
// Reference semantics: contrastive = embedding_net/losses_and_accuracies.py:4-11 over all pairs with the Siamese clamp
// of embedding_net/models.py:225; batch-all = Moindrot's batch_all_triplet_loss (README.md:116), not in the reference.
#include "common.cuh"
#include "tc_engine.cuh"

namespace en {
namespace ptc {

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
constexpr int ET_STAGES = 2;            // GEMM2 B ring: ET_hi | ET_lo, each [128 gradient columns x 64 rows j] BF16
constexpr int ET_STAGE_BYTES = 2 * TILE_BYTES;
constexpr int JB = 64;                  // rows j per GEMM2 k-block (one 128-byte swizzle row of BF16)
// Epilogue: 16 warps = 4 per SM sub-partition (TMEM lane quarter = warp id % 4), each owning ONE 32-column chunk of
// its 32 rows.  ncu r2 (8 warps x 64 columns): 1.7 warp instructions per cycle per SM, i.e. 0.43 per scheduler with
// two warps each -- the per-element chains (FADD -> FSET -> FADD, MUFU) were latency-, not issue-bound, and the
// epilogue, which sits between GEMM1 and GEMM2 of a tile, set the tile period.
constexpr int EPI_Q = 4;                // column chunks (32 wide) per tile row = epilogue warps per lane quarter
constexpr int EPI_WARPS = 4 * EPI_Q;
constexpr int CTRL_WARPS = 3;            // warp 0: GEMM1 operand TMA, warp 1: MMA issuer, warp 2: E^T (GEMM2) TMA
constexpr int NUM_THREADS = 32 * (CTRL_WARPS + EPI_WARPS);
constexpr int MAXP = 8;                 // positives per anchor held in registers per pass over a tile
constexpr int MAXP_BIG = 64;            // kBig: sorted lists of up to 64 slots, binary-searched per element
constexpr int WARP_SCR = 128 + 128 + 32 * MAXP * 4;  // per warp, for its 32 columns: norms | labels | positives lists
constexpr int SMEM_BYTES = G1_STAGES * G1_STAGE_BYTES + ET_STAGES * ET_STAGE_BYTES + 256 + EPI_WARPS * WARP_SCR +
                           BM * EPI_Q * 4;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
// TMEM (512 columns): S | C_hi | C_lo | G
constexpr uint32_t TM_S = 0;       // 128 columns (single buffer: the epilogue copies it to registers and releases it)
constexpr uint32_t TM_CH = 128;    // 64 columns: BF16 high parts of the coefficient tile, two per column
constexpr uint32_t TM_CL = 192;    // 64 columns: BF16 low parts
constexpr uint32_t TM_G = 256;     // 256 columns: the gradient accumulator (NSUB x DN)

struct Bars {
  uint64_t g1_full[G1_STAGES], g1_empty[G1_STAGES];
  uint64_t et_full[ET_STAGES], et_empty[ET_STAGES];
  uint64_t s_full, s_empty;
  uint64_t c_full;     // epilogue -> MMA: C_hi and C_lo of this tile are in TMEM
  uint64_t c_empty;    // MMA -> epilogue: GEMM2 has consumed them
  uint64_t g_full, g_empty;
  uint32_t tmem_base;
};

struct Params {
  const int32_t* labels;
  const float* norms;
  const float* pos_d;     // [B rounded up to whole tiles][cap] (batch-all), -inf past each list's end
  const double* pos_pre;  // kBig: [..][cap] float64 prefix sums of the lists, which are sorted by decreasing distance
  const int32_t* pos_n;   // [B]
  int32_t* pos_cnt;       // [B][cap] out: active negatives per (anchor, positive slot)
  int cap;                // list capacity: MAXP, or a multiple of 8 up to MAXP_BIG (kBig)
  float* gemb;            // out: -(C.E) over the centred rows, unscaled by gloss; ZEROED by the caller when n_jparts > 1
                          // (items that share rows add into it: two addends per element commute, still deterministic)
  float* rowsum;          // out [B], zeroed by the caller: sum_k C_ik (what the MMAs saw); pair_finish_kernel turns
                          // the two into the gradient with coalesced accesses
  PairPartial* partial;   // kLoss: [B][n_jparts][EPI_Q] loss partial sums (+ positive-term counts)
  int64_t B;
  int d, tiles, n_wide, n_jparts, tiles_per_part, kblocks;
  int squared;
  float margin;
  float coef_scale;       // multiplies every pair coefficient: contrastive 4 / (B (B-1)); batch-all 1
  const double* stats;    // optional (batch-all, forward already done): coefficients are also divided by stats[1] =
                          // #positive triplets; when null the caller rescales the finished gradient
};

__device__ __forceinline__ float rsqrt_ftz(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// {hi16 = bf16_rn(a), lo16 = bf16_rn(b)}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void tmem_st_32x4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}

struct Item {
  int ti, wide, part, j0, j1;
};
// work item -> (row tile, 256-column group, range of column tiles); row tile fastest so that concurrently running
// CTAs stream the same E_J / E^T tiles
__device__ __forceinline__ Item decode_item(const Params& p, int item) {
  Item it;
  it.ti = item % p.tiles;
  const int rest = item / p.tiles;
  it.wide = rest % p.n_wide;
  it.part = rest / p.n_wide;
  it.j0 = it.part * p.tiles_per_part;
  it.j1 = min(it.j0 + p.tiles_per_part, p.tiles);
  return it;
}

// kMode: 0 = batch-all, 1 = contrastive (template parameter so that each instantiation carries only its own
// coefficient code: both together overflowed the instruction cache, ncu r1: stall_no_instruction 5.4 / issue).
// kLoss: also accumulate the forward loss.  kG1Bf16: GEMM1 on BF16 planes (backward-only contrastive).
// kBig (batch-all only): classes with more than MAXP positives per anchor.  The positives lists (sorted by
// decreasing distance) are binary-searched per element; the per-(anchor, positive) counts live in a shared-memory
// histogram that takes the place of the second E^T stage (the epilogue is longer than GEMM2 here, so GEMM2's
// operand ring can be one deep).
template <int kMode, bool kLoss, bool kG1Bf16, bool kBig = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pair_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const __grid_constant__ CUtensorMap tm_et_hi, const __grid_constant__ CUtensorMap tm_et_lo,
               const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* g1 = smem;
  uint8_t* et = smem + G1_STAGES * G1_STAGE_BYTES;
  Bars* bars = reinterpret_cast<Bars*>(et + ET_STAGES * ET_STAGE_BYTES);
  uint8_t* warp_scr = reinterpret_cast<uint8_t*>(bars) + 256;
  float* rowsum_x = reinterpret_cast<float*>(warp_scr + EPI_WARPS * WARP_SCR);  // [EPI_Q][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.tiles * p.n_wide * p.n_jparts;
  static_assert(!kBig || kMode == 0, "kBig is a batch-all variant");
  constexpr int kEtStages = kBig ? 1 : ET_STAGES;
  // kBig shared-memory plan: the second E^T stage holds the row anchors' sorted lists, transposed [slot][row in tile]
  // (conflict-free for the per-lane binary search); the per-warp scratch shrinks to its norms / labels part (256 B
  // per warp: the column lists are read from global memory) and its remainder holds the histogram of prefix lengths,
  // [slot][row pair] with two 16-bit counters per word (a counter sees at most the item's columns: B <= 65535).
  float* row_lists = reinterpret_cast<float*>(et + ET_STAGE_BYTES);
  uint32_t* hist = reinterpret_cast<uint32_t*>(warp_scr + EPI_WARPS * 256);
  static_assert(MAXP_BIG * BM * 4 <= ET_STAGE_BYTES, "the row lists replace one E^T stage");
  static_assert(EPI_WARPS * 256 + MAXP_BIG * (BM / 2) * 4 <= EPI_WARPS * WARP_SCR, "histogram fits the freed scratch");

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_hi);
    ptx::prefetch_tmap(&tm_lo);
    ptx::prefetch_tmap(&tm_et_hi);
    ptx::prefetch_tmap(&tm_et_lo);
    for (int s = 0; s < G1_STAGES; ++s) { ptx::mbar_init(&bars->g1_full[s], 1); ptx::mbar_init(&bars->g1_empty[s], 1); }
    for (int s = 0; s < ET_STAGES; ++s) { ptx::mbar_init(&bars->et_full[s], 1); ptx::mbar_init(&bars->et_empty[s], 1); }
    ptx::mbar_init(&bars->s_full, 1);
    ptx::mbar_init(&bars->s_empty, EPI_WARPS);
    ptx::mbar_init(&bars->c_full, EPI_WARPS);
    ptx::mbar_init(&bars->c_empty, 1);
    ptx::mbar_init(&bars->g_full, 1);
    ptx::mbar_init(&bars->g_empty, EPI_WARPS);
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
    if (lane == 0) {
      int gs = 0;
      uint32_t gph = 0;
      constexpr int kBk1 = kG1Bf16 ? tc::BK16 : BK;  // 128-byte k-blocks: 64 BF16 or 32 TF32 elements
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        for (int J = it.j0; J < it.j1; ++J) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(&bars->g1_empty[gs], gph ^ 1);
            uint8_t* st = g1 + gs * G1_STAGE_BYTES;
            ptx::mbar_arrive_expect_tx(&bars->g1_full[gs], G1_STAGE_BYTES);
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
    // per 64-row k-block of the column tile and 128-column sub-slice of the gradient: both BF16 planes
    if (lane == 0) {
      int es = 0;
      uint32_t eph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        for (int J = it.j0; J < it.j1; ++J) {
          for (int kb2 = 0; kb2 < BN / JB; ++kb2) {
            for (int sub = 0; sub < NSUB; ++sub) {
              ptx::mbar_wait(&bars->et_empty[es], eph ^ 1);
              uint8_t* st = et + es * ET_STAGE_BYTES;
              ptx::mbar_arrive_expect_tx(&bars->et_full[es], ET_STAGE_BYTES);
              const int rowc = (it.wide * NSUB + sub) * DN;
              ptx::tma_load_2d(&tm_et_hi, &bars->et_full[es], st, J * BN + kb2 * JB, rowc);
              ptx::tma_load_2d(&tm_et_lo, &bars->et_full[es], st + TILE_BYTES, J * BN + kb2 * JB, rowc);
              if (++es == kEtStages) { es = 0; eph ^= 1; }
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
      uint32_t gph = 0, eph = 0, s_it = 0, c_it = 0, item_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
        const Item it = decode_item(p, item);
        auto gemm1 = [&]() {
          ptx::mbar_wait(&bars->s_empty, (s_it & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tm = tmem + TM_S;
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
              // one accumulator for all three products: the operands are mean-centred, so the dot products have mixed
              // signs and the accumulator's truncation does not build a one-sided bias (measured 1-3e-7 relative)
              if (kG1Bf16) {
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
          ptx::mma_commit(&bars->s_full);
          ++s_it;
        };
        ptx::mbar_wait(&bars->g_empty, (item_it & 1) ^ 1);
        ptx::tc_fence_after();
        gemm1();
        for (int J = it.j0; J < it.j1; ++J) {
          if (J + 1 < it.j1) gemm1();  // keeps the tensor pipe busy while the epilogue builds C_J
          ptx::mbar_wait(&bars->c_full, c_it & 1);
          ++c_it;
          ptx::tc_fence_after();
          for (int kb2 = 0; kb2 < BN / JB; ++kb2) {
            for (int sub = 0; sub < NSUB; ++sub) {
              ptx::mbar_wait(&bars->et_full[es], eph);
              ptx::tc_fence_after();
              const uint32_t st = ptx::smem_u32(et + es * ET_STAGE_BYTES);
              const uint64_t e_hi = ptx::make_kmajor_sw128_desc(st), e_lo = ptx::make_kmajor_sw128_desc(st + TILE_BYTES);
              const uint32_t acc = tmem + TM_G + sub * DN;
#pragma unroll
              for (int k = 0; k < JB / 16; ++k) {
                const uint64_t koff = static_cast<uint64_t>(k * 16 * 2 / 16);
                // 16 rows j of the coefficient tile = 8 packed columns of each plane
                const uint32_t a_h = tmem + TM_CH + kb2 * (JB / 2) + k * 8;
                const uint32_t a_l = tmem + TM_CL + kb2 * (JB / 2) + k * 8;
                ptx::mma_bf16_ts(acc, a_l, e_hi + koff, idesc16, (J != it.j0) | kb2 | k);
                ptx::mma_bf16_ts(acc, a_h, e_lo + koff, idesc16, 1);
                ptx::mma_bf16_ts(acc, a_h, e_hi + koff, idesc16, 1);
              }
              ptx::mma_commit(&bars->et_empty[es]);
              if (++es == kEtStages) { es = 0; eph ^= 1; }
            }
          }
          ptx::mma_commit(&bars->c_empty);
        }
        ptx::mma_commit(&bars->g_full);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps: loss, C, then the gradient slice
    const int quarter = warp & 3, cq = (warp - CTRL_WARPS) >> 2;  // TMEM lane quarter = warp id % 4; column chunk
    uint8_t* ws = warp_scr + (warp - CTRL_WARPS) * (kBig ? 256 : WARP_SCR);
    float* wf = reinterpret_cast<float*>(ws);
    int32_t* wi = reinterpret_cast<int32_t*>(ws + 128);
    float* wpos = reinterpret_cast<float*>(ws + 256);  // [32 columns][MAXP], margin added, -inf padded
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t e_it = 0, item_it = 0;
    const float cs = p.coef_scale * (p.stats ? static_cast<float>(1.0 / (p.stats[1] + 1e-16)) : 1.f);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
      const Item it = decode_item(p, item);
      const int64_t row = static_cast<int64_t>(it.ti) * BM + quarter * 32 + lane;
      const bool row_ok = row < p.B;
      const int32_t la = row_ok ? p.labels[row] : -1;
      const float na = row_ok ? p.norms[row] : 0.f;
      float pi[MAXP];
      float cnt_s[MAXP];  // counts as floats (exact far beyond the 4096 columns of an item): FSET + FADD per slot
      int npi = 0;
      unsigned long long np_big = 0;  // kBig: this thread's share of the positive-triplet count
      if (kMode == 0) {
        npi = row_ok ? p.pos_n[row] : 0;
        if (!kBig) {
#pragma unroll
          for (int s = 0; s < MAXP; ++s) {
            pi[s] = (s < npi) ? p.pos_d[row * MAXP + s] + p.margin : -INFINITY;
            cnt_s[s] = 0.f;
          }
        }
      }
      // loss terms and per-slot counts are reported by the first column group only (the tiles are visited once per
      // 256 gradient columns): the other groups skip that arithmetic
      const bool full = it.wide == 0;
      if (kBig) {
        // clear the histogram; the four warps that share a row (one per column chunk) stage its list, a quarter each
        for (int i = (warp - CTRL_WARPS) * 32 + lane; i < MAXP_BIG * (BM / 2); i += EPI_WARPS * 32) hist[i] = 0u;
        for (int s = cq; s < MAXP_BIG; s += EPI_Q)  // (+ margin: one add less per search step; -inf stays -inf)
          row_lists[s * BM + quarter * 32 + lane] = row_ok ? __ldg(p.pos_d + row * MAXP_BIG + s) + p.margin : -INFINITY;
        ptx::named_bar_sync(1, EPI_WARPS * 32);
      }
      double rowsum = 0.0, loss_sum = 0.0;
      for (int J = it.j0; J < it.j1; ++J, ++e_it) {
        ptx::mbar_wait(&bars->s_full, e_it & 1);
        ptx::tc_fence_after();
        // pull this thread's 32 columns of S into registers and hand the accumulator straight back to the MMA warp
        float w[32];
        ptx::tmem_ld_32x32(tmem + lane_base + TM_S + cq * 32, w);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->s_empty);
        // stage this warp's 32 columns' norms / labels / positives lists (one round of global loads per tile)
        __syncwarp();
        {
          const int64_t cc = static_cast<int64_t>(J) * BN + cq * 32 + lane;
          const bool ok = cc < p.B;
          wf[lane] = ok ? __ldg(&p.norms[cc]) : 0.f;
          wi[lane] = ok ? __ldg(&p.labels[cc]) : -2;
          if (kMode == 0 && !kBig) {
            const int npk = ok ? p.pos_n[cc] : 0;
#pragma unroll
            for (int s = 0; s < MAXP; ++s)
              wpos[lane * MAXP + s] = (s < npk) ? p.pos_d[cc * MAXP + s] + p.margin : -INFINITY;
          }
        }
        __syncwarp();
        if constexpr (kBig) {
          // ---- lists longer than eight (collect_positives_kernel sorted them by DECREASING distance and padded them
          // with -inf): for a negative at distance dn the active hinges D_ap + m - dn > 0 are a PREFIX of the
          // anchor's list, so a binary search (<= 6 steps over <= 64 slots) replaces one compare per slot -- once in
          // the row anchor's list (shared memory, [slot][row]: conflict free) and once in the column anchor's (global
          // memory: one list per warp and element).  (A two-level 4 x 16 search -- two dependent loads instead of
          // six, but 2.5x the instructions -- measured slower: 1.12 vs 1.01 ms per 64 x 64 step.)  The prefix length L feeds everything: the pair coefficient (L_row + L_col), the loss
          // (float64 prefix sum of the list: pre[L-1] + L (m - dn)), and the per-(anchor, positive) counts (slots
          // 0 .. L-1 each gain one: a histogram over L in shared memory, suffix-summed when the item ends).
          // Eight elements per trip of a rolled loop, as in the short-list path.
          const int64_t col0 = static_cast<int64_t>(J) * BN + cq * 32;
          const bool interior = row_ok && (col0 + 32 <= p.B) && (col0 != row - lane);
          constexpr int cap = MAXP_BIG;  // long lists always have 64 slots (-inf padded): constant strides, no bound checks
          const float* row_list = row_lists + quarter * 32 + lane;  // [slot][row in tile]: stride BM between slots
          const double* row_pre = p.pos_pre + (row_ok ? row : 0) * cap;
          float chunk_sum = 0.f;
          double chunk_loss = 0.0;
          unsigned np_tile = 0;
#pragma unroll 1
          for (int jj = 0; jj < 32; jj += 8) {
            const float* col_lists = p.pos_d + (col0 + jj) * cap;  // lists cover whole tiles (empty past B)
            float cv8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = jj + u;
              const bool ok = interior || (row_ok && col0 + j < p.B && col0 + j != row);
              const float d2 = fmaxf(na + wf[j] - 2.f * w[u], 0.f);
              const float rs = d2 > 1e-30f ? rsqrt_ftz(d2) : 0.f;
              const bool isneg = ok && wi[j] != la;
              const float dn = isneg ? (p.squared ? d2 : d2 * rs) : INFINITY;
              const float* cl = col_lists + u * cap;
              int lr = 0, lc = 0;  // prefix lengths: slots [0, lr) of the row anchor, [0, lc) of the column anchor
#pragma unroll
              for (int step = 32; step >= 1; step >>= 1) {  // at most 63 positives: slot 63 is always padding
                const float tr = row_list[(lr + step - 1) * BM];           // margin already added
                const float tcn = __ldg(cl + lc + step - 1) + p.margin;
                lr += (tr - dn > 1e-16f) ? step : 0;
                lc += (tcn - dn > 1e-16f) ? step : 0;
              }
              if (full && lr > 0) {
                atomicAdd(&hist[(lr - 1) * (BM / 2) + ((quarter * 32 + lane) >> 1)], (lane & 1) ? 65536u : 1u);
                np_tile += static_cast<unsigned>(lr);
                if (kLoss)
                  chunk_loss += __ldg(row_pre + lr - 1) +
                                static_cast<double>(lr) * (static_cast<double>(p.margin) - static_cast<double>(dn));
              }
              cv8[u] = -static_cast<float>(lr + lc) * cs * (p.squared ? 2.f : rs);  // 0 where the pair is no negative pair
            }
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              hw[u] = pack_bf16x2(cv8[2 * u + 1], cv8[2 * u]);
              const float h0 = __uint_as_float(hw[u] << 16), h1 = __uint_as_float(hw[u] & 0xFFFF0000u);
              lw[u] = pack_bf16x2(cv8[2 * u + 1] - h1, cv8[2 * u] - h0);
              const float l0 = __uint_as_float(lw[u] << 16), l1 = __uint_as_float(lw[u] & 0xFFFF0000u);
              chunk_sum += (h0 + l0) + (h1 + l1);
            }
#pragma unroll
            for (int i = 0; i < 24; ++i) w[i] = w[i + 8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              w[24 + u] = __uint_as_float(hw[u]);
              w[28 + u] = __uint_as_float(lw[u]);
            }
          }
          np_big += np_tile;
          rowsum += static_cast<double>(chunk_sum);
          if (kLoss) loss_sum += chunk_loss;
        } else {
          const int c = cq;
          const int64_t col0 = static_cast<int64_t>(J) * BN + c * 32;
          const float* wfc = wf;
          const int32_t* wic = wi;
          const float* wposc = wpos;
          float chunk_sum = 0.f, chunk_loss = 0.f;
          // interior chunks (no ragged edge, no diagonal) skip the per-element index checks
          const bool interior = row_ok && (col0 + 32 <= p.B) && (col0 != row - lane);
          // Eight elements per trip of a ROLLED loop: the trip consumes w[0..7], produces four packed high words and
          // four packed low words, and rotates the register array so that after four trips w holds, per trip, the
          // words {h0..h3, l0..l3} in order.  Fully unrolled, the coefficient code of one tile was ~60 KB and the
          // kernel stalled on instruction fetch (ncu r1: stall_no_instruction 5.4 per issue, the top stall).
#pragma unroll 1
          for (int jj = 0; jj < 32; jj += 8) {
            float cv8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = jj + u;
              const bool ok = interior || (row_ok && col0 + j < p.B && col0 + j != row);
              const float d2 = fmaxf(na + wfc[j] - 2.f * w[u], 0.f);
              float cv = 0.f;
              if (kMode == 1) {
                // Siamese clamp (models.py:225): d = sqrt(max(d2, 1e-7)); inside the clamp the slope is zero.
                // 1/sqrt via the SFU (relative error ~1e-7): 1/d and d = d2/d from one MUFU instead of an IEEE sqrt
                // plus an IEEE divide per element
                const float d2c = fmaxf(d2, 1e-7f);
                const float rs = rsqrt_ftz(d2c);
                const bool same = wic[j] == la;
                // t'(d2) = 1 (same label) or -max(1 - d, 0) / d = -max(1/d - 1, 0)
                const float diff = -fmaxf(rs - 1.f, 0.f);
                cv = (ok && d2 >= 1e-7f) ? cs * (same ? 1.f : diff) : 0.f;
                if (kLoss) {
                  const float m = fmaxf(1.f - d2c * rs, 0.f);
                  chunk_loss += ok ? (same ? d2c : m * m) : 0.f;
                }
              } else {
                // squared distances below 1e-30 count as zero (the .ftz MUFU skips the denormal rescaling code)
                const float rs = d2 > 1e-30f ? rsqrt_ftz(d2) : 0.f;
                const bool isneg = ok && wic[j] != la;
                // +inf where the column is not a negative of this anchor: every hinge below is then inactive without
                // a mask per slot (thresholds are finite or -inf, so no NaN arises)
                const float dn = isneg ? (p.squared ? d2 : d2 * rs) : INFINITY;
                float cnt = 0.f;
                if (full) {
#pragma unroll
                  for (int s = 0; s < MAXP; ++s) {
                    const float t = pi[s] - dn;       // D_ap + margin - D_an
                    const float act = t > 1e-16f ? 1.f : 0.f;
                    cnt += act;
                    cnt_s[s] += act;
                    if (kLoss) chunk_loss += fmaxf(t, 0.f);  // terms in (0, 1e-16] add < 1e-15 in total
                  }
                } else {
#pragma unroll
                  for (int s = 0; s < MAXP; ++s) cnt += (pi[s] - dn) > 1e-16f ? 1.f : 0.f;
                }
                const float4 q0 = *reinterpret_cast<const float4*>(wposc + j * MAXP);
                const float4 q1 = *reinterpret_cast<const float4*>(wposc + j * MAXP + 4);
                cnt += ((q0.x - dn > 1e-16f ? 1.f : 0.f) + (q0.y - dn > 1e-16f ? 1.f : 0.f)) +
                       ((q0.z - dn > 1e-16f ? 1.f : 0.f) + (q0.w - dn > 1e-16f ? 1.f : 0.f)) +
                       ((q1.x - dn > 1e-16f ? 1.f : 0.f) + (q1.y - dn > 1e-16f ? 1.f : 0.f)) +
                       ((q1.z - dn > 1e-16f ? 1.f : 0.f) + (q1.w - dn > 1e-16f ? 1.f : 0.f));
                const float sfac = p.squared ? 2.f : rs;
                cv = -cnt * cs * sfac;            // cnt = 0 where the pair is not a negative pair
              }
              cv8[u] = cv;
            }
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              // element 2u (even row j of the tile) in the low half, 2u+1 in the high half
              hw[u] = pack_bf16x2(cv8[2 * u + 1], cv8[2 * u]);
              const float h0 = __uint_as_float(hw[u] << 16), h1 = __uint_as_float(hw[u] & 0xFFFF0000u);
              lw[u] = pack_bf16x2(cv8[2 * u + 1] - h1, cv8[2 * u] - h0);
              const float l0 = __uint_as_float(lw[u] << 16), l1 = __uint_as_float(lw[u] & 0xFFFF0000u);
              chunk_sum += (h0 + l0) + (h1 + l1);   // exactly what the MMAs will see
            }
#pragma unroll
            for (int i = 0; i < 24; ++i) w[i] = w[i + 8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              w[24 + u] = __uint_as_float(hw[u]);
              w[28 + u] = __uint_as_float(lw[u]);
            }
          }
          rowsum += static_cast<double>(chunk_sum);
          if (kLoss) loss_sum += static_cast<double>(chunk_loss);
        }
        // the C region is free once GEMM2 of the previous tile has retired
        ptx::mbar_wait(&bars->c_empty, (e_it & 1) ^ 1);
        ptx::tc_fence_after();
        {
          const uint32_t* wu = reinterpret_cast<const uint32_t*>(w);
          // this thread's 32 tile columns cq*32 .. +31 = packed columns cq*16 .. +15
          const uint32_t pc0 = cq * 16;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            tmem_st_32x4(tmem + lane_base + TM_CH + pc0 + t * 4, wu + t * 8);
            tmem_st_32x4(tmem + lane_base + TM_CL + pc0 + t * 4, wu + t * 8 + 4);
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->c_full);
      }
      // ---- loss partial of this (row, J range, column chunk); only the first column group reports it
      if (kLoss && it.wide == 0 && row_ok) {
        unsigned long long np = 0;
        if (kMode == 0 && !kBig) {
#pragma unroll
          for (int s = 0; s < MAXP; ++s) np += static_cast<unsigned long long>(cnt_s[s]);  // exact integers
        }
        if (kBig) np = np_big;
        p.partial[(row * p.n_jparts + it.part) * EPI_Q + cq] = PairPartial{loss_sum, np};
      }
      // ---- this item's share of the gradient.  grad_i = rowsum_i (e_i - mu) - (C.E)_i: the kernel leaves the two
      // ingredients -- rowsum_i (first column group only) and -(C.E)_i over its J range, added into the zeroed gemb
      // with 128-bit reductions (two J ranges per row: the two additions commute, the result stays deterministic) --
      // and pair_finish_kernel combines them with coalesced row accesses.  Doing it here, one float per lane and row,
      // made every load / atomic of a warp touch 32 different rows: ncu r2, 44 % of the kernel's stall samples.
      rowsum_x[cq * BM + quarter * 32 + lane] = static_cast<float>(rowsum);
      ptx::named_bar_sync(1, EPI_WARPS * 32);
      if (kBig && it.wide == 0 && row_ok && cq == 0) {
        // every warp's shared-memory adds of this item are done.  hist[L-1] counted the negatives whose active
        // prefix has length L: positive s is active for every L > s, i.e. its count is the suffix sum from s on.
        unsigned run = 0;
        for (int s = npi - 1; s >= 0; --s) {
          const uint32_t two = hist[s * (BM / 2) + ((quarter * 32 + lane) >> 1)];
          run += (lane & 1) ? (two >> 16) : (two & 0xFFFFu);
          if (run != 0u) atomicAdd(&p.pos_cnt[row * p.cap + s], static_cast<int>(run));
        }
      }
      if (it.wide == 0 && cq == 0 && row_ok) {
        float rs = 0.f;
#pragma unroll
        for (int q = 0; q < EPI_Q; ++q) rs += rowsum_x[q * BM + quarter * 32 + lane];  // fixed order: deterministic
        atomicAdd(&p.rowsum[row], rs);
      }
      ptx::mbar_wait(&bars->g_full, item_it & 1);
      ptx::tc_fence_after();
      const bool vec = (p.d & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gemb) & 15) == 0;
      for (int c = cq * (DW / 32 / EPI_Q); c < (cq + 1) * (DW / 32 / EPI_Q); ++c) {  // this thread's 64 of the 256 columns
        float v[32];
        ptx::tmem_ld_32x32(tmem + lane_base + TM_G + c * 32, v);
        ptx::tmem_ld_wait();
        if (row_ok) {
          const int col0 = it.wide * DW + c * 32;
          float* gr = p.gemb + row * p.d + col0;
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (col0 + j < p.d) {
                if (p.n_jparts == 1)
                  *reinterpret_cast<float4*>(gr + j) = make_float4(-v[j], -v[j + 1], -v[j + 2], -v[j + 3]);
                else
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gr + j), "f"(-v[j]), "f"(-v[j + 1]),
                               "f"(-v[j + 2]), "f"(-v[j + 3])
                               : "memory");
              }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.d) {
                if (p.n_jparts == 1) gr[j] = -v[j];
                else atomicAdd(&gr[j], -v[j]);
              }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->g_empty);
      ptx::named_bar_sync(1, EPI_WARPS * 32);  // rowsum_x may be rewritten by the next item
      if (kMode == 0 && !kBig && it.wide == 0 && row_ok) {
#pragma unroll
        for (int s = 0; s < MAXP; ++s)
          if (s < npi && cnt_s[s] != 0.f) atomicAdd(&p.pos_cnt[row * MAXP + s], static_cast<int>(cnt_s[s]));
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

// E (B x d) -> E^T BF16 planes: ET_hi/lo [rows_t = n_wide*256][bpad], zero padded, centred (K-major B operand of GEMM2)
__global__ void transpose_split_bf16_kernel(const float* __restrict__ e, const float* __restrict__ mu, int64_t B, int d,
                                            int rows_t, int64_t bpad, uint16_t* __restrict__ et_hi,
                                            uint16_t* __restrict__ et_lo) {
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
      const uint16_t h = tc::to_bf16_bits(v);
      et_hi[static_cast<int64_t>(c) * bpad + j] = h;
      et_lo[static_cast<int64_t>(c) * bpad + j] = tc::to_bf16_bits(v - __uint_as_float(static_cast<uint32_t>(h) << 16));
    }
  }
}

// Closes the step, one warp per row, every access coalesced:
//   grad_i = scale * ( rowsum_i (e_i - mu) - (C.E)_i  +  sum_{j in P(i)} (cnt_ij + cnt_ji) s(D_ij) (e_i - e_j) )
// The last sum is batch-all's sparse positive-pair part (G_ij = +#{k : D_ij + m - D_ik > 0}; cnt_ji is looked up in
// j's own list -- "same label" is symmetric and the lists are complete, the forward rejects overflowing classes);
// each row has a single writer, no atomics.  scale = gloss (contrastive; the pair coefficients already carry
// 4 / (B (B-1))) or gloss / #positive triplets (batch-all; the count is only known after the last tile).
__global__ void pair_finish_kernel(const float* __restrict__ emb, const float* __restrict__ mu,
                                   const float* __restrict__ rowsum, int64_t B, int d, int cap, int squared,
                                   const float* __restrict__ pos_d, const int32_t* __restrict__ pos_j,
                                   const int32_t* __restrict__ pos_n, const int32_t* __restrict__ pos_cnt,
                                   const double* __restrict__ stats, const float* __restrict__ gloss,
                                   float* __restrict__ gemb) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= B) return;
  float scale = gloss ? gloss[0] : 1.f;
  if (stats) scale = static_cast<float>(static_cast<double>(scale) / (stats[1] + 1e-16));
  const float rs = rowsum[i];
  // lane s (and s + 32) prepares the weight of positive s
  const int n = pos_n ? pos_n[i] : 0;
  float wgt[2] = {0.f, 0.f};
  int jj[2] = {0, 0};
  if (n > 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int s = lane + 32 * h;
      if (s < n) {
        const int64_t j = pos_j[i * cap + s];
        int cnt = pos_cnt[i * cap + s];
        const int nj = pos_n[j];
        for (int t = 0; t < nj; ++t)
          if (pos_j[j * cap + t] == static_cast<int32_t>(i)) {
            cnt += pos_cnt[j * cap + t];
            break;
          }
        const float dij = pos_d[i * cap + s];
        wgt[h] = static_cast<float>(cnt) * (squared ? 2.f : (dij > 0.f ? 1.f / dij : 0.f));
        jj[h] = static_cast<int>(j);
      }
    }
  }
  const float* ei = emb + i * d;
  float* gi = gemb + i * d;
  const bool vec = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(gemb) & 15) == 0 && (reinterpret_cast<uintptr_t>(mu) & 15) == 0;
  auto ld4 = [&](const float* base, int c) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec) {
      if (c < d) r = *reinterpret_cast<const float4*>(base + c);
    } else {
      r.x = c < d ? base[c] : 0.f; r.y = c + 1 < d ? base[c + 1] : 0.f;
      r.z = c + 2 < d ? base[c + 2] : 0.f; r.w = c + 3 < d ? base[c + 3] : 0.f;
    }
    return r;
  };
  for (int c0 = 0; c0 < d; c0 += 128) {
    const int c = c0 + 4 * lane;
    const float4 x = ld4(ei, c), m = ld4(mu, c), g = ld4(gi, c);
    float4 acc = make_float4(fmaf(rs, x.x - m.x, g.x), fmaf(rs, x.y - m.y, g.y), fmaf(rs, x.z - m.z, g.z),
                             fmaf(rs, x.w - m.w, g.w));
    for (int s = 0; s < n; ++s) {
      const float w = __shfl_sync(0xffffffffu, wgt[s >> 5], s & 31);
      const int j = __shfl_sync(0xffffffffu, jj[s >> 5], s & 31);
      if (w == 0.f) continue;  // warp-uniform
      const float4 y = ld4(emb + static_cast<int64_t>(j) * d, c);
      acc.x = fmaf(w, x.x - y.x, acc.x);
      acc.y = fmaf(w, x.y - y.y, acc.y);
      acc.z = fmaf(w, x.z - y.z, acc.z);
      acc.w = fmaf(w, x.w - y.w, acc.w);
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    if (vec) {
      if (c < d) *reinterpret_cast<float4*>(gi + c) = acc;
    } else {
      if (c < d) gi[c] = acc.x;
      if (c + 1 < d) gi[c + 1] = acc.y;
      if (c + 2 < d) gi[c + 2] = acc.z;
      if (c + 3 < d) gi[c + 3] = acc.w;
    }
  }
}

struct Geometry {
  int dpad32, n_wide, rows_t, tiles, n_jparts, tiles_per_part;
  int64_t bpad;
};
static Geometry geometry(int64_t B, int d, int sms) {
  Geometry g;
  g.dpad32 = tc::dpad_for(d, 0);
  g.n_wide = (d + DW - 1) / DW;
  g.rows_t = g.n_wide * DW;
  g.bpad = (B + JB - 1) / JB * JB;
  g.tiles = static_cast<int>((B + BM - 1) / BM);
  // column-tile ranges per (row tile, column group): as many as it takes to give every SM an item (B = 4096,
  // d = 512: 32 x 2 x 2 = 128 items); each range re-reads nothing, the partial gradients are summed in gemb
  int jparts = sms / (g.tiles * g.n_wide);
  if (jparts < 1) jparts = 1;
  if (jparts > g.tiles) jparts = g.tiles;
  g.tiles_per_part = (g.tiles + jparts - 1) / jparts;
  g.n_jparts = (g.tiles + g.tiles_per_part - 1) / g.tiles_per_part;
  return g;
}

}  // namespace ptc

// ---------------------------------------------------------------------------------------------- host entry
size_t pair_tc_ws_bytes(int64_t B, int d) {
  const ptc::Geometry g = ptc::geometry(B, d, 148);
  return 2 * align_up(static_cast<size_t>(B) * g.dpad32 * 4) + 2 * align_up(static_cast<size_t>(B) * 4) +
         2 * align_up(static_cast<size_t>(g.rows_t) * g.bpad * 2) + align_up(static_cast<size_t>(d) * 4);
}

// Number of loss partials per row the fused kernel writes (PairPartial[B][n][2]); depends on the SM count only
// through the J-range split, which is capped by the tile count.
int pair_tc_partials_per_row(int64_t B, int d) {
  const int sms = device_sm_count();
  return ptc::geometry(B, d, sms > 0 ? sms : 148).n_jparts * ptc::EPI_Q;
}

// mode 0 = batch-all (pos_* describe lists with capacity `cap`: 8, or 64 for large classes -- sorted, -inf padded),
// mode 1 = all-pairs contrastive.
// partial != nullptr: also accumulate the forward loss (PairPartial[B][pair_tc_partials_per_row()]).
// coef_scale multiplies every pair coefficient (contrastive: 4 / (B (B-1)), batch-all: 1); stats (optional, device):
// batch-all coefficients are also divided by stats[1] = #positive triplets -- known when the forward already ran;
// the fused step passes null and rescales the finished gradient.
// gemb receives -(C.E) and `fin` the pointers pair_tc_finish() needs to turn it into the gradient (the caller may
// reduce the loss partials in between: batch-all's scale 1 / #positive triplets comes out of that reduction).
int pair_tc_launch(const float* emb, const int32_t* labels, int64_t B, int d, int mode, int squared, float margin,
                   float coef_scale, const float* pos_d, const double* pos_pre, const int32_t* pos_n, int32_t* pos_cnt,
                   int cap, const double* stats, PairPartial* partial, float* gemb, PairTcFinish* fin, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  if (int rc = check_sm100()) return rc;
  if (mode == 0 && cap > ptc::MAXP && pos_pre == nullptr)
    return fail(EN_ERR_ARG, "pair kernel: lists of more than %d slots need their prefix sums", ptc::MAXP);
  if (mode == 0 && cap > ptc::MAXP && B > 65535)
    return fail(EN_ERR_ARG, "pair kernel: classes of more than %d rows are supported up to 65535 rows per batch",
                ptc::MAXP + 1);
  if (mode == 0 && cap != ptc::MAXP && cap != ptc::MAXP_BIG)
    return fail(EN_ERR_ARG, "pair kernel: list capacity %d (must be %d or %d)", cap, ptc::MAXP, ptc::MAXP_BIG);
  if (!ws || ws_bytes < pair_tc_ws_bytes(B, d)) return fail(EN_ERR_WORKSPACE, "pair kernel: workspace too small");
  Workspace w(ws, ws_bytes);
  const int sms = device_sm_count();
  const ptc::Geometry g = ptc::geometry(B, d, sms);
  // Contrastive: BF16 GEMM1 planes (3 kind::f16 MMAs per k-step at twice the TF32 rate).  S feeds the smooth loss
  // terms and 1/D factors only -- no hinge between two distances is decided on it -- and the measured error of the
  // split-BF16 product on centred rows (~1e-6, random sign) averages out over the B (B-1) terms of the loss.
  const bool g1_bf16 = mode == 1;
  const int dpad = tc::dpad_for(d, g1_bf16);
  float* hi = w.take<float>(static_cast<size_t>(B) * g.dpad32);
  float* lo = w.take<float>(static_cast<size_t>(B) * g.dpad32);
  float* norms = w.take<float>(B);
  float* rowsum = w.take<float>(B);
  uint16_t* et_hi = w.take<uint16_t>(static_cast<size_t>(g.rows_t) * g.bpad);
  uint16_t* et_lo = w.take<uint16_t>(static_cast<size_t>(g.rows_t) * g.bpad);
  float* mu = w.take<float>(d);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "pair kernel: workspace too small or misaligned");
  tc::launch_column_mean(emb, B, d, mu, st);
  EN_LAUNCHED("column_mean_kernel");
  // GEMM1 also runs on the centred rows (norms are the centred norms): ||a-b|| is unchanged, S loses its
  // one-sided truncation bias, and with it the hinge-activity flips against the float64 oracle
  if (g1_bf16) EN_CUDA(tc::launch_split_bf16(emb, B, d, d, dpad, hi, lo, norms, st, nullptr, nullptr, mu));
  else EN_CUDA(tc::launch_split(emb, B, d, d, dpad, hi, lo, norms, st, mu));
  ++launch_counter();
  dim3 tb(32, 8), tg(static_cast<unsigned>(g.bpad / 32), static_cast<unsigned>(g.rows_t / 32));
  ptc::transpose_split_bf16_kernel<<<tg, tb, 0, st>>>(emb, mu, B, d, g.rows_t, g.bpad, et_hi, et_lo);
  EN_LAUNCHED("transpose_split_bf16_kernel");
  CUtensorMap th, tl, teh, tel;
  if ((g1_bf16 ? (tc::make_plane_tmap_bf16(&th, hi, B, dpad) || tc::make_plane_tmap_bf16(&tl, lo, B, dpad))
               : (tc::make_plane_tmap(&th, hi, B, dpad) || tc::make_plane_tmap(&tl, lo, B, dpad))) ||
      tc::make_plane_tmap_bf16(&teh, et_hi, g.rows_t, g.bpad) || tc::make_plane_tmap_bf16(&tel, et_lo, g.rows_t, g.bpad))
    return fail(EN_ERR_DRIVER, "pair kernel: cuTensorMapEncodeTiled failed");
  ptc::Params p;
  p.labels = labels; p.norms = norms; p.pos_d = pos_d; p.pos_pre = pos_pre; p.pos_n = pos_n; p.pos_cnt = pos_cnt; p.cap = cap;
  p.gemb = gemb; p.rowsum = rowsum; p.partial = partial; p.B = B; p.d = d;
  p.tiles = g.tiles; p.n_wide = g.n_wide; p.n_jparts = g.n_jparts; p.tiles_per_part = g.tiles_per_part;
  p.kblocks = dpad / (g1_bf16 ? tc::BK16 : tc::BK); p.squared = squared; p.margin = margin; p.coef_scale = coef_scale; p.stats = stats;
  const int items = p.tiles * p.n_wide * p.n_jparts;
  const int grid = items < sms ? items : sms;
  if (p.n_jparts > 1) EN_CUDA(cudaMemsetAsync(gemb, 0, static_cast<size_t>(B) * d * sizeof(float), st));
  EN_CUDA(cudaMemsetAsync(rowsum, 0, static_cast<size_t>(B) * sizeof(float), st));
  fin->mu = mu;
  fin->rowsum = rowsum;
#define EN_PAIR_LAUNCH(MODE, LOSS, G1B, BIG)                                                                         \
  do {                                                                                                              \
    EN_CUDA(cudaFuncSetAttribute(ptc::pair_tc_kernel<MODE, LOSS, G1B, BIG>,                                          \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, ptc::SMEM_BYTES));                    \
    prof_begin(st);                                                                                                 \
    ptc::pair_tc_kernel<MODE, LOSS, G1B, BIG><<<grid, ptc::NUM_THREADS, ptc::SMEM_BYTES, st>>>(th, tl, teh, tel, p);  \
    prof_end(st);                                                                                                   \
  } while (0)
  const bool big = mode == 0 && cap > ptc::MAXP;
  if (big && partial) EN_PAIR_LAUNCH(0, true, false, true);
  else if (big) EN_PAIR_LAUNCH(0, false, false, true);
  else if (mode == 0 && partial) EN_PAIR_LAUNCH(0, true, false, false);
  else if (mode == 0) EN_PAIR_LAUNCH(0, false, false, false);
  else if (partial) EN_PAIR_LAUNCH(1, true, true, false);
  else EN_PAIR_LAUNCH(1, false, true, false);
#undef EN_PAIR_LAUNCH
  EN_LAUNCHED("pair_tc_kernel");
  return EN_OK;
}

// pos_* may be null (contrastive); stats (device, optional): divide by stats[1]; gloss (device, optional)
int pair_tc_finish(const PairTcFinish& fin, const float* emb, int64_t B, int d, int cap, int squared, const float* pos_d,
                   const int32_t* pos_j, const int32_t* pos_n, const int32_t* pos_cnt, const double* stats,
                   const float* gloss, float* gemb, cudaStream_t st) {
  ptc::pair_finish_kernel<<<static_cast<unsigned>((B * 32 + 127) / 128), 128, 0, st>>>(
      emb, fin.mu, fin.rowsum, B, d, cap, squared, pos_d, pos_j, pos_n, pos_cnt, stats, gloss, gemb);
  EN_LAUNCHED("pair_finish_kernel");
  return EN_OK;
}

}  // namespace en
