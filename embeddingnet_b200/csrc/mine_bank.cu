// Offline hard-negative mining over an encoding bank with the reference's selection strategies
// (embedding_net/datagenerators.py:188-199 applied at bank scale -- BASELINE.json config 4):
// for every (anchor, positive) pair
//     loss_n = (d_ap - d_an) + margin            over all bank rows n of another class           (dg:235)
//     random_hard : uniform choice among { n : loss_n > 0 }                                       (dg:192-194)
//     semihard    : uniform choice among { n : 0 < loss_n < margin }                              (dg:196-199)
// The reference draws with np.random.choice(candidates), candidates in ascending row order; as in the in-batch path
// the GPU therefore returns candidate COUNTS, the host draws the rank from the legacy NumPy RNG, and the GPU returns
// the rank-th candidate in ascending row id.  Two passes of the tcgen05 distance GEMM (anchors x bank shard):
//   pass 1  en_mine_bank_count  : per (anchor, positive slot) the number of random-hard and semi-hard candidates;
//   pass 2  en_mine_bank_select : walks the shard in ascending row order with running counts and emits the row at
//                                 which the count reaches the requested rank.
// Exactness: the scan's distance carries a rounding error E (cert_bound()); an element whose loss lies within E of a
// predicate boundary (0 or margin) is re-evaluated on the spot from the fp32 rows (float64 sum (a-b)^2 -> float32
// sqrt, the loss in float32 exactly as dg:235) -- about 1e-5 of the elements.  Counts and selected ids therefore equal
// the float64 oracle's bit for bit, for either operand format and any sharding.
// (hardest = label-excluded nearest neighbour: en_knn_shard_topk with query_labels, then the loss > 0 test.)
#include "common.cuh"
#include "tc_engine.cuh"

namespace en {
namespace {

constexpr int MS = EN_MINE_MAX_SLOTS;  // positives per anchor

__device__ __forceinline__ float mining_loss(float d_ap, float d_an, float margin) {
  return __fadd_rn(__fsub_rn(d_ap, d_an), margin);  // float32, left to right (dg:235)
}

// one thread, float64 accumulation: the distance this path is defined by
__device__ __noinline__ float exact_dist(const float* __restrict__ a, const float* __restrict__ b, int d) {
  double acc = 0.0;
  for (int c = 0; c < d; ++c) {
    const double t = static_cast<double>(a[c]) - static_cast<double>(b[c]);
    acc = fma(t, t, acc);
  }
  return sqrtf(static_cast<float>(acc));
}

__global__ void pair_dist_exact_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, int d,
                                       float* __restrict__ out) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  // same summation as exact_dist would be ideal; a warp-parallel float64 sum differs from the serial one by ~1e-16
  // relative, far below the float32 rounding of the result, and this value is an INPUT (d_ap) of the predicates
  double acc = 0.0;
  for (int c = lane; c < d; c += 32) {
    const double t = static_cast<double>(a[i * d + c]) - static_cast<double>(b[i * d + c]);
    acc = fma(t, t, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[i] = sqrtf(static_cast<float>(acc));
}

struct MineParams {
    const float* anchors;        // (A, d) fp32
    const float* bank;           // (n, d) fp32 (this shard)
    const float* anchor_norms;   // (A,)
    const float* bank_norms;     // (n,)
    const int32_t* anchor_labels;
    const int32_t* bank_labels;  // (n,) labels of this shard's rows
    const float* pos_d;          // (A, MS) d_ap per slot, < 0 = unused
    int32_t* counts;             // (A, MS, 2) [random_hard, semihard]          (count pass)
    const int32_t* target;       // (A, MS) requested rank inside this shard, < 0 = none   (select pass)
    int32_t* running;            // (A, MS) candidates seen so far, carried between launches (select pass)
    int64_t* selected;           // (A, MS) global id                                        (select pass)
    int64_t A, n_bank, id_offset;
    int d, semihard;             // select pass: which predicate
    float margin, c_err;         // c_err: |d2~ - d2| <= c_err (|a|^2 + |b|^2)
};

// NS = positive slots actually in use (1, 2, 4 or 8): the per-element work is proportional to it
template <bool kSelect, int NS>
struct EpMine {
  using Params = MineParams;
  struct Row {
    float dap[NS];
    float na, dmax;
    int32_t la;
    int32_t cnt_h[kSelect ? 1 : NS], cnt_s[kSelect ? 1 : NS];
    unsigned long long mask[kSelect ? NS : 1];
    int32_t run[kSelect ? NS : 1], tgt[kSelect ? NS : 1];
  };
  static constexpr int kSmemBytes = kSelect ? tc::EPI_H * tc::BM * NS * 4 : 0;

  static __device__ void item_begin(const Params& p, Row& r, const tc::Ctx&, int64_t row, bool valid, int, int) {
    r.na = valid ? p.anchor_norms[row] : 0.f;
    r.la = valid ? p.anchor_labels[row] : 0;
    r.dmax = 0.f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float v = valid ? p.pos_d[row * MS + s] : -1.f;
      r.dap[s] = v >= 0.f ? v : -INFINITY;  // unused slot: loss = -inf, no predicate holds
      r.dmax = fmaxf(r.dmax, v);
      if (kSelect) {
        r.mask[s] = 0ull;
        r.run[s] = valid ? p.running[row * MS + s] : 0;
        r.tgt[s] = valid ? p.target[row * MS + s] : -1;
      } else {
        r.cnt_h[s] = 0;
        r.cnt_s[s] = 0;
      }
    }
    r.dmax += p.margin;
  }

  static __device__ void chunk(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int64_t col0,
                               const float (&dot)[32]) {
    if (col0 >= p.n_bank) return;  // warp-uniform
    tc::stage_columns(ctx, p.bank_norms, p.bank_labels, col0, p.n_bank);
    const int ncols = static_cast<int>(p.n_bank - col0 < 32 ? p.n_bank - col0 : 32);
    const int bit0 = static_cast<int>(col0 % tc::BN) - ctx.half * tc::COLS_PER_EPI_WARP;  // 0 or 32 inside the half
    // Chunk-level reject in the proxy domain (two instructions per element): with T = d_max + A, A >= the error
    // bound of any column of this chunk whose distance exceeds d_max (A uses the chunk's largest norm; e2 / d~ <=
    // e2 / d_max there), every element with  |a|^2 + t > T^2 (1 + 2e-6)  is beyond every slot's d_ap + margin by
    // more than its own bound: no predicate holds and none is uncertain.  ncu (round 1): without it the count pass
    // ran 413 M warp instructions per launch with the tensor pipe 22 % active.
    {
      const float nbmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(ctx.wf[ctx.lane])));
      // (a non-positive d_max -- only with a non-positive margin and no slot in use -- disables the reject)
      const float A = r.dmax > 0.f ? p.c_err * (r.na + nbmax) / r.dmax + 1e-6f * r.dmax : INFINITY;
      const float T = r.dmax + A;
      const float tthr = T * T * 1.000002f - r.na;
      float tmin = INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) tmin = fminf(tmin, fmaf(-2.f, dot[j], ctx.wf[j]));
      if (tmin > tthr) return;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float nb = ctx.wf[j];
      const bool cand = valid && j < ncols && ctx.wi[j] != r.la;
      const float d2 = fmaxf(r.na + nb - 2.f * dot[j], 0.f);
      const float e2 = p.c_err * (r.na + nb);
      float rs;
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(d2, 1e-30f)));
      const float dn = d2 * rs;
      // |d~ - d| <= e2 / d~ (+ the approximate sqrt and the float32 roundings of the loss); tiny distances: always exact
      const float eb = d2 > 4.f * e2 ? fmaf(e2, rs, 6e-7f * (dn + r.dmax)) : INFINITY;
      // beyond every slot's d_ap + margin by more than the error bound: no predicate can hold (nor be uncertain) --
      // the common case for a bank, and the reason this epilogue keeps up with the tensor pipe
      if (!(dn <= r.dmax + eb)) continue;
      unsigned hard = 0, semi = 0;
      bool unc = false;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float l = (r.dap[s] - dn) + p.margin;
        hard |= (l > 0.f ? 1u : 0u) << s;
        semi |= ((l > 0.f && l < p.margin) ? 1u : 0u) << s;
        unc = unc || (isfinite(l) && (fabsf(l) <= eb || fabsf(l - p.margin) <= eb));
      }
      if (cand && unc) {  // re-evaluate from the fp32 rows: ~1e-5 of the elements
        const float de = exact_dist(p.anchors + row * p.d, p.bank + (col0 + j) * p.d, p.d);
        hard = semi = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float l = mining_loss(r.dap[s], de, p.margin);
          hard |= (l > 0.f ? 1u : 0u) << s;
          semi |= ((l > 0.f && l < p.margin) ? 1u : 0u) << s;
        }
      }
      if (!cand) hard = semi = 0;
      if (kSelect) {
        const unsigned pick = p.semihard ? semi : hard;
#pragma unroll
        for (int s = 0; s < NS; ++s) r.mask[s] |= static_cast<unsigned long long>((pick >> s) & 1u) << (bit0 + j);
      } else {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          r.cnt_h[s] += (hard >> s) & 1u;
          r.cnt_s[s] += (semi >> s) & 1u;
        }
      }
    }
  }

  // select pass: candidates are ranked in ascending row id, i.e. column half 0 of a tile before half 1: the two
  // threads of a row exchange their per-tile counts through shared memory
  static __device__ void tile_end(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int tile_n) {
    if (!kSelect) return;
    int32_t* sm = reinterpret_cast<int32_t*>(ctx.smem);
#pragma unroll
    for (int s = 0; s < NS; ++s) sm[(ctx.half * tc::BM + ctx.erow) * NS + s] = __popcll(r.mask[s]);
    ptx::named_bar_sync(2, tc::EPI_WARPS * 32);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int c0 = sm[ctx.erow * NS + s], c1 = sm[(tc::BM + ctx.erow) * NS + s];
      const int before = r.run[s] + (ctx.half ? c0 : 0);
      const int own = ctx.half ? c1 : c0;
      const int want = r.tgt[s] - before;
      if (valid && r.tgt[s] >= 0 && want >= 0 && want < own) {
        unsigned long long m = r.mask[s];
        for (int k = 0; k < want; ++k) m &= m - 1;  // drop the `want` lowest candidates
        const int bit = __ffsll(static_cast<long long>(m)) - 1;
        p.selected[row * MS + s] = p.id_offset + static_cast<int64_t>(tile_n) * tc::BN +
                                   ctx.half * tc::COLS_PER_EPI_WARP + bit;
      }
      r.run[s] += c0 + c1;
      r.mask[s] = 0ull;
    }
    ptx::named_bar_sync(2, tc::EPI_WARPS * 32);  // the exchange area is rewritten at the next tile
  }

  static __device__ void item_end(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int, int) {
    if (!valid) return;
    if (kSelect) {
      if (ctx.half == 0)
#pragma unroll
        for (int s = 0; s < NS; ++s) p.running[row * MS + s] = r.run[s];
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (r.cnt_h[s]) atomicAdd(&p.counts[(row * MS + s) * 2 + 0], r.cnt_h[s]);
        if (r.cnt_s[s]) atomicAdd(&p.counts[(row * MS + s) * 2 + 1], r.cnt_s[s]);
      }
    }
  }
};

constexpr int kMineChunkTiles = 64;  // bank tiles per launch and column range: planes stay L2 resident

struct MineOperands {
  CUtensorMap ah, al, bh, bl;
  float* an;
  int dpad, bf16;
};

int mine_prepare(const float* anchors, int64_t A, int d, const void* bank_hi, const void* bank_lo, int64_t n_bank,
                 int precision, Workspace& w, cudaStream_t st, MineOperands& o, const char* who) {
  o.bf16 = precision == EN_PREC_BF16X3;
  o.dpad = tc::dpad_for(d, o.bf16);
  const size_t dpad32 = static_cast<size_t>(tc::dpad_for(d, 0));
  float* ahi = w.take<float>(static_cast<size_t>(A) * dpad32);
  float* alo = w.take<float>(static_cast<size_t>(A) * dpad32);
  o.an = w.take<float>(A);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "%s: workspace too small or misaligned", who);
  int bad;
  if (o.bf16) {
    EN_CUDA(tc::launch_split_bf16(anchors, A, d, d, o.dpad, ahi, alo, o.an, st));
    bad = tc::make_plane_tmap_bf16(&o.ah, ahi, A, o.dpad) || tc::make_plane_tmap_bf16(&o.al, alo, A, o.dpad) ||
          tc::make_plane_tmap_bf16(&o.bh, bank_hi, n_bank, o.dpad) || tc::make_plane_tmap_bf16(&o.bl, bank_lo, n_bank, o.dpad);
  } else {
    EN_CUDA(tc::launch_split(anchors, A, d, d, o.dpad, ahi, alo, o.an, st));
    bad = tc::make_plane_tmap(&o.ah, ahi, A, o.dpad) || tc::make_plane_tmap(&o.al, alo, A, o.dpad) ||
          tc::make_plane_tmap(&o.bh, static_cast<const float*>(bank_hi), n_bank, o.dpad) ||
          tc::make_plane_tmap(&o.bl, static_cast<const float*>(bank_lo), n_bank, o.dpad);
  }
  ++launch_counter();
  if (bad) return fail(EN_ERR_DRIVER, "%s: cuTensorMapEncodeTiled failed", who);
  return EN_OK;
}

template <bool kSelect, int NS>
int mine_scan(const MineOperands& o, const MineParams& ep, int64_t A, int64_t n_bank, int d,
              cudaStream_t st) {
  const int sms = device_sm_count();
  const int tiles_total = static_cast<int>((n_bank + tc::BN - 1) / tc::BN);
  const int tiles_m = static_cast<int>((A + tc::BM - 1) / tc::BM);
  // count pass: any number of column ranges per launch (counts are summed with atomics); select pass: ONE range per
  // launch, so that each row's thread pair meets its candidates in ascending row order
  int splits = 1;
  if (!kSelect) {
    splits = (2 * sms + tiles_m - 1) / tiles_m;
    if (splits < 1) splits = 1;
  }
  const int chunk_tiles = splits * kMineChunkTiles;
  prof_begin(st);
  for (int base = 0; base < tiles_total; base += chunk_tiles) {
    const int tiles_here = tiles_total - base < chunk_tiles ? tiles_total - base : chunk_tiles;
    tc::Shape sh = tc::make_shape(A, static_cast<int64_t>(tiles_here) * tc::BN, d, splits, 3, o.bf16);
    sh.nt_base = base;
    EN_CUDA((tc::launch<EpMine<kSelect, NS>>(o.ah, o.al, o.bh, o.bl, sh, ep, sms, st)));
    ++launch_counter();
  }
  prof_end(st);
  return EN_OK;
}

}  // namespace
}  // namespace en

using namespace en;

extern "C" {

int en_pair_dist_exact(const float* a, const float* b, int64_t n, int d, float* dist, void* stream) {
  EN_REQUIRE(a && b && dist && n >= 0 && d > 0, "en_pair_dist_exact: bad arguments");
  if (n == 0) return EN_OK;
  pair_dist_exact_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(a, b, n, d, dist);
  EN_LAUNCHED("pair_dist_exact_kernel");
  return EN_OK;
}

size_t en_ws_bytes_mine_bank(int64_t A, int d) {
  if (A <= 0 || d <= 0) return 0;
  return 2 * align_up(static_cast<size_t>(A) * tc::dpad_for(d, 0) * 4) + align_up(static_cast<size_t>(A) * 4) +
         align_up(static_cast<size_t>(A) * MS * 4);
}

int en_mine_bank_count(const float* anchors, const int32_t* anchor_labels, const float* pos_d, int64_t A, int d,
                       int n_slots, float margin, const float* bank, const void* bank_hi, const void* bank_lo,
                       const float* bank_norms, const int32_t* bank_labels, int64_t n_bank, int precision,
                       int32_t* counts, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(anchors && anchor_labels && pos_d && bank && bank_hi && bank_lo && bank_norms && bank_labels && counts &&
                 A > 0 && d > 0 && n_bank > 0,
             "en_mine_bank_count: bad arguments");
  EN_REQUIRE(precision == EN_PREC_TF32X3 || precision == EN_PREC_BF16X3, "en_mine_bank_count: unknown precision");
  EN_REQUIRE(n_slots >= 1 && n_slots <= MS, "en_mine_bank_count: n_slots must be in [1, %d]", MS);
  EN_REQUIRE(n_bank < (int64_t(1) << 31), "en_mine_bank_count: shard too large");
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_mine_bank(A, d)) return fail(EN_ERR_WORKSPACE, "en_mine_bank_count: workspace too small");
  cudaStream_t st = as_stream(stream);
  Workspace w(ws, ws_bytes);
  MineOperands o;
  if (int rc = mine_prepare(anchors, A, d, bank_hi, bank_lo, n_bank, precision, w, st, o, "en_mine_bank_count")) return rc;
  EN_CUDA(cudaMemsetAsync(counts, 0, static_cast<size_t>(A) * MS * 2 * sizeof(int32_t), st));
  MineParams ep{anchors, bank, o.an, bank_norms, anchor_labels, bank_labels, pos_d, counts, nullptr,
                              nullptr, nullptr, A, n_bank, 0, d, 0, margin,
                              static_cast<float>(cert_bound(precision, o.dpad) + 3e-7)};
  if (n_slots <= 1) return mine_scan<false, 1>(o, ep, A, n_bank, d, st);
  if (n_slots <= 2) return mine_scan<false, 2>(o, ep, A, n_bank, d, st);
  if (n_slots <= 4) return mine_scan<false, 4>(o, ep, A, n_bank, d, st);
  return mine_scan<false, 8>(o, ep, A, n_bank, d, st);
}

int en_mine_bank_select(const float* anchors, const int32_t* anchor_labels, const float* pos_d, int64_t A, int d,
                        int n_slots, float margin, int mode, const int32_t* rank, const float* bank, const void* bank_hi,
                        const void* bank_lo, const float* bank_norms, const int32_t* bank_labels, int64_t n_bank,
                        int64_t id_offset, int precision, int64_t* selected, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(anchors && anchor_labels && pos_d && rank && bank && bank_hi && bank_lo && bank_norms && bank_labels &&
                 selected && A > 0 && d > 0 && n_bank > 0,
             "en_mine_bank_select: bad arguments");
  EN_REQUIRE(mode == EN_MODE_SEMIHARD || mode == EN_MODE_RANDOM_HARD,
             "en_mine_bank_select: mode must be EN_MODE_SEMIHARD or EN_MODE_RANDOM_HARD (got %d)", mode);
  EN_REQUIRE(precision == EN_PREC_TF32X3 || precision == EN_PREC_BF16X3, "en_mine_bank_select: unknown precision");
  EN_REQUIRE(n_slots >= 1 && n_slots <= MS, "en_mine_bank_select: n_slots must be in [1, %d]", MS);
  EN_REQUIRE(n_bank < (int64_t(1) << 31), "en_mine_bank_select: shard too large");
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_mine_bank(A, d)) return fail(EN_ERR_WORKSPACE, "en_mine_bank_select: workspace too small");
  cudaStream_t st = as_stream(stream);
  Workspace w(ws, ws_bytes);
  MineOperands o;
  if (int rc = mine_prepare(anchors, A, d, bank_hi, bank_lo, n_bank, precision, w, st, o, "en_mine_bank_select")) return rc;
  int32_t* running = w.take<int32_t>(static_cast<size_t>(A) * MS);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_mine_bank_select: workspace too small or misaligned");
  EN_CUDA(cudaMemsetAsync(running, 0, static_cast<size_t>(A) * MS * sizeof(int32_t), st));
  EN_CUDA(cudaMemsetAsync(selected, 0xFF, static_cast<size_t>(A) * MS * sizeof(int64_t), st));  // -1
  MineParams ep{anchors, bank, o.an, bank_norms, anchor_labels, bank_labels, pos_d, nullptr, rank, running,
                             selected, A, n_bank, id_offset, d, mode == EN_MODE_SEMIHARD ? 1 : 0, margin,
                             static_cast<float>(cert_bound(precision, o.dpad) + 3e-7)};
  if (n_slots <= 1) return mine_scan<true, 1>(o, ep, A, n_bank, d, st);
  if (n_slots <= 2) return mine_scan<true, 2>(o, ep, A, n_bank, d, st);
  if (n_slots <= 4) return mine_scan<true, 4>(o, ep, A, n_bank, d, st);
  return mine_scan<true, 8>(o, ep, A, n_bank, d, st);
}

}  // extern "C"
