// Fused in-batch losses on the tcgen05 distance GEMM: batch-hard triplet, batch-all triplet and all-pairs
// contrastive, forward and backward.  The B x B distance matrix is consumed tile by tile out of TMEM and never
// written to memory.
//
// None of the three exists in the reference: BASELINE.json's north_star asks for them behind the
// reference's `factory(margin) -> fn(y_true, y_pred)` shape (embedding_net/losses_and_accuracies.py:14,26); the
// formulas are the ones the reference README cites (README.md:112 Hermans et al., README.md:116 Moindrot).
// Contrastive keeps losses_and_accuracies.py:4-11 (margin 1, label 1 = same) with the Siamese head's distance
// clamp from embedding_net/models.py:225.
//
// Numerics: the tensor-core pass (3xTF32; split-BF16 for batch-hard) only *selects* (arg-max positive, arg-min negative) or feeds sums whose
// terms are O(1); every distance that reaches a loss value or a gradient of the batch-hard path is re-evaluated
// exactly (float64 sum (a-b)^2) for the one or two candidates per anchor that matter.
#include <cstdlib>
#include "common.cuh"
#include "tc_engine.cuh"
#include "tc_engine_wide.cuh"

namespace en {

// csrc/pair_tc.cu: fused tensor-core kernel of the pair losses (two chained tcgen05 GEMMs; loss and / or gradient)
size_t pair_tc_ws_bytes(int64_t B, int d);
int pair_tc_partials_per_row(int64_t B, int d);
int pair_tc_launch(const float* emb, const int32_t* labels, int64_t B, int d, int mode, int squared, float margin,
                   float coef_scale, const float* pos_d, const double* pos_pre, const int32_t* pos_n, int32_t* pos_cnt,
                   int cap, const double* stats, PairPartial* partial, float* gemb, PairTcFinish* fin, void* ws,
                   size_t ws_bytes, cudaStream_t st);
int pair_tc_finish(const PairTcFinish& fin, const float* emb, int64_t B, int d, int cap, int squared, const float* pos_d,
                   const int32_t* pos_j, const int32_t* pos_n, const int32_t* pos_cnt, const double* stats,
                   const float* gloss, float* gemb, cudaStream_t st);

namespace {

constexpr float kBig = 3.0e38f;

// Developer aid, compiled in only with -DEN_FIN_TRACE (tools/trace_bh.py builds such a variant; in the production
// build these are empty -- the stamps cost ~1 us per step, measured A/B): globaltimer stamps of the two finalize
// kernels.  fin_stamp: first entry (atomic min) / last exit (atomic max); fin_sample: plain per-phase stamps of every
// 64th anchor's warp at [8 + 2048 + 4 * (row / 64) + phase].
__device__ unsigned long long* g_fin_trace = nullptr;
__device__ __forceinline__ void fin_stamp(int slot, bool is_min) {
#ifdef EN_FIN_TRACE
  unsigned long long* t = g_fin_trace;
  if (t == nullptr) return;
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (is_min) atomicMin(&t[slot], now);
  else atomicMax(&t[slot], now);
#endif
}
__device__ __forceinline__ void fin_sample(int64_t row, int lane, int phase) {
#ifdef EN_FIN_TRACE
  unsigned long long* t = g_fin_trace;
  if (t == nullptr || lane != 0 || (row & 63) != 0 || row >= 64 * 64) return;
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  t[8 + 2048 + 4 * (row >> 6) + phase] = now;
#endif
}
constexpr int kTcBwdMaxPos = 8;  // pair_tc_kernel keeps eight positives per anchor in registers / scratch per pass
// list capacity of the tensor-core pair kernel for a class bound: 8, or 64 -- longer lists are sorted, padded to 64
// slots and binary-searched per element with constant strides (pair_tc_kernel<..., kBig>)
static int tc_list_cap(int max_positives) { return max_positives <= kTcBwdMaxPos ? kTcBwdMaxPos : 64; }
// EN_BATCH_ALL_CUDA_CORE=1 sends classes with more than 8 positives per anchor to the CUDA-core tile kernel of
// round 1 (kept as an independent implementation for the tests to compare against)
static bool cuda_core_bwd_requested() {
  const char* e = getenv("EN_BATCH_ALL_CUDA_CORE");
  return e && e[0] == '1';
}

// warp-cooperative exact squared distance between rows i and j (float64 accumulate); result in every lane
// (not inlined: it is called from many sites of kernels whose warps run the code once, where instruction fetch,
// not issue, is the cost -- ncu round 1: stall_no_instruction 6.4 per issue in the finalize kernel)
__device__ __noinline__ double exact_d2(const float* __restrict__ e, int d, int64_t i, int64_t j, int lane) {
  const float* a = e + i * d;
  const float* b = e + j * d;
  double acc = 0.0;
  if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(e) & 15) == 0) {
    for (int c = lane * 4; c < d; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(a + c), y = *reinterpret_cast<const float4*>(b + c);
      double t;
      t = static_cast<double>(x.x) - static_cast<double>(y.x); acc = fma(t, t, acc);
      t = static_cast<double>(x.y) - static_cast<double>(y.y); acc = fma(t, t, acc);
      t = static_cast<double>(x.z) - static_cast<double>(y.z); acc = fma(t, t, acc);
      t = static_cast<double>(x.w) - static_cast<double>(y.w); acc = fma(t, t, acc);
    }
  } else {
    for (int c = lane; c < d; c += 32) {
      const double t = static_cast<double>(a[c]) - static_cast<double>(b[c]);
      acc = fma(t, t, acc);
    }
  }
  return warp_sum(acc);
}

// =====================================================================================================
// Batch-hard
// =====================================================================================================
// Candidate record per (anchor row, other tile T, slot): the two largest same-label and the two smallest
// other-label proxies t = |b|^2 - 2 a.b (monotone in the distance for a fixed anchor).  Each float carries the
// candidate's index INSIDE tile T (0..127) in its 7 lowest mantissa bits, so the running top-2 is three FMNMX per
// element instead of compare + select chains on (value, index) pairs, and a record is 16 bytes.  The truncation
// (< 2^-16 |t|) is part of the error band the finalize kernel re-evaluates exactly.
struct BhCand {
  float p1, p2, n1, n2;
};
constexpr unsigned kKeyMask = 0xFFFFFF80u;
__device__ __forceinline__ float bh_pack(float t, int in_tile) {
  return __uint_as_float((__float_as_uint(t) & kKeyMask) | static_cast<unsigned>(in_tile));
}
__device__ __forceinline__ bool bh_valid(float key) { return fabsf(key) < 1.0e38f; }

// running top-2 of packed keys; masked elements carry -kBig / +kBig and never win
__device__ __forceinline__ void top2_max(float& v1, float& v2, float k) {
  v2 = fmaxf(v2, fminf(v1, k));
  v1 = fmaxf(v1, k);
}
__device__ __forceinline__ void top2_min(float& v1, float& v2, float k) {
  v2 = fminf(v2, fmaxf(v1, k));
  v1 = fminf(v1, k);
}

// Symmetric schedule: only tiles (I, J) with J >= I are computed.  A strictly-upper tile serves both the anchors
// of row tile I (row view: candidates among the columns) and the anchors of row tile J (column view: candidates
// among the rows), so every dot product is computed once.  Record slots per (anchor, other tile T):
//   T >= tile(anchor): slots 0,1 = the two column halves of the row view;
//   T <  tile(anchor): slots 0..3 = the four 32-row quarters of the column view.
constexpr int BH_SLOTS = 4;

struct EpBatchHard {
  struct Params {
    const int32_t* labels;
    const float* norms;
    BhCand* cand;  // [B][tiles_n][BH_SLOTS]
    int64_t B;
    int tiles_n;
  };
  struct Row {
    int32_t la;
    int32_t lmin, lmax;  // label range of the 32 rows this warp serves
    float na;
    int tile_m;
    BhCand c;
  };
  // per-warp 32x32 transposition scratch for the column view
  static constexpr int kSmemBytes = tc::EPI_WARPS * 32 * 32 * 4;
  static __device__ void reset(BhCand& c) {
    c.p1 = c.p2 = -kBig;
    c.n1 = c.n2 = kBig;
  }
  static __device__ void store(BhCand* out, const BhCand& c) {
    *reinterpret_cast<float4*>(out) = make_float4(c.p1, c.p2, c.n1, c.n2);
  }
  static __device__ void item_begin(const Params& p, Row& r, const tc::Ctx&, int64_t row, bool valid, int tile_m,
                                    int) {
    r.la = valid ? p.labels[row] : 0;
    r.na = valid ? p.norms[row] : 0.f;
    r.lmin = __reduce_min_sync(0xffffffffu, valid ? r.la : 0x7fffffff);
    r.lmax = __reduce_max_sync(0xffffffffu, valid ? r.la : static_cast<int32_t>(0x80000000));
    r.tile_m = tile_m;
    reset(r.c);
  }
  static __device__ void chunk(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int64_t col0,
                               const float (&dot)[32]) {
    if (col0 >= p.B) return;  // warp-uniform
    tc::stage_columns(ctx, p.norms, p.labels, col0, p.B);
    const float4* n4 = reinterpret_cast<const float4*>(ctx.wf);
    const int4* l4 = reinterpret_cast<const int4*>(ctx.wi);
    const int jt0 = static_cast<int>(col0 % tc::BN);  // first column of the chunk inside its tile
    const int64_t row0 = row - ctx.lane;              // first row served by this warp
    // Rows and columns are tiled in aligned groups of 32, so the diagonal can only fall into the chunk whose first
    // column equals this warp's first row.
    const bool edge = (col0 + 32 > p.B) || (col0 == row0) || (row0 + 32 > p.B);
    // No label shared between this warp's rows and the chunk's columns (always true off the diagonal when the
    // batch is class-major, as P x K batches are): every element is a negative, no label compares needed.
    const int32_t lcol = ctx.wi[ctx.lane];
    const int32_t cmin = __reduce_min_sync(0xffffffffu, lcol), cmax = __reduce_max_sync(0xffffffffu, lcol);
    const bool disjoint = !edge && (cmax < r.lmin || cmin > r.lmax);
    const int tile_n = static_cast<int>(col0 / tc::BN);
    const bool col_view = tile_n > r.tile_m;  // strictly-upper tiles also serve the column anchors
    float* sc = reinterpret_cast<float*>(ctx.smem) + (ctx.quarter + 4 * ctx.half) * 1024;
    if (col_view) __syncwarp();
    if (disjoint) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 nb = n4[g];
        const float nbv[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = g * 4 + u;
          top2_min(r.c.n1, r.c.n2, bh_pack(fmaf(-2.f, dot[j], nbv[u]), jt0 + j));
          // column view: transpose through shared memory (skewed: conflict free), already packed with the ROW index
          if (col_view) sc[ctx.lane * 32 + ((j + ctx.lane) & 31)] = bh_pack(fmaf(-2.f, dot[j], r.na), ctx.erow);
        }
      }
    } else if (!edge) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 nb = n4[g];
        const int4 lb = l4[g];
        const float nbv[4] = {nb.x, nb.y, nb.z, nb.w};
        const int lbv[4] = {lb.x, lb.y, lb.z, lb.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = g * 4 + u;
          const float k = bh_pack(fmaf(-2.f, dot[j], nbv[u]), jt0 + j);
          const bool same = lbv[u] == r.la;
          top2_max(r.c.p1, r.c.p2, same ? k : -kBig);
          top2_min(r.c.n1, r.c.n2, same ? kBig : k);
          if (col_view) sc[ctx.lane * 32 + ((j + ctx.lane) & 31)] = bh_pack(fmaf(-2.f, dot[j], r.na), ctx.erow);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t c = col0 + j;
        const bool ok = c < p.B && c != row;
        const float k = bh_pack(fmaf(-2.f, dot[j], ctx.wf[j]), jt0 + j);
        const bool same = ctx.wi[j] == r.la;
        top2_max(r.c.p1, r.c.p2, (ok && same) ? k : -kBig);
        top2_min(r.c.n1, r.c.n2, (ok && !same) ? k : kBig);
        if (col_view) sc[ctx.lane * 32 + ((j + ctx.lane) & 31)] = bh_pack(fmaf(-2.f, dot[j], r.na), ctx.erow);
      }
    }
    // ---- column view: lane c owns column col0 + c and scans this warp's 32 rows.
    if (col_view) {
      __syncwarp();
      BhCand cc;
      reset(cc);
      if (disjoint) {
#pragma unroll
        for (int i = 0; i < 32; ++i) top2_min(cc.n1, cc.n2, sc[i * 32 + ((ctx.lane + i) & 31)]);
      } else {
        const bool col_ok = col0 + ctx.lane < p.B;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          const float v = sc[i * 32 + ((ctx.lane + i) & 31)];
          const int32_t li = __shfl_sync(0xffffffffu, r.la, i);
          const bool ok = col_ok && (row0 + i < p.B);
          const bool same = li == lcol;
          top2_max(cc.p1, cc.p2, (ok && same) ? v : -kBig);
          top2_min(cc.n1, cc.n2, (ok && !same) ? v : kBig);
        }
      }
      if (col0 + ctx.lane < p.B)
        store(p.cand + ((col0 + ctx.lane) * p.tiles_n + r.tile_m) * BH_SLOTS + ctx.quarter, cc);
    }
  }
  static __device__ void tile_end(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid,
                                  int tile_n) {
    if (valid) store(p.cand + (row * p.tiles_n + tile_n) * BH_SLOTS + ctx.half, r.c);
    reset(r.c);
  }
  static __device__ void item_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int, int) {}
};

// One warp per anchor: pick the winners among the per-tile candidates.  Every candidate whose proxy lies within the
// tensor-core error band of the best one is re-evaluated exactly (float64); ties resolve to the lowest index.
struct BhPick {
  double d2;
  int idx;
};

// Contender thresholds.  |dot~ - dot| <= c |a||b|, so a proxy t = |b|^2 - 2 a.b is off by at most c (|a|^2 + |b|^2)
// (band_c already holds 2c for "best and contender both carry it" plus the index-packing truncation).  The
// candidate's norm is bounded from the proxy itself (t >= |b|^2 - 2|a||b|  =>  |b| <= |a| + sqrt(|a|^2 + t)), and
// that bound grows with t, so ONE threshold per side covers the best entry and every contender: positives have
// t <= bp; negatives have t <= bn + band_n, evaluated at that (slightly inflated) upper end.
struct BhThr {
  float p, n;
};
__device__ __forceinline__ BhThr bh_thresholds(float na, float bp, float bn, float band_c) {
  const float sa = sqrtf(na);
  auto band_at = [&](float t) {
    const float sb = sa + sqrtf(fmaxf(na + t, 0.f)) * 1.0001f;
    return band_c * (na + sb * sb) + 1e-30f;
  };
  BhThr r;
  r.p = bp - band_at(bp);
  r.n = bn + band_at(bn + band_at(bn) * 1.5f) * 1.0001f;
  return r;
}

__device__ __forceinline__ void bh_update_max(BhPick& win, double d2, int ci) {
  if (win.idx < 0 || d2 > win.d2 || (d2 == win.d2 && ci < win.idx)) win = BhPick{d2, ci};
}
__device__ __forceinline__ void bh_update_min(BhPick& win, double d2, int ci) {
  if (win.idx < 0 || d2 < win.d2 || (d2 == win.d2 && ci < win.idx)) win = BhPick{d2, ci};
}

// Squared distance between rows i and j by ONE lane in float32 (four partial sums).  All terms are non-negative, so
// the result is within (d/4 + 8) * 2^-23 RELATIVE of the true value: a filter five orders of magnitude sharper than
// the tensor-core proxy, at a fraction of the cost of the float64 evaluation.  Lanes of a warp work on different
// rows j of the same anchor i (the anchor's loads are broadcasts).
__device__ __forceinline__ float lane_d2_f32(const float* __restrict__ e, int d, int64_t i, int64_t j) {
  const float* a = e + i * d;
  const float* b = e + j * d;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(e) & 15) == 0) {
#pragma unroll 4
    for (int c = 0; c < d; c += 4) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a + c)), y = __ldg(reinterpret_cast<const float4*>(b + c));
      float t;
      t = x.x - y.x; s0 = fmaf(t, t, s0);
      t = x.y - y.y; s1 = fmaf(t, t, s1);
      t = x.z - y.z; s2 = fmaf(t, t, s2);
      t = x.w - y.w; s3 = fmaf(t, t, s3);
    }
  } else {
    for (int c = 0; c < d; ++c) {
      const float t = a[c] - b[c];
      s0 = fmaf(t, t, s0);
    }
  }
  return (s0 + s1) + (s2 + s3);
}

// float32 squared distances of the anchor row to four rows at once, warp-cooperative (d % 128 == 0, 16-byte aligned)
__device__ __forceinline__ void warp_d2_f32x4(const float* __restrict__ e, int d, int64_t row, const int (&j)[4],
                                              int lane, float (&out)[4]) {
  const float4* a = reinterpret_cast<const float4*>(e + row * d) + lane;
  const float4* b0 = reinterpret_cast<const float4*>(e + static_cast<int64_t>(j[0]) * d) + lane;
  const float4* b1 = reinterpret_cast<const float4*>(e + static_cast<int64_t>(j[1]) * d) + lane;
  const float4* b2 = reinterpret_cast<const float4*>(e + static_cast<int64_t>(j[2]) * d) + lane;
  const float4* b3 = reinterpret_cast<const float4*>(e + static_cast<int64_t>(j[3]) * d) + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  auto acc = [](float s, const float4& x, const float4& y) {
    float t;
    t = x.x - y.x; s = fmaf(t, t, s);
    t = x.y - y.y; s = fmaf(t, t, s);
    t = x.z - y.z; s = fmaf(t, t, s);
    t = x.w - y.w; s = fmaf(t, t, s);
    return s;
  };
#pragma unroll 4
  for (int i = 0; i < d / 128; ++i) {
    const float4 x = __ldg(a + 32 * i), y0 = __ldg(b0 + 32 * i), y1 = __ldg(b1 + 32 * i), y2 = __ldg(b2 + 32 * i),
                 y3 = __ldg(b3 + 32 * i);
    s0 = acc(s0, x, y0);
    s1 = acc(s1, x, y1);
    s2 = acc(s2, x, y2);
    s3 = acc(s3, x, y3);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  out[0] = s0; out[1] = s1; out[2] = s2; out[3] = s3;
}

// Re-scan of ONE record slot: every row the slot covers (64 columns of the row view, 32 rows of the column view)
// with the wanted label relation is measured in float32 (four rows per trip, all loads of a trip in flight; rows
// of other shapes: one row per lane); the rows within the float32 error of the slot's best are then re-evaluated in
// float64 by the same exact_d2() every other candidate goes through (so exact duplicates compare equal bit for bit
// and resolve to the lowest index).  Used when the slot is SATURATED: its second entry is itself a contender, so
// the slot may hide further contenders behind its top-2 (three duplicates of the hardest negative in adjacent rows
// -- the reference's sampler draws with replacement, embedding_net/datagenerators.py:205 -- or three near-ties).
// A hidden entry's packed key is never better than the slot's second one, so "second entry outside the band"
// proves that nothing hidden matters.
__device__ __noinline__ BhPick bh_rescan_slot(const float* __restrict__ emb, const int32_t* __restrict__ labels,
                                              int64_t B, int d, int64_t row, int32_t la, int t, int my_tile,
                                              bool want_same, int lane, BhPick win) {
  const int tile = t >> 2, slot = t & 3;
  const bool col_view = tile < my_tile;
  const int len = col_view ? 32 : 64;
  const int64_t j0 = static_cast<int64_t>(tile) * tc::BN + slot * len;
  const bool coop = (d & 127) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0;
  const float rel = static_cast<float>(d / 4 + 16) * 2.4e-7f;  // both the best and the contender carry the error
  for (int64_t jb = j0; jb < j0 + len; jb += 32) {
    const int64_t j = jb + lane;
    const bool ok = j < B && j != row;
    const bool want = ok && ((__ldg(&labels[ok ? j : 0]) == la) == want_same);
    unsigned wm = __ballot_sync(0xffffffffu, want);
    if (wm == 0) continue;
    float f = want_same ? -1.f : kBig;
    if (coop) {
      while (wm) {
        int l[4], jj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          l[u] = wm ? __ffs(wm) - 1 : l[0];
          if (wm) wm &= wm - 1;
          jj[u] = static_cast<int>(jb) + l[u];
        }
        float f4[4];
        warp_d2_f32x4(emb, d, row, jj, lane, f4);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (lane == l[u]) f = f4[u];
      }
    } else if (want) {
      f = lane_d2_f32(emb, d, row, j);
    }
    float best = f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, best, o);
      best = want_same ? fmaxf(best, other) : fminf(best, other);
    }
    const bool surv = want && (want_same ? f >= best - best * rel - 1e-30f : f <= best + best * rel + 1e-30f);
    unsigned m = __ballot_sync(0xffffffffu, surv);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int ci = static_cast<int>(jb) + src;
      const double d2 = exact_d2(emb, d, row, ci, lane);
      if (want_same) bh_update_max(win, d2, ci);
      else bh_update_min(win, d2, ci);
    }
  }
  return win;
}

// General resolver (one warp per anchor): every record entry inside the band is re-evaluated exactly, saturated
// slots are re-scanned.  Ties resolve to the lowest index because every candidate that can tie is evaluated and the
// (d2, index) order is applied on the exact values -- the order of the packed keys (which for negative proxies
// favours the HIGHER in-tile index among equal values) never decides anything.
struct BhPair {
  BhPick pos, neg;
};
__device__ __noinline__ BhPair bh_resolve(const float* __restrict__ emb, const int32_t* __restrict__ labels,
                                          const BhCand* __restrict__ mine, int n_cand, int my_tile, int64_t row,
                                          int64_t B, int d, BhThr thr, int lane) {
  BhPair r{BhPick{-1.0, -1}, BhPick{1e300, -1}};
  const int32_t la = labels[row];
#pragma unroll 1
  for (int t0 = 0; t0 < n_cand; t0 += 32) {
    const int t = t0 + lane;
    float4 v = make_float4(-kBig, -kBig, kBig, kBig);
    // slots 2,3 exist only for tiles left of the anchor's own (column view); see EpBatchHard
    if (t < n_cand && ((t & 3) < 2 || (t >> 2) < my_tile)) v = __ldcg(reinterpret_cast<const float4*>(mine + t));
    const bool cp1 = bh_valid(v.x) && v.x >= thr.p, cp2 = bh_valid(v.y) && v.y >= thr.p;
    const bool cn1 = bh_valid(v.z) && v.z <= thr.n, cn2 = bh_valid(v.w) && v.w <= thr.n;
    const int ip = (t >> 2) * tc::BN + static_cast<int>(__float_as_uint(v.x) & 0x7Fu);
    const int in = (t >> 2) * tc::BN + static_cast<int>(__float_as_uint(v.z) & 0x7Fu);
    unsigned m = __ballot_sync(0xffffffffu, cp1 && !cp2);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int ci = __shfl_sync(0xffffffffu, ip, src);
      bh_update_max(r.pos, exact_d2(emb, d, row, ci, lane), ci);
    }
    m = __ballot_sync(0xffffffffu, cn1 && !cn2);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int ci = __shfl_sync(0xffffffffu, in, src);
      bh_update_min(r.neg, exact_d2(emb, d, row, ci, lane), ci);
    }
    m = __ballot_sync(0xffffffffu, cp2);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      r.pos = bh_rescan_slot(emb, labels, B, d, row, la, t0 + src, my_tile, true, lane, r.pos);
    }
    m = __ballot_sync(0xffffffffu, cn2);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      r.neg = bh_rescan_slot(emb, labels, B, d, row, la, t0 + src, my_tile, false, lane, r.neg);
    }
  }
  return r;
}

// No other-label row at all.  Moindrot's min(D + rowmax * (1 - mask_neg)) then degenerates to the row maximum; a
// degenerate batch, handled exactly by a brute-force scan (never on a training path).
__device__ __noinline__ BhPick bh_row_maximum(const float* __restrict__ emb, int64_t B, int d, int64_t row, int lane) {
  BhPick rmx{-1.0, -1};
  for (int64_t j = 0; j < B; ++j) {
    if (j == row) continue;
    const double d2 = exact_d2(emb, d, row, j, lane);
    if (rmx.idx < 0 || d2 > rmx.d2) { rmx.d2 = d2; rmx.idx = static_cast<int>(j); }
  }
  return rmx.idx >= 0 ? rmx : BhPick{0.0, -1};
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// gemb += s * (e_i - e_j) on row `dst` (atomically: several anchors may select the same row)
__device__ __noinline__ void red_axpy_diff(float* __restrict__ gemb, const float* __restrict__ emb, int d,
                                              int64_t dst, int64_t i, int64_t j, float s, int lane) {
  const float* a = emb + i * d;
  const float* b = emb + j * d;
  float* g = gemb + dst * d;
  if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0 && (reinterpret_cast<uintptr_t>(gemb) & 15) == 0) {
    for (int c = lane * 4; c < d; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(a + c), y = *reinterpret_cast<const float4*>(b + c);
      red_add_v4(g + c, s * (x.x - y.x), s * (x.y - y.y), s * (x.z - y.z), s * (x.w - y.w));
    }
  } else {
    for (int c = lane; c < d; c += 32) atomicAdd(g + c, s * (a[c] - b[c]));
  }
}

// Deterministic mean over all anchors: per-block partials, the last block to finish adds them in a fixed-shape
// tree (the result does not depend on which block finishes last).  Called by all 256 threads of every block.
__device__ __forceinline__ void block_mean(double hinge, double* sh, double* __restrict__ partial,
                                           unsigned* __restrict__ counter, float* __restrict__ loss, int64_t B) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  __shared__ bool is_last;
  if (lane == 0) sh[warp] = hinge;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += sh[w];
    partial[blockIdx.x] = s;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double tot = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) tot += __ldcg(&partial[b]);
    tot = warp_sum(tot);
    __syncthreads();
    if (lane == 0) sh[warp] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += sh[w];
      loss[0] = static_cast<float>(s / static_cast<double>(B));
      *counter = 0;  // re-armed for the next call on this workspace
    }
  }
}

// kGrad: also accumulate d loss / d emb into a ZEROED gemb (fused loss + gradient: the rows of the selected
// positive / negative are already hot from the exact re-evaluation).
template <bool kGrad>
__global__ void batch_hard_finalize_kernel(const float* __restrict__ emb, const int32_t* __restrict__ labels,
                                           const float* __restrict__ norms, const BhCand* __restrict__ cand,
                                           int64_t B, int d, int tiles_n, float margin, int squared, int soft,
                                           float band_c, int32_t* __restrict__ hp_idx, int32_t* __restrict__ hn_idx,
                                           float* __restrict__ hp_out, float* __restrict__ hn_out,
                                           float* __restrict__ coef, double* __restrict__ partial,
                                           unsigned* __restrict__ counter, float* __restrict__ loss,
                                           const float* __restrict__ gloss, float* __restrict__ gemb) {
  __shared__ double sh[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  double hinge = 0.0;
  if (row < B) {
    const int n_cand = tiles_n * BH_SLOTS;
    const int my_tile = static_cast<int>(row / tc::BM);
    const BhCand* mine = cand + row * n_cand;
    const float na = norms[row];
    // slots 2,3 exist only for tiles left of the anchor's own (column view); see EpBatchHard
    auto slot_valid = [&](int t) { return t < n_cand && ((t & 3) < 2 || (t >> 2) < my_tile); };
    // pass 1: best proxies over all records
    float bp = -kBig, bn = kBig;
    for (int t = lane; t < n_cand; t += 32) {
      if (!slot_valid(t)) continue;
      const float4 v = *reinterpret_cast<const float4*>(mine + t);
      bp = fmaxf(bp, v.x);  // p1 >= p2
      bn = fminf(bn, v.z);  // n1 <= n2
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bp = fmaxf(bp, __shfl_xor_sync(0xffffffffu, bp, o));
      bn = fminf(bn, __shfl_xor_sync(0xffffffffu, bn, o));
    }
    // pass 2: exact re-evaluation of everything inside the band, saturated slots re-scanned
    const BhPair won = bh_resolve(emb, labels, mine, n_cand, my_tile, row, B, d, bh_thresholds(na, bp, bn, band_c), lane);
    BhPick pos = won.pos, neg = won.neg;
    if (neg.idx < 0) neg = bh_row_maximum(emb, B, d, row, lane);
    const double hp = pos.idx >= 0 ? (squared ? pos.d2 : sqrt(pos.d2)) : 0.0;
    const double hn = neg.idx >= 0 ? (squared ? neg.d2 : sqrt(neg.d2)) : 0.0;
    const double z = hp - hn;
    double g;
    if (soft) {
      hinge = z > 0 ? z + log1p(exp(-z)) : log1p(exp(z));
      g = 1.0 / (1.0 + exp(-z));
    } else {
      hinge = fmax(z + static_cast<double>(margin), 0.0);
      g = (z + static_cast<double>(margin)) >= 0.0 ? 1.0 : 0.0;
    }
    if (lane == 0) {
      hp_idx[row] = pos.idx;
      hn_idx[row] = neg.idx;
      hp_out[row] = static_cast<float>(hp);
      hn_out[row] = static_cast<float>(hn);
      coef[row] = static_cast<float>(g / static_cast<double>(B));
    }
    if (kGrad) {
      const float gg = static_cast<float>(g / static_cast<double>(B)) * (gloss ? gloss[0] : 1.0f);
      if (gg != 0.f) {
        const float hpf = static_cast<float>(hp), hnf = static_cast<float>(hn);
        const float sp = pos.idx >= 0 ? (squared ? 2.f * gg : (hpf > 0.f ? gg / hpf : 0.f)) : 0.f;
        const float sn = neg.idx >= 0 ? (squared ? 2.f * gg : (hnf > 0.f ? gg / hnf : 0.f)) : 0.f;
        if (sp != 0.f) {
          red_axpy_diff(gemb, emb, d, row, row, pos.idx, sp, lane);
          red_axpy_diff(gemb, emb, d, pos.idx, row, pos.idx, -sp, lane);
        }
        if (sn != 0.f) {
          red_axpy_diff(gemb, emb, d, row, row, neg.idx, -sn, lane);
          red_axpy_diff(gemb, emb, d, neg.idx, row, neg.idx, sn, lane);
        }
      }
    }
  }
  block_mean(hinge, sh, partial, counter, loss, B);
}

// ---- fast finalize: d = 128 * DV, at most 128 candidate records per anchor (B <= 4096) ----------------------
// The generic kernel above is a chain of dependent round trips per warp (records, records again, norms, rows for
// each exact distance, rows again for each gradient term; ncu round 1: 8.6 long-scoreboard stalls per issue, 31 us
// at B = 4096 with one wave of warps).  Here every stage is ONE batch of independent loads: all records into
// registers, the 16 candidate norms, then the anchor / positive / negative rows (kept in registers and reused for
// the two exact distances AND the gradient).  Anything unusual -- several contenders inside the error band, no
// negative at all -- takes the generic per-candidate path on the same registers.
constexpr int FF_WARPS = 4;  // anchors per block: small blocks so that one slow warp holds back few others
// One warp, one anchor.  Returns true when the anchor is COMPLEX (a saturated slot, three contenders queued on one
// lane, or no negative at all) and was left untouched for the block-level resolver.
template <bool kGrad, int DV>
__device__ __forceinline__ bool bh_fast_anchor(const float* __restrict__ emb, const int32_t* __restrict__ labels,
                                               const float* __restrict__ norms, const BhCand* __restrict__ cand,
                                               int64_t B, int tiles_n, float margin, int squared, int soft,
                                               float band_c, int32_t* __restrict__ hp_idx,
                                               int32_t* __restrict__ hn_idx, float* __restrict__ hp_out,
                                               float* __restrict__ hn_out, float* __restrict__ coef,
                                               double* __restrict__ hinge_all, const float* __restrict__ gloss,
                                               float* __restrict__ gemb, int64_t row, int lane) {
  constexpr int d = 128 * DV;
  double hinge = 0.0;
  {
    const int n_cand = tiles_n * BH_SLOTS;  // <= 128
    const int my_tile = static_cast<int>(row / tc::BM);
    const BhCand* mine = cand + row * n_cand;
    const float na = norms[row];
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = lane + 32 * u;
      v[u] = make_float4(-kBig, -kBig, kBig, kBig);
      if (t < n_cand && ((t & 3) < 2 || (t >> 2) < my_tile))  // slots 2,3: column view, tiles left of the anchor's
        v[u] = __ldcg(reinterpret_cast<const float4*>(mine + t));
    }
    float bp = fmaxf(fmaxf(v[0].x, v[1].x), fmaxf(v[2].x, v[3].x));  // p1 >= p2, n1 <= n2 inside a record
    float bn = fminf(fminf(v[0].z, v[1].z), fminf(v[2].z, v[3].z));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bp = fmaxf(bp, __shfl_xor_sync(0xffffffffu, bp, o));
      bn = fminf(bn, __shfl_xor_sync(0xffffffffu, bn, o));
    }
    // Contenders inside the error band of the best proxy (bh_thresholds(): the band needs the candidate's norm, which
    // the proxy itself bounds -- no dependent gather of 16 norms per lane, three square roots per anchor instead of
    // sixteen IEEE ones per lane; ncu, round 1: the kernel is issue bound, ~1000 instructions per anchor).  The index
    // inside the record's tile rides in the low mantissa bits.  Each lane queues up to two contenders per kind.  A
    // record whose SECOND entry is a contender marks a saturated slot (it may hide more): general resolver.
    int pc = 0, nc = 0, pi0 = -1, pi1 = -1, ni0 = -1, ni1 = -1;
    fin_sample(row, lane, 1);  // (trace builds) this warp has its records
    const BhThr thr = bh_thresholds(na, bp, bn, band_c);
    const float thr_p = thr.p, thr_n = thr.n;
    bool sat = false;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      sat = sat || (bh_valid(v[u].y) && v[u].y >= thr_p) || (bh_valid(v[u].w) && v[u].w <= thr_n);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int tile = (lane + 32 * u) >> 2;
      const float key[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // branch free (selects): the divergence bookkeeping of the branchy form was 16 % of the stall samples
        const int idx = tile * tc::BN + static_cast<int>(__float_as_uint(key[e]) & 0x7Fu);
        if (e < 2) {
          const bool c = bh_valid(key[e]) && key[e] >= thr_p;
          pi1 = c ? pi0 : pi1;
          pi0 = c ? idx : pi0;
          pc += c ? 1 : 0;
        } else {
          const bool c = bh_valid(key[e]) && key[e] <= thr_n;
          ni1 = c ? ni0 : ni1;
          ni0 = c ? idx : ni0;
          nc += c ? 1 : 0;
        }
      }
    }
    BhPick pos{-1.0, -1}, neg{1e300, -1};
    float4 a[DV], pr[DV], nr[DV];
    int rounds = 0;
    const bool overflow = __any_sync(0xffffffffu, sat || pc > 2 || nc > 2);
    const bool no_neg = !__any_sync(0xffffffffu, nc > 0);
    // A saturated slot, three contenders queued on one lane, or no negative at all: the anchor goes on the work list
    // of the block-level resolver (a whole block per anchor).  Resolving it here, one warp per anchor (tried twice in
    // round 2, the second time with the four-rows-per-trip float32 filter of bh_rescan_slot), made the ~5 % such
    // anchors the critical path of this kernel: 51 us instead of 12 + 17 us for the two kernels.
    if (overflow || no_neg) return true;
    {
      // Each round re-evaluates one positive and one negative contender exactly, all row loads of the round in one
      // batch.  The usual case is a single round (one contender each; none for an anchor alone in its class).
      const float4* arow = reinterpret_cast<const float4*>(emb + row * d) + lane;
#pragma unroll
      for (int i = 0; i < DV; ++i) a[i] = arow[32 * i];
#pragma unroll 1
      for (;;) {
        const unsigned pm = __ballot_sync(0xffffffffu, pc > 0), nm = __ballot_sync(0xffffffffu, nc > 0);
        if ((pm | nm) == 0) break;
        int p_idx = -1, n_idx = -1;
        if (pm) {
          const int src = __ffs(pm) - 1;
          p_idx = __shfl_sync(0xffffffffu, pi0, src);
          if (lane == src) { pi0 = pi1; --pc; }
        }
        if (nm) {
          const int src = __ffs(nm) - 1;
          n_idx = __shfl_sync(0xffffffffu, ni0, src);
          if (lane == src) { ni0 = ni1; --nc; }
        }
        const float4* prow = reinterpret_cast<const float4*>(emb + static_cast<int64_t>(p_idx >= 0 ? p_idx : row) * d) + lane;
        const float4* nrow = reinterpret_cast<const float4*>(emb + static_cast<int64_t>(n_idx >= 0 ? n_idx : row) * d) + lane;
#pragma unroll
        for (int i = 0; i < DV; ++i) pr[i] = prow[32 * i];
#pragma unroll
        for (int i = 0; i < DV; ++i) nr[i] = nrow[32 * i];
        double dp = 0.0, dn = 0.0;
#pragma unroll
        for (int i = 0; i < DV; ++i) {
          // same element order as exact_d2(): c = lane * 4 + 128 * i + {0, 1, 2, 3}
          double t;
          t = static_cast<double>(a[i].x) - static_cast<double>(pr[i].x); dp = fma(t, t, dp);
          t = static_cast<double>(a[i].y) - static_cast<double>(pr[i].y); dp = fma(t, t, dp);
          t = static_cast<double>(a[i].z) - static_cast<double>(pr[i].z); dp = fma(t, t, dp);
          t = static_cast<double>(a[i].w) - static_cast<double>(pr[i].w); dp = fma(t, t, dp);
          t = static_cast<double>(a[i].x) - static_cast<double>(nr[i].x); dn = fma(t, t, dn);
          t = static_cast<double>(a[i].y) - static_cast<double>(nr[i].y); dn = fma(t, t, dn);
          t = static_cast<double>(a[i].z) - static_cast<double>(nr[i].z); dn = fma(t, t, dn);
          t = static_cast<double>(a[i].w) - static_cast<double>(nr[i].w); dn = fma(t, t, dn);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          dp += __shfl_xor_sync(0xffffffffu, dp, o);
          dn += __shfl_xor_sync(0xffffffffu, dn, o);
        }
        if (p_idx >= 0 && (pos.idx < 0 || dp > pos.d2 || (dp == pos.d2 && p_idx < pos.idx))) pos = BhPick{dp, p_idx};
        if (n_idx >= 0 && (neg.idx < 0 || dn < neg.d2 || (dn == neg.d2 && n_idx < neg.idx))) neg = BhPick{dn, n_idx};
        ++rounds;
      }
    }
    // a single round leaves the winners' rows in registers for the gradient
    const bool rows_cached = rounds == 1;
    fin_sample(row, lane, 2);  // (trace builds) this warp has its exact distances
    const double hp = pos.idx >= 0 ? (squared ? pos.d2 : sqrt(pos.d2)) : 0.0;
    const double hn = neg.idx >= 0 ? (squared ? neg.d2 : sqrt(neg.d2)) : 0.0;
    const double z = hp - hn;
    double g;
    if (soft) {
      hinge = z > 0 ? z + log1p(exp(-z)) : log1p(exp(z));
      g = 1.0 / (1.0 + exp(-z));
    } else {
      hinge = fmax(z + static_cast<double>(margin), 0.0);
      g = (z + static_cast<double>(margin)) >= 0.0 ? 1.0 : 0.0;
    }
    if (lane == 0) {
      hp_idx[row] = pos.idx;
      hn_idx[row] = neg.idx;
      hp_out[row] = static_cast<float>(hp);
      hn_out[row] = static_cast<float>(hn);
      coef[row] = static_cast<float>(g / static_cast<double>(B));
      hinge_all[row] = hinge;
    }
    if (kGrad) {
      const float gg = static_cast<float>(g / static_cast<double>(B)) * (gloss ? gloss[0] : 1.0f);
      if (gg != 0.f) {
        const float hpf = static_cast<float>(hp), hnf = static_cast<float>(hn);
        const float sp = pos.idx >= 0 ? (squared ? 2.f * gg : (hpf > 0.f ? gg / hpf : 0.f)) : 0.f;
        const float sn = neg.idx >= 0 ? (squared ? 2.f * gg : (hnf > 0.f ? gg / hnf : 0.f)) : 0.f;
        if (rows_cached && (reinterpret_cast<uintptr_t>(gemb) & 15) == 0) {
          // same products and the same four atomic adds per element as red_axpy_diff(), from registers
          float* ga = gemb + row * d + 4 * lane;
          float* gp = gemb + static_cast<int64_t>(pos.idx >= 0 ? pos.idx : row) * d + 4 * lane;
          float* gn = gemb + static_cast<int64_t>(neg.idx) * d + 4 * lane;
#pragma unroll
          for (int i = 0; i < DV; ++i) {
            if (sp != 0.f) {
              const float x = a[i].x - pr[i].x, y = a[i].y - pr[i].y, zz = a[i].z - pr[i].z, w = a[i].w - pr[i].w;
              red_add_v4(ga + 128 * i, sp * x, sp * y, sp * zz, sp * w);
              red_add_v4(gp + 128 * i, -sp * x, -sp * y, -sp * zz, -sp * w);
            }
            if (sn != 0.f) {
              const float x = a[i].x - nr[i].x, y = a[i].y - nr[i].y, zz = a[i].z - nr[i].z, w = a[i].w - nr[i].w;
              red_add_v4(ga + 128 * i, -sn * x, -sn * y, -sn * zz, -sn * w);
              red_add_v4(gn + 128 * i, sn * x, sn * y, sn * zz, sn * w);
            }
          }
        } else {
          if (sp != 0.f) {
            red_axpy_diff(gemb, emb, d, row, row, pos.idx, sp, lane);
            red_axpy_diff(gemb, emb, d, pos.idx, row, pos.idx, -sp, lane);
          }
          if (sn != 0.f) {
            red_axpy_diff(gemb, emb, d, row, row, neg.idx, -sn, lane);
            red_axpy_diff(gemb, emb, d, neg.idx, row, neg.idx, sn, lane);
          }
        }
      }
    }
  }
  return false;
}

template <bool kGrad, int DV>
__global__ void __launch_bounds__(FF_WARPS * 32, 7)  // 72 registers: 7 blocks = 28 warps per SM, all 4096 anchors of the headline shape resident in ONE wave
batch_hard_finalize_fast_kernel(const float* __restrict__ emb, const int32_t* __restrict__ labels,
                                const float* __restrict__ norms, const BhCand* __restrict__ cand, int64_t B,
                                int tiles_n, float margin, int squared, int soft, float band_c,
                                int32_t* __restrict__ hp_idx, int32_t* __restrict__ hn_idx,
                                float* __restrict__ hp_out, float* __restrict__ hn_out, float* __restrict__ coef,
                                double* __restrict__ hinge_all, int32_t* __restrict__ work_list,
                                unsigned* __restrict__ counters, const float* __restrict__ gloss,
                                float* __restrict__ gemb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * FF_WARPS + warp;
  if (row >= B) return;
  if (lane == 0) fin_stamp(0, true);
  fin_sample(row, lane, 0);
  if (bh_fast_anchor<kGrad, DV>(emb, labels, norms, cand, B, tiles_n, margin, squared, soft, band_c, hp_idx, hn_idx,
                                hp_out, hn_out, coef, hinge_all, gloss, gemb, row, lane) &&
      lane == 0)
    work_list[atomicAdd(&counters[0], 1u)] = static_cast<int32_t>(row);
  if (lane == 0) fin_stamp(1, false);
  fin_sample(row, lane, 3);
}

// ---- slow finalize: the anchors the fast kernel put on the work list, one BLOCK per anchor -------------------
// Eight warps: four of them hold the anchor's (at most 128) records; a saturated slot's candidates are split over all,
// measured in float32 (four rows per trip, all loads of a trip in flight together), and the ones within the float32
// error of the slot's best are re-evaluated by exact_d2() -- the arithmetic every other candidate goes through, so
// exact duplicates compare equal bit for bit and resolve to the lowest index.  The last block to finish adds the
// per-anchor hinge values of BOTH kernels in a fixed order (deterministic mean).
constexpr int FS_WARPS = 8;

// hinge / outputs / gradient of one anchor from its resolved picks (one warp); returns the hinge value
template <bool kGrad>
__device__ __forceinline__ double bh_finish(const float* __restrict__ emb, int d, int64_t B, int64_t row, BhPick pos,
                                            BhPick neg, float margin, int squared, int soft,
                                            int32_t* __restrict__ hp_idx, int32_t* __restrict__ hn_idx,
                                            float* __restrict__ hp_out, float* __restrict__ hn_out,
                                            float* __restrict__ coef, const float* __restrict__ gloss,
                                            float* __restrict__ gemb, int lane, int term = -1) {
  // term: -1 = outputs and all four gradient terms (one warp does everything); 0..3 = only that gradient term, and
  // the outputs with term 0 (the four terms of one anchor spread over four warps)
  const double hp = pos.idx >= 0 ? (squared ? pos.d2 : sqrt(pos.d2)) : 0.0;
  const double hn = neg.idx >= 0 ? (squared ? neg.d2 : sqrt(neg.d2)) : 0.0;
  const double z = hp - hn;
  double g, hinge;
  if (soft) {
    hinge = z > 0 ? z + log1p(exp(-z)) : log1p(exp(z));
    g = 1.0 / (1.0 + exp(-z));
  } else {
    hinge = fmax(z + static_cast<double>(margin), 0.0);
    g = (z + static_cast<double>(margin)) >= 0.0 ? 1.0 : 0.0;
  }
  if (lane == 0 && term <= 0) {
    hp_idx[row] = pos.idx;
    hn_idx[row] = neg.idx;
    hp_out[row] = static_cast<float>(hp);
    hn_out[row] = static_cast<float>(hn);
    coef[row] = static_cast<float>(g / static_cast<double>(B));
  }
  if (kGrad) {
    const float gg = static_cast<float>(g / static_cast<double>(B)) * (gloss ? gloss[0] : 1.0f);
    if (gg != 0.f) {
      const float hpf = static_cast<float>(hp), hnf = static_cast<float>(hn);
      const float sp = pos.idx >= 0 ? (squared ? 2.f * gg : (hpf > 0.f ? gg / hpf : 0.f)) : 0.f;
      const float sn = neg.idx >= 0 ? (squared ? 2.f * gg : (hnf > 0.f ? gg / hnf : 0.f)) : 0.f;
      if (sp != 0.f) {
        if (term < 0 || term == 0) red_axpy_diff(gemb, emb, d, row, row, pos.idx, sp, lane);
        if (term < 0 || term == 1) red_axpy_diff(gemb, emb, d, pos.idx, row, pos.idx, -sp, lane);
      }
      if (sn != 0.f) {
        if (term < 0 || term == 2) red_axpy_diff(gemb, emb, d, row, row, neg.idx, -sn, lane);
        if (term < 0 || term == 3) red_axpy_diff(gemb, emb, d, neg.idx, row, neg.idx, sn, lane);
      }
    }
  }
  return hinge;
}

// Shared state of one block resolving one listed anchor (NW warps)
template <int NW>
struct BhBlockShared {
  float bp[NW], bn[NW];
  double d2[2][NW];
  int idx[2][NW];
  int nsat;
  int sat[2 * 128];
  float f[64];
};

// All NW warps of a block resolve ONE anchor: records over the first four warps, a saturated slot's candidates over
// all of them; outputs, hinge value and gradient are written by warps 0..3.  Contains block-wide barriers.
template <bool kGrad, int NW>
__device__ __forceinline__ void bh_block_resolve(BhBlockShared<NW>& sm, const float* __restrict__ emb,
                                                 const int32_t* __restrict__ labels, const float* __restrict__ norms,
                                                 const BhCand* __restrict__ cand, int64_t B, int d, int tiles_n,
                                                 float margin, int squared, int soft, float band_c, int64_t row,
                                                 int32_t* __restrict__ hp_idx, int32_t* __restrict__ hn_idx,
                                                 float* __restrict__ hp_out, float* __restrict__ hn_out,
                                                 float* __restrict__ coef, double* __restrict__ hinge_all,
                                                 const float* __restrict__ gloss, float* __restrict__ gemb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cand = tiles_n * BH_SLOTS;  // <= 128: the records live in warps 0..3
  const int my_tile = static_cast<int>(row / tc::BM);
  const BhCand* mine = cand + row * n_cand;
  const float na = norms[row];
  const int32_t la = labels[row];
  const int t = warp * 32 + lane;
  float4 v = make_float4(-kBig, -kBig, kBig, kBig);
  if (t < n_cand && ((t & 3) < 2 || (t >> 2) < my_tile)) v = __ldcg(reinterpret_cast<const float4*>(mine + t));
  float bp = v.x, bn = v.z;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bp = fmaxf(bp, __shfl_xor_sync(0xffffffffu, bp, o));
    bn = fminf(bn, __shfl_xor_sync(0xffffffffu, bn, o));
  }
  if (lane == 0) { sm.bp[warp] = bp; sm.bn[warp] = bn; }
  if (threadIdx.x == 0) sm.nsat = 0;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < NW; ++w) { bp = fmaxf(bp, sm.bp[w]); bn = fminf(bn, sm.bn[w]); }
  const BhThr thr = bh_thresholds(na, bp, bn, band_c);
  const bool cp1 = bh_valid(v.x) && v.x >= thr.p, cp2 = bh_valid(v.y) && v.y >= thr.p;
  const bool cn1 = bh_valid(v.z) && v.z <= thr.n, cn2 = bh_valid(v.w) && v.w <= thr.n;
  const int ip = (t >> 2) * tc::BN + static_cast<int>(__float_as_uint(v.x) & 0x7Fu);
  const int in = (t >> 2) * tc::BN + static_cast<int>(__float_as_uint(v.z) & 0x7Fu);
  BhPick pos{-1.0, -1}, neg{1e300, -1};
  // entries of unsaturated slots: exact, one by one
  unsigned m = __ballot_sync(0xffffffffu, cp1 && !cp2);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const int ci = __shfl_sync(0xffffffffu, ip, src);
    bh_update_max(pos, exact_d2(emb, d, row, ci, lane), ci);
  }
  m = __ballot_sync(0xffffffffu, cn1 && !cn2);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const int ci = __shfl_sync(0xffffffffu, in, src);
    bh_update_min(neg, exact_d2(emb, d, row, ci, lane), ci);
  }
  // saturated slots (second entry inside the band): queued for the whole block
  if (cp2) sm.sat[atomicAdd(&sm.nsat, 1)] = t * 2 + 1;
  if (cn2) sm.sat[atomicAdd(&sm.nsat, 1)] = t * 2;
  __syncthreads();
  const int n_sat = sm.nsat;
  const float rel = static_cast<float>(d / 32 + 16) * 2.4e-7f;  // float32 error of both values being compared
  for (int s = 0; s < n_sat; ++s) {
    const int code = sm.sat[s];
    const bool want_same = (code & 1) != 0;
    const int tt = code >> 1, tile = tt >> 2, slot = tt & 3;
    const int len = tile < my_tile ? 32 : 64;  // column view: 32-row quarters; row view: 64-column halves
    const int j0 = tile * tc::BN + slot * len;
    const int per = len / NW;                  // candidates per warp
    const int64_t jq = static_cast<int64_t>(j0) + warp * per + lane;
    const bool ok = lane < per && jq < B && jq != row;
    const bool want = ok && ((__ldg(&labels[ok ? jq : 0]) == la) == want_same);
    unsigned wm = __ballot_sync(0xffffffffu, want);
    float myf = want_same ? -1.f : kBig;       // "not a candidate"
    while (wm) {
      int l[4], jj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        l[u] = wm ? __ffs(wm) - 1 : l[0];
        if (wm) wm &= wm - 1;
        jj[u] = j0 + warp * per + l[u];
      }
      float f4[4];
      warp_d2_f32x4(emb, d, row, jj, lane, f4);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (lane == l[u]) myf = f4[u];
    }
    if (lane < per) sm.f[warp * per + lane] = myf;
    __syncthreads();
    const float f0 = lane < len ? sm.f[lane] : (want_same ? -1.f : kBig);
    const float f1 = lane + 32 < len ? sm.f[lane + 32] : (want_same ? -1.f : kBig);
    float best = want_same ? fmaxf(f0, f1) : fminf(f0, f1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, best, o);
      best = want_same ? fmaxf(best, other) : fminf(best, other);
    }
    const float lim = want_same ? best - best * rel - 1e-30f : best + best * rel + 1e-30f;
    const bool v0 = want_same ? f0 >= 0.f : f0 < 1e38f, v1 = want_same ? f1 >= 0.f : f1 < 1e38f;
    const unsigned m0 = __ballot_sync(0xffffffffu, v0 && (want_same ? f0 >= lim : f0 <= lim));
    const unsigned m1 = __ballot_sync(0xffffffffu, v1 && (want_same ? f1 >= lim : f1 <= lim));
    int n_surv = 0;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      unsigned mm = h ? m1 : m0;
      while (mm) {
        const int src = __ffs(mm) - 1;
        mm &= mm - 1;
        if ((n_surv++ % NW) == warp) {  // survivors round-robin over the warps
          const int ci = j0 + h * 32 + src;
          const double d2 = exact_d2(emb, d, row, ci, lane);
          if (want_same) bh_update_max(pos, d2, ci);
          else bh_update_min(neg, d2, ci);
        }
      }
    }
    __syncthreads();  // sm.f is rewritten by the next slot
  }
  // merge the warps' picks ((d2, index) order: associative and commutative, so the result is deterministic)
  if (lane == 0) {
    sm.d2[0][warp] = pos.d2; sm.idx[0][warp] = pos.idx;
    sm.d2[1][warp] = neg.d2; sm.idx[1][warp] = neg.idx;
  }
  __syncthreads();
  BhPick P{-1.0, -1}, N{1e300, -1};
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    if (sm.idx[0][w] >= 0) bh_update_max(P, sm.d2[0][w], sm.idx[0][w]);
    if (sm.idx[1][w] >= 0) bh_update_min(N, sm.d2[1][w], sm.idx[1][w]);
  }
  if (N.idx < 0) {  // block-uniform; degenerate batch without any other label
    __syncthreads();
    if (warp == 0) {
      N = bh_row_maximum(emb, B, d, row, lane);
      if (lane == 0) { sm.d2[1][0] = N.d2; sm.idx[1][0] = N.idx; }
    }
    __syncthreads();
    N = BhPick{sm.d2[1][0], sm.idx[1][0]};
  }
  if (warp < 4) {  // outputs + the four gradient terms, one per warp
    const double hinge = bh_finish<kGrad>(emb, d, B, row, P, N, margin, squared, soft, hp_idx, hn_idx, hp_out, hn_out,
                                          coef, gloss, gemb, lane, warp);
    if (warp == 0 && lane == 0) hinge_all[row] = hinge;
  }
  __syncthreads();  // shared state is reused by the next anchor
}

// hinge_all[0..B) -> loss, in a fixed order (deterministic), by the calling block
template <int NW>
__device__ __forceinline__ void bh_block_mean(const double* __restrict__ hinge_all, int64_t B, double* s_red,
                                              float* __restrict__ loss) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double tot = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) tot += __ldcg(&hinge_all[i]);
  tot = warp_sum(tot);
  if (lane == 0) s_red[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    for (int w = 0; w < NW; ++w) sum += s_red[w];
    loss[0] = static_cast<float>(sum / static_cast<double>(B));
  }
}

template <bool kGrad>
__global__ void __launch_bounds__(FS_WARPS * 32)
batch_hard_finalize_slow_kernel(const float* __restrict__ emb, const int32_t* __restrict__ labels,
                                const float* __restrict__ norms, const BhCand* __restrict__ cand, int64_t B, int d,
                                int tiles_n, float margin, int squared, int soft, float band_c,
                                int32_t* __restrict__ hp_idx, int32_t* __restrict__ hn_idx,
                                float* __restrict__ hp_out, float* __restrict__ hn_out, float* __restrict__ coef,
                                double* __restrict__ hinge_all, const int32_t* __restrict__ work_list,
                                unsigned* __restrict__ counters, float* __restrict__ loss,
                                const float* __restrict__ gloss, float* __restrict__ gemb) {
  __shared__ BhBlockShared<FS_WARPS> sm;
  __shared__ double s_red[FS_WARPS];
  __shared__ bool s_last;
  if (threadIdx.x == 0) fin_stamp(2, true);
  const unsigned n_work = *reinterpret_cast<volatile unsigned*>(&counters[0]);
  for (unsigned k = blockIdx.x; k < n_work; k += gridDim.x) {
#ifdef EN_FIN_TRACE
    unsigned long long t_in = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_in));
#endif
    bh_block_resolve<kGrad, FS_WARPS>(sm, emb, labels, norms, cand, B, d, tiles_n, margin, squared, soft, band_c,
                                      work_list[k], hp_idx, hn_idx, hp_out, hn_out, coef, hinge_all, gloss, gemb);
#ifdef EN_FIN_TRACE
    if (g_fin_trace != nullptr && threadIdx.x == 0 && k < 1024) {  // per listed anchor: resolve ns, start stamp
      unsigned long long t_out;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_out));
      g_fin_trace[8 + 2 * k] = t_out - t_in;
      g_fin_trace[9 + 2 * k] = t_in;
    }
#endif
  }
  if (threadIdx.x == 0 && blockIdx.x < n_work) fin_stamp(3, false);
  // ---- mean over all anchors, in a fixed order, by the last WORKING block (blocks without an anchor leave at
  // once) -- or by block 0 when the list is empty
  const unsigned n_working = n_work < gridDim.x ? n_work : gridDim.x;
  if (blockIdx.x >= n_working && !(n_working == 0 && blockIdx.x == 0)) return;
  if (threadIdx.x == 0) {
    if (n_working == 0) {
      s_last = true;
    } else {
      __threadfence();
      s_last = atomicAdd(&counters[1], 1u) == n_working - 1;
    }
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    bh_block_mean<FS_WARPS>(hinge_all, B, s_red, loss);
    if (threadIdx.x == 0) {
      counters[0] = 0;  // re-armed for the next call on this workspace
      counters[1] = 0;
      fin_stamp(4, false);
#ifdef EN_FIN_TRACE
      if (g_fin_trace) g_fin_trace[5] = n_work;
#endif
    }
  }
}

// Backward, stage 1: the anchor's own row, overwritten (no zero-fill pass needed).
//   d mean / d e_i (direct) = coef_i * gl * ( s_p (e_i - e_p) - s_n (e_i - e_n) ),  s = 2 (squared) or 1 / D.
__global__ void batch_hard_bwd_own_kernel(const float* __restrict__ emb, int64_t B, int d, int squared,
                                          const int32_t* __restrict__ hp_idx, const int32_t* __restrict__ hn_idx,
                                          const float* __restrict__ hp, const float* __restrict__ hn,
                                          const float* __restrict__ coef, const float* __restrict__ gloss,
                                          float* __restrict__ gemb) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float g = coef[row] * gloss[0];
  const int p = hp_idx[row], n = hn_idx[row];
  float sp = 0.f, sn = 0.f;
  if (g != 0.f) {
    if (p >= 0) sp = squared ? 2.f * g : (hp[row] > 0.f ? g / hp[row] : 0.f);
    if (n >= 0) sn = squared ? 2.f * g : (hn[row] > 0.f ? g / hn[row] : 0.f);
  }
  const float* ei = emb + row * d;
  const float* ep = emb + static_cast<int64_t>(p >= 0 ? p : 0) * d;
  const float* en_ = emb + static_cast<int64_t>(n >= 0 ? n : 0) * d;
  float* gi = gemb + row * d;
  for (int c = lane; c < d; c += 32) {
    const float v = ei[c];
    gi[c] = sp * (v - ep[c]) - sn * (v - en_[c]);
  }
}


// Backward, stage 2: scatter the mirrored terms into the selected positive / negative rows.
__global__ void batch_hard_bwd_scatter_kernel(const float* __restrict__ emb, int64_t B, int d, int squared,
                                              const int32_t* __restrict__ hp_idx, const int32_t* __restrict__ hn_idx,
                                              const float* __restrict__ hp, const float* __restrict__ hn,
                                              const float* __restrict__ coef, const float* __restrict__ gloss,
                                              float* __restrict__ gemb) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float g = coef[row] * gloss[0];
  if (g == 0.f) return;
  const int p = hp_idx[row], n = hn_idx[row];
  float sp = 0.f, sn = 0.f;
  if (p >= 0) sp = squared ? 2.f * g : (hp[row] > 0.f ? g / hp[row] : 0.f);
  if (n >= 0) sn = squared ? 2.f * g : (hn[row] > 0.f ? g / hn[row] : 0.f);
  const float* ei = emb + row * d;
  const bool vec = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(gemb) & 15) == 0;
  if (sp != 0.f) {
    const float* ep = emb + static_cast<int64_t>(p) * d;
    float* gp = gemb + static_cast<int64_t>(p) * d;
    if (vec) {
      for (int c = lane * 4; c < d; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(ei + c), b = *reinterpret_cast<const float4*>(ep + c);
        red_add_v4(gp + c, -sp * (a.x - b.x), -sp * (a.y - b.y), -sp * (a.z - b.z), -sp * (a.w - b.w));
      }
    } else {
      for (int c = lane; c < d; c += 32) atomicAdd(gp + c, -sp * (ei[c] - ep[c]));
    }
  }
  if (sn != 0.f) {
    const float* en_ = emb + static_cast<int64_t>(n) * d;
    float* gn = gemb + static_cast<int64_t>(n) * d;
    if (vec) {
      for (int c = lane * 4; c < d; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(ei + c), b = *reinterpret_cast<const float4*>(en_ + c);
        red_add_v4(gn + c, sn * (a.x - b.x), sn * (a.y - b.y), sn * (a.z - b.z), sn * (a.w - b.w));
      }
    } else {
      for (int c = lane; c < d; c += 32) atomicAdd(gn + c, sn * (ei[c] - en_[c]));
    }
  }
}

// =====================================================================================================
// Batch-all / all-pairs contrastive: shared pieces
// =====================================================================================================
constexpr int kMaxPos = 63;  // largest supported (class size - 1); bounded by the epilogue's shared-memory budget

// exact_d2() for four rows at once: the same per-row arithmetic (element order, fma chain, butterfly), so the
// results are bit-identical; the loads of all four rows are in flight together.
__device__ __forceinline__ void exact_d2x4(const float* __restrict__ e, int d, int64_t i, const int (&j)[4], int lane,
                                           double (&out)[4]) {
  const float* a = e + i * d;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(e) & 15) == 0) {
    for (int c = lane * 4; c < d; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(a + c);
      float4 y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) y[u] = *reinterpret_cast<const float4*>(e + static_cast<int64_t>(j[u]) * d + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double t;
        t = static_cast<double>(x.x) - static_cast<double>(y[u].x); acc[u] = fma(t, t, acc[u]);
        t = static_cast<double>(x.y) - static_cast<double>(y[u].y); acc[u] = fma(t, t, acc[u]);
        t = static_cast<double>(x.z) - static_cast<double>(y[u].z); acc[u] = fma(t, t, acc[u]);
        t = static_cast<double>(x.w) - static_cast<double>(y[u].w); acc[u] = fma(t, t, acc[u]);
      }
    }
  } else {
    for (int c = lane; c < d; c += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double t = static_cast<double>(a[c]) - static_cast<double>(e[static_cast<int64_t>(j[u]) * d + c]);
        acc[u] = fma(t, t, acc[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) out[u] = warp_sum(acc[u]);
}

// One warp per anchor: list its positives (same label, j != i, ascending j) with exact distances.
// Two phases so that a warp's chain of dependent memory round trips is short (the kernel is a fraction of one wave:
// its duration is ONE warp's latency; round 1/2 launch lists: 23 us, more than a third of the distance GEMM it
// prepares): (1) the label scan, 512 labels per trip with all four 128-bit loads in flight; (2) the distances, four
// positives per trip.
// Rows B .. rows_padded-1 (the launch covers them when the lists are allocated for a whole number of 128-row tiles:
// the tensor-core pair kernel reads the lists of every column of a tile) get empty lists.
__global__ void collect_positives_kernel(const float* __restrict__ emb, const int32_t* __restrict__ labels,
                                         int64_t B, int d, int squared, int cap, float* __restrict__ pos_d,
                                         int32_t* __restrict__ pos_j, int32_t* __restrict__ pos_n,
                                         int32_t* __restrict__ status, int64_t rows_padded,
                                         double* __restrict__ pos_pre = nullptr) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) {
    if (row < rows_padded) {
      for (int s = lane; s < cap; s += 32) pos_d[row * cap + s] = -INFINITY;
      if (lane == 0) pos_n[row] = 0;
    }
    return;
  }
  const int32_t la = labels[row];
  int count = 0;
  int32_t* mine_j = pos_j + row * cap;
  auto emit = [&](int64_t jj) {
    if (count < cap && lane == 0) mine_j[count] = static_cast<int32_t>(jj);
    ++count;
  };
  int64_t j0 = 0;
  if ((reinterpret_cast<uintptr_t>(labels) & 15) == 0) {
    for (; j0 + 512 <= B; j0 += 512) {
      int4 l4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) l4[u] = __ldg(reinterpret_cast<const int4*>(labels + j0 + 128 * u) + lane);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t jb = j0 + 128 * u + 4 * lane;
        const unsigned f = (l4[u].x == la && jb != row ? 1u : 0u) | (l4[u].y == la && jb + 1 != row ? 2u : 0u) |
                           (l4[u].z == la && jb + 2 != row ? 4u : 0u) | (l4[u].w == la && jb + 3 != row ? 8u : 0u);
        unsigned m = __ballot_sync(0xffffffffu, f != 0);
        while (m) {  // ascending j: lanes in order, then the four labels of a lane in order
          const int src = __ffs(m) - 1;
          m &= m - 1;
          unsigned fs = __shfl_sync(0xffffffffu, f, src);
          while (fs) {
            const int sub = __ffs(fs) - 1;
            fs &= fs - 1;
            emit(j0 + 128 * u + 4 * src + sub);
          }
        }
      }
    }
  }
  for (; j0 < B; j0 += 32) {
    const int64_t j = j0 + lane;
    const bool same = j < B && j != row && labels[j] == la;
    unsigned m = __ballot_sync(0xffffffffu, same);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      emit(j0 + src);
    }
  }
  const int n = count < cap ? count : cap;
  __syncwarp();  // lane 0's index stores are visible to the warp
  for (int s0 = 0; s0 < n; s0 += 4) {
    int jj[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) jj[u] = mine_j[s0 + u < n ? s0 + u : s0];
    double d2[4];
    exact_d2x4(emb, d, row, jj, lane, d2);
    if (lane < 4 && s0 + lane < n) {
      const double v = lane == 0 ? d2[0] : lane == 1 ? d2[1] : lane == 2 ? d2[2] : d2[3];
      pos_d[row * cap + s0 + lane] = static_cast<float>(squared ? v : sqrt(v));
    }
  }
  // Lists for the large-class pair kernel (pos_pre given): ordered by DECREASING distance (ties: ascending slot), so
  // that the hinges D_ap + m - D_an > 0 of a given negative are active for a PREFIX of the list, whose length a
  // binary search finds; pos_pre[s] = D_ap(0) + ... + D_ap(s) in float64 turns the hinge sum of that prefix into
  // one lookup.  Slot order carries no other meaning (pos_j / pos_cnt follow; "i in j's list" lookups search).
  if (pos_pre != nullptr) {
    __syncwarp();
    float dv[2];
    int jv[2], rank[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int s = lane + 32 * h;
      dv[h] = s < n ? pos_d[row * cap + s] : -INFINITY;
      jv[h] = s < n ? mine_j[s] : -1;
    }
    for (int t = 0; t < n; ++t) {
      const float dt = __shfl_sync(0xffffffffu, t < 32 ? dv[0] : dv[1], t & 31);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int s = lane + 32 * h;
        rank[h] += (dt > dv[h] || (dt == dv[h] && t < s)) ? 1 : 0;
      }
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (lane + 32 * h < n) {
        pos_d[row * cap + rank[h]] = dv[h];
        mine_j[rank[h]] = jv[h];
      }
    __syncwarp();
    if (lane == 0) {  // n <= 64 sequential float64 additions: a fixed order
      double run = 0.0;
      for (int s = 0; s < n; ++s) {
        run += static_cast<double>(pos_d[row * cap + s]);
        pos_pre[row * cap + s] = run;
      }
    }
  }
  // unused slots hold -inf ("never the harder positive"): readers that walk whole lists need no count check
  for (int s = n + lane; s < cap; s += 32) pos_d[row * cap + s] = -INFINITY;
  if (lane == 0) {
    pos_n[row] = n;
    if (count > cap) atomicMax(status, count);  // caller's max_positives was too small
  }
}

// ---------------------------------------------------------------- batch-all forward epilogue
struct EpBatchAll {
  struct Params {
    const int32_t* labels;
    const float* norms;
    const float* pos_d;    // [B][cap]
    const int32_t* pos_n;  // [B]
    PairPartial* partial;  // [B][n_splits][EPI_H]
    int64_t B;
    int cap, n_splits, squared;
    float margin;
  };
  struct Row {
    int32_t la;
    float na;
    int npos;
    float tile_sum;
    unsigned tile_cnt;
    double sum;
    unsigned long long cnt;
  };
  static constexpr int kSmemBytes = kMaxPos * tc::BM * 4;  // positives + margin, transposed: [slot][row in tile]
  static __device__ void item_begin(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int, int) {
    r.la = valid ? p.labels[row] : 0;
    r.na = valid ? p.norms[row] : 0.f;
    r.npos = valid ? p.pos_n[row] : 0;
    // both column-half threads of a row write the same values to the same slots (idempotent); each thread only
    // ever reads its own row's column, after its own writes
    // (slots past the row's own count hold -inf: the hinge below is then inactive without a per-row trip count)
    float* sm = reinterpret_cast<float*>(ctx.smem);
    for (int s = 0; s < p.cap; ++s)
      sm[s * tc::BM + ctx.erow] = s < r.npos ? p.pos_d[row * p.cap + s] + p.margin : -INFINITY;
    r.sum = 0.0;
    r.cnt = 0;
    r.tile_sum = 0.f;
    r.tile_cnt = 0;
  }
  static __device__ void chunk(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int64_t col0,
                               const float (&dot)[32]) {
    if (col0 >= p.B) return;
    tc::stage_columns(ctx, p.norms, p.labels, col0, p.B);
    const float* sm = reinterpret_cast<const float*>(ctx.smem) + ctx.erow;
    const int ncols = static_cast<int>(p.B - col0 < 32 ? p.B - col0 : 32);
    // Branch free: the negatives' distances first (+inf where the column is not a negative of this anchor), then one
    // uniform pass per positive slot.  The divergent form (per-row trip count, a branch per hinge, an IEEE sqrt per
    // element) ran at twice the time of the contrastive epilogue (ncu launch list, round 1: 201 vs 101 us).
    // d = d2 * rsqrt(d2) as in the backward kernel, so both passes see the same hinge decisions.
    float dn[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float d2 = fmaxf(r.na + ctx.wf[j] - 2.f * dot[j], 0.f);
      float rs;
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(d2, 1e-30f)));
      const float d = p.squared ? d2 : d2 * rs;
      dn[j] = (j < ncols && ctx.wi[j] != r.la) ? d : INFINITY;
    }
    float acc = 0.f;
    unsigned cnt = 0;
#pragma unroll 1
    for (int s = 0; s < p.cap; ++s) {
      const float th = sm[s * tc::BM];  // D_ap + margin (or -inf)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float t = th - dn[j];     // D_ap + margin - D_an
        const bool act = t > 1e-16f;
        acc += act ? t : 0.f;
        cnt += act ? 1u : 0u;
      }
    }
    r.tile_sum += acc;
    r.tile_cnt += cnt;
  }
  static __device__ void tile_end(const Params&, Row& r, const tc::Ctx&, int64_t, bool, int) {
    r.sum += static_cast<double>(r.tile_sum);
    r.cnt += r.tile_cnt;
    r.tile_sum = 0.f;
    r.tile_cnt = 0;
  }
  static __device__ void item_end(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int,
                                  int split) {
    if (valid) p.partial[(row * p.n_splits + split) * tc::EPI_H + ctx.half] = PairPartial{r.sum, r.cnt};
  }
};

// ---------------------------------------------------------------- all-pairs contrastive forward epilogue
struct EpContrastive {
  struct Params {
    const int32_t* labels;
    const float* norms;
    PairPartial* partial;  // [B][n_splits][EPI_H]
    int64_t B;
    int n_splits;
  };
  struct Row {
    int32_t la;
    float na;
    float tile_sum;
    double sum;
  };
  static constexpr int kSmemBytes = 0;
  static __device__ void item_begin(const Params& p, Row& r, const tc::Ctx&, int64_t row, bool valid, int, int) {
    r.la = valid ? p.labels[row] : 0;
    r.na = valid ? p.norms[row] : 0.f;
    r.sum = 0.0;
    r.tile_sum = 0.f;
  }
  static __device__ void chunk(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int64_t col0,
                               const float (&dot)[32]) {
    if (col0 >= p.B) return;
    tc::stage_columns(ctx, p.norms, p.labels, col0, p.B);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int64_t c = col0 + j;
      const float d2 = fmaxf(r.na + ctx.wf[j] - 2.f * dot[j], 1e-7f);  // models.py:225 clamp
      const float m = fmaxf(1.f - sqrtf(d2), 0.f);
      const float term = (ctx.wi[j] == r.la) ? d2 : m * m;
      r.tile_sum += (c < p.B && c != row) ? term : 0.f;
    }
  }
  static __device__ void tile_end(const Params&, Row& r, const tc::Ctx&, int64_t, bool, int) {
    r.sum += static_cast<double>(r.tile_sum);
    r.tile_sum = 0.f;
  }
  static __device__ void item_end(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int,
                                  int split) {
    if (valid) p.partial[(row * p.n_splits + split) * tc::EPI_H + ctx.half] = PairPartial{r.sum, 0ull};
  }
};

// Deterministic reduction of the per-(row, split) partials: one block, fixed order.
// mode 0: batch-all  -> out[0] = sum / (npos + 1e-16), out[1] = npos / (nvalid + 1e-16), stats = {sum,npos,nvalid}
// mode 1: contrastive -> out[0] = sum / (B (B-1))
// Stage 1 (many blocks): block b folds its contiguous chunk of the per-(row, split, half) partials into the chunk's
// first slot, in place and in a fixed order (deterministic).  One block reading 4 MB took 46 us (ncu, round 1).
constexpr int kReduceBlocks = 256;
__global__ void pair_reduce_stage1_kernel(PairPartial* __restrict__ partial, int64_t n_partials, int64_t chunk) {
  __shared__ double s_sum[8];
  __shared__ unsigned long long s_cnt[8];
  const int64_t lo = static_cast<int64_t>(blockIdx.x) * chunk;
  const int64_t hi = lo + chunk < n_partials ? lo + chunk : n_partials;
  double sum = 0.0;
  unsigned long long cnt = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const PairPartial x = partial[i];
    sum += x.sum;
    cnt += x.npos;
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; }
  __syncthreads();  // also: every read of this chunk has completed before its first slot is overwritten
  if (threadIdx.x == 0 && lo < n_partials) {
    double a = 0.0;
    unsigned long long c = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += s_sum[w]; c += s_cnt[w]; }
    partial[lo] = PairPartial{a, c};
  }
}

// Stage 2 (one block): the chunk heads (stride `chunk`), plus the valid-triplet count of batch-all.
__global__ void pair_reduce_kernel(const PairPartial* __restrict__ partial, int64_t n_partials, int64_t chunk,
                                   const int32_t* __restrict__ pos_n, int64_t B, int mode, float* __restrict__ out,
                                   double* __restrict__ stats, const int32_t* __restrict__ overflow) {
  __shared__ double s_sum[32], s_cnt[32], s_val[32];
  double sum = 0.0, cnt = 0.0, nvalid = 0.0;
  for (int64_t i = static_cast<int64_t>(threadIdx.x) * chunk; i < n_partials; i += blockDim.x * chunk) {
    sum += partial[i].sum;
    cnt += static_cast<double>(partial[i].npos);
  }
  if (mode == 0) {
    // valid triplets: sum_i |P_i| * |N_i|, |N_i| = B - |P_i| - 1
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
      const double np = pos_n[i];
      nvalid += np * (static_cast<double>(B) - np - 1.0);
    }
  }
  sum = warp_sum(sum);
  cnt = warp_sum(cnt);
  nvalid = warp_sum(nvalid);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; s_val[warp] = nvalid; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += s_sum[w]; b += s_cnt[w]; c += s_val[w]; }
    // a class with more positives per anchor than the lists hold would silently drop triplets: poison the results
    // (loss, fraction, and through stats[1] the gradient) instead -- the fused step does not read the flag back
    if (overflow != nullptr && *overflow > 0) a = b = c = __longlong_as_double(0x7ff8000000000000ll);
    if (mode == 0) {
      out[0] = static_cast<float>(a / (b + 1e-16));
      out[1] = static_cast<float>(b / (c + 1e-16));
      stats[0] = a; stats[1] = b; stats[2] = c;
    } else {
      out[0] = static_cast<float>(a / (static_cast<double>(B) * static_cast<double>(B - 1)));
    }
  }
}

// =====================================================================================================
// Pair-coefficient backward (batch-all negatives, all-pairs contrastive): CUDA-core version.
//   grad_i = sum_k c_ik (e_i - e_k),  c_ik symmetric-ised pair coefficient regenerated on the fly from the
//   distance tile; nothing of size B x B is stored.  CTA = 32 anchors, streams all column tiles of 32 rows.
// =====================================================================================================
constexpr int PT = 32;        // tile edge (anchors and columns)
constexpr int PDMAX = 512;    // embedding columns kept in registers per pass (16 chunks of 32)

struct CoefBatchAll {
  const float* pos_d;   // [B][cap] distances (squared or not) of each row's positives
  const int32_t* pos_n;
  int cap;
  float margin;
  int squared;
  double inv_np;        // 1 / (#positive triplets + 1e-16)
};

template <int kMode>  // 0 = batch-all, 1 = contrastive
__global__ void __launch_bounds__(256)
pair_bwd_kernel(const float* __restrict__ emb, const int32_t* __restrict__ labels, int64_t B, int d, int d_off,
                CoefBatchAll ba, const double* __restrict__ stats, float scale_c, const float* __restrict__ gloss,
                int32_t* __restrict__ pos_cnt /*[B][cap] batch-all only*/, float* __restrict__ gemb) {
  __shared__ float Ei[PT][PT + 1];
  __shared__ float Ej[PT][PT + 1];
  __shared__ float Ct[PT][PT + 4];
  __shared__ int32_t lab_i[PT], lab_j[PT];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * PT;
  const int dcols = min(PDMAX, d - d_off);        // columns of the gradient produced by this pass
  const int nchunk = (dcols + 31) / 32;
  float acc[4][PDMAX / 32];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < PDMAX / 32; ++c) acc[r][c] = 0.f;
  float rs[4] = {0.f, 0.f, 0.f, 0.f};

  float inv_np = 0.f;
  if (kMode == 0) inv_np = static_cast<float>(1.0 / (stats[1] + 1e-16));
  const float gl = gloss[0];
  if (t < PT) lab_i[t] = (i0 + t < B) ? labels[i0 + t] : -1;

  for (int64_t j0 = 0; j0 < B; j0 += PT) {
    if (t < PT) lab_j[t] = (j0 + t < B) ? labels[j0 + t] : -2;
    // ---- phase 1: dot-product tile over the full embedding dimension
    float dsum[4] = {0.f, 0.f, 0.f, 0.f};  // thread owns (row = w*4 + r, col = lane)
    float ni[4] = {0.f, 0.f, 0.f, 0.f};
    float nj = 0.f;
    for (int k0 = 0; k0 < d; k0 += PT) {
      __syncthreads();
      for (int e = t; e < PT * PT; e += 256) {
        const int r = e >> 5, k = e & 31;
        Ei[r][k] = (i0 + r < B && k0 + k < d) ? emb[(i0 + r) * d + k0 + k] : 0.f;
        Ej[r][k] = (j0 + r < B && k0 + k < d) ? emb[(j0 + r) * d + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < PT; ++k) {
        const float b = Ej[lane][k];
        nj = fmaf(b, b, nj);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float a = Ei[w * 4 + r][k];
          dsum[r] = fmaf(a, b, dsum[r]);
          ni[r] = fmaf(a, a, ni[r]);
        }
      }
    }
    // ---- phase 2: pair coefficients
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int li = w * 4 + r;
      const int64_t gi = i0 + li, gj = j0 + lane;
      float c = 0.f;
      if (gi < B && gj < B && gi != gj) {
        const float d2 = fmaxf(ni[r] + nj - 2.f * dsum[r], 0.f);
        if (kMode == 1) {
          // L = 1/Z sum_{i != j} t(d2); both (i,j) and (j,i) appear => c = 4 t'(d2) / Z, t' w.r.t. d2
          if (d2 >= 1e-7f) {
            if (lab_i[li] == lab_j[lane]) c = 4.f * scale_c;
            else {
              const float dd = sqrtf(d2);
              c = -4.f * scale_c * fmaxf(1.f - dd, 0.f) / dd;
            }
          }
        }
      }
      if (kMode == 0) {
        // negative pair: G_ik = -#{s in P_i : D_is + m - D_ik > 0} / np ; c = (G_ik + G_ki) * s_ik.
        // All 32 lanes share the anchor gi (warp-uniform loop bound); lanes that are not a negative pair idle.
        const bool isneg = gi < B && gj < B && lab_i[li] != lab_j[lane];
        float dn = 0.f;
        if (isneg) {
          const float d2 = fmaxf(ni[r] + nj - 2.f * dsum[r], 0.f);
          dn = ba.squared ? d2 : sqrtf(d2);
        }
        int cnt = 0;
        const int npi = gi < B ? ba.pos_n[gi] : 0;
        for (int s = 0; s < npi; ++s) {
          const bool act = isneg && (ba.pos_d[gi * ba.cap + s] + ba.margin - dn) > 1e-16f;
          cnt += act;
          // number of active negatives of (anchor gi, positive slot s) inside this column tile
          const unsigned m = __ballot_sync(0xffffffffu, act);
          if (lane == 0 && m && d_off == 0) atomicAdd(&pos_cnt[gi * ba.cap + s], __popc(m));
        }
        if (isneg) {
          const int npj = ba.pos_n[gj];
          for (int s = 0; s < npj; ++s) cnt += (ba.pos_d[gj * ba.cap + s] + ba.margin - dn) > 1e-16f;
          const float sfac = ba.squared ? 2.f : (dn > 0.f ? 1.f / dn : 0.f);
          c = -static_cast<float>(cnt) * inv_np * sfac;
        }
      }
      Ct[li][lane] = c;
      float rsum = c;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
      rs[r] += rsum;
    }
    __syncthreads();
    // ---- phase 3: acc[row][col] -= sum_k C[row][k] * E_j[k][col], streamed over column chunks
#pragma unroll
    for (int ch = 0; ch < PDMAX / 32; ++ch) {
      if (ch < nchunk) {  // block-uniform
        __syncthreads();
        for (int e = t; e < PT * PT; e += 256) {
          const int r = e >> 5, k = e & 31;
          const int col = d_off + ch * 32 + k;
          Ej[r][k] = (j0 + r < B && col < d) ? emb[(j0 + r) * d + col] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < PT; ++k) {
          const float ej = Ej[k][lane];
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[r][ch] = fmaf(-Ct[w * 4 + r][k], ej, acc[r][ch]);
        }
      }
    }
  }
  // ---- write: grad_i = gl * (rowsum_i * e_i + acc_i), added atomically (positive-pair terms land separately)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t gi = i0 + w * 4 + r;
    if (gi >= B) continue;
#pragma unroll
    for (int cc = 0; cc < PDMAX / 32; ++cc) {
      const int col = d_off + cc * 32 + lane;
      if (cc < nchunk && col < d) atomicAdd(&gemb[gi * d + col], gl * (rs[r] * emb[gi * d + col] + acc[r][cc]));
    }
  }
}

// Batch-all positive pairs: G_ij = +#{k in N_i : D_ij + m - D_ik > 0} / np, sparse.  One warp per anchor i:
//   grad_i += sum_{j in P(i)} (c_ij + c_ji) (e_i - e_j),  c_ij = cnt(anchor i, slot of j) * s(D_ij) / np
// (the mirrored term of pair (j, i) lands on row i with the opposite sign of e_j - e_i, so it is the same vector).
// Each row has a single writer: no atomics (the one-warp-per-(anchor, slot) version issued 33 M float atomics and
// took 34 us at B = 4096, d = 512; ncu launch list r2).  c_ji is looked up in j's own list (j lists i because the
// relation "same label" is symmetric and the lists are complete -- the forward pass rejects overflowing classes).
__global__ void batch_all_bwd_pos_kernel(const float* __restrict__ emb, int64_t B, int d, int cap, int squared,
                                         const float* __restrict__ pos_d, const int32_t* __restrict__ pos_j,
                                         const int32_t* __restrict__ pos_n, const int32_t* __restrict__ pos_cnt,
                                         const double* __restrict__ stats, const float* __restrict__ gloss,
                                         float* __restrict__ gemb, int unscaled) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= B) return;
  const int n = pos_n[i];
  if (n == 0) return;
  // unscaled: the fused step divides the finished gradient by #positive triplets (pair_scale_kernel)
  const float scale = unscaled ? 1.f : gloss[0] * static_cast<float>(1.0 / (stats[1] + 1e-16));
  // lane s (and s + 32) prepares the weight of positive s: w_s = (cnt_ij + cnt_ji) * s(D_ij) * scale
  float wgt[2] = {0.f, 0.f};
  int jj[2] = {0, 0};
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
      const float sfac = squared ? 2.f : (dij > 0.f ? 1.f / dij : 0.f);
      wgt[h] = static_cast<float>(cnt) * sfac * scale;
      jj[h] = static_cast<int>(j);
    }
  }
  const float* ei = emb + i * d;
  float* gi = gemb + i * d;
  const bool vec = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(gemb) & 15) == 0;
  for (int c0 = 0; c0 < d; c0 += 128) {
    const int c = c0 + 4 * lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && c < d) x = *reinterpret_cast<const float4*>(ei + c);
    else if (!vec) {
      x.x = c < d ? ei[c] : 0.f; x.y = c + 1 < d ? ei[c + 1] : 0.f;
      x.z = c + 2 < d ? ei[c + 2] : 0.f; x.w = c + 3 < d ? ei[c + 3] : 0.f;
    }
    for (int s = 0; s < n; ++s) {
      const float w = __shfl_sync(0xffffffffu, wgt[s >> 5], s & 31);
      const int j = __shfl_sync(0xffffffffu, jj[s >> 5], s & 31);
      if (w == 0.f) continue;  // warp-uniform
      const float* ej = emb + static_cast<int64_t>(j) * d;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec && c < d) y = __ldg(reinterpret_cast<const float4*>(ej + c));
      else if (!vec) {
        y.x = c < d ? ej[c] : 0.f; y.y = c + 1 < d ? ej[c + 1] : 0.f;
        y.z = c + 2 < d ? ej[c + 2] : 0.f; y.w = c + 3 < d ? ej[c + 3] : 0.f;
      }
      acc.x = fmaf(w, x.x - y.x, acc.x);
      acc.y = fmaf(w, x.y - y.y, acc.y);
      acc.z = fmaf(w, x.z - y.z, acc.z);
      acc.w = fmaf(w, x.w - y.w, acc.w);
    }
    if (vec) {
      if (c < d) {
        float4 g = *reinterpret_cast<float4*>(gi + c);
        g.x += acc.x; g.y += acc.y; g.z += acc.z; g.w += acc.w;
        *reinterpret_cast<float4*>(gi + c) = g;
      }
    } else {
      if (c < d) gi[c] += acc.x;
      if (c + 1 < d) gi[c + 1] += acc.y;
      if (c + 2 < d) gi[c + 2] += acc.z;
      if (c + 3 < d) gi[c + 3] += acc.w;
    }
  }
}

int launch_pair_reduce(PairPartial* partial, int64_t n_partials, const int32_t* pos_n, int64_t B, int mode,
                       float* out, double* stats, cudaStream_t st, const int32_t* overflow = nullptr) {
  const int64_t chunk = (n_partials + kReduceBlocks - 1) / kReduceBlocks;
  const unsigned blocks = static_cast<unsigned>((n_partials + chunk - 1) / chunk);
  pair_reduce_stage1_kernel<<<blocks, 256, 0, st>>>(partial, n_partials, chunk);
  EN_LAUNCHED("pair_reduce_stage1_kernel");
  pair_reduce_kernel<<<1, 256, 0, st>>>(partial, n_partials, chunk, pos_n, B, mode, out, stats, overflow);
  EN_LAUNCHED("pair_reduce_kernel");
  return EN_OK;
}

struct TcOperands {
  float *hi, *lo, *norms;
  int dpad;
  CUtensorMap th, tl;
};

// `centre`: subtract the column mean before the TF32 split (norms become the centred norms).  Distances do not
// change; the tensor core's accumulation truncates toward zero, which biases all-positive dot products (post-ReLU
// embeddings) by ~6e-6 -- enough to flip ~1e-5 of the batch-all hinge decisions against float64.  Centred dot
// products have mixed signs and the bias averages out (measured: all-pairs contrastive gradient 4e-6 -> 3e-7).
// Batch-hard does not need it: its finalize kernel re-evaluates the candidates exactly.
// `bf16`: BF16 planes instead (selection-only paths; they fit in the same buffers).
int prepare_operands(const float* emb, int64_t B, int d, Workspace& w, cudaStream_t st, TcOperands& o,
                     bool centre = false, bool bf16 = false, float* zero_rows = nullptr,
                     unsigned* zero_word = nullptr) {
  o.dpad = tc::dpad_for(d, bf16);
  const size_t dpad32 = static_cast<size_t>(tc::dpad_for(d, 0));
  o.hi = w.take<float>(static_cast<size_t>(B) * dpad32);
  o.lo = w.take<float>(static_cast<size_t>(B) * dpad32);
  o.norms = w.take<float>(B);
  float* mu = w.take<float>(d);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "workspace too small or misaligned");
  if (centre) {
    tc::launch_column_mean(emb, B, d, mu, st);
    EN_LAUNCHED("column_mean_kernel");
  }
  if (bf16) {
    EN_CUDA(tc::launch_split_bf16(emb, B, d, d, o.dpad, o.hi, o.lo, o.norms, st, zero_rows, zero_word));
    ++launch_counter();
    if (tc::make_plane_tmap_bf16(&o.th, o.hi, B, o.dpad) || tc::make_plane_tmap_bf16(&o.tl, o.lo, B, o.dpad))
      return fail(EN_ERR_DRIVER, "cuTensorMapEncodeTiled failed");
    return EN_OK;
  }
  EN_CUDA(tc::launch_split(emb, B, d, d, o.dpad, o.hi, o.lo, o.norms, st, centre ? mu : nullptr));
  ++launch_counter();
  if (tc::make_plane_tmap(&o.th, o.hi, B, o.dpad) || tc::make_plane_tmap(&o.tl, o.lo, B, o.dpad))
    return fail(EN_ERR_DRIVER, "cuTensorMapEncodeTiled failed");
  return EN_OK;
}

size_t operand_bytes(int64_t B, int d) {
  const size_t dpad = static_cast<size_t>((d + tc::BK - 1) / tc::BK * tc::BK);
  return 2 * align_up(static_cast<size_t>(B) * dpad * 4) + align_up(static_cast<size_t>(B) * 4) +
         align_up(static_cast<size_t>(d) * 4);
}

int splits_for(int64_t B, int sms) {
  // enough (row tile, column range) items to give every SM several, without making the per-item partial lists long
  // one column tile per item balances best (1024 items over 148 SMs at B = 4096); cap the item count for huge B
  (void)sms;
  const int tiles = static_cast<int>((B + tc::BM - 1) / tc::BM);
  int s = tiles;
  if (static_cast<int64_t>(s) * tiles > 65536) s = 65536 / tiles;
  if (s < 1) s = 1;
  return s;
}

}  // namespace
}  // namespace en

using namespace en;

// developer aid (tools/trace_bh.py; not part of the C ABI): device buffer of 64 stamps per CTA for the batch-hard GEMM
static unsigned long long* g_bh_trace = nullptr;
// layout: 64 stamps per GEMM CTA (csrc/tc_engine.cuh trace_stamp), then the words of the finalize kernels (fin_stamp,
// fin_sample, per listed anchor; written only by a -DEN_FIN_TRACE build)
extern "C" void en_debug_set_bh_trace(void* device_buffer, int n_ctas) {
  g_bh_trace = static_cast<unsigned long long*>(device_buffer);
  unsigned long long* fin = g_bh_trace ? g_bh_trace + static_cast<size_t>(n_ctas) * 64 : nullptr;
  cudaMemcpyToSymbol(g_fin_trace, &fin, sizeof(fin));
}

extern "C" {

// ------------------------------------------------------------------------------------------ batch-hard
size_t en_ws_bytes_batch_hard(int64_t B, int d) {
  if (B <= 0 || d <= 0) return 0;
  const size_t tiles_n = static_cast<size_t>((B + tc::BN - 1) / tc::BN);
  return operand_bytes(B, d) + align_up(static_cast<size_t>(B) * tiles_n * BH_SLOTS * sizeof(BhCand)) +
         align_up(static_cast<size_t>(B) * sizeof(double)) + align_up(static_cast<size_t>(B) * sizeof(int32_t)) +
         align_up(4 * sizeof(unsigned));
}

static int batch_hard_core(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                           int soft, float* loss, int32_t* hp_idx, int32_t* hn_idx, float* hp, float* hn,
                           float* coef, const float* gloss, float* gemb, void* ws, size_t ws_bytes, void* stream,
                           const char* who) {
  EN_REQUIRE(emb && labels && loss && hp_idx && hn_idx && hp && hn && coef && B > 0 && d > 0,
             "%s: bad arguments (B=%lld d=%d)", who, (long long)B, d);
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_batch_hard(B, d))
    return fail(EN_ERR_WORKSPACE, "%s: workspace too small (%zu < %zu)", who, ws_bytes, en_ws_bytes_batch_hard(B, d));
  cudaStream_t st = as_stream(stream);
  Workspace w(ws, ws_bytes);
  TcOperands o;
  // The GEMM only SELECTS (top-2 candidates per anchor and tile slot); the finalize kernels re-evaluate everything
  // inside the error band exactly.  Split-BF16 operands run the tensor pipe at twice the TF32 rate.
  const int tiles_n = static_cast<int>((B + tc::BN - 1) / tc::BN);
  BhCand* cand = w.take<BhCand>(static_cast<size_t>(B) * tiles_n * BH_SLOTS);
  // per-anchor hinge values (fast path) / per-block partial sums (generic path: one per 8 anchors)
  double* hinge_all = w.take<double>(B);
  int32_t* work_list = w.take<int32_t>(B);
  unsigned* counters = w.take<unsigned>(4);  // [0] work-list length, [1] finished blocks
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "%s: workspace too small or misaligned", who);
  // the operand split also zeroes the gradient buffer and the two counters (no memset nodes in the step)
  prof_mark(st, 0);
  if (int rc = prepare_operands(emb, B, d, w, st, o, false, true, gemb, counters)) return rc;
  prof_mark(st, 1);
  // |dot~ - dot| <= c |a||b| with c = 3 * 2^-16 (dropped lo*lo and residual products of the BF16 split) +
  // 2^-22 (d/16 + 1) (accumulator truncation): 5.4e-5 at d = 512 (measured maximum: 4e-6).  Proxy = |b|^2 - 2 dot,
  // |a||b| <= (|a|^2 + |b|^2) / 2, best and contender both off by it: band = 2 c (|a|^2 + |b|^2).
  // The epilogue also truncates 7 mantissa bits of each proxy (index packing): < 2^-16 |t|, |t| <= 2 (|a|^2 +
  // |b|^2), on both sides: + 2^-14.
  // From four column tiles on, a CTA takes one row tile and a PAIR of column tiles (csrc/tc_engine_wide.cuh: the A
  // tiles are fetched once for both; the kernel is bound by the L2 -> SM operand stream).  Its three products share
  // one accumulator, so the truncation term counts 3 d/16 links.  EN_BH_NARROW=1 keeps the 128 x 128 schedule (tests
  // compare the two).
  const char* narrow_env = getenv("EN_BH_NARROW");
  const bool wide = tiles_n >= 4 && !(narrow_env && narrow_env[0] == '1');
  const int links = wide ? 3 * (o.dpad / 16) + 1 : o.dpad / 16 + 1;
  const float band_c = 2.0f * (3.0f / 65536.0f + links / 4194304.0f) + 1.0f / 16384.0f;
  tc::Shape sh = tc::make_shape_symmetric(B, d, 3, 1);  // upper-triangular tiles, one per work item; BF16 planes
  sh.trace = g_bh_trace;
  EpBatchHard::Params ep{labels, o.norms, cand, B, tiles_n};
  const int sms = device_sm_count();
  if (wide) {
    CUtensorMap bh, bl;  // the same planes with 256-row boxes (both column tiles of a pair in one load)
    if (tc::make_plane_tmap_bf16(&bh, o.hi, B, o.dpad, 2 * tc::BN) || tc::make_plane_tmap_bf16(&bl, o.lo, B, o.dpad, 2 * tc::BN))
      return fail(EN_ERR_DRIVER, "%s: cuTensorMapEncodeTiled failed", who);
    prof_begin(st);
    EN_CUDA(tc::wide::launch<EpBatchHard>(o.th, o.tl, bh, bl, sh, ep, sms, st));
    prof_end(st);
  } else {
    prof_begin(st);
    EN_CUDA(tc::launch<EpBatchHard>(o.th, o.tl, o.th, o.tl, sh, ep, sms, st));
    prof_end(st);
  }
  prof_mark(st, 2);
  ++launch_counter();
  const bool fast = d % 128 == 0 && d <= 512 && tiles_n * BH_SLOTS <= 128 &&
                    (reinterpret_cast<uintptr_t>(emb) & 15) == 0;
  if (fast) {
    const unsigned fblocks = static_cast<unsigned>((B + FF_WARPS - 1) / FF_WARPS);
#define EN_BH_FAST(G, DV)                                                                                         \
  batch_hard_finalize_fast_kernel<G, DV><<<fblocks, FF_WARPS * 32, 0, st>>>(                                        \
      emb, labels, o.norms, cand, B, tiles_n, margin, squared, soft, band_c, hp_idx, hn_idx, hp, hn, coef, hinge_all, \
      work_list, counters, G ? gloss : nullptr, G ? gemb : nullptr)
    if (gemb) {
      if (d == 128) EN_BH_FAST(true, 1);
      else if (d == 256) EN_BH_FAST(true, 2);
      else if (d == 384) EN_BH_FAST(true, 3);
      else EN_BH_FAST(true, 4);
    } else {
      if (d == 128) EN_BH_FAST(false, 1);
      else if (d == 256) EN_BH_FAST(false, 2);
      else if (d == 384) EN_BH_FAST(false, 3);
      else EN_BH_FAST(false, 4);
    }
#undef EN_BH_FAST
    EN_LAUNCHED("batch_hard_finalize_fast_kernel");
    prof_mark(st, 3);
    // the anchors on the work list (a block each; enough blocks that each takes one: the kernel's duration is one
    // anchor's dependent chain of loads) and the deterministic mean
    const unsigned sblocks = static_cast<unsigned>(sms > 0 ? sms : 148) * 4;
    if (gemb)
      batch_hard_finalize_slow_kernel<true><<<sblocks, FS_WARPS * 32, 0, st>>>(
          emb, labels, o.norms, cand, B, d, tiles_n, margin, squared, soft, band_c, hp_idx, hn_idx, hp, hn, coef,
          hinge_all, work_list, counters, loss, gloss, gemb);
    else
      batch_hard_finalize_slow_kernel<false><<<sblocks, FS_WARPS * 32, 0, st>>>(
          emb, labels, o.norms, cand, B, d, tiles_n, margin, squared, soft, band_c, hp_idx, hn_idx, hp, hn, coef,
          hinge_all, work_list, counters, loss, nullptr, nullptr);
    EN_LAUNCHED("batch_hard_finalize_slow_kernel");
    prof_mark(st, 4);
    return EN_OK;
  }
  const unsigned blocks = static_cast<unsigned>((B + 7) / 8);
  if (gemb)
    batch_hard_finalize_kernel<true><<<blocks, 256, 0, st>>>(emb, labels, o.norms, cand, B, d, tiles_n, margin,
                                                             squared, soft, band_c, hp_idx, hn_idx, hp, hn, coef,
                                                             hinge_all, counters + 1, loss, gloss, gemb);
  else
    batch_hard_finalize_kernel<false><<<blocks, 256, 0, st>>>(emb, labels, o.norms, cand, B, d, tiles_n, margin,
                                                              squared, soft, band_c, hp_idx, hn_idx, hp, hn, coef,
                                                              hinge_all, counters + 1, loss, nullptr, nullptr);
  EN_LAUNCHED("batch_hard_finalize_kernel");
  return EN_OK;
}

int en_batch_hard_fwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared, int soft,
                      float* loss, int32_t* hp_idx, int32_t* hn_idx, float* hp, float* hn, float* coef, void* ws,
                      size_t ws_bytes, void* stream) {
  return batch_hard_core(emb, labels, B, d, margin, squared, soft, loss, hp_idx, hn_idx, hp, hn, coef, nullptr,
                         nullptr, ws, ws_bytes, stream, "en_batch_hard_fwd");
}

int en_batch_hard_fwd_bwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                          int soft, float* loss, int32_t* hp_idx, int32_t* hn_idx, float* hp, float* hn, float* coef,
                          const float* gloss, float* gemb, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(gemb != nullptr, "en_batch_hard_fwd_bwd: gemb is null");
  return batch_hard_core(emb, labels, B, d, margin, squared, soft, loss, hp_idx, hn_idx, hp, hn, coef, gloss, gemb,
                         ws, ws_bytes, stream, "en_batch_hard_fwd_bwd");
}

int en_batch_hard_bwd(const float* emb, int64_t B, int d, int squared, const int32_t* hp_idx, const int32_t* hn_idx,
                      const float* hp, const float* hn, const float* coef, const float* gloss, float* gemb,
                      void* stream) {
  EN_REQUIRE(emb && hp_idx && hn_idx && hp && hn && coef && gloss && gemb && B > 0 && d > 0,
             "en_batch_hard_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  const unsigned blocks = static_cast<unsigned>((B * 32 + 255) / 256);
  batch_hard_bwd_own_kernel<<<blocks, 256, 0, st>>>(emb, B, d, squared, hp_idx, hn_idx, hp, hn, coef, gloss, gemb);
  EN_LAUNCHED("batch_hard_bwd_own_kernel");
  batch_hard_bwd_scatter_kernel<<<blocks, 256, 0, st>>>(emb, B, d, squared, hp_idx, hn_idx, hp, hn, coef, gloss,
                                                        gemb);
  EN_LAUNCHED("batch_hard_bwd_scatter_kernel");
  return EN_OK;
}

// ------------------------------------------------------------------------------------------ batch-all
static size_t pos_bytes(int64_t B, int cap) {
  return align_up(static_cast<size_t>(B) * cap * 4) * 3 + align_up(static_cast<size_t>(B) * 4) + align_up(4) +
         (cap > kTcBwdMaxPos ? align_up(static_cast<size_t>(B) * cap * 8) : 0);  // + float64 prefix sums (large classes)
}

size_t en_ws_bytes_batch_all(int64_t B, int d, int max_positives) {
  if (B <= 0 || d <= 0 || max_positives <= 0 || max_positives > kMaxPos) return 0;
  const size_t tiles = static_cast<size_t>((B + tc::BM - 1) / tc::BM);
  const size_t fwd = operand_bytes(B, d) + pos_bytes(B, max_positives) +
                     align_up(static_cast<size_t>(B) * tiles * tc::EPI_H * sizeof(PairPartial));
  const size_t bwd = pos_bytes((B + tc::BM - 1) / tc::BM * tc::BM, tc_list_cap(max_positives)) +
                     align_up(static_cast<size_t>(B) * pair_tc_partials_per_row(B, d) * sizeof(PairPartial)) +
                     pair_tc_ws_bytes(B, d);
  return fwd > bwd ? fwd : bwd;
}

struct PosLists {
  float* pos_d;
  int32_t* pos_j;
  int32_t* pos_cnt;
  int32_t* pos_n;
  int32_t* status;
  double* pos_pre;  // large classes only (cap > 8): prefix sums of the sorted distances
};

static PosLists take_pos(Workspace& w, int64_t B, int cap) {
  PosLists p;
  p.pos_d = w.take<float>(static_cast<size_t>(B) * cap);
  p.pos_j = w.take<int32_t>(static_cast<size_t>(B) * cap);
  p.pos_cnt = w.take<int32_t>(static_cast<size_t>(B) * cap);
  p.pos_n = w.take<int32_t>(B);
  p.status = w.take<int32_t>(1);
  p.pos_pre = cap > kTcBwdMaxPos ? w.take<double>(static_cast<size_t>(B) * cap) : nullptr;
  return p;
}

int en_batch_all_fwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                     int max_positives, float* out, double* stats, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(emb && labels && out && stats && B > 1 && d > 0, "en_batch_all_fwd: bad arguments");
  EN_REQUIRE(max_positives > 0 && max_positives <= kMaxPos,
             "en_batch_all_fwd: max_positives must be in [1, %d] (largest class size - 1); got %d", kMaxPos,
             max_positives);
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_batch_all(B, d, max_positives))
    return fail(EN_ERR_WORKSPACE, "en_batch_all_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  Workspace w(ws, ws_bytes);
  TcOperands o;
  if (int rc = prepare_operands(emb, B, d, w, st, o, true)) return rc;
  PosLists pl = take_pos(w, B, max_positives);
  const int sms = device_sm_count();
  tc::Shape sh = tc::make_shape(B, B, d, splits_for(B, sms), 3);
  PairPartial* partial = w.take<PairPartial>(static_cast<size_t>(B) * sh.n_splits * tc::EPI_H);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_batch_all_fwd: workspace too small or misaligned");
  EN_CUDA(cudaMemsetAsync(pl.status, 0, 4, st));
  collect_positives_kernel<<<static_cast<unsigned>((B * 32 + 255) / 256), 256, 0, st>>>(
      emb, labels, B, d, squared, max_positives, pl.pos_d, pl.pos_j, pl.pos_n, pl.status, B);
  EN_LAUNCHED("collect_positives_kernel");
  EpBatchAll::Params ep{labels, o.norms, pl.pos_d, pl.pos_n, partial, B, max_positives, sh.n_splits, squared, margin};
  prof_begin(st);
  EN_CUDA(tc::launch<EpBatchAll>(o.th, o.tl, o.th, o.tl, sh, ep, sms, st));
  prof_end(st);
  ++launch_counter();
  if (int rc = launch_pair_reduce(partial, B * sh.n_splits * tc::EPI_H, pl.pos_n, B, 0, out, stats, st)) return rc;
  // a class larger than max_positives + 1 would silently drop triplets: surface it (one 4-byte read-back)
  int32_t status_h = 0;
  EN_CUDA(cudaMemcpyAsync(&status_h, pl.status, 4, cudaMemcpyDeviceToHost, st));
  EN_CUDA(cudaStreamSynchronize(st));
  if (status_h > 0)
    return fail(EN_ERR_ARG, "en_batch_all_fwd: a class has %d positives per anchor but max_positives = %d", status_h,
                max_positives);
  return EN_OK;
}

int en_batch_all_bwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                     int max_positives, const double* stats, const float* gloss, float* gemb, void* ws,
                     size_t ws_bytes, void* stream) {
  EN_REQUIRE(emb && labels && stats && gloss && gemb && B > 1 && d > 0, "en_batch_all_bwd: bad arguments");
  EN_REQUIRE(max_positives > 0 && max_positives <= kMaxPos, "en_batch_all_bwd: bad max_positives %d", max_positives);
  if (!ws || ws_bytes < en_ws_bytes_batch_all(B, d, max_positives))
    return fail(EN_ERR_WORKSPACE, "en_batch_all_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  Workspace w(ws, ws_bytes);
  const bool tensor = max_positives <= kTcBwdMaxPos || !cuda_core_bwd_requested();
  const int cap = tensor ? tc_list_cap(max_positives) : max_positives;
  const int64_t Bp = (B + tc::BM - 1) / tc::BM * tc::BM;  // lists for whole tiles (empty past B)
  PosLists pl = take_pos(w, Bp, cap);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_batch_all_bwd: workspace too small or misaligned");
  EN_CUDA(cudaMemsetAsync(pl.status, 0, 4, st));
  collect_positives_kernel<<<static_cast<unsigned>((Bp * 32 + 255) / 256), 256, 0, st>>>(
      emb, labels, B, d, squared, cap, pl.pos_d, pl.pos_j, pl.pos_n, pl.status, Bp, tensor ? pl.pos_pre : nullptr);
  EN_LAUNCHED("collect_positives_kernel");
  EN_CUDA(cudaMemsetAsync(pl.pos_cnt, 0, static_cast<size_t>(B) * cap * 4, st));
  if (tensor) {
    // negatives: two chained tcgen05 GEMMs (csrc/pair_tc.cu), coefficients unscaled; the finishing kernel adds the
    // row-sum term and the sparse positive pairs and applies gloss / #positive triplets
    void* rest = w.base + w.off;
    PairTcFinish fin;
    if (int rc = pair_tc_launch(emb, labels, B, d, 0, squared, margin, 1.f, pl.pos_d, pl.pos_pre, pl.pos_n, pl.pos_cnt, cap,
                                nullptr, nullptr, gemb, &fin, rest, ws_bytes - w.off, st))
      return rc;
    return pair_tc_finish(fin, emb, B, d, cap, squared, pl.pos_d, pl.pos_j, pl.pos_n, pl.pos_cnt, stats, gloss, gemb, st);
  } else {
    // EN_BATCH_ALL_CUDA_CORE=1: CUDA-core tile kernel (independent implementation, for comparison)
    EN_CUDA(cudaMemsetAsync(gemb, 0, static_cast<size_t>(B) * d * 4, st));
    CoefBatchAll ba{pl.pos_d, pl.pos_n, cap, margin, squared, 0.0};
    const unsigned blocks = static_cast<unsigned>((B + PT - 1) / PT);
    for (int d_off = 0; d_off < d; d_off += PDMAX) {
      pair_bwd_kernel<0><<<blocks, 256, 0, st>>>(emb, labels, B, d, d_off, ba, stats, 0.f, gloss, pl.pos_cnt, gemb);
      EN_LAUNCHED("pair_bwd_kernel<batch_all>");
    }
  }
  // positives: sparse, one warp per anchor
  batch_all_bwd_pos_kernel<<<static_cast<unsigned>((B * 32 + 127) / 128), 128, 0, st>>>(
      emb, B, d, cap, squared, pl.pos_d, pl.pos_j, pl.pos_n, pl.pos_cnt, stats, gloss, gemb, 0);
  EN_LAUNCHED("batch_all_bwd_pos_kernel");
  return EN_OK;
}

// Loss AND gradient of batch-all from ONE pass over the distance tiles (csrc/pair_tc.cu): positives lists, centred
// operand planes and S tiles are built once; the gradient is accumulated unscaled and divided by the number of
// positive triplets at the end (that count is only known once every tile has been seen).
int en_batch_all_fwd_bwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                         int max_positives, float* out, double* stats, const float* gloss, float* gemb,
                         int32_t* overflow, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(emb && labels && out && stats && gemb && B > 1 && d > 0, "en_batch_all_fwd_bwd: bad arguments");
  EN_REQUIRE(max_positives > 0 && max_positives <= kMaxPos,
             "en_batch_all_fwd_bwd: max_positives must be in [1, %d] (largest class size - 1); got %d", kMaxPos,
             max_positives);
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_batch_all(B, d, max_positives))
    return fail(EN_ERR_WORKSPACE, "en_batch_all_fwd_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  if (max_positives > kTcBwdMaxPos && cuda_core_bwd_requested()) {
    // comparison path: forward kernel, then the CUDA-core backward (needs a device gloss)
    EN_REQUIRE(gloss != nullptr, "en_batch_all_fwd_bwd: gloss is required when max_positives > %d", kTcBwdMaxPos);
    if (overflow) EN_CUDA(cudaMemsetAsync(overflow, 0, 4, st));  // the forward pass below checks synchronously
    if (int rc = en_batch_all_fwd(emb, labels, B, d, margin, squared, max_positives, out, stats, ws, ws_bytes, stream))
      return rc;
    return en_batch_all_bwd(emb, labels, B, d, margin, squared, max_positives, stats, gloss, gemb, ws, ws_bytes, stream);
  }
  Workspace w(ws, ws_bytes);
  const int cap = tc_list_cap(max_positives);
  const int64_t Bp = (B + tc::BM - 1) / tc::BM * tc::BM;  // lists for whole tiles (empty past B)
  PosLists pl = take_pos(w, Bp, cap);
  const int ppr = pair_tc_partials_per_row(B, d);
  PairPartial* partial = w.take<PairPartial>(static_cast<size_t>(B) * ppr);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_batch_all_fwd_bwd: workspace too small or misaligned");
  EN_CUDA(cudaMemsetAsync(pl.status, 0, 4, st));
  collect_positives_kernel<<<static_cast<unsigned>((Bp * 32 + 255) / 256), 256, 0, st>>>(
      emb, labels, B, d, squared, cap, pl.pos_d, pl.pos_j, pl.pos_n, pl.status, Bp, pl.pos_pre);
  EN_LAUNCHED("collect_positives_kernel");
  EN_CUDA(cudaMemsetAsync(pl.pos_cnt, 0, static_cast<size_t>(B) * cap * 4, st));
  EN_CUDA(cudaMemsetAsync(partial, 0, static_cast<size_t>(B) * ppr * sizeof(PairPartial), st));
  void* rest = w.base + w.off;
  PairTcFinish fin;
  if (int rc = pair_tc_launch(emb, labels, B, d, 0, squared, margin, 1.f, pl.pos_d, pl.pos_pre, pl.pos_n, pl.pos_cnt, cap, nullptr,
                              partial, gemb, &fin, rest, ws_bytes - w.off, st))
    return rc;
  if (int rc = launch_pair_reduce(partial, static_cast<int64_t>(B) * ppr, pl.pos_n, B, 0, out, stats, st, pl.status))
    return rc;
  if (int rc = pair_tc_finish(fin, emb, B, d, cap, squared, pl.pos_d, pl.pos_j, pl.pos_n, pl.pos_cnt, stats, gloss, gemb,
                              st))
    return rc;
  // A class larger than the lists (8 positives per anchor here) would silently drop triplets.  No host read-back on
  // this path (a training step must not stall the stream): the results are poisoned with NaN above, and the flag --
  // the offending positives count, 0 when fine -- is left in `overflow` for the caller to inspect when convenient.
  if (overflow) EN_CUDA(cudaMemcpyAsync(overflow, pl.status, 4, cudaMemcpyDeviceToDevice, st));
  return EN_OK;
}

// ------------------------------------------------------------------------------------------ all-pairs contrastive
size_t en_ws_bytes_contrastive_allpairs(int64_t B, int d) {
  if (B <= 0 || d <= 0) return 0;
  const size_t tiles = static_cast<size_t>((B + tc::BM - 1) / tc::BM);
  const size_t fwd = operand_bytes(B, d) + align_up(static_cast<size_t>(B) * tiles * tc::EPI_H * sizeof(PairPartial));
  const size_t bwd = align_up(static_cast<size_t>(B) * pair_tc_partials_per_row(B, d) * sizeof(PairPartial)) +
                     pair_tc_ws_bytes(B, d);
  return fwd > bwd ? fwd : bwd;
}

int en_contrastive_allpairs_fwd(const float* emb, const int32_t* labels, int64_t B, int d, float* loss, void* ws,
                                size_t ws_bytes, void* stream) {
  EN_REQUIRE(emb && labels && loss && B > 1 && d > 0, "en_contrastive_allpairs_fwd: bad arguments");
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_contrastive_allpairs(B, d))
    return fail(EN_ERR_WORKSPACE, "en_contrastive_allpairs_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  Workspace w(ws, ws_bytes);
  TcOperands o;
  if (int rc = prepare_operands(emb, B, d, w, st, o, true)) return rc;
  const int sms = device_sm_count();
  tc::Shape sh = tc::make_shape(B, B, d, splits_for(B, sms), 3);
  PairPartial* partial = w.take<PairPartial>(static_cast<size_t>(B) * sh.n_splits * tc::EPI_H);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_contrastive_allpairs_fwd: workspace too small or misaligned");
  EpContrastive::Params ep{labels, o.norms, partial, B, sh.n_splits};
  prof_begin(st);
  EN_CUDA(tc::launch<EpContrastive>(o.th, o.tl, o.th, o.tl, sh, ep, sms, st));
  prof_end(st);
  ++launch_counter();
  if (int rc = launch_pair_reduce(partial, B * sh.n_splits * tc::EPI_H, nullptr, B, 1, loss, nullptr, st)) return rc;
  return EN_OK;
}

int en_contrastive_allpairs_bwd(const float* emb, const int32_t* labels, int64_t B, int d, const float* gloss,
                                float* gemb, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(emb && labels && gloss && gemb && B > 1 && d > 0, "en_contrastive_allpairs_bwd: bad arguments");
  if (!ws || ws_bytes < en_ws_bytes_contrastive_allpairs(B, d))
    return fail(EN_ERR_WORKSPACE, "en_contrastive_allpairs_bwd: workspace too small");
  const float scale = static_cast<float>(4.0 / (static_cast<double>(B) * static_cast<double>(B - 1)));
  PairTcFinish fin;
  if (int rc = pair_tc_launch(emb, labels, B, d, 1, 0, 0.f, scale, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, gemb,
                              &fin, ws, ws_bytes, as_stream(stream)))
    return rc;
  return pair_tc_finish(fin, emb, B, d, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, gloss, gemb,
                        as_stream(stream));
}

// Loss AND gradient of the all-pairs contrastive loss from one pass over the distance tiles (csrc/pair_tc.cu).
int en_contrastive_allpairs_fwd_bwd(const float* emb, const int32_t* labels, int64_t B, int d, float* loss,
                                    const float* gloss, float* gemb, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(emb && labels && loss && gemb && B > 1 && d > 0, "en_contrastive_allpairs_fwd_bwd: bad arguments");
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_contrastive_allpairs(B, d))
    return fail(EN_ERR_WORKSPACE, "en_contrastive_allpairs_fwd_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  Workspace w(ws, ws_bytes);
  const int ppr = pair_tc_partials_per_row(B, d);
  PairPartial* partial = w.take<PairPartial>(static_cast<size_t>(B) * ppr);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_contrastive_allpairs_fwd_bwd: workspace too small or misaligned");
  EN_CUDA(cudaMemsetAsync(partial, 0, static_cast<size_t>(B) * ppr * sizeof(PairPartial), st));
  const float scale = static_cast<float>(4.0 / (static_cast<double>(B) * static_cast<double>(B - 1)));
  void* rest = w.base + w.off;
  PairTcFinish fin;
  if (int rc = pair_tc_launch(emb, labels, B, d, 1, 0, 0.f, scale, nullptr, nullptr, nullptr, nullptr, 0, nullptr, partial, gemb,
                              &fin, rest, ws_bytes - w.off, st))
    return rc;
  if (int rc = pair_tc_finish(fin, emb, B, d, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, gloss, gemb, st))
    return rc;
  return launch_pair_reduce(partial, static_cast<int64_t>(B) * ppr, nullptr, B, 1, loss, nullptr, st);
}

}  // extern "C"
