// Pairwise Euclidean distance matrix with scikit-learn's float32 semantics, and the per-pair negative-selection
// scans of the reference's in-batch triplet miner.
//
//   en_pairwise_dist      <- sklearn.metrics.pairwise_distances(all_embeddings)   embedding_net/datagenerators.py:219
//   en_mine_batch_scan    <- hardest / random_hard / semihard predicates           embedding_net/datagenerators.py:188-199
//   en_mine_batch_select  <- np.random.choice(candidates) made RNG-faithful        embedding_net/datagenerators.py:194,199
//
// The mining batch is small (k_classes*k_samples rows; 60 in the shipped config, 256 at BASELINE config 1), and the
// reference decides on float32 values that sklearn obtains from float64 arithmetic.  To return the same indices
// the "exact" path therefore accumulates in float64 on the CUDA cores; the tcgen05 path is offered for large n.
#include "common.cuh"
#include "tc_engine.cuh"

namespace en {
namespace {

// ---------------------------------------------------------------- float64 row norms
__global__ void row_norms_f64_kernel(const float* __restrict__ x, int64_t n, int d, double* __restrict__ xx) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double acc = 0.0;
  for (int c = lane; c < d; c += 32) {
    const double v = x[row * d + c];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if (lane == 0) xx[row] = acc;
}

// ---------------------------------------------------------------- float64 tiled distance kernel
constexpr int XT = 64;   // tile edge
constexpr int XK = 16;   // k chunk

__global__ void __launch_bounds__(256)
pairwise_exact_kernel(const float* __restrict__ x, const double* __restrict__ xx, int64_t n, int d, int squared,
                      float* __restrict__ out) {
  __shared__ double As[XK][XT + 1];
  __shared__ double Bs[XK][XT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = static_cast<int64_t>(blockIdx.y) * XT, col0 = static_cast<int64_t>(blockIdx.x) * XT;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < d; k0 += XK) {
    // 64 rows x 16 k per operand = 1024 elements, 4 per thread; k fastest so global reads are contiguous
    for (int e = threadIdx.x; e < XT * XK; e += 256) {
      const int r = e / XK, k = e % XK;
      const int64_t ra = row0 + r, rb = col0 + r;
      As[k][r] = (ra < n && k0 + k < d) ? static_cast<double>(x[ra * d + k0 + k]) : 0.0;
      Bs[k][r] = (rb < n && k0 + k < d) ? static_cast<double>(x[rb * d + k0 + k]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < XK; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = col0 + tx * 4 + j;
      if (c >= n) continue;
      // sklearn: d = -2 x.y ; d += |x|^2 ; d += |y|^2  (float64), cast, clamp, zero diagonal, sqrt (float32)
      double t = -2.0 * acc[i][j];
      t += xx[r];
      t += xx[c];
      float f = fmaxf(static_cast<float>(t), 0.f);
      if (r == c) f = 0.f;
      out[r * n + c] = squared ? f : sqrtf(f);
    }
  }
}

// ---------------------------------------------------------------- tensor-core variant (diagnostics / large n)
struct EpStoreDist {
  struct Params {
    float* out;
    const float* norms;
    int64_t n;
    int squared;
  };
  struct Row {
    float na;
  };
  static constexpr int kSmemBytes = 0;
  static __device__ void item_begin(const Params& p, Row& r, const tc::Ctx&, int64_t row, bool valid, int, int) {
    r.na = valid ? p.norms[row] : 0.f;
  }
  static __device__ void chunk(const Params& p, Row& r, const tc::Ctx&, int64_t row, bool valid, int64_t col0,
                               const float (&dot)[32]) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int64_t c = col0 + j;
      if (c < p.n) {
        float f = fmaxf(r.na + __ldg(&p.norms[c]) - 2.f * dot[j], 0.f);  // diagnostics path: plain loads
        if (c == row) f = 0.f;
        p.out[row * p.n + c] = p.squared ? f : sqrtf(f);
      }
    }
  }
  static __device__ void tile_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int) {}
  static __device__ void item_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int, int) {}
};

// ---------------------------------------------------------------- mining scans (one warp per (anchor, positive))
// loss = (D[a,p] - D[a,n]) + margin, evaluated in float32 left to right exactly as datagenerators.py:235 does.
__device__ __forceinline__ float mining_loss(float d_ap, float d_an, float margin) {
  return __fadd_rn(__fsub_rn(d_ap, d_an), margin);
}

__global__ void mine_scan_kernel(const float* __restrict__ D, const int32_t* __restrict__ labels, int64_t n,
                                 const int32_t* __restrict__ pairs, int64_t n_pairs, float margin,
                                 int32_t* __restrict__ hardest, int32_t* __restrict__ n_hard,
                                 int32_t* __restrict__ n_semi) {
  const int64_t p = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  const int a = pairs[2 * p], pos = pairs[2 * p + 1];
  const float d_ap = D[static_cast<int64_t>(a) * n + pos];
  const int32_t la = labels[a];
  const float* row = D + static_cast<int64_t>(a) * n;
  float best = -INFINITY;
  int best_idx = 0x7fffffff;
  int cnt_hard = 0, cnt_semi = 0;
  for (int64_t j = lane; j < n; j += 32) {
    if (labels[j] == la) continue;
    const float l = mining_loss(d_ap, row[j], margin);
    if (l > best) {  // strict: first maximum wins inside a lane (ascending j)
      best = l;
      best_idx = static_cast<int>(j);
    }
    cnt_hard += l > 0.f;
    cnt_semi += (l > 0.f) && (l < margin);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
    if (ob > best || (ob == best && oi < best_idx)) {  // np.argmax: lowest index among equal maxima
      best = ob;
      best_idx = oi;
    }
    cnt_hard += __shfl_xor_sync(0xffffffffu, cnt_hard, o);
    cnt_semi += __shfl_xor_sync(0xffffffffu, cnt_semi, o);
  }
  if (lane == 0) {
    hardest[p] = (best_idx != 0x7fffffff && best > 0.f) ? best_idx : -1;
    n_hard[p] = cnt_hard;
    n_semi[p] = cnt_semi;
  }
}

__global__ void mine_select_kernel(const float* __restrict__ D, const int32_t* __restrict__ labels, int64_t n,
                                   const int32_t* __restrict__ pairs, int64_t n_pairs, float margin, int mode,
                                   const int32_t* __restrict__ rank, int32_t* __restrict__ selected) {
  const int64_t p = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  const int a = pairs[2 * p], pos = pairs[2 * p + 1];
  int want = rank[p];
  int result = -1;
  if (want >= 0) {
    const float d_ap = D[static_cast<int64_t>(a) * n + pos];
    const int32_t la = labels[a];
    const float* row = D + static_cast<int64_t>(a) * n;
    for (int64_t j0 = 0; j0 < n; j0 += 32) {
      const int64_t j = j0 + lane;
      bool cand = false;
      if (j < n && labels[j] != la) {
        const float l = mining_loss(d_ap, row[j], margin);
        cand = (mode == EN_MODE_SEMIHARD) ? (l > 0.f && l < margin) : (l > 0.f);
      }
      const unsigned m = __ballot_sync(0xffffffffu, cand);
      const int c = __popc(m);
      if (want < c) {
        // the want-th set bit of m
        unsigned mm = m;
        for (int t = 0; t < want; ++t) mm &= mm - 1;
        result = static_cast<int>(j0) + (__ffs(mm) - 1);
        break;
      }
      want -= c;
    }
  }
  if (lane == 0) selected[p] = result;
}


// The reference's three selection callables applied to ONE loss vector (datagenerators.py:188-199), for callers
// that use them directly.  Single block; out = {first arg-max if its loss > 0 else -1, #(loss > 0),
// #(0 < loss < margin)}.
__global__ void loss_scan_kernel(const float* __restrict__ loss, int64_t n, float margin, int32_t* __restrict__ out) {
  __shared__ float sb[32];
  __shared__ int si[32], sh[32], ss[32];
  float best = -INFINITY;
  int best_idx = 0x7fffffff, ch = 0, cs = 0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const float l = loss[j];
    if (l > best) { best = l; best_idx = static_cast<int>(j); }
    ch += l > 0.f;
    cs += (l > 0.f) && (l < margin);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
    if (ob > best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
    ch += __shfl_xor_sync(0xffffffffu, ch, o);
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
  }
  if (lane == 0) { sb[warp] = best; si[warp] = best_idx; sh[warp] = ch; ss[warp] = cs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) {
      if (sb[w] > best || (sb[w] == best && si[w] < best_idx)) { best = sb[w]; best_idx = si[w]; }
      ch += sh[w];
      cs += ss[w];
    }
    out[0] = (best_idx != 0x7fffffff && best > 0.f) ? best_idx : -1;
    out[1] = ch;
    out[2] = cs;
  }
}

// rank-th (0-based, ascending index) element satisfying the mode's predicate; one warp.
__global__ void loss_select_kernel(const float* __restrict__ loss, int64_t n, float margin, int mode, int rank,
                                   int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int want = rank, result = -1;
  if (want >= 0) {
    for (int64_t j0 = 0; j0 < n; j0 += 32) {
      const int64_t j = j0 + lane;
      bool cand = false;
      if (j < n) {
        const float l = loss[j];
        cand = (mode == EN_MODE_SEMIHARD) ? (l > 0.f && l < margin) : (l > 0.f);
      }
      const unsigned m = __ballot_sync(0xffffffffu, cand);
      const int c = __popc(m);
      if (want < c) {
        unsigned mm = m;
        for (int t = 0; t < want; ++t) mm &= mm - 1;
        result = static_cast<int>(j0) + (__ffs(mm) - 1);
        break;
      }
      want -= c;
    }
  }
  if (lane == 0) out[0] = result;
}

}  // namespace
}  // namespace en

using namespace en;

extern "C" {

size_t en_ws_bytes_pairwise(int64_t n, int d, int exact) {
  if (n < 0 || d <= 0) return 0;
  if (exact) return align_up(static_cast<size_t>(n) * sizeof(double));
  const size_t dpad = static_cast<size_t>((d + tc::BK - 1) / tc::BK * tc::BK);
  return 2 * align_up(static_cast<size_t>(n) * dpad * 4) + align_up(static_cast<size_t>(n) * 4);
}

int en_pairwise_dist(const float* x, int64_t n, int d, int squared, int exact, float* out, void* ws, size_t ws_bytes,
                     void* stream) {
  EN_REQUIRE(x && out && n >= 0 && d > 0, "en_pairwise_dist: bad arguments (n=%lld d=%d)", (long long)n, d);
  if (n == 0) return EN_OK;
  if (!ws || ws_bytes < en_ws_bytes_pairwise(n, d, exact))
    return fail(EN_ERR_WORKSPACE, "en_pairwise_dist: workspace too small (%zu < %zu)", ws_bytes,
                en_ws_bytes_pairwise(n, d, exact));
  Workspace w(ws, ws_bytes);
  cudaStream_t st = as_stream(stream);
  if (exact) {
    double* xx = w.take<double>(n);
    if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_pairwise_dist: workspace misaligned");
    row_norms_f64_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, st>>>(x, n, d, xx);
    EN_LAUNCHED("row_norms_f64_kernel");
    dim3 grid(static_cast<unsigned>((n + XT - 1) / XT), static_cast<unsigned>((n + XT - 1) / XT));
    pairwise_exact_kernel<<<grid, 256, 0, st>>>(x, xx, n, d, squared, out);
    EN_LAUNCHED("pairwise_exact_kernel");
    return EN_OK;
  }
  if (int rc = check_sm100()) return rc;
  const int dpad = (d + tc::BK - 1) / tc::BK * tc::BK;
  float* hi = w.take<float>(static_cast<size_t>(n) * dpad);
  float* lo = w.take<float>(static_cast<size_t>(n) * dpad);
  float* norms = w.take<float>(n);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_pairwise_dist: workspace misaligned");
  EN_CUDA(tc::launch_split(x, n, d, d, dpad, hi, lo, norms, st));
  ++launch_counter();
  CUtensorMap th, tl;
  if (tc::make_plane_tmap(&th, hi, n, dpad) || tc::make_plane_tmap(&tl, lo, n, dpad))
    return fail(EN_ERR_DRIVER, "en_pairwise_dist: cuTensorMapEncodeTiled failed");
  tc::Shape sh = tc::make_shape(n, n, d, 1 << 30, 3);
  EpStoreDist::Params ep{out, norms, n, squared};
  EN_CUDA(tc::launch<EpStoreDist>(th, tl, th, tl, sh, ep, device_sm_count(), st));
  ++launch_counter();
  return EN_OK;
}

int en_mine_batch_scan(const float* D, const int32_t* labels, int64_t n, const int32_t* pairs, int64_t n_pairs,
                       float margin, int32_t* hardest, int32_t* n_hard, int32_t* n_semi, void* stream) {
  EN_REQUIRE(D && labels && pairs && hardest && n_hard && n_semi && n > 0 && n_pairs >= 0,
             "en_mine_batch_scan: bad arguments");
  if (n_pairs == 0) return EN_OK;
  mine_scan_kernel<<<static_cast<unsigned>((n_pairs * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
      D, labels, n, pairs, n_pairs, margin, hardest, n_hard, n_semi);
  EN_LAUNCHED("mine_scan_kernel");
  return EN_OK;
}

int en_mine_batch_select(const float* D, const int32_t* labels, int64_t n, const int32_t* pairs, int64_t n_pairs,
                         float margin, int mode, const int32_t* rank, int32_t* selected, void* stream) {
  EN_REQUIRE(D && labels && pairs && rank && selected && n > 0 && n_pairs >= 0, "en_mine_batch_select: bad arguments");
  EN_REQUIRE(mode == EN_MODE_SEMIHARD || mode == EN_MODE_RANDOM_HARD,
             "en_mine_batch_select: mode must be EN_MODE_SEMIHARD or EN_MODE_RANDOM_HARD (got %d)", mode);
  if (n_pairs == 0) return EN_OK;
  mine_select_kernel<<<static_cast<unsigned>((n_pairs * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
      D, labels, n, pairs, n_pairs, margin, mode, rank, selected);
  EN_LAUNCHED("mine_select_kernel");
  return EN_OK;
}

int en_loss_scan(const float* loss_values, int64_t n, float margin, int32_t* out3, void* stream) {
  EN_REQUIRE(loss_values && out3 && n > 0, "en_loss_scan: bad arguments");
  loss_scan_kernel<<<1, 256, 0, as_stream(stream)>>>(loss_values, n, margin, out3);
  EN_LAUNCHED("loss_scan_kernel");
  return EN_OK;
}

int en_loss_select(const float* loss_values, int64_t n, float margin, int mode, int rank, int32_t* out1,
                   void* stream) {
  EN_REQUIRE(loss_values && out1 && n > 0, "en_loss_select: bad arguments");
  EN_REQUIRE(mode == EN_MODE_SEMIHARD || mode == EN_MODE_RANDOM_HARD, "en_loss_select: bad mode %d", mode);
  loss_select_kernel<<<1, 32, 0, as_stream(stream)>>>(loss_values, n, margin, mode, rank, out1);
  EN_LAUNCHED("loss_select_kernel");
  return EN_OK;
}

}  // extern "C"
