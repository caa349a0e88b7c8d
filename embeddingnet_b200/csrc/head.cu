// Embedding head: Dense(n_out, activation="relu") followed (optionally) by K.l2_normalize -- the last two layers of
// every backbone the reference builds (embedding_net/backbones.py:114-119; also :36-38, :75-77), i.e. the step right
// before the distance / mining / kNN path.  One tcgen05 3xTF32 GEMM  x . W  whose epilogue adds
// the bias, applies the ReLU, stores the row and accumulates its sum of squares; the row is rescaled in place when
// the last column tile has passed, so the un-normalised activations never make a second trip through a kernel.
#include "common.cuh"
#include "tc_engine.cuh"

namespace en {
namespace {

// W (n_in, n_out) row-major -- the layout of a Keras Dense kernel -- -> W^T TF32 planes (n_out, kpad), zero padded.
__global__ void dense_weight_planes_kernel(const float* __restrict__ w, int n_in, int n_out, int kpad,
                                           float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, o = o0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < n_in && o < n_out) ? w[static_cast<int64_t>(k) * n_out + o] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int o = o0 + r, k = k0 + threadIdx.x;
    if (o < n_out && k < kpad) {
      const float v = tile[threadIdx.x][r];
      const float h = tc::to_tf32(v);
      hi[static_cast<int64_t>(o) * kpad + k] = h;
      lo[static_cast<int64_t>(o) * kpad + k] = tc::to_tf32(v - h);
    }
  }
}

// Work item = one 128-row tile with ALL column tiles (n_splits = 1), so a row's sum of squares is complete at
// item_end.  The two threads that share a row (column halves) exchange their partial sums through shared memory.
struct EpDense {
  struct Params {
    const float* bias;  // (n_out,), may be null
    float* out;         // (B, n_out)
    float* inv_norm;    // optional (B,): the row scale 1 / sqrt(max(sum y^2, 1e-12)) the backward pass needs
    int64_t B;
    int n_out;
    int normalize;
  };
  struct Row {
    float ss;
  };
  static constexpr int kSmemBytes = tc::EPI_H * tc::BM * 4;
  static __device__ void item_begin(const Params&, Row& r, const tc::Ctx&, int64_t, bool, int, int) { r.ss = 0.f; }
  static __device__ void chunk(const Params& p, Row& r, const tc::Ctx&, int64_t row, bool valid, int64_t col0,
                               const float (&dot)[32]) {
    if (!valid || col0 >= p.n_out) return;
    float* o = p.out + row * p.n_out + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (col0 + j < p.n_out) {
        const float y = fmaxf(dot[j] + (p.bias ? __ldg(&p.bias[col0 + j]) : 0.f), 0.f);
        o[j] = y;
        r.ss = fmaf(y, y, r.ss);
      }
    }
  }
  static __device__ void tile_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int) {}
  static __device__ void item_end(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int, int) {
    if (!p.normalize) return;  // uniform over the grid
    float* sm = reinterpret_cast<float*>(ctx.smem);
    sm[ctx.half * tc::BM + ctx.erow] = r.ss;
    ptx::named_bar_sync(2, tc::EPI_WARPS * 32);
    const float ss = sm[ctx.erow] + sm[tc::BM + ctx.erow];
    ptx::named_bar_sync(2, tc::EPI_WARPS * 32);  // the next item may overwrite the exchange area
    if (!valid) return;
    const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));  // K.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))
    if (p.inv_norm != nullptr && ctx.half == 0) p.inv_norm[row] = ss >= 1e-12f ? inv : -inv;  // < 0: clamped row
    float* o = p.out + row * p.n_out;
    // this thread wrote columns [tile * 128 + half * 64, + 64) of every column tile: rescale exactly those
    for (int c0 = ctx.half * tc::COLS_PER_EPI_WARP; c0 < p.n_out; c0 += tc::BN)
      for (int j = 0; j < tc::COLS_PER_EPI_WARP && c0 + j < p.n_out; ++j) o[c0 + j] *= inv;
  }
};

// ---------------------------------------------------------------- backward (training through the head)
// y = [l2norm](relu(x W + b)).  With g = dL/dy:
//   gz   = g                                   (no normalisation)
//        = (g - y (y.g)) * inv                 (normalised rows: y = z * inv; clamped rows: g * inv)
//   gpre = gz where y > 0, else 0              (TF's ReLU gradient: passes where the input is > 0)
//   db = column sums of gpre,  dW = x^T gpre,  dx = gpre W^T   -- two 3xTF32 tcgen05 GEMMs with a store epilogue.
__global__ void dense_gpre_kernel(const float* __restrict__ y, const float* __restrict__ gy,
                                  const float* __restrict__ inv_norm, int64_t B, int n_out, int normalize,
                                  float* __restrict__ gpre) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* yr = y + row * n_out;
  const float* gr = gy + row * n_out;
  float* o = gpre + row * n_out;
  float inv = 1.f, dot = 0.f;
  if (normalize) {
    inv = inv_norm[row];
    if (inv > 0.f) {
      for (int c = lane; c < n_out; c += 32) dot = fmaf(yr[c], gr[c], dot);
      dot = warp_sum(dot);
    } else {
      inv = -inv;  // sum of squares below the clamp: y = z * 1e6, a plain scale
    }
  }
  for (int c = lane; c < n_out; c += 32) {
    const float yv = yr[c];
    const float gz = normalize ? (gr[c] - yv * dot) * inv : gr[c];
    o[c] = yv > 0.f ? gz : 0.f;
  }
}

// a (rows, cols) row-major -> a^T TF32 planes (cols, rpad), zero padded: K-major operands with K = rows
__global__ void transpose_planes_kernel(const float* __restrict__ a, int64_t rows, int cols, int64_t rpad,
                                        float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t rr = r0 + r;
    const int c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (rr < rows && c < cols) ? a[rr * cols + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const int64_t rr = r0 + threadIdx.x;
    if (c < cols && rr < rpad) {
      const float v = tile[threadIdx.x][r];
      const float h = tc::to_tf32(v);
      hi[static_cast<int64_t>(c) * rpad + rr] = h;
      lo[static_cast<int64_t>(c) * rpad + rr] = tc::to_tf32(v - h);
    }
  }
}

// column sums in float64, fixed order (deterministic): one block per 8 columns
__global__ void column_sum_kernel(const float* __restrict__ a, int64_t rows, int cols, float* __restrict__ out) {
  __shared__ double part[128][9];
  const int c = blockIdx.x * 8 + threadIdx.x;
  double acc = 0.0;
  if (c < cols)
    for (int64_t r = threadIdx.y; r < rows; r += 128) acc += static_cast<double>(a[r * cols + c]);
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  for (int h = 64; h > 0; h >>= 1) {
    if (static_cast<int>(threadIdx.y) < h) part[threadIdx.y][threadIdx.x] += part[threadIdx.y + h][threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.y == 0 && c < cols) out[c] = static_cast<float>(part[0][threadIdx.x]);
}

// plain store epilogue: out[row, col] = dot
struct EpStoreTile {
  struct Params {
    float* out;
    int64_t M;
    int N;
  };
  struct Row {};
  static constexpr int kSmemBytes = 0;
  static __device__ void item_begin(const Params&, Row&, const tc::Ctx&, int64_t, bool, int, int) {}
  static __device__ void chunk(const Params& p, Row&, const tc::Ctx&, int64_t row, bool valid, int64_t col0,
                               const float (&dot)[32]) {
    if (!valid || col0 >= p.N) return;
    float* o = p.out + row * p.N + col0;
    if ((p.N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && col0 + 32 <= p.N) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(dot[j], dot[j + 1], dot[j + 2], dot[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) o[j] = dot[j];
    }
  }
  static __device__ void tile_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int) {}
  static __device__ void item_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int, int) {}
};

// C (M, N) = A (M, K) . B (N, K)^T from TF32 planes (both K-major, K padded to kpad)
int gemm_store(const float* a_hi, const float* a_lo, int64_t M, const float* b_hi, const float* b_lo, int64_t N, int K,
               int kpad, float* out, cudaStream_t st) {
  CUtensorMap tah, tal, tbh, tbl;
  if (tc::make_plane_tmap(&tah, a_hi, M, kpad) || tc::make_plane_tmap(&tal, a_lo, M, kpad) ||
      tc::make_plane_tmap(&tbh, b_hi, N, kpad) || tc::make_plane_tmap(&tbl, b_lo, N, kpad))
    return fail(EN_ERR_DRIVER, "dense backward: cuTensorMapEncodeTiled failed");
  const int sms = device_sm_count();
  const int tiles_m = static_cast<int>((M + tc::BM - 1) / tc::BM);
  int splits = (2 * sms + tiles_m - 1) / tiles_m;
  tc::Shape sh = tc::make_shape(M, N, K, splits, 3);
  EpStoreTile::Params ep{out, M, static_cast<int>(N)};
  EN_CUDA(tc::launch<EpStoreTile>(tah, tal, tbh, tbl, sh, ep, sms, st));
  ++launch_counter();
  return EN_OK;
}

}  // namespace
}  // namespace en

using namespace en;

extern "C" {

size_t en_dense_plane_bytes(int n_in, int n_out) {
  if (n_in <= 0 || n_out <= 0) return 0;
  return static_cast<size_t>(n_out) * tc::dpad_for(n_in, 0) * 4;
}

int en_dense_prepare(const float* w, int n_in, int n_out, float* w_hi, float* w_lo, void* stream) {
  EN_REQUIRE(w && w_hi && w_lo && n_in > 0 && n_out > 0, "en_dense_prepare: bad arguments");
  const int kpad = tc::dpad_for(n_in, 0);
  dim3 grid(static_cast<unsigned>(kpad / 32), static_cast<unsigned>((n_out + 31) / 32));
  dense_weight_planes_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(w, n_in, n_out, kpad, w_hi, w_lo);
  EN_LAUNCHED("dense_weight_planes_kernel");
  return EN_OK;
}

size_t en_ws_bytes_dense(int64_t B, int n_in) {
  if (B <= 0 || n_in <= 0) return 0;
  return 2 * align_up(static_cast<size_t>(B) * tc::dpad_for(n_in, 0) * 4);
}

int en_dense_relu_fwd(const float* x, int64_t B, int n_in, const float* w_hi, const float* w_lo, const float* bias,
                      int n_out, int normalize, float* out, float* inv_norm, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(x && w_hi && w_lo && out && B > 0 && n_in > 0 && n_out > 0, "en_dense_relu_fwd: bad arguments");
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_dense(B, n_in))
    return fail(EN_ERR_WORKSPACE, "en_dense_relu_fwd: workspace too small (%zu < %zu)", ws_bytes,
                en_ws_bytes_dense(B, n_in));
  cudaStream_t st = as_stream(stream);
  Workspace w(ws, ws_bytes);
  const int kpad = tc::dpad_for(n_in, 0);
  float* xhi = w.take<float>(static_cast<size_t>(B) * kpad);
  float* xlo = w.take<float>(static_cast<size_t>(B) * kpad);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_dense_relu_fwd: workspace too small or misaligned");
  EN_CUDA(tc::launch_split(x, B, n_in, n_in, kpad, xhi, xlo, nullptr, st));
  ++launch_counter();
  CUtensorMap txh, txl, twh, twl;
  if (tc::make_plane_tmap(&txh, xhi, B, kpad) || tc::make_plane_tmap(&txl, xlo, B, kpad) ||
      tc::make_plane_tmap(&twh, w_hi, n_out, kpad) || tc::make_plane_tmap(&twl, w_lo, n_out, kpad))
    return fail(EN_ERR_DRIVER, "en_dense_relu_fwd: cuTensorMapEncodeTiled failed");
  tc::Shape sh = tc::make_shape(B, n_out, n_in, 1, 3);  // one work item per row tile: all column tiles in turn
  EpDense::Params ep{bias, out, normalize ? inv_norm : nullptr, B, n_out, normalize};
  prof_begin(st);
  EN_CUDA(tc::launch<EpDense>(txh, txl, twh, twl, sh, ep, device_sm_count(), st));
  prof_end(st);
  ++launch_counter();
  return EN_OK;
}

size_t en_ws_bytes_dense_bwd(int64_t B, int n_in, int n_out) {
  if (B <= 0 || n_in <= 0 || n_out <= 0) return 0;
  const size_t bpad = static_cast<size_t>((B + tc::BK - 1) / tc::BK * tc::BK);
  const size_t opad = static_cast<size_t>(tc::dpad_for(n_out, 0));
  return align_up(static_cast<size_t>(B) * n_out * 4) +            // gpre
         2 * align_up(static_cast<size_t>(B) * opad * 4) +         // gpre planes (K = n_out)
         2 * align_up(static_cast<size_t>(n_in) * opad * 4) +      // W planes in the Keras layout (K = n_out)
         2 * align_up(static_cast<size_t>(n_in) * bpad * 4) +      // x^T planes (K = B)
         2 * align_up(static_cast<size_t>(n_out) * bpad * 4);      // gpre^T planes (K = B)
}

// Backward of en_dense_relu_fwd.  y = the forward output, inv_norm = the row scales it stored (normalize != 0),
// gy = dL/dy; any of gx (B, n_in), gw (n_in, n_out: the Keras kernel layout), gb (n_out,) may be null.
int en_dense_relu_bwd(const float* x, int64_t B, int n_in, const float* w, int n_out, int normalize, const float* y,
                      const float* inv_norm, const float* gy, float* gx, float* gw, float* gb, void* ws,
                      size_t ws_bytes, void* stream) {
  EN_REQUIRE(x && w && y && gy && B > 0 && n_in > 0 && n_out > 0, "en_dense_relu_bwd: bad arguments");
  EN_REQUIRE(!normalize || inv_norm, "en_dense_relu_bwd: inv_norm (from the forward pass) is required");
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_dense_bwd(B, n_in, n_out))
    return fail(EN_ERR_WORKSPACE, "en_dense_relu_bwd: workspace too small (%zu < %zu)", ws_bytes,
                en_ws_bytes_dense_bwd(B, n_in, n_out));
  cudaStream_t st = as_stream(stream);
  Workspace wsp(ws, ws_bytes);
  const int64_t bpad = (B + tc::BK - 1) / tc::BK * tc::BK;
  const int opad = tc::dpad_for(n_out, 0);
  float* gpre = wsp.take<float>(static_cast<size_t>(B) * n_out);
  float* g_hi = wsp.take<float>(static_cast<size_t>(B) * opad);
  float* g_lo = wsp.take<float>(static_cast<size_t>(B) * opad);
  float* w_hi = wsp.take<float>(static_cast<size_t>(n_in) * opad);
  float* w_lo = wsp.take<float>(static_cast<size_t>(n_in) * opad);
  float* xt_hi = wsp.take<float>(static_cast<size_t>(n_in) * bpad);
  float* xt_lo = wsp.take<float>(static_cast<size_t>(n_in) * bpad);
  float* gt_hi = wsp.take<float>(static_cast<size_t>(n_out) * bpad);
  float* gt_lo = wsp.take<float>(static_cast<size_t>(n_out) * bpad);
  if (!wsp.ok()) return fail(EN_ERR_WORKSPACE, "en_dense_relu_bwd: workspace too small or misaligned");
  dense_gpre_kernel<<<static_cast<unsigned>((B * 32 + 255) / 256), 256, 0, st>>>(y, gy, inv_norm, B, n_out, normalize,
                                                                                 gpre);
  EN_LAUNCHED("dense_gpre_kernel");
  if (gb) {
    column_sum_kernel<<<static_cast<unsigned>((n_out + 7) / 8), dim3(8, 128), 0, st>>>(gpre, B, n_out, gb);
    EN_LAUNCHED("column_sum_kernel");
  }
  if (gx) {  // dx (B, n_in) = gpre (B, n_out) . W (n_in, n_out)^T
    EN_CUDA(tc::launch_split(gpre, B, n_out, n_out, opad, g_hi, g_lo, nullptr, st));
    EN_CUDA(tc::launch_split(w, n_in, n_out, n_out, opad, w_hi, w_lo, nullptr, st));
    launch_counter() += 2;
    if (int rc = gemm_store(g_hi, g_lo, B, w_hi, w_lo, n_in, n_out, opad, gx, st)) return rc;
  }
  if (gw) {  // dW (n_in, n_out) = x^T (n_in, B) . gpre^T (n_out, B)^T
    dim3 tb(32, 8);
    transpose_planes_kernel<<<dim3(static_cast<unsigned>(bpad / 32), static_cast<unsigned>((n_in + 31) / 32)), tb, 0, st>>>(
        x, B, n_in, bpad, xt_hi, xt_lo);
    EN_LAUNCHED("transpose_planes_kernel");
    transpose_planes_kernel<<<dim3(static_cast<unsigned>(bpad / 32), static_cast<unsigned>((n_out + 31) / 32)), tb, 0, st>>>(
        gpre, B, n_out, bpad, gt_hi, gt_lo);
    EN_LAUNCHED("transpose_planes_kernel");
    if (int rc = gemm_store(xt_hi, xt_lo, n_in, gt_hi, gt_lo, n_out, static_cast<int>(B), static_cast<int>(bpad), gw, st))
      return rc;
  }
  return EN_OK;
}

}  // extern "C"
