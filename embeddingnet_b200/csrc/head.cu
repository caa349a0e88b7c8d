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
    float* o = p.out + row * p.n_out;
    // this thread wrote columns [tile * 128 + half * 64, + 64) of every column tile: rescale exactly those
    for (int c0 = ctx.half * tc::COLS_PER_EPI_WARP; c0 < p.n_out; c0 += tc::BN)
      for (int j = 0; j < tc::COLS_PER_EPI_WARP && c0 + j < p.n_out; ++j) o[c0 + j] *= inv;
  }
};

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
                      int n_out, int normalize, float* out, void* ws, size_t ws_bytes, void* stream) {
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
  EpDense::Params ep{bias, out, B, n_out, normalize};
  prof_begin(st);
  EN_CUDA(tc::launch<EpDense>(txh, txl, twh, twl, sh, ep, device_sm_count(), st));
  prof_end(st);
  ++launch_counter();
  return EN_OK;
}

}  // extern "C"
