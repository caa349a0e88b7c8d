// Memory-bound row-wise kernels: L2 normalisation, the pre-mined [a|p|n] triplet hinge, element-wise contrastive
// loss / accuracy, and the Siamese distance heads -- forward and backward.  One warp per row, 128-bit loads when the
// row is 16-byte aligned, FP32 math with IEEE sqrt/div (no fast-math), reductions by warp shuffle.
//
// Reference call sites (under /root/reference): embedding_net/backbones.py:38,77,118 (l2_normalize),
// embedding_net/losses_and_accuracies.py:4-11,26-42,47-50, embedding_net/models.py:217-228.
#include <initializer_list>
#include "common.cuh"

namespace en {

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int THREADS = WARPS_PER_BLOCK * 32;

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------- l2 normalise
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int d) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * d;
  float* yr = y + row * d;
  const bool vec = (d & 3) == 0 && aligned16(xr) && aligned16(yr);
  float ss = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    for (int c = lane; c < d / 4; c += 32) {
      float4 v = x4[c];
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (int c = lane; c < d; c += 32) ss += xr[c] * xr[c];
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    float4* y4 = reinterpret_cast<float4*>(yr);
    for (int c = lane; c < d / 4; c += 32) {
      float4 v = x4[c];
      y4[c] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    }
  } else {
    for (int c = lane; c < d; c += 32) yr[c] = xr[c] * inv;
  }
}

// y = x*inv, inv = max(ss,eps)^-1/2  =>  gx = g*inv - [ss >= eps] * x * (g.x) * inv^3
__global__ void l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                  int64_t rows, int d) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * d;
  const float* gr = gy + row * d;
  float* or_ = gx + row * d;
  float ss = 0.f, dot = 0.f;
  for (int c = lane; c < d; c += 32) {
    float v = xr[c];
    ss += v * v;
    dot += v * gr[c];
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
  const float k = ss >= 1e-12f ? dot * inv * inv * inv : 0.f;
  for (int c = lane; c < d; c += 32) or_[c] = gr[c] * inv - xr[c] * k;
}

// ---------------------------------------------------------------- triplet [a|p|n]
__global__ void triplet_apn_fwd_kernel(const float* __restrict__ y, int64_t B, int third, float margin,
                                       float* __restrict__ loss) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* a = y + row * 3 * third;
  const float* p = a + third;
  const float* n = p + third;
  float pos = 0.f, neg = 0.f;
  if ((third & 3) == 0 && aligned16(a)) {
    const float4 *a4 = reinterpret_cast<const float4*>(a), *p4 = reinterpret_cast<const float4*>(p),
                 *n4 = reinterpret_cast<const float4*>(n);
    for (int c = lane; c < third / 4; c += 32) {
      float4 va = a4[c], vp = p4[c], vn = n4[c];
      float t;
      t = va.x - vp.x; pos += t * t; t = va.y - vp.y; pos += t * t;
      t = va.z - vp.z; pos += t * t; t = va.w - vp.w; pos += t * t;
      t = va.x - vn.x; neg += t * t; t = va.y - vn.y; neg += t * t;
      t = va.z - vn.z; neg += t * t; t = va.w - vn.w; neg += t * t;
    }
  } else {
    for (int c = lane; c < third; c += 32) {
      float t = a[c] - p[c];
      pos += t * t;
      t = a[c] - n[c];
      neg += t * t;
    }
  }
  pos = warp_sum(pos);
  neg = warp_sum(neg);
  if (lane == 0) loss[row] = fmaxf(__fadd_rn(__fsub_rn(pos, neg), margin), 0.f);
}

__global__ void triplet_apn_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gloss, int64_t B,
                                       int third, float margin, float* __restrict__ gy) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* a = y + row * 3 * third;
  const float* p = a + third;
  const float* n = p + third;
  float pos = 0.f, neg = 0.f;
  for (int c = lane; c < third; c += 32) {
    float t = a[c] - p[c];
    pos += t * t;
    t = a[c] - n[c];
    neg += t * t;
  }
  pos = warp_sum(pos);
  neg = warp_sum(neg);
  // TF routes maximum(x, 0)'s gradient to x when x >= 0.
  const float g = (__fadd_rn(__fsub_rn(pos, neg), margin) >= 0.f) ? gloss[row] : 0.f;
  float* ga = gy + row * 3 * third;
  float* gp = ga + third;
  float* gn = gp + third;
  for (int c = lane; c < third; c += 32) {
    const float ap = a[c] - p[c], an = a[c] - n[c];
    ga[c] = g * 2.f * (ap - an);
    gp[c] = -g * 2.f * ap;
    gn[c] = g * 2.f * an;
  }
}

// ---------------------------------------------------------------- register-cached vector variants
// Rows of up to 1024 floats (16-byte aligned, d % 4 == 0): each lane keeps its NV float4 of every input row in
// registers, so ALL loads of a warp are issued back to back before the first dependent instruction (NV x 512 bytes in
// flight per warp and input), the row is read from memory exactly once, and outputs are written with 128-bit
// streaming stores.  Bound: HBM; algorithmic bytes per row are listed with each kernel.
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ float sq4(float4 v) { return v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 scale4(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

template <int NV>
__device__ __forceinline__ void load_row(float4 (&v)[NV], const float* row, int d4, int lane) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < d4 ? ld_stream(r4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int NV>
__device__ __forceinline__ void store_row(float* row, const float4 (&v)[NV], int d4, int lane) {
  float4* r4 = reinterpret_cast<float4*>(row);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < d4) st_stream(r4 + c, v[i]);
  }
}

// bytes per row: 4 d read + 4 d written
template <int NV>
__global__ void __launch_bounds__(THREADS) l2norm_fwd_vec_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                 int64_t rows, int d) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4 v[NV];
  load_row<NV>(v, x + row * d, d / 4, lane);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) ss += sq4(v[i]);
  ss = warp_sum(ss);
  const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = scale4(v[i], inv);
  store_row<NV>(y + row * d, v, d / 4, lane);
}

// bytes per row: 8 d read + 4 d written
template <int NV>
__global__ void __launch_bounds__(THREADS) l2norm_bwd_vec_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ gy,
                                                                 float* __restrict__ gx, int64_t rows, int d) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4 v[NV], g[NV];
  load_row<NV>(v, x + row * d, d / 4, lane);
  load_row<NV>(g, gy + row * d, d / 4, lane);
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    ss += sq4(v[i]);
    dot += v[i].x * g[i].x + v[i].y * g[i].y + v[i].z * g[i].z + v[i].w * g[i].w;
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
  const float k = ss >= 1e-12f ? dot * inv * inv * inv : 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    g[i] = make_float4(g[i].x * inv - v[i].x * k, g[i].y * inv - v[i].y * k, g[i].z * inv - v[i].z * k,
                       g[i].w * inv - v[i].w * k);
  store_row<NV>(gx + row * d, g, d / 4, lane);
}

// bytes per row: 12 third read (+ 4 written)
template <int NV>
__global__ void __launch_bounds__(THREADS) triplet_apn_fwd_vec_kernel(const float* __restrict__ y, int64_t B, int third,
                                                                      float margin, float* __restrict__ loss) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* a = y + row * 3 * third;
  float4 va[NV], vp[NV], vn[NV];
  load_row<NV>(va, a, third / 4, lane);
  load_row<NV>(vp, a + third, third / 4, lane);
  load_row<NV>(vn, a + 2 * third, third / 4, lane);
  float pos = 0.f, neg = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 dp = sub4(va[i], vp[i]), dn = sub4(va[i], vn[i]);
    pos += dp.x * dp.x; pos += dp.y * dp.y; pos += dp.z * dp.z; pos += dp.w * dp.w;
    neg += dn.x * dn.x; neg += dn.y * dn.y; neg += dn.z * dn.z; neg += dn.w * dn.w;
  }
  pos = warp_sum(pos);
  neg = warp_sum(neg);
  if (lane == 0) loss[row] = fmaxf(__fadd_rn(__fsub_rn(pos, neg), margin), 0.f);
}

// bytes per row: 12 third read + 12 third written
template <int NV>
__global__ void __launch_bounds__(THREADS) triplet_apn_bwd_vec_kernel(const float* __restrict__ y,
                                                                      const float* __restrict__ gloss, int64_t B,
                                                                      int third, float margin, float* __restrict__ gy) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* a = y + row * 3 * third;
  float4 va[NV], vp[NV], vn[NV];
  load_row<NV>(va, a, third / 4, lane);
  load_row<NV>(vp, a + third, third / 4, lane);
  load_row<NV>(vn, a + 2 * third, third / 4, lane);
  float pos = 0.f, neg = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    vp[i] = sub4(va[i], vp[i]);  // a - p
    vn[i] = sub4(va[i], vn[i]);  // a - n
    pos += vp[i].x * vp[i].x; pos += vp[i].y * vp[i].y; pos += vp[i].z * vp[i].z; pos += vp[i].w * vp[i].w;
    neg += vn[i].x * vn[i].x; neg += vn[i].y * vn[i].y; neg += vn[i].z * vn[i].z; neg += vn[i].w * vn[i].w;
  }
  pos = warp_sum(pos);
  neg = warp_sum(neg);
  // TF routes maximum(x, 0)'s gradient to x when x >= 0.
  const float g2 = 2.f * ((__fadd_rn(__fsub_rn(pos, neg), margin) >= 0.f) ? gloss[row] : 0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    va[i] = scale4(sub4(vp[i], vn[i]), g2);
    vp[i] = scale4(vp[i], -g2);
    vn[i] = scale4(vn[i], g2);
  }
  float* ga = gy + row * 3 * third;
  store_row<NV>(ga, va, third / 4, lane);
  store_row<NV>(ga + third, vp, third / 4, lane);
  store_row<NV>(ga + 2 * third, vn, third / 4, lane);
}

// bytes per row: 8 d read (+ 4 written).  kClamp: the Siamese head's sqrt(max(s, 1e-7)) (models.py:225); without it
// the plain Euclidean distance, and e2 may be ONE row broadcast against all rows of e1 (stride2 = 0): the
// ``calculate_distances(encoding)`` that ``EmbeddingNet.predict`` calls (models.py:123) over the bank.
template <int NV, bool kClamp>
__global__ void __launch_bounds__(THREADS) row_dist_vec_kernel(const float* __restrict__ e1, const float* __restrict__ e2,
                                                               int64_t B, int d, int64_t stride2,
                                                               float* __restrict__ dist) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float4 a[NV], b[NV];
  load_row<NV>(a, e1 + row * d, d / 4, lane);
  if (stride2 == 0) {
    const float4* q4 = reinterpret_cast<const float4*>(e2);
#pragma unroll
    for (int i = 0; i < NV; ++i) b[i] = lane + 32 * i < d / 4 ? __ldg(q4 + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    load_row<NV>(b, e2 + row * stride2, d / 4, lane);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 t = sub4(a[i], b[i]);
    s += t.x * t.x; s += t.y * t.y; s += t.z * t.z; s += t.w * t.w;
  }
  s = warp_sum(s);
  if (lane == 0) dist[row] = sqrtf(kClamp ? fmaxf(s, 1e-7f) : s);
}

// bytes per row: 8 d read + 8 d written
template <int NV>
__global__ void __launch_bounds__(THREADS) siamese_l2_bwd_vec_kernel(const float* __restrict__ e1,
                                                                     const float* __restrict__ e2,
                                                                     const float* __restrict__ gdist, int64_t B, int d,
                                                                     float* __restrict__ g1, float* __restrict__ g2) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float4 a[NV], b[NV];
  load_row<NV>(a, e1 + row * d, d / 4, lane);
  load_row<NV>(b, e2 + row * d, d / 4, lane);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    a[i] = sub4(a[i], b[i]);
    s += a[i].x * a[i].x; s += a[i].y * a[i].y; s += a[i].z * a[i].z; s += a[i].w * a[i].w;
  }
  s = warp_sum(s);
  const float k = s >= 1e-7f ? gdist[row] / sqrtf(s) : 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    a[i] = scale4(a[i], k);
    b[i] = scale4(a[i], -1.f);
  }
  store_row<NV>(g1 + row * d, a, d / 4, lane);
  store_row<NV>(g2 + row * d, b, d / 4, lane);
}

// x *= s[0], in place (the upstream gradient of a scalar loss applied to a stored d loss / d emb)
__global__ void scale_inplace_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ s) {
  const float k = __ldg(s);
  if (k == 1.0f) return;  // plain loss.backward(): nothing to do, no traffic
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    float4* x4 = reinterpret_cast<float4*>(x);
    for (int64_t c = i; c < n / 4; c += stride) x4[c] = scale4(x4[c], k);
    for (int64_t c = (n / 4) * 4 + i; c < n; c += stride) x[c] *= k;
  } else {
    for (; i < n; i += stride) x[i] *= k;
  }
}

// NV = float4 per lane needed for a row of `len` floats; 0 = not eligible for the vector kernels
inline int vec_nv(int len, std::initializer_list<const void*> ptrs) {
  if ((len & 3) != 0 || len > 1024) return 0;
  for (const void* p : ptrs)
    if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) return 0;
  const int need = (len / 4 + 31) / 32;
  return need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : 8;
}
#define EN_VEC_DISPATCH(nv, KERNEL, ...)                                                   \
  do {                                                                                     \
    switch (nv) {                                                                          \
      case 1: KERNEL<1> __VA_ARGS__; break;                                                \
      case 2: KERNEL<2> __VA_ARGS__; break;                                                \
      case 4: KERNEL<4> __VA_ARGS__; break;                                                \
      default: KERNEL<8> __VA_ARGS__; break;                                               \
    }                                                                                      \
  } while (0)

// ---------------------------------------------------------------- element-wise contrastive / accuracy
// (B,1)-shaped inputs: a single block with a deterministic tree reduction in double.
__device__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

__global__ void contrastive_fwd_kernel(const float* __restrict__ yt, const float* __restrict__ yp, int64_t n,
                                       float* __restrict__ loss) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = yp[i], t = yt[i];
    const float m = fmaxf(1.0f - d, 0.f);
    acc += static_cast<double>(t * (d * d) + (1.0f - t) * (m * m));
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) loss[0] = static_cast<float>(acc / static_cast<double>(n));
}

__global__ void contrastive_bwd_kernel(const float* __restrict__ yt, const float* __restrict__ yp,
                                       const float* __restrict__ gloss, int64_t n, float* __restrict__ g) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = yp[i], t = yt[i];
  const float scale = gloss[0] / static_cast<float>(n);
  g[i] = scale * (t * 2.f * d - (1.f - t) * 2.f * fmaxf(1.f - d, 0.f));
}

__global__ void pair_accuracy_kernel(const float* __restrict__ yt, const float* __restrict__ yp, int64_t n,
                                     float* __restrict__ acc_out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += (yt[i] == (yp[i] < 0.5f ? 1.0f : 0.0f)) ? 1.0 : 0.0;
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) acc_out[0] = static_cast<float>(acc / static_cast<double>(n));
}

// ---------------------------------------------------------------- Siamese heads
__global__ void siamese_l2_fwd_kernel(const float* __restrict__ e1, const float* __restrict__ e2, int64_t B, int d,
                                      float* __restrict__ dist) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* a = e1 + row * d;
  const float* b = e2 + row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float t = a[c] - b[c];
    s += t * t;
  }
  s = warp_sum(s);
  if (lane == 0) dist[row] = sqrtf(fmaxf(s, 1e-7f));
}

// plain Euclidean distance of every row of `bank` to ONE query row (generic-shape fallback of row_dist_vec_kernel)
__global__ void row_dist_query_kernel(const float* __restrict__ bank, const float* __restrict__ q, int64_t n, int d,
                                      float* __restrict__ dist) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* a = bank + row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float t = a[c] - __ldg(&q[c]);
    s += t * t;
  }
  s = warp_sum(s);
  if (lane == 0) dist[row] = sqrtf(s);
}

__global__ void siamese_l2_bwd_kernel(const float* __restrict__ e1, const float* __restrict__ e2,
                                      const float* __restrict__ gdist, int64_t B, int d, float* __restrict__ g1,
                                      float* __restrict__ g2) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* a = e1 + row * d;
  const float* b = e2 + row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float t = a[c] - b[c];
    s += t * t;
  }
  s = warp_sum(s);
  // d/ds sqrt(max(s, eps)) = [s >= eps] / (2 sqrt(s));  ds/de1 = 2 (e1 - e2)
  const float k = s >= 1e-7f ? gdist[row] / sqrtf(s) : 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = k * (a[c] - b[c]);
    g1[row * d + c] = v;
    g2[row * d + c] = -v;
  }
}

__global__ void siamese_l1_fwd_kernel(const float* __restrict__ e1, const float* __restrict__ e2, int64_t n,
                                      float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fabsf(e1[i] - e2[i]);
}
__global__ void siamese_l1_bwd_kernel(const float* __restrict__ e1, const float* __restrict__ e2,
                                      const float* __restrict__ go, int64_t n, float* __restrict__ g1,
                                      float* __restrict__ g2) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = e1[i] - e2[i];
  const float s = t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f);
  g1[i] = s * go[i];
  g2[i] = -s * go[i];
}

inline unsigned row_blocks(int64_t rows) { return static_cast<unsigned>((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }
inline unsigned elem_blocks(int64_t n) { return static_cast<unsigned>((n + THREADS - 1) / THREADS); }

// Gather of the mined triplets' rows (datagenerators.py:241-243,252-256: A, P, N batches are the rows of the sampled
// set picked by the mined indices).  One block per (triplet, slot); 128-bit copies when rows are 16-byte multiples.
__global__ void gather_triplet_rows_kernel(const float* __restrict__ src, int64_t n_rows, int64_t row_len,
                                           const int64_t* __restrict__ trip, int64_t T, float* __restrict__ a,
                                           float* __restrict__ p, float* __restrict__ n) {
  const int64_t t = blockIdx.x;
  const int slot = blockIdx.y;
  const int64_t r = trip[t * 3 + slot];
  float* dst = (slot == 0 ? a : (slot == 1 ? p : n)) + t * row_len;
  if (r < 0 || r >= n_rows) {  // never produced by the miner; keep the output defined
    for (int64_t c = threadIdx.x; c < row_len; c += blockDim.x) dst[c] = 0.f;
    return;
  }
  const float* from = src + r * row_len;
  if ((row_len & 3) == 0 && ((reinterpret_cast<uintptr_t>(from) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const float4* f4 = reinterpret_cast<const float4*>(from);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int64_t c = threadIdx.x; c < row_len / 4; c += blockDim.x) d4[c] = __ldg(f4 + c);
  } else {
    for (int64_t c = threadIdx.x; c < row_len; c += blockDim.x) dst[c] = from[c];
  }
}

}  // namespace
}  // namespace en

using namespace en;

extern "C" {

int en_l2_normalize_fwd(const float* x, float* y, int64_t rows, int d, void* stream) {
  EN_REQUIRE(x && y && rows >= 0 && d > 0, "en_l2_normalize_fwd: bad arguments (rows=%lld d=%d)", (long long)rows, d);
  if (rows == 0) return EN_OK;
  if (const int nv = vec_nv(d, {x, y}))
    EN_VEC_DISPATCH(nv, l2norm_fwd_vec_kernel, <<<row_blocks(rows), THREADS, 0, as_stream(stream)>>>(x, y, rows, d));
  else
    l2norm_fwd_kernel<<<row_blocks(rows), THREADS, 0, as_stream(stream)>>>(x, y, rows, d);
  EN_LAUNCHED("l2norm_fwd_kernel");
  return EN_OK;
}

int en_l2_normalize_bwd(const float* x, const float* gy, float* gx, int64_t rows, int d, void* stream) {
  EN_REQUIRE(x && gy && gx && rows >= 0 && d > 0, "en_l2_normalize_bwd: bad arguments");
  if (rows == 0) return EN_OK;
  if (const int nv = vec_nv(d, {x, gy, gx}))
    EN_VEC_DISPATCH(nv, l2norm_bwd_vec_kernel,
                    <<<row_blocks(rows), THREADS, 0, as_stream(stream)>>>(x, gy, gx, rows, d));
  else
    l2norm_bwd_kernel<<<row_blocks(rows), THREADS, 0, as_stream(stream)>>>(x, gy, gx, rows, d);
  EN_LAUNCHED("l2norm_bwd_kernel");
  return EN_OK;
}

int en_triplet_apn_fwd(const float* y_pred, int64_t B, int total_len, float margin, float* loss, void* stream) {
  EN_REQUIRE(y_pred && loss && B >= 0 && total_len > 0, "en_triplet_apn_fwd: bad arguments");
  EN_REQUIRE(total_len % 3 == 0,
             "en_triplet_apn_fwd: y_pred last dimension (%d) must be a multiple of 3 ([a|p|n] thirds, lac:29-31)",
             total_len);
  if (B == 0) return EN_OK;
  if (const int nv = vec_nv(total_len / 3, {y_pred}))
    EN_VEC_DISPATCH(nv, triplet_apn_fwd_vec_kernel,
                    <<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(y_pred, B, total_len / 3, margin, loss));
  else
    triplet_apn_fwd_kernel<<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(y_pred, B, total_len / 3, margin, loss);
  EN_LAUNCHED("triplet_apn_fwd_kernel");
  return EN_OK;
}

int en_triplet_apn_bwd(const float* y_pred, const float* gloss, int64_t B, int total_len, float margin,
                       float* gy_pred, void* stream) {
  EN_REQUIRE(y_pred && gloss && gy_pred && B >= 0 && total_len > 0 && total_len % 3 == 0,
             "en_triplet_apn_bwd: bad arguments");
  if (B == 0) return EN_OK;
  if (const int nv = vec_nv(total_len / 3, {y_pred, gy_pred}))
    EN_VEC_DISPATCH(nv, triplet_apn_bwd_vec_kernel, <<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(
                                                        y_pred, gloss, B, total_len / 3, margin, gy_pred));
  else
    triplet_apn_bwd_kernel<<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(y_pred, gloss, B, total_len / 3, margin,
                                                                             gy_pred);
  EN_LAUNCHED("triplet_apn_bwd_kernel");
  return EN_OK;
}

int en_contrastive_fwd(const float* y_true, const float* y_pred, int64_t n, float* loss, void* stream) {
  EN_REQUIRE(y_true && y_pred && loss && n > 0, "en_contrastive_fwd: bad arguments");
  contrastive_fwd_kernel<<<1, 1024, 0, as_stream(stream)>>>(y_true, y_pred, n, loss);
  EN_LAUNCHED("contrastive_fwd_kernel");
  return EN_OK;
}

int en_contrastive_bwd(const float* y_true, const float* y_pred, const float* gloss, int64_t n, float* gy_pred,
                       void* stream) {
  EN_REQUIRE(y_true && y_pred && gloss && gy_pred && n > 0, "en_contrastive_bwd: bad arguments");
  contrastive_bwd_kernel<<<elem_blocks(n), THREADS, 0, as_stream(stream)>>>(y_true, y_pred, gloss, n, gy_pred);
  EN_LAUNCHED("contrastive_bwd_kernel");
  return EN_OK;
}

int en_pair_accuracy(const float* y_true, const float* y_pred, int64_t n, float* acc, void* stream) {
  EN_REQUIRE(y_true && y_pred && acc && n > 0, "en_pair_accuracy: bad arguments");
  pair_accuracy_kernel<<<1, 1024, 0, as_stream(stream)>>>(y_true, y_pred, n, acc);
  EN_LAUNCHED("pair_accuracy_kernel");
  return EN_OK;
}

int en_siamese_l2_fwd(const float* e1, const float* e2, int64_t B, int d, float* dist, void* stream) {
  EN_REQUIRE(e1 && e2 && dist && B >= 0 && d > 0, "en_siamese_l2_fwd: bad arguments");
  if (B == 0) return EN_OK;
  if (const int nv = vec_nv(d, {e1, e2})) {
    switch (nv) {
      case 1: row_dist_vec_kernel<1, true><<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(e1, e2, B, d, d, dist); break;
      case 2: row_dist_vec_kernel<2, true><<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(e1, e2, B, d, d, dist); break;
      case 4: row_dist_vec_kernel<4, true><<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(e1, e2, B, d, d, dist); break;
      default: row_dist_vec_kernel<8, true><<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(e1, e2, B, d, d, dist); break;
    }
  } else {
    siamese_l2_fwd_kernel<<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(e1, e2, B, d, dist);
  }
  EN_LAUNCHED("siamese_l2_fwd_kernel");
  return EN_OK;
}

int en_siamese_l2_bwd(const float* e1, const float* e2, const float* gdist, int64_t B, int d, float* g1, float* g2,
                      void* stream) {
  EN_REQUIRE(e1 && e2 && gdist && g1 && g2 && B >= 0 && d > 0, "en_siamese_l2_bwd: bad arguments");
  if (B == 0) return EN_OK;
  if (const int nv = vec_nv(d, {e1, e2, g1, g2}))
    EN_VEC_DISPATCH(nv, siamese_l2_bwd_vec_kernel,
                    <<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(e1, e2, gdist, B, d, g1, g2));
  else
    siamese_l2_bwd_kernel<<<row_blocks(B), THREADS, 0, as_stream(stream)>>>(e1, e2, gdist, B, d, g1, g2);
  EN_LAUNCHED("siamese_l2_bwd_kernel");
  return EN_OK;
}

int en_siamese_l1_fwd(const float* e1, const float* e2, int64_t n, float* out, void* stream) {
  EN_REQUIRE(e1 && e2 && out && n >= 0, "en_siamese_l1_fwd: bad arguments");
  if (n == 0) return EN_OK;
  siamese_l1_fwd_kernel<<<elem_blocks(n), THREADS, 0, as_stream(stream)>>>(e1, e2, n, out);
  EN_LAUNCHED("siamese_l1_fwd_kernel");
  return EN_OK;
}

int en_siamese_l1_bwd(const float* e1, const float* e2, const float* gout, int64_t n, float* g1, float* g2,
                      void* stream) {
  EN_REQUIRE(e1 && e2 && gout && g1 && g2 && n >= 0, "en_siamese_l1_bwd: bad arguments");
  if (n == 0) return EN_OK;
  siamese_l1_bwd_kernel<<<elem_blocks(n), THREADS, 0, as_stream(stream)>>>(e1, e2, gout, n, g1, g2);
  EN_LAUNCHED("siamese_l1_bwd_kernel");
  return EN_OK;
}

int en_query_distances(const float* bank, const float* query, int64_t n, int d, float* dist, void* stream) {
  EN_REQUIRE(bank && query && dist && n >= 0 && d > 0, "en_query_distances: bad arguments");
  if (n == 0) return EN_OK;
  cudaStream_t st = as_stream(stream);
  if (const int nv = vec_nv(d, {bank, query})) {
    switch (nv) {
      case 1: row_dist_vec_kernel<1, false><<<row_blocks(n), THREADS, 0, st>>>(bank, query, n, d, 0, dist); break;
      case 2: row_dist_vec_kernel<2, false><<<row_blocks(n), THREADS, 0, st>>>(bank, query, n, d, 0, dist); break;
      case 4: row_dist_vec_kernel<4, false><<<row_blocks(n), THREADS, 0, st>>>(bank, query, n, d, 0, dist); break;
      default: row_dist_vec_kernel<8, false><<<row_blocks(n), THREADS, 0, st>>>(bank, query, n, d, 0, dist); break;
    }
  } else {
    row_dist_query_kernel<<<row_blocks(n), THREADS, 0, st>>>(bank, query, n, d, dist);
  }
  EN_LAUNCHED("row_dist_query_kernel");
  return EN_OK;
}

int en_scale_inplace(float* x, int64_t n, const float* scale, void* stream) {
  EN_REQUIRE(x && scale && n >= 0, "en_scale_inplace: bad arguments");
  if (n == 0) return EN_OK;
  int64_t blocks = (n / 4 + THREADS - 1) / THREADS + 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  scale_inplace_kernel<<<static_cast<unsigned>(blocks), THREADS, 0, as_stream(stream)>>>(x, n, scale);
  EN_LAUNCHED("scale_inplace_kernel");
  return EN_OK;
}

int en_gather_triplet_rows(const float* src, int64_t n_rows, int64_t row_len, const int64_t* triplets, int64_t T,
                           float* a, float* p, float* n, void* stream) {
  EN_REQUIRE(src && triplets && a && p && n && n_rows > 0 && row_len > 0 && T >= 0,
             "en_gather_triplet_rows: bad arguments");
  EN_REQUIRE(T < (int64_t(1) << 31), "en_gather_triplet_rows: too many triplets");
  if (T == 0) return EN_OK;
  const int threads = row_len >= 4096 ? 256 : 128;
  gather_triplet_rows_kernel<<<dim3(static_cast<unsigned>(T), 3), threads, 0, as_stream(stream)>>>(
      src, n_rows, row_len, triplets, T, a, p, n);
  EN_LAUNCHED("gather_triplet_rows_kernel");
  return EN_OK;
}

}  // extern "C"
