// Encoding-bank nearest neighbours: the scan the reference leaves to scikit-learn's brute-force
// KNeighborsClassifier (embedding_net/models.py:15,58,136-138) and to the undefined `calculate_distances` + np.argmin
// of EmbeddingNet.predict (embedding_net/models.py:123-124).
//
//   stage 1a  en_knn_shard_topk : tcgen05 distance GEMM (queries x bank shard; split-BF16 or 3xTF32 planes) with a
//                                 per-query running top-(k+slack) kept in registers by the epilogue thread that owns
//                                 the query row;
//   stage 1b  en_knn_stream_topk: for a handful of queries (the reference's one-image-per-call pattern) a CUDA-core
//                                 fp32 streaming scan bounded by HBM bandwidth;
//   stage 2   exact re-rank     : the surviving candidates are re-evaluated as float64 sum (q-b)^2 and ordered by
//                                 (distance, global id) -- lowest id wins ties, shard-invariant by construction;
//                                 a CERTIFICATE per query proves that no row the scan rejected can beat the k-th
//                                 result (rigorous bound on the scan's arithmetic error); queries without one are
//                                 flagged and redone by
//   stage 3   en_knn_exact_topk : float64 brute force over the shard for those few queries;
//   merge / vote / accuracy     : k-way merge of per-shard lists (after the NCCL all-gather), majority vote
//                                 (KNeighborsClassifier.predict), top-1 / top-5 tallies (models.py:144-161).
#include "common.cuh"
#include "tc_engine.cuh"

namespace en {
namespace {

constexpr float kInf = 3.0e38f;
constexpr int kChunkTilesPerSplit = 32;  // bank tiles per (launch, column range): 16 MB of planes at d = 512

struct Cand {
  float t;      // ranking proxy, ascending = nearer
  int32_t idx;  // row inside the shard, -1 = empty
};

// sorted insert into a register-resident list (ascending by t; arrival order breaks ties, and rows arrive in
// ascending id order, so equal proxies keep the lower id first)
template <int KC>
__device__ __forceinline__ void list_insert(float (&v)[KC], int32_t (&id)[KC], float t, int32_t i) {
  v[KC - 1] = t;
  id[KC - 1] = i;
#pragma unroll
  for (int q = KC - 1; q > 0; --q) {
    if (v[q] < v[q - 1]) {
      const float tv = v[q]; v[q] = v[q - 1]; v[q - 1] = tv;
      const int32_t ti = id[q]; id[q] = id[q - 1]; id[q - 1] = ti;
    }
  }
}

// ---------------------------------------------------------------- stage 1a: tensor-core scan epilogue
template <int KC>
struct EpTopK {
  struct Params {
    const float* bank_norms;
    const int32_t* bank_labels;   // may be null
    const int32_t* query_labels;  // may be null (then no exclusion)
    Cand* lists;                  // [Q][n_lists][KC]
    int64_t n_bank;
    int n_lists;
    int resume;                   // != 0: continue from the lists left by the previous bank chunk
  };
  struct Row {
    float v[KC];
    int32_t id[KC];
    int32_t qlabel;
    bool exclude;
  };
  static constexpr int kSmemBytes = 0;
  static __device__ void item_begin(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int,
                                    int split) {
    if (p.resume && valid) {
      const Cand* in = p.lists + (row * p.n_lists + split * tc::EPI_H + ctx.half) * KC;
#pragma unroll
      for (int q = 0; q < KC; ++q) {
        const Cand c = in[q];
        r.v[q] = c.t;
        r.id[q] = c.idx;
      }
    } else {
#pragma unroll
      for (int q = 0; q < KC; ++q) {
        r.v[q] = kInf;
        r.id[q] = -1;
      }
    }
    r.exclude = p.query_labels != nullptr && p.bank_labels != nullptr;
    r.qlabel = (r.exclude && valid) ? p.query_labels[row] : 0;
  }
  static __device__ void chunk(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int64_t col0,
                               const float (&dot)[32]) {
    if (col0 >= p.n_bank) return;
    tc::stage_columns(ctx, p.bank_norms, nullptr, col0, p.n_bank);
    const int ncols = static_cast<int>(p.n_bank - col0 < 32 ? p.n_bank - col0 : 32);
    const float4* n4 = reinterpret_cast<const float4*>(ctx.wf);
    // fast reject: most chunks contain nothing below the current k-th best once the list has warmed up
    float tmin = kInf;
    float t[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 nb = n4[g];
      t[4 * g + 0] = fmaf(-2.f, dot[4 * g + 0], nb.x);
      t[4 * g + 1] = fmaf(-2.f, dot[4 * g + 1], nb.y);
      t[4 * g + 2] = fmaf(-2.f, dot[4 * g + 2], nb.z);
      t[4 * g + 3] = fmaf(-2.f, dot[4 * g + 3], nb.w);
    }
    if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) tmin = fminf(tmin, t[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) tmin = fminf(tmin, j < ncols ? t[j] : kInf);
    }
    if (!__any_sync(0xffffffffu, tmin < r.v[KC - 1])) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncols && t[j] < r.v[KC - 1]) {
        const int64_t c = col0 + j;
        if (!r.exclude || __ldg(&p.bank_labels[c]) != r.qlabel) list_insert<KC>(r.v, r.id, t[j], static_cast<int32_t>(c));
      }
    }
  }
  static __device__ void tile_end(const Params&, Row&, const tc::Ctx&, int64_t, bool, int) {}
  static __device__ void item_end(const Params& p, Row& r, const tc::Ctx& ctx, int64_t row, bool valid, int,
                                  int split) {
    if (!valid) return;
    Cand* out = p.lists + (row * p.n_lists + split * tc::EPI_H + ctx.half) * KC;
#pragma unroll
    for (int q = 0; q < KC; ++q) out[q] = Cand{r.v[q], r.id[q]};
  }
};

// ---------------------------------------------------------------- stage 1b: HBM-bound streaming scan
// The reference's call pattern is one query per predict() (models.py:122,135): the scan is then bounded by HBM
// bandwidth, not by math.  One warp streams two bank rows at a time (all 128-bit loads of both rows are issued
// before any arithmetic, evict-first), the (<= 8, zero padded to QT) queries sit in shared memory, lane q keeps
// query q's running list, and the eight warps of a block merge their lists in shared memory before writing out.
// With bank norms available (a fitted bank) the proxy is |b|^2 - 2 q.b: one FMA per (element, query) instead of a
// subtract + FMA, which keeps the 8-query case on the HBM roofline instead of the FP32 pipe.
template <int KC, int DV, int QT, bool NORM>  // DV = float4 per lane per row (d == 128*DV), 0 = any d; QT = padded Q
__global__ void __launch_bounds__(256)
knn_stream_kernel(const float* __restrict__ queries, int Q, int d, const float* __restrict__ bank,
                  const float* __restrict__ bank_norms, int64_t n_bank, int64_t rows_per_warp,
                  Cand* __restrict__ lists) {
  extern __shared__ float qs[];  // [QT][d], then the block's merge area
  for (int i = threadIdx.x; i < QT * d; i += blockDim.x) qs[i] = i < Q * d ? queries[i] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gwarp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
  const int64_t r0 = gwarp * rows_per_warp;
  const int64_t r1 = min(r0 + rows_per_warp, n_bank);
  float v[KC];
  int32_t id[KC];
#pragma unroll
  for (int q = 0; q < KC; ++q) {
    v[q] = kInf;
    id[q] = -1;
  }
  for (int64_t r = r0; r < r1; r += 2) {
    const bool two = r + 1 < r1;
    float s0[QT], s1[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) s0[q] = s1[q] = 0.f;
    if (DV > 0) {
      const float4* b0 = reinterpret_cast<const float4*>(bank + r * d) + lane;
      const float4* b1 = reinterpret_cast<const float4*>(bank + (two ? r + 1 : r) * d) + lane;
      float4 x0[DV > 0 ? DV : 1], x1[DV > 0 ? DV : 1];
#pragma unroll
      for (int i = 0; i < DV; ++i) x0[i] = __ldcs(b0 + 32 * i);  // streamed once: evict-first
#pragma unroll
      for (int i = 0; i < DV; ++i) x1[i] = __ldcs(b1 + 32 * i);
#pragma unroll
      for (int i = 0; i < DV; ++i) {
#pragma unroll
        for (int q = 0; q < QT; ++q) {
          const float4 qv = *reinterpret_cast<const float4*>(qs + q * d + 4 * (lane + 32 * i));
          if (NORM) {
            s0[q] = fmaf(qv.x, x0[i].x, s0[q]); s0[q] = fmaf(qv.y, x0[i].y, s0[q]);
            s0[q] = fmaf(qv.z, x0[i].z, s0[q]); s0[q] = fmaf(qv.w, x0[i].w, s0[q]);
            s1[q] = fmaf(qv.x, x1[i].x, s1[q]); s1[q] = fmaf(qv.y, x1[i].y, s1[q]);
            s1[q] = fmaf(qv.z, x1[i].z, s1[q]); s1[q] = fmaf(qv.w, x1[i].w, s1[q]);
          } else {
            float t;
            t = qv.x - x0[i].x; s0[q] = fmaf(t, t, s0[q]);
            t = qv.y - x0[i].y; s0[q] = fmaf(t, t, s0[q]);
            t = qv.z - x0[i].z; s0[q] = fmaf(t, t, s0[q]);
            t = qv.w - x0[i].w; s0[q] = fmaf(t, t, s0[q]);
            t = qv.x - x1[i].x; s1[q] = fmaf(t, t, s1[q]);
            t = qv.y - x1[i].y; s1[q] = fmaf(t, t, s1[q]);
            t = qv.z - x1[i].z; s1[q] = fmaf(t, t, s1[q]);
            t = qv.w - x1[i].w; s1[q] = fmaf(t, t, s1[q]);
          }
        }
      }
    } else {
      const float* b0 = bank + r * d;
      const float* b1 = bank + (two ? r + 1 : r) * d;
      for (int c = lane; c < d; c += 32) {
        const float y0 = b0[c], y1 = b1[c];
#pragma unroll
        for (int q = 0; q < QT; ++q) {
          const float qv = qs[q * d + c];
          if (NORM) {
            s0[q] = fmaf(qv, y0, s0[q]);
            s1[q] = fmaf(qv, y1, s1[q]);
          } else {
            float t = qv - y0;
            s0[q] = fmaf(t, t, s0[q]);
            t = qv - y1;
            s1[q] = fmaf(t, t, s1[q]);
          }
        }
      }
    }
    float m0 = kInf, m1 = kInf;
    float nb0 = 0.f, nb1 = 0.f;
    if (NORM) {
      nb0 = __ldg(&bank_norms[r]);
      nb1 = __ldg(&bank_norms[two ? r + 1 : r]);
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      const float t0 = warp_sum(s0[q]), t1 = warp_sum(s1[q]);
      if (lane == q) {
        m0 = NORM ? fmaf(-2.f, t0, nb0) : t0;
        m1 = NORM ? fmaf(-2.f, t1, nb1) : t1;
      }
    }
    if (lane < Q) {
      if (m0 < v[KC - 1]) list_insert<KC>(v, id, m0, static_cast<int32_t>(r));
      if (two && m1 < v[KC - 1]) list_insert<KC>(v, id, m1, static_cast<int32_t>(r + 1));
    }
  }
  // block-level merge: warps cover ascending row ranges, so appending warp after warp keeps ties in id order
  __syncthreads();
  Cand* merge = reinterpret_cast<Cand*>(qs);  // [8 warps][QT][KC]
  if (lane < QT) {
#pragma unroll
    for (int q = 0; q < KC; ++q) merge[(warp * QT + lane) * KC + q] = Cand{v[q], id[q]};
  }
  __syncthreads();
  if (threadIdx.x < Q) {
    const int qq = threadIdx.x;
#pragma unroll
    for (int q = 0; q < KC; ++q) {
      v[q] = kInf;
      id[q] = -1;
    }
    for (int w = 0; w < 8; ++w) {
      for (int q = 0; q < KC; ++q) {
        const Cand c = merge[(w * QT + qq) * KC + q];
        if (c.idx >= 0 && c.t < v[KC - 1]) list_insert<KC>(v, id, c.t, c.idx);
      }
    }
    Cand* out = lists + (static_cast<int64_t>(qq) * gridDim.x + blockIdx.x) * KC;
#pragma unroll
    for (int q = 0; q < KC; ++q) out[q] = Cand{v[q], id[q]};
  }
}

// ---------------------------------------------------------------- stage 2: exact re-rank
// One block per query (32 threads when the candidate lists are short, 256 when the streaming scan left thousands).
// Extract the KC best proxies over all lists in (t, idx) order, re-evaluate them exactly in float64, order by
// (d2, global id), emit the first k.
//
// Certificate.  Let t~ be the scan's proxy of a bank row and t its exact value, |t~ - t| <= E for every row (E from
// the operand split and the accumulation of the scan arithmetic, see cert_bound()).  Every row that is NOT among
// the KC re-evaluated candidates was rejected against a list threshold or sorts after the KC-th extracted proxy
// tau, so t~ >= tau and hence t >= tau - E.  If the exact k-th result satisfies t_k < tau - E, no rejected row can
// precede it: the k results are THE k nearest rows.  Otherwise uncertified[q] = 1 and the caller re-does the query
// with en_knn_exact_topk.  (Lists that ran dry before KC extractions mean every admissible row was re-evaluated.)
struct CertParams {
  int form;               // 0: proxy t = |b|^2 - 2 q.b, absolute bound;  1: proxy = d2 itself, relative bound
  double c;               // form 0: |dot~ - dot| <= c |q| |b|;  form 1: |d2~ - d2| <= c d2
  const unsigned* bmax2;  // form 0: bit pattern of max_b |b|^2 over the shard (device)
  int32_t* uncertified;   // (Q,) flags, may be null
};

template <int KC>
__global__ void knn_rerank_kernel(const float* __restrict__ queries, int64_t Q, int d,
                                  const float* __restrict__ bank, int64_t id_offset, const Cand* __restrict__ lists,
                                  int n_lists, int k, double* __restrict__ d2_out, int64_t* __restrict__ ids_out,
                                  const CertParams cert) {
  __shared__ float s_t[8];
  __shared__ int32_t s_i[8];
  const int64_t q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const Cand* mine = lists + q * n_lists * KC;
  const int total = n_lists * KC;
  float last_t = -kInf;
  int32_t last_i = -1;
  double my_d2 = 1e300;   // warp 0: lane r holds the r-th extracted candidate
  int32_t my_idx = -1;
  int extracted = 0;
  for (int r = 0; r < KC; ++r) {
    float bt = kInf;
    int32_t bi = 0x7fffffff;
    for (int c = threadIdx.x; c < total; c += blockDim.x) {
      const Cand x = mine[c];
      if (x.idx < 0) continue;
      const bool after = x.t > last_t || (x.t == last_t && x.idx > last_i);
      if (after && (x.t < bt || (x.t == bt && x.idx < bi))) {
        bt = x.t;
        bi = x.idx;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ot = __shfl_xor_sync(0xffffffffu, bt, o);
      const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ot < bt || (ot == bt && oi < bi)) {
        bt = ot;
        bi = oi;
      }
    }
    if (nwarps > 1) {
      __syncthreads();
      if (lane == 0) {
        s_t[warp] = bt;
        s_i[warp] = bi;
      }
      __syncthreads();
      bt = s_t[0];
      bi = s_i[0];
      for (int w = 1; w < nwarps; ++w) {
        if (s_t[w] < bt || (s_t[w] == bt && s_i[w] < bi)) {
          bt = s_t[w];
          bi = s_i[w];
        }
      }
    }
    if (bi == 0x7fffffff) break;  // lists exhausted (block-uniform)
    last_t = bt;
    last_i = bi;
    ++extracted;
    if (warp == 0) {
      const float* a = queries + q * d;
      const float* b = bank + static_cast<int64_t>(bi) * d;
      double acc = 0.0;
      for (int c = lane; c < d; c += 32) {
        const double t = static_cast<double>(a[c]) - static_cast<double>(b[c]);
        acc = fma(t, t, acc);
      }
      acc = warp_sum(acc);
      if (lane == r) {
        my_d2 = acc;
        my_idx = bi;
      }
    }
  }
  if (warp != 0) return;
  // rank by (d2, idx) among the KC (<= 32) extracted; lanes >= KC hold sentinels
  int rank = 0;
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    const double od = __shfl_sync(0xffffffffu, my_d2, o);
    const int32_t oi = __shfl_sync(0xffffffffu, my_idx, o);
    if (oi >= 0 && (od < my_d2 || (od == my_d2 && oi < my_idx))) ++rank;
  }
  if (my_idx >= 0 && rank < k) {
    d2_out[q * k + rank] = my_d2;
    ids_out[q * k + rank] = id_offset + my_idx;
  }
  // fewer than k candidates: pad
  const int found = __popc(__ballot_sync(0xffffffffu, my_idx >= 0));
  if (lane >= found && lane < k) {
    d2_out[q * k + lane] = INFINITY;
    ids_out[q * k + lane] = -1;
  }
  if (cert.uncertified != nullptr) {
    bool certified = extracted < KC;  // lists ran dry: every admissible row was re-evaluated exactly
    if (!certified) {
      const unsigned kth = __ballot_sync(0xffffffffu, my_idx >= 0 && rank == k - 1);
      const double dk = __shfl_sync(0xffffffffu, my_d2, kth ? __ffs(kth) - 1 : 0);
      const double tau = static_cast<double>(last_t);  // the KC-th extracted proxy
      if (kth == 0) {
        certified = false;  // cannot happen (extracted == KC > k), be safe
      } else if (cert.form == 1) {
        certified = dk < tau * (1.0 - cert.c);
      } else {
        const float* a = queries + q * d;
        double qn2 = 0.0;
        for (int c = lane; c < d; c += 32) qn2 = fma(static_cast<double>(a[c]), static_cast<double>(a[c]), qn2);
        qn2 = warp_sum(qn2);
        const double bm2 = static_cast<double>(__uint_as_float(*cert.bmax2));
        const double qb = sqrt(qn2 * bm2);
        // dot error doubled by the proxy, plus the fp32 roundings of |b|^2 and of the final fma
        const double E = 2.0 * cert.c * qb + 2.4e-7 * (bm2 + 2.0 * qb);
        certified = (dk - qn2) + E < tau;
      }
    }
    if (lane == 0) cert.uncertified[q] = certified ? 0 : 1;
  }
}

// max over the shard of the squared row norms (non-negative floats order like their bit patterns)
__global__ void max_norm_kernel(const float* __restrict__ norms, int64_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, norms[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// (cert_bound(): common.cuh)

// ---------------------------------------------------------------- merge of per-shard lists
// part_stride: elements between consecutive parts (Q * k for separate arrays, 2 * Q * k for the packed records of one
// all-gather); only_flagged (optional): leave the outputs of queries whose flag is 0 untouched.
__global__ void knn_merge_kernel(const double* __restrict__ d2p, const int64_t* __restrict__ idp, int P, int64_t Q,
                                 int k, double* __restrict__ d2, int64_t* __restrict__ ids, int64_t part_stride,
                                 const int32_t* __restrict__ only_flagged) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  if (only_flagged != nullptr && only_flagged[q] == 0) return;
  double last_d = -1.0;
  int64_t last_i = -1;
  for (int r = 0; r < k; ++r) {
    double bd = INFINITY;
    int64_t bi = -1;
    for (int p = 0; p < P; ++p) {
      const double* dd = d2p + static_cast<int64_t>(p) * part_stride + q * k;
      const int64_t* ii = idp + static_cast<int64_t>(p) * part_stride + q * k;
      for (int c = 0; c < k; ++c) {
        const int64_t i = ii[c];
        if (i < 0) continue;
        const double x = dd[c];
        const bool after = x > last_d || (x == last_d && i > last_i);
        if (after && (bi < 0 || x < bd || (x == bd && i < bi))) {
          bd = x;
          bi = i;
        }
      }
    }
    d2[q * k + r] = bi >= 0 ? bd : INFINITY;
    ids[q * k + r] = bi;
    if (bi >= 0) {
      last_d = bd;
      last_i = bi;
    } else {
      last_d = INFINITY;
    }
  }
}

__global__ void sqrt_kernel(const double* __restrict__ d2, int64_t n, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(sqrt(d2[i]));
}

// majority vote, ties -> smallest label id (sklearn: classes_[argmax(counts)], classes_ sorted)
__global__ void knn_vote_kernel(const int64_t* __restrict__ ids, int64_t Q, int k, const int32_t* __restrict__ labels,
                                int64_t n_total, int32_t* __restrict__ pred) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int best_cnt = 0;
  int32_t best_lab = 0x7fffffff;
  for (int a = 0; a < k; ++a) {
    const int64_t ia = ids[q * k + a];
    if (ia < 0 || ia >= n_total) continue;
    const int32_t la = labels[ia];
    int cnt = 0;
    for (int b = 0; b < k; ++b) {
      const int64_t ib = ids[q * k + b];
      if (ib >= 0 && ib < n_total && labels[ib] == la) ++cnt;
    }
    if (cnt > best_cnt || (cnt == best_cnt && la < best_lab)) {
      best_cnt = cnt;
      best_lab = la;
    }
  }
  pred[q] = best_cnt > 0 ? best_lab : -1;
}

__global__ void knn_accuracy_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ pred,
                                    const int32_t* __restrict__ qlabels, int64_t Q, int k_ids,
                                    const int32_t* __restrict__ labels, int64_t n_total,
                                    unsigned long long* __restrict__ counts) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int top1 = 0, top5 = 0;
  if (q < Q) {
    const int32_t want = qlabels[q];
    top1 = pred[q] == want;
    const int kk = k_ids < 5 ? k_ids : 5;
    for (int a = 0; a < kk; ++a) {
      const int64_t ia = ids[q * k_ids + a];
      if (ia >= 0 && ia < n_total && labels[ia] == want) top5 = 1;
    }
  }
  const unsigned m1 = __ballot_sync(0xffffffffu, top1), m5 = __ballot_sync(0xffffffffu, top5);
  if ((threadIdx.x & 31) == 0) {
    if (m1) atomicAdd(&counts[0], static_cast<unsigned long long>(__popc(m1)));
    if (m5) atomicAdd(&counts[1], static_cast<unsigned long long>(__popc(m5)));
  }
}

inline int kc_for(int k) {
  const int need = k + EN_KNN_SLACK;
  return need <= 8 ? 8 : (need <= 16 ? 16 : 32);
}

// column ranges per launch: enough (query tile, range) items to give every SM a few
inline int knn_splits(int64_t Q, int64_t n_bank, int sms) {
  const int tiles_m = static_cast<int>((Q + tc::BM - 1) / tc::BM);
  const int tiles_n = static_cast<int>((n_bank + tc::BN - 1) / tc::BN);
  int s = (2 * sms + tiles_m - 1) / tiles_m;
  if (s < 2) s = 2;
  const int max_s = (tiles_n + kChunkTilesPerSplit - 1) / kChunkTilesPerSplit;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}

// The bank is walked in chunks small enough to stay L2-resident while every query tile passes over them, one
// launch per chunk; the per-query lists carry over between launches (item_begin reloads them).  Without this the
// 148 CTAs drift apart along a 10M-row bank and each streams it from HBM on its own (ncu, round 1: 30.9 TB of DRAM
// reads for one 100k x 10M scan, i.e. HBM-bound at 4.9 TB/s instead of tensor-bound).

template <int KC>
int run_scan(const CUtensorMap& qh, const CUtensorMap& ql, const CUtensorMap& bh, const CUtensorMap& bl, int64_t Q,
             int64_t n_bank, int d, int bf16, int splits_per_launch, const float* bank_norms,
             const int32_t* bank_labels, const int32_t* query_labels, Cand* lists, int sms, cudaStream_t st) {
  const int tiles_total = static_cast<int>((n_bank + tc::BN - 1) / tc::BN);
  const int chunk_tiles = splits_per_launch * kChunkTilesPerSplit;
  prof_begin(st);
  for (int base = 0, launch = 0; base < tiles_total; base += chunk_tiles, ++launch) {
    const int tiles_here = tiles_total - base < chunk_tiles ? tiles_total - base : chunk_tiles;
    tc::Shape sh = tc::make_shape(Q, static_cast<int64_t>(tiles_here) * tc::BN, d, splits_per_launch, 3, bf16);
    // keep the split count (= list slots per query) fixed across launches, even for a short last chunk
    sh.n_splits = splits_per_launch;
    sh.tiles_per_split = (tiles_here + splits_per_launch - 1) / splits_per_launch;
    sh.num_items = sh.tiles_m * sh.n_splits;
    sh.nt_base = base;
    typename EpTopK<KC>::Params ep{bank_norms, bank_labels, query_labels, lists, n_bank,
                                   splits_per_launch * tc::EPI_H, launch > 0 ? 1 : 0};
    EN_CUDA(tc::launch<EpTopK<KC>>(qh, ql, bh, bl, sh, ep, sms, st));
    ++launch_counter();
  }
  prof_end(st);
  return EN_OK;
}

template <int KC>
int run_rerank(const float* queries, int64_t Q, int d, const float* bank, int64_t id_offset, const Cand* lists,
               int n_lists, int k, double* d2, int64_t* ids, const CertParams& cert, cudaStream_t st) {
  const int threads = n_lists * KC > 1024 ? 256 : 32;
  knn_rerank_kernel<KC><<<static_cast<unsigned>(Q), threads, 0, st>>>(queries, Q, d, bank, id_offset, lists, n_lists,
                                                                       k, d2, ids, cert);
  EN_LAUNCHED("knn_rerank_kernel");
  return EN_OK;
}

int dispatch_rerank(int KC, const float* queries, int64_t Q, int d, const float* bank, int64_t id_offset,
                    const Cand* lists, int n_lists, int k, double* d2, int64_t* ids, const CertParams& cert,
                    cudaStream_t st) {
  if (KC == 8) return run_rerank<8>(queries, Q, d, bank, id_offset, lists, n_lists, k, d2, ids, cert, st);
  if (KC == 16) return run_rerank<16>(queries, Q, d, bank, id_offset, lists, n_lists, k, d2, ids, cert, st);
  return run_rerank<32>(queries, Q, d, bank, id_offset, lists, n_lists, k, d2, ids, cert, st);
}

// bmax2 <- max squared row norm of the shard (for the certificate); one pass over 4 bytes per row
int launch_max_norm(const float* norms, int64_t n, unsigned* bmax2, cudaStream_t st) {
  EN_CUDA(cudaMemsetAsync(bmax2, 0, sizeof(unsigned), st));
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 1184) blocks = 1184;
  max_norm_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(norms, n, bmax2);
  EN_LAUNCHED("max_norm_kernel");
  return EN_OK;
}

// ---------------------------------------------------------------- stage 3: float64 brute force (rare path)
// For the few queries the certificate could not cover (near-ties at the candidate cut-off, duplicated bank rows).
// Exact by construction: every admissible row's float64 sum (q-b)^2 (the re-rank's arithmetic, same loop order) is
// compared by (d2, id).  grid = (bank blocks, query groups of EX_QT); each warp streams a contiguous row range and
// keeps one sorted k-list per query in shared memory; per-block lists go to (P, Q, k) partials for knn_merge_kernel.
constexpr int EX_QT = 4;
constexpr int EX_KMAX = 32;
constexpr int EX_WARPS = 8;

__device__ __forceinline__ bool pair_less(double da, int32_t ia, double db, int32_t ib) {
  return da < db || (da == db && ia < ib);
}

__global__ void __launch_bounds__(EX_WARPS * 32)
knn_exact_kernel(const float* __restrict__ queries, int Q, int d, const float* __restrict__ bank, int64_t n_bank,
                 int64_t id_offset, int k, const int32_t* __restrict__ query_labels,
                 const int32_t* __restrict__ bank_labels, int64_t rows_per_warp, double* __restrict__ pd2,
                 int64_t* __restrict__ pid, const int32_t* __restrict__ only_flagged) {
  extern __shared__ __align__(16) uint8_t ex_smem[];
  if (only_flagged != nullptr) {  // device-side decision: nothing to redo for this group of queries
    bool any = false;
    for (int q = blockIdx.y * EX_QT; q < min(blockIdx.y * EX_QT + EX_QT, Q); ++q) any = any || only_flagged[q] != 0;
    if (!any) return;  // block-uniform
  }
  float* qs = reinterpret_cast<float*>(ex_smem);                                       // [EX_QT][d]
  double* ld = reinterpret_cast<double*>(ex_smem + ((static_cast<size_t>(EX_QT) * d * 4 + 15) / 16) * 16);
  int32_t* li = reinterpret_cast<int32_t*>(ld + EX_WARPS * EX_QT * EX_KMAX);           // same shape as ld
  const int q0 = blockIdx.y * EX_QT;
  const int nq = min(EX_QT, Q - q0);
  for (int i = threadIdx.x; i < EX_QT * d; i += blockDim.x) qs[i] = i < nq * d ? queries[static_cast<int64_t>(q0) * d + i] : 0.f;
  for (int i = threadIdx.x; i < EX_WARPS * EX_QT * EX_KMAX; i += blockDim.x) {
    ld[i] = INFINITY;
    li[i] = -1;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gwarp = static_cast<int64_t>(blockIdx.x) * EX_WARPS + warp;
  const int64_t r0 = gwarp * rows_per_warp;
  const int64_t r1 = min(r0 + rows_per_warp, n_bank);
  int32_t qlab = 0;
  const bool exclude = query_labels != nullptr && bank_labels != nullptr;
  if (exclude && lane < nq) qlab = query_labels[q0 + lane];
  double* myd = ld + (warp * EX_QT + lane) * EX_KMAX;  // lane q < nq owns query q's list of this warp
  int32_t* myi = li + (warp * EX_QT + lane) * EX_KMAX;
  for (int64_t r = r0; r < r1; ++r) {
    const float* b = bank + r * d;
    double acc[EX_QT];
#pragma unroll
    for (int q = 0; q < EX_QT; ++q) acc[q] = 0.0;
    for (int c = lane; c < d; c += 32) {
      const double y = static_cast<double>(b[c]);
#pragma unroll
      for (int q = 0; q < EX_QT; ++q) {
        const double t = static_cast<double>(qs[q * d + c]) - y;
        acc[q] = fma(t, t, acc[q]);
      }
    }
    double mine = 0.0;
#pragma unroll
    for (int q = 0; q < EX_QT; ++q) {
      const double v = warp_sum(acc[q]);
      if (lane == q) mine = v;
    }
    if (lane < nq) {
      const int32_t idx = static_cast<int32_t>(r);
      if (!(exclude && __ldg(&bank_labels[r]) == qlab) && pair_less(mine, idx, myd[k - 1], myi[k - 1] < 0 ? 0x7fffffff : myi[k - 1])) {
        int pos = k - 1;
        while (pos > 0 && pair_less(mine, idx, myd[pos - 1], myi[pos - 1] < 0 ? 0x7fffffff : myi[pos - 1])) {
          myd[pos] = myd[pos - 1];
          myi[pos] = myi[pos - 1];
          --pos;
        }
        myd[pos] = mine;
        myi[pos] = idx;
      }
    }
  }
  __syncthreads();
  // block merge: thread q selects the k best of the 8 warps' lists in (d2, id) order
  if (threadIdx.x < nq) {
    const int q = threadIdx.x;
    double last_d = -1.0;
    int32_t last_i = -1;
    double* od = pd2 + (static_cast<int64_t>(blockIdx.x) * Q + q0 + q) * k;
    int64_t* oi = pid + (static_cast<int64_t>(blockIdx.x) * Q + q0 + q) * k;
    for (int r = 0; r < k; ++r) {
      double bd = INFINITY;
      int32_t bi = -1;
      for (int w = 0; w < EX_WARPS; ++w) {
        const double* wd = ld + (w * EX_QT + q) * EX_KMAX;
        const int32_t* wi = li + (w * EX_QT + q) * EX_KMAX;
        for (int c = 0; c < k; ++c) {
          const int32_t i = wi[c];
          if (i < 0) break;
          const double x = wd[c];
          const bool after = x > last_d || (x == last_d && i > last_i);
          if (after && (bi < 0 || pair_less(x, i, bd, bi))) {
            bd = x;
            bi = i;
          }
        }
      }
      od[r] = bi >= 0 ? bd : INFINITY;
      oi[r] = bi >= 0 ? id_offset + bi : -1;
      if (bi < 0) {
        for (int rr = r + 1; rr < k; ++rr) {
          od[rr] = INFINITY;
          oi[rr] = -1;
        }
        break;
      }
      last_d = bd;
      last_i = bi;
    }
  }
}

inline void exact_geometry(int64_t n_bank, int sms, int64_t* rows_per_warp, int* blocks) {
  int64_t warps = static_cast<int64_t>(sms) * 2 * EX_WARPS;
  if (warps > n_bank) warps = n_bank;
  const int64_t rpw = (n_bank + warps - 1) / warps;
  *rows_per_warp = rpw;
  const int64_t used = (n_bank + rpw - 1) / rpw;
  *blocks = static_cast<int>((used + EX_WARPS - 1) / EX_WARPS);
}

constexpr int STREAM_WARPS = 8;

inline void stream_geometry(int64_t n_bank, int sms, int64_t* rows_per_warp, int* blocks) {
  int64_t warps = static_cast<int64_t>(sms) * 4 * STREAM_WARPS;  // 4 resident CTAs per SM
  if (warps * 2 > n_bank) warps = n_bank > 1 ? n_bank / 2 : 1;
  int64_t rpw = (n_bank + warps - 1) / warps;
  rpw = (rpw + 1) / 2 * 2;  // rows are streamed in pairs
  *rows_per_warp = rpw;
  const int64_t used = (n_bank + rpw - 1) / rpw;
  *blocks = static_cast<int>((used + STREAM_WARPS - 1) / STREAM_WARPS);
}

template <int KC, int DV, int QT, bool NORM>
int launch_stream(const float* queries, int64_t Q, int d, const float* bank, const float* norms, int64_t n_bank,
                  int64_t rpw, int blocks, Cand* lists, cudaStream_t st) {
  size_t smem = static_cast<size_t>(QT) * d * 4;
  const size_t merge = static_cast<size_t>(STREAM_WARPS) * QT * KC * sizeof(Cand);
  if (merge > smem) smem = merge;
  if (smem > 48 * 1024)
    EN_CUDA(cudaFuncSetAttribute(knn_stream_kernel<KC, DV, QT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  prof_begin(st);
  knn_stream_kernel<KC, DV, QT, NORM><<<blocks, STREAM_WARPS * 32, smem, st>>>(queries, static_cast<int>(Q), d, bank,
                                                                                norms, n_bank, rpw, lists);
  prof_end(st);
  EN_LAUNCHED("knn_stream_kernel");
  return EN_OK;
}

template <int KC, int DV, bool NORM>
int launch_stream_q(const float* queries, int64_t Q, int d, const float* bank, const float* norms, int64_t n_bank,
                    int64_t rpw, int blocks, Cand* lists, cudaStream_t st) {
  if (Q == 1) return launch_stream<KC, DV, 1, NORM>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
  if (Q == 2) return launch_stream<KC, DV, 2, NORM>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
  if (Q <= 4) return launch_stream<KC, DV, 4, NORM>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
  return launch_stream<KC, DV, 8, NORM>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
}

template <int KC, bool NORM>
int launch_stream_n(const float* queries, int64_t Q, int d, const float* bank, const float* norms, int64_t n_bank,
                    int64_t rpw, int blocks, Cand* lists, cudaStream_t st) {
  const bool aligned = (reinterpret_cast<uintptr_t>(bank) & 15) == 0;
  if (aligned && d == 128) return launch_stream_q<KC, 1, NORM>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
  if (aligned && d == 256) return launch_stream_q<KC, 2, NORM>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
  if (aligned && d == 512) return launch_stream_q<KC, 4, NORM>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
  return launch_stream_q<KC, 0, NORM>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
}

template <int KC>
int launch_stream_d(const float* queries, int64_t Q, int d, const float* bank, const float* norms, int64_t n_bank,
                    int64_t rpw, int blocks, Cand* lists, cudaStream_t st) {
  if (norms) return launch_stream_n<KC, true>(queries, Q, d, bank, norms, n_bank, rpw, blocks, lists, st);
  return launch_stream_n<KC, false>(queries, Q, d, bank, nullptr, n_bank, rpw, blocks, lists, st);
}

// ---------------------------------------------------------------- stage 1c: small query sets on the tensor cores
// 5 <= Q <= 64 queries per call.  The engine above makes the QUERIES the 128-row operand: with a handful of queries
// it still issues full 128 x 128 MMAs per bank tile and re-fetches the query tile for every bank tile, and ran at
// 0.61 of the HBM roofline at one GPU (0.49 at eight; driver SCALE_r01 tails).  Here the roles are swapped:
//   A (M = 128) = a BANK tile, streamed once from HBM through a TMA ring (both BF16 planes);
//   B (N = QP)  = the query set padded to 16 / 32 / 64 rows, loaded into shared memory ONCE and kept there;
//   D (TMEM)    = 128 bank rows x QP queries, main (hi.hi) and cross (lo.hi + hi.lo) accumulators, double buffered.
// Tensor time per tile shrinks with QP / 128 and the kernel is bounded by the bank stream: algorithmic bytes per
// call = n_bank * d * 4 (the two BF16 planes are exactly as large as the fp32 rows).  Each epilogue thread owns one
// bank row of the tile and QP proxies; a warp keeps one sorted list per query in shared memory (inserts are rare
// once the lists have warmed up: a tile is skipped with one vote), the CTA merges its four warps' lists at the end
// and the exact re-rank + certificate of stage 2 take over.  Same arithmetic as the engine (split-BF16, separate
// cross accumulator), hence the same certificate bound.
namespace sq {
constexpr int STAGE_BYTES = 2 * tc::TILE_BYTES;  // A_hi | A_lo of one k-block (64 BF16 columns x 128 bank rows)
constexpr int EPI_WARPS = 4;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int KC = 8;                            // k <= 5 with the usual slack of 3
constexpr int MAX_STAGES = 6;

struct Bars {
  uint64_t q_full;
  uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

inline size_t smem_bytes(int qp, int kblocks, int stages) {
  return static_cast<size_t>(2) * kblocks * qp * 128 + static_cast<size_t>(stages) * STAGE_BYTES + 256 +
         static_cast<size_t>(EPI_WARPS) * qp * KC * 8;
}
// ring depth that fits beside the resident queries (0 = this shape does not fit: use the engine)
inline int pick_stages(int qp, int kblocks) {
  for (int s = MAX_STAGES; s >= 2; --s)
    if (smem_bytes(qp, kblocks, s) <= 227 * 1024) return s;
  return 0;
}

template <int QP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
knn_smallq_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                  const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                  const float* __restrict__ bank_norms, int64_t n_bank, int Q, int kblocks, int stages,
                  Cand* __restrict__ lists /*[Q][gridDim.x][KC]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  const int q_plane = kblocks * QP * 128;          // bytes of one query plane
  uint8_t* q_hi = smem;
  uint8_t* q_lo = smem + q_plane;
  uint8_t* ring = smem + 2 * q_plane;
  Bars* bars = reinterpret_cast<Bars*>(ring + stages * STAGE_BYTES);
  float* l_t = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [warp][QP][KC]
  int32_t* l_i = reinterpret_cast<int32_t*>(l_t + EPI_WARPS * QP * KC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_total = static_cast<int>((n_bank + tc::BM - 1) / tc::BM);
  const int t0 = static_cast<int>(static_cast<int64_t>(tiles_total) * blockIdx.x / gridDim.x);
  const int t1 = static_cast<int>(static_cast<int64_t>(tiles_total) * (blockIdx.x + 1) / gridDim.x);
  constexpr int ACC = 2 * QP;                      // main | cross
  constexpr uint32_t TMEM_COLS = 2 * ACC <= 32 ? 32 : (2 * ACC <= 64 ? 64 : (2 * ACC <= 128 ? 128 : 256));

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_q_hi);
    ptx::prefetch_tmap(&tm_q_lo);
    ptx::prefetch_tmap(&tm_b_hi);
    ptx::prefetch_tmap(&tm_b_lo);
    ptx::mbar_init(&bars->q_full, 1);
    for (int s = 0; s < stages; ++s) { ptx::mbar_init(&bars->full[s], 1); ptx::mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&bars->tmem_full[a], 1); ptx::mbar_init(&bars->tmem_empty[a], EPI_WARPS); }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  for (int i = threadIdx.x; i < EPI_WARPS * QP * KC; i += blockDim.x) {
    l_t[i] = kInf;
    l_i[i] = -1;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: queries once, then the bank stream
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->q_full, 2 * q_plane);
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::tma_load_2d(&tm_q_hi, &bars->q_full, q_hi + kb * QP * 128, kb * tc::BK16, 0);
        ptx::tma_load_2d(&tm_q_lo, &bars->q_full, q_lo + kb * QP * 128, kb * tc::BK16, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t0; t < t1; ++t) {
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* st = ring + stage * STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&bars->full[stage], STAGE_BYTES);
          ptx::tma_load_2d(&tm_b_hi, &bars->full[stage], st, kb * tc::BK16, t * tc::BM);
          ptx::tma_load_2d(&tm_b_lo, &bars->full[stage], st + tc::TILE_BYTES, kb * tc::BK16, t * tc::BM);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(tc::BM, QP);
      ptx::mbar_wait(&bars->q_full, 0);
      ptx::tc_fence_after();
      int stage = 0;
      uint32_t phase = 0, acc_it = 0;
      for (int t = t0; t < t1; ++t, ++acc_it) {
        const uint32_t acc = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
        ptx::mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tm_d = tmem_base + acc * ACC, tm_x = tm_d + QP;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t st = ptx::smem_u32(ring + stage * STAGE_BYTES);
          const uint64_t a_hi = ptx::make_kmajor_sw128_desc(st), a_lo = ptx::make_kmajor_sw128_desc(st + tc::TILE_BYTES);
          const uint64_t b_hi = ptx::make_kmajor_sw128_desc(ptx::smem_u32(q_hi + kb * QP * 128));
          const uint64_t b_lo = ptx::make_kmajor_sw128_desc(ptx::smem_u32(q_lo + kb * QP * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t koff = static_cast<uint64_t>(k * 2);  // 16 BF16 = 32 bytes along the swizzle row
            ptx::mma_bf16_ss(tm_x, a_lo + koff, b_hi + koff, idesc, (kb | k) != 0);
            ptx::mma_bf16_ss(tm_x, a_hi + koff, b_lo + koff, idesc, 1);
            ptx::mma_bf16_ss(tm_d, a_hi + koff, b_hi + koff, idesc, (kb | k) != 0);
          }
          ptx::mma_commit(&bars->empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&bars->tmem_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: one bank row per thread, QP proxies
    const int ew = warp - 2;          // list owner index 0..3
    const int quarter = warp & 3;     // TMEM lane quarter this warp may read
    float* my_t = l_t + ew * QP * KC;
    int32_t* my_i = l_i + ew * QP * KC;
    uint32_t acc_it = 0;
    for (int t = t0; t < t1; ++t, ++acc_it) {
      const uint32_t acc = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
      const int64_t row = static_cast<int64_t>(t) * tc::BM + quarter * 32 + lane;
      const float nb = row < n_bank ? __ldg(&bank_norms[row]) : kInf;   // rows past the end never qualify
      ptx::mbar_wait(&bars->tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * ACC;
      constexpr int CH = QP < 32 ? QP : 32;       // query columns per TMEM load
#pragma unroll
      for (int c0 = 0; c0 < QP; c0 += CH) {
        float dot[CH], cross[CH];
        if (CH == 16) {
          ptx::tmem_ld_32x16(taddr + c0, reinterpret_cast<float (&)[16]>(dot));
          ptx::tmem_ld_32x16(taddr + QP + c0, reinterpret_cast<float (&)[16]>(cross));
        } else {
          ptx::tmem_ld_32x32(taddr + c0, reinterpret_cast<float (&)[32]>(dot));
          ptx::tmem_ld_32x32(taddr + QP + c0, reinterpret_cast<float (&)[32]>(cross));
        }
        ptx::tmem_ld_wait();
        unsigned mask = 0;                        // queries (of this chunk) for which this row beats the list's worst
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          dot[j] = fmaf(-2.f, dot[j] + cross[j], nb);
          mask |= (dot[j] < my_t[(c0 + j) * KC + KC - 1] ? 1u : 0u) << j;
        }
        unsigned any = __reduce_or_sync(0xffffffffu, mask);
        // rare once the lists are warm: per query with a hit, the hitting lanes in ascending row order
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (any & (1u << j)) {                  // warp-uniform
            unsigned m = __ballot_sync(0xffffffffu, (mask >> j) & 1u);
            while (m) {
              const int src = __ffs(m) - 1;
              m &= m - 1;
              const float tv = __shfl_sync(0xffffffffu, dot[j], src);
              if (lane == 0) {
                float* lt = my_t + (c0 + j) * KC;
                int32_t* li = my_i + (c0 + j) * KC;
                if (tv < lt[KC - 1]) {            // an earlier insert of this tile may have raised the bar
                  int pos = KC - 1;
                  while (pos > 0 && tv < lt[pos - 1]) {  // strict: equal proxies keep the earlier (lower) row first
                    lt[pos] = lt[pos - 1];
                    li[pos] = li[pos - 1];
                    --pos;
                  }
                  lt[pos] = tv;
                  li[pos] = static_cast<int32_t>(static_cast<int64_t>(t) * tc::BM + quarter * 32 + src);
                }
              }
            }
            __syncwarp();                          // the list (shared memory) is read again by all lanes
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
    }
    // ---- merge the four warps' lists per query and publish (t, idx) in ascending order
    ptx::named_bar_sync(1, EPI_WARPS * 32);
    const int tid = threadIdx.x - 64;
    for (int q = tid; q < Q; q += EPI_WARPS * 32) {
      int head[EPI_WARPS] = {0, 0, 0, 0};
      Cand* out = lists + (static_cast<int64_t>(q) * gridDim.x + blockIdx.x) * KC;
      for (int e = 0; e < KC; ++e) {
        float bt = kInf;
        int32_t bi = 0x7fffffff;
        int bw = -1;
#pragma unroll
        for (int w = 0; w < EPI_WARPS; ++w) {
          if (head[w] < KC) {
            const float tv = l_t[(w * QP + q) * KC + head[w]];
            const int32_t iv = l_i[(w * QP + q) * KC + head[w]];
            if (iv >= 0 && (tv < bt || (tv == bt && iv < bi))) { bt = tv; bi = iv; bw = w; }
          }
        }
        if (bw >= 0) {
          ++head[bw];
          out[e] = Cand{bt, bi};
        } else {
          out[e] = Cand{kInf, -1};
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

inline int pad_q(int64_t Q) { return Q <= 16 ? 16 : (Q <= 32 ? 32 : 64); }

template <int QP>
int launch(const CUtensorMap& qh, const CUtensorMap& ql, const CUtensorMap& bh, const CUtensorMap& bl,
           const float* norms, int64_t n_bank, int Q, int kblocks, int stages, int grid, Cand* lists, cudaStream_t st) {
  const size_t smem = smem_bytes(QP, kblocks, stages);
  EN_CUDA(cudaFuncSetAttribute(knn_smallq_kernel<QP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  prof_begin(st);
  knn_smallq_kernel<QP><<<grid, NUM_THREADS, smem, st>>>(qh, ql, bh, bl, norms, n_bank, Q, kblocks, stages, lists);
  prof_end(st);
  EN_LAUNCHED("knn_smallq_kernel");
  return EN_OK;
}
}  // namespace sq

}  // namespace
}  // namespace en

using namespace en;

extern "C" {

int en_bank_dpad(int d, int precision) { return d > 0 ? tc::dpad_for(d, precision == EN_PREC_BF16X3) : 0; }

size_t en_bank_plane_bytes(int64_t n, int d, int precision) {
  if (n <= 0 || d <= 0) return 0;
  return static_cast<size_t>(n) * en_bank_dpad(d, precision) * (precision == EN_PREC_BF16X3 ? 2 : 4);
}

int en_bank_prepare(const float* bank, int64_t n, int d, int precision, void* hi, void* lo, float* norms,
                    void* stream) {
  EN_REQUIRE(bank && hi && lo && norms && n >= 0 && d > 0, "en_bank_prepare: bad arguments");
  EN_REQUIRE(precision == EN_PREC_TF32X3 || precision == EN_PREC_BF16X3, "en_bank_prepare: unknown precision %d",
             precision);
  if (n == 0) return EN_OK;
  if (precision == EN_PREC_BF16X3)
    EN_CUDA(tc::launch_split_bf16(bank, n, d, d, en_bank_dpad(d, precision), hi, lo, norms, as_stream(stream)));
  else
    EN_CUDA(tc::launch_split(bank, n, d, d, en_bank_dpad(d, precision), static_cast<float*>(hi),
                             static_cast<float*>(lo), norms, as_stream(stream)));
  ++launch_counter();
  return EN_OK;
}

size_t en_ws_bytes_knn(int64_t Q, int64_t n_bank, int d, int k) {
  if (Q <= 0 || n_bank <= 0 || d <= 0 || k <= 0 || k > EN_KNN_MAX_K) return 0;
  const size_t dpad = static_cast<size_t>(en_bank_dpad(d, EN_PREC_TF32X3));  // the larger of the two plane formats
  const int s = knn_splits(Q, n_bank, 160);  // sized for the largest SM count
  return 2 * align_up(static_cast<size_t>(Q) * dpad * 4) + align_up(static_cast<size_t>(Q) * 4) + align_up(4) +
         align_up(static_cast<size_t>(Q) * s * tc::EPI_H * kc_for(k) * sizeof(Cand));
}

int en_knn_shard_topk(const float* queries, int64_t Q, int d, const float* bank, const void* bank_hi,
                      const void* bank_lo, const float* bank_norms, int64_t n_bank, int64_t id_offset, int k,
                      int precision, const int32_t* query_labels, const int32_t* bank_labels, double* d2,
                      int64_t* ids, int32_t* uncertified, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(queries && bank && bank_hi && bank_lo && bank_norms && d2 && ids && Q > 0 && n_bank > 0 && d > 0,
             "en_knn_shard_topk: bad arguments");
  EN_REQUIRE(k > 0 && k <= EN_KNN_MAX_K, "en_knn_shard_topk: k must be in [1, %d] (got %d)", EN_KNN_MAX_K, k);
  EN_REQUIRE((query_labels == nullptr) == (bank_labels == nullptr) || query_labels == nullptr,
             "en_knn_shard_topk: query_labels requires bank_labels");
  EN_REQUIRE(n_bank < (int64_t(1) << 31), "en_knn_shard_topk: shard too large (%lld rows)", (long long)n_bank);
  EN_REQUIRE(precision == EN_PREC_TF32X3 || precision == EN_PREC_BF16X3, "en_knn_shard_topk: unknown precision %d",
             precision);
  if (int rc = check_sm100()) return rc;
  if (!ws || ws_bytes < en_ws_bytes_knn(Q, n_bank, d, k))
    return fail(EN_ERR_WORKSPACE, "en_knn_shard_topk: workspace too small (%zu < %zu)", ws_bytes,
                en_ws_bytes_knn(Q, n_bank, d, k));
  cudaStream_t st = as_stream(stream);
  const int sms = device_sm_count();
  const int bf16 = precision == EN_PREC_BF16X3;
  const int dpad = en_bank_dpad(d, precision);
  const int KC = kc_for(k);
  Workspace w(ws, ws_bytes);
  float* qhi = w.take<float>(static_cast<size_t>(Q) * en_bank_dpad(d, EN_PREC_TF32X3));
  float* qlo = w.take<float>(static_cast<size_t>(Q) * en_bank_dpad(d, EN_PREC_TF32X3));
  float* qn = w.take<float>(Q);
  unsigned* bmax2 = w.take<unsigned>(1);
  const int splits = knn_splits(Q, n_bank, sms);
  Cand* lists = w.take<Cand>(static_cast<size_t>(Q) * splits * tc::EPI_H * KC);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_knn_shard_topk: workspace too small or misaligned");
  CUtensorMap tqh, tql, tbh, tbl;
  if (bf16) {
    EN_CUDA(tc::launch_split_bf16(queries, Q, d, d, dpad, qhi, qlo, qn, st));
    if (tc::make_plane_tmap_bf16(&tqh, qhi, Q, dpad) || tc::make_plane_tmap_bf16(&tql, qlo, Q, dpad) ||
        tc::make_plane_tmap_bf16(&tbh, bank_hi, n_bank, dpad) || tc::make_plane_tmap_bf16(&tbl, bank_lo, n_bank, dpad))
      return fail(EN_ERR_DRIVER, "en_knn_shard_topk: cuTensorMapEncodeTiled failed");
  } else {
    EN_CUDA(tc::launch_split(queries, Q, d, d, dpad, qhi, qlo, qn, st));
    if (tc::make_plane_tmap(&tqh, qhi, Q, dpad) || tc::make_plane_tmap(&tql, qlo, Q, dpad) ||
        tc::make_plane_tmap(&tbh, static_cast<const float*>(bank_hi), n_bank, dpad) ||
        tc::make_plane_tmap(&tbl, static_cast<const float*>(bank_lo), n_bank, dpad))
      return fail(EN_ERR_DRIVER, "en_knn_shard_topk: cuTensorMapEncodeTiled failed");
  }
  ++launch_counter();
  const int32_t* ql = bank_labels ? query_labels : nullptr;
  int rc;
  if (KC == 8) rc = run_scan<8>(tqh, tql, tbh, tbl, Q, n_bank, d, bf16, splits, bank_norms, bank_labels, ql, lists, sms, st);
  else if (KC == 16) rc = run_scan<16>(tqh, tql, tbh, tbl, Q, n_bank, d, bf16, splits, bank_norms, bank_labels, ql, lists, sms, st);
  else rc = run_scan<32>(tqh, tql, tbh, tbl, Q, n_bank, d, bf16, splits, bank_norms, bank_labels, ql, lists, sms, st);
  if (rc) return rc;
  CertParams cert{0, cert_bound(precision, dpad), bmax2, uncertified};
  if (uncertified != nullptr)
    if (int rc2 = launch_max_norm(bank_norms, n_bank, bmax2, st)) return rc2;
  return dispatch_rerank(KC, queries, Q, d, bank, id_offset, lists, splits * tc::EPI_H, k, d2, ids, cert, st);
}

size_t en_ws_bytes_knn_stream(int64_t Q, int64_t n_bank, int d, int k) {
  if (Q <= 0 || Q > EN_KNN_STREAM_MAX_Q || n_bank <= 0 || d <= 0 || k <= 0 || k > EN_KNN_MAX_K) return 0;
  int64_t rpw;
  int blocks;
  stream_geometry(n_bank, 160, &rpw, &blocks);  // sized for the largest SM count
  return align_up(static_cast<size_t>(Q) * (blocks + 8) * kc_for(k) * sizeof(Cand)) + align_up(4);
}

int en_knn_stream_topk(const float* queries, int64_t Q, int d, const float* bank, const float* bank_norms,
                       int64_t n_bank, int64_t id_offset, int k, double* d2, int64_t* ids, int32_t* uncertified,
                       void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(queries && bank && d2 && ids && Q > 0 && n_bank > 0 && d > 0, "en_knn_stream_topk: bad arguments");
  EN_REQUIRE(Q <= EN_KNN_STREAM_MAX_Q, "en_knn_stream_topk: at most %d queries per call (got %lld)",
             EN_KNN_STREAM_MAX_Q, (long long)Q);
  EN_REQUIRE(k > 0 && k <= EN_KNN_MAX_K, "en_knn_stream_topk: k must be in [1, %d]", EN_KNN_MAX_K);
  EN_REQUIRE(n_bank < (int64_t(1) << 31), "en_knn_stream_topk: shard too large");
  EN_REQUIRE(static_cast<size_t>(8) * d * 4 <= 160 * 1024, "en_knn_stream_topk: d too large for shared memory");
  if (!ws || ws_bytes < en_ws_bytes_knn_stream(Q, n_bank, d, k))
    return fail(EN_ERR_WORKSPACE, "en_knn_stream_topk: workspace too small");
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(EN_ERR_WORKSPACE, "workspace misaligned");
  cudaStream_t st = as_stream(stream);
  const int sms = device_sm_count();
  int64_t rpw;
  int blocks;
  stream_geometry(n_bank, sms, &rpw, &blocks);
  const int KC = kc_for(k);
  Cand* lists = static_cast<Cand*>(ws);
  unsigned* bmax2 = reinterpret_cast<unsigned*>(
      static_cast<uint8_t*>(ws) + align_up(static_cast<size_t>(Q) * (blocks + 8) * KC * sizeof(Cand)));
  int rc;
  if (KC == 8) rc = launch_stream_d<8>(queries, Q, d, bank, bank_norms, n_bank, rpw, blocks, lists, st);
  else if (KC == 16) rc = launch_stream_d<16>(queries, Q, d, bank, bank_norms, n_bank, rpw, blocks, lists, st);
  else rc = launch_stream_d<32>(queries, Q, d, bank, bank_norms, n_bank, rpw, blocks, lists, st);
  if (rc) return rc;
  // fp32 CUDA-core arithmetic: dot form (absolute bound) with norms, direct sum (q-b)^2 (relative bound) without
  CertParams cert{bank_norms ? 0 : 1, bank_norms ? cert_bound(-1, d) : 1.5 * (d / 32 + 12) / 16777216.0, bmax2,
                  uncertified};
  if (uncertified != nullptr && bank_norms != nullptr)
    if (int rc2 = launch_max_norm(bank_norms, n_bank, bmax2, st)) return rc2;
  return dispatch_rerank(KC, queries, Q, d, bank, id_offset, lists, blocks, k, d2, ids, cert, st);
}

size_t en_ws_bytes_knn_exact(int64_t Q, int64_t n_bank, int d, int k) {
  if (Q <= 0 || Q > EN_KNN_EXACT_MAX_Q || n_bank <= 0 || d <= 0 || k <= 0 || k > EN_KNN_MAX_K) return 0;
  int64_t rpw;
  int blocks;
  exact_geometry(n_bank, 160, &rpw, &blocks);  // sized for the largest SM count
  return 2 * align_up(static_cast<size_t>(blocks) * Q * k * 8);
}

static int knn_exact_impl(const float* queries, int64_t Q, int d, const float* bank, int64_t n_bank, int64_t id_offset,
                          int k, const int32_t* query_labels, const int32_t* bank_labels, const int32_t* only_flagged,
                          double* d2, int64_t* ids, void* ws, size_t ws_bytes, void* stream, const char* who) {
  (void)who;
  EN_REQUIRE(queries && bank && d2 && ids && Q > 0 && n_bank > 0 && d > 0, "en_knn_exact_topk: bad arguments");
  EN_REQUIRE(Q <= EN_KNN_EXACT_MAX_Q, "en_knn_exact_topk: at most %d queries per call (got %lld)",
             EN_KNN_EXACT_MAX_Q, (long long)Q);
  EN_REQUIRE(k > 0 && k <= EN_KNN_MAX_K, "en_knn_exact_topk: k must be in [1, %d]", EN_KNN_MAX_K);
  EN_REQUIRE(n_bank < (int64_t(1) << 31), "en_knn_exact_topk: shard too large");
  EN_REQUIRE((query_labels == nullptr) == (bank_labels == nullptr) || query_labels == nullptr,
             "en_knn_exact_topk: query_labels requires bank_labels");
  const size_t smem = (static_cast<size_t>(EX_QT) * d * 4 + 15) / 16 * 16 +
                      static_cast<size_t>(EX_WARPS) * EX_QT * EX_KMAX * 12;
  EN_REQUIRE(smem <= 200 * 1024, "en_knn_exact_topk: d too large for shared memory");
  if (!ws || ws_bytes < en_ws_bytes_knn_exact(Q, n_bank, d, k) || (reinterpret_cast<uintptr_t>(ws) & 255) != 0)
    return fail(EN_ERR_WORKSPACE, "en_knn_exact_topk: workspace too small or misaligned");
  cudaStream_t st = as_stream(stream);
  int64_t rpw;
  int blocks;
  exact_geometry(n_bank, device_sm_count(), &rpw, &blocks);
  double* pd2 = static_cast<double*>(ws);
  int64_t* pid = reinterpret_cast<int64_t*>(static_cast<uint8_t*>(ws) + align_up(static_cast<size_t>(blocks) * Q * k * 8));
  if (smem > 48 * 1024)
    EN_CUDA(cudaFuncSetAttribute(knn_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>((Q + EX_QT - 1) / EX_QT));
  const int32_t* ql = bank_labels ? query_labels : nullptr;
  knn_exact_kernel<<<grid, EX_WARPS * 32, smem, st>>>(queries, static_cast<int>(Q), d, bank, n_bank, id_offset, k, ql,
                                                      bank_labels, rpw, pd2, pid, only_flagged);
  EN_LAUNCHED("knn_exact_kernel");
  knn_merge_kernel<<<static_cast<unsigned>((Q + 127) / 128), 128, 0, st>>>(pd2, pid, blocks, Q, k, d2, ids, Q * k,
                                                                           only_flagged);
  EN_LAUNCHED("knn_merge_kernel");
  return EN_OK;
}

int en_knn_exact_topk(const float* queries, int64_t Q, int d, const float* bank, int64_t n_bank, int64_t id_offset,
                      int k, const int32_t* query_labels, const int32_t* bank_labels, double* d2, int64_t* ids,
                      void* ws, size_t ws_bytes, void* stream) {
  return knn_exact_impl(queries, Q, d, bank, n_bank, id_offset, k, query_labels, bank_labels, nullptr, d2, ids, ws,
                        ws_bytes, stream, "en_knn_exact_topk");
}

// The same brute force, decided on the device: only the queries whose flag (the `uncertified` output of a scan) is
// non-zero are recomputed and overwritten in d2 / ids; with no flag set the two launches return at once.  Lets a
// small-batch predict call finish without reading the certificate back to the host.
int en_knn_exact_redo(const float* queries, int64_t Q, int d, const float* bank, int64_t n_bank, int64_t id_offset,
                      int k, const int32_t* query_labels, const int32_t* bank_labels, const int32_t* flags, double* d2,
                      int64_t* ids, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(flags != nullptr, "en_knn_exact_redo: flags is null");
  return knn_exact_impl(queries, Q, d, bank, n_bank, id_offset, k, query_labels, bank_labels, flags, d2, ids, ws,
                        ws_bytes, stream, "en_knn_exact_redo");
}

size_t en_ws_bytes_knn_smallq(int64_t Q, int64_t n_bank, int d, int k) {
  if (Q <= 0 || Q > EN_KNN_SMALLQ_MAX_Q || n_bank <= 0 || d <= 0 || k <= 0 || k + EN_KNN_SLACK > sq::KC) return 0;
  const int dpad = en_bank_dpad(d, EN_PREC_BF16X3);
  if (sq::pick_stages(sq::pad_q(Q), dpad / tc::BK16) == 0) return 0;  // the resident query tile does not fit
  const size_t qp = static_cast<size_t>(sq::pad_q(Q));
  return 2 * align_up(qp * dpad * 2) + align_up(qp * 4) + align_up(4) +
         align_up(static_cast<size_t>(Q) * 160 * sq::KC * sizeof(Cand));  // sized for the largest SM count
}

// 5 <= Q <= 64 queries against the BF16 planes of the bank (en_bank_prepare with EN_PREC_BF16X3): bank-stationary
// tensor-core scan bounded by the bank stream, then the exact re-rank / certificate of en_knn_shard_topk.
int en_knn_smallq_topk(const float* queries, int64_t Q, int d, const float* bank, const void* bank_hi,
                       const void* bank_lo, const float* bank_norms, int64_t n_bank, int64_t id_offset, int k,
                       double* d2, int64_t* ids, int32_t* uncertified, void* ws, size_t ws_bytes, void* stream) {
  EN_REQUIRE(queries && bank && bank_hi && bank_lo && bank_norms && d2 && ids && Q > 0 && n_bank > 0 && d > 0,
             "en_knn_smallq_topk: bad arguments");
  EN_REQUIRE(Q <= EN_KNN_SMALLQ_MAX_Q, "en_knn_smallq_topk: at most %d queries per call (got %lld)",
             EN_KNN_SMALLQ_MAX_Q, (long long)Q);
  EN_REQUIRE(k > 0 && k + EN_KNN_SLACK <= sq::KC, "en_knn_smallq_topk: k must be in [1, %d]", sq::KC - EN_KNN_SLACK);
  EN_REQUIRE(n_bank < (int64_t(1) << 31), "en_knn_smallq_topk: shard too large");
  if (int rc = check_sm100()) return rc;
  const size_t need = en_ws_bytes_knn_smallq(Q, n_bank, d, k);
  EN_REQUIRE(need != 0, "en_knn_smallq_topk: d = %d does not fit the resident query tile (use en_knn_shard_topk)", d);
  if (!ws || ws_bytes < need) return fail(EN_ERR_WORKSPACE, "en_knn_smallq_topk: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int sms = device_sm_count();
  const int dpad = en_bank_dpad(d, EN_PREC_BF16X3);
  const int kblocks = dpad / tc::BK16;
  const int qp = sq::pad_q(Q);
  const int stages = sq::pick_stages(qp, kblocks);
  Workspace w(ws, ws_bytes);
  uint16_t* qhi = w.take<uint16_t>(static_cast<size_t>(qp) * dpad);
  uint16_t* qlo = w.take<uint16_t>(static_cast<size_t>(qp) * dpad);
  float* qn = w.take<float>(qp);
  unsigned* bmax2 = w.take<unsigned>(1);
  const int tiles_total = static_cast<int>((n_bank + tc::BM - 1) / tc::BM);
  const int grid = tiles_total < sms ? tiles_total : sms;
  Cand* lists = w.take<Cand>(static_cast<size_t>(Q) * grid * sq::KC);
  if (!w.ok()) return fail(EN_ERR_WORKSPACE, "en_knn_smallq_topk: workspace too small or misaligned");
  EN_CUDA(tc::launch_split_bf16(queries, Q, d, d, dpad, qhi, qlo, qn, st));
  ++launch_counter();
  CUtensorMap tqh, tql, tbh, tbl;
  if (tc::make_plane_tmap_bf16(&tqh, qhi, Q, dpad, qp) || tc::make_plane_tmap_bf16(&tql, qlo, Q, dpad, qp) ||
      tc::make_plane_tmap_bf16(&tbh, bank_hi, n_bank, dpad) || tc::make_plane_tmap_bf16(&tbl, bank_lo, n_bank, dpad))
    return fail(EN_ERR_DRIVER, "en_knn_smallq_topk: cuTensorMapEncodeTiled failed");
  int rc;
  if (qp == 16) rc = sq::launch<16>(tqh, tql, tbh, tbl, bank_norms, n_bank, static_cast<int>(Q), kblocks, stages, grid, lists, st);
  else if (qp == 32) rc = sq::launch<32>(tqh, tql, tbh, tbl, bank_norms, n_bank, static_cast<int>(Q), kblocks, stages, grid, lists, st);
  else rc = sq::launch<64>(tqh, tql, tbh, tbl, bank_norms, n_bank, static_cast<int>(Q), kblocks, stages, grid, lists, st);
  if (rc) return rc;
  CertParams cert{0, cert_bound(EN_PREC_BF16X3, dpad), bmax2, uncertified};
  if (uncertified != nullptr)
    if (int rc2 = launch_max_norm(bank_norms, n_bank, bmax2, st)) return rc2;
  return dispatch_rerank(sq::KC, queries, Q, d, bank, id_offset, lists, grid, k, d2, ids, cert, st);
}

int en_knn_merge(const double* d2_parts, const int64_t* id_parts, int n_parts, int64_t Q, int k, double* d2,
                 int64_t* ids, void* stream) {
  EN_REQUIRE(d2_parts && id_parts && d2 && ids && n_parts > 0 && Q > 0 && k > 0, "en_knn_merge: bad arguments");
  knn_merge_kernel<<<static_cast<unsigned>((Q + 127) / 128), 128, 0, as_stream(stream)>>>(
      d2_parts, id_parts, n_parts, Q, k, d2, ids, Q * k, nullptr);
  EN_LAUNCHED("knn_merge_kernel");
  return EN_OK;
}

// The same merge over the packed records of ONE all-gather: parts is (n_parts, 2, Q, k) 8-byte words, per part the
// (Q, k) float64 squared distances followed by the (Q, k) int64 ids.
int en_knn_merge_packed(const void* parts, int n_parts, int64_t Q, int k, double* d2, int64_t* ids, void* stream) {
  EN_REQUIRE(parts && d2 && ids && n_parts > 0 && Q > 0 && k > 0, "en_knn_merge_packed: bad arguments");
  const double* dp = static_cast<const double*>(parts);
  const int64_t* ip = static_cast<const int64_t*>(parts) + Q * k;
  knn_merge_kernel<<<static_cast<unsigned>((Q + 127) / 128), 128, 0, as_stream(stream)>>>(dp, ip, n_parts, Q, k, d2,
                                                                                          ids, 2 * Q * k, nullptr);
  EN_LAUNCHED("knn_merge_kernel");
  return EN_OK;
}

int en_knn_finalize_dist(const double* d2, int64_t n, float* dist, void* stream) {
  if (n == 0) return EN_OK;  // empty query set: nothing to do (the pointers of empty buffers may be null)
  EN_REQUIRE(d2 && dist && n > 0, "en_knn_finalize_dist: bad arguments");
  sqrt_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream)>>>(d2, n, dist);
  EN_LAUNCHED("sqrt_kernel");
  return EN_OK;
}

int en_knn_vote(const int64_t* ids, int64_t Q, int k, const int32_t* labels, int64_t n_total, int32_t* pred,
                void* stream) {
  EN_REQUIRE(ids && labels && pred && Q > 0 && k > 0 && n_total > 0, "en_knn_vote: bad arguments");
  knn_vote_kernel<<<static_cast<unsigned>((Q + 127) / 128), 128, 0, as_stream(stream)>>>(ids, Q, k, labels, n_total,
                                                                                         pred);
  EN_LAUNCHED("knn_vote_kernel");
  return EN_OK;
}

int en_knn_accuracy(const int64_t* ids, const int32_t* pred, const int32_t* query_labels, int64_t Q, int k_ids,
                    const int32_t* labels, int64_t n_total, int64_t* counts, void* stream) {
  EN_REQUIRE(ids && pred && query_labels && labels && counts && Q > 0 && k_ids > 0, "en_knn_accuracy: bad arguments");
  knn_accuracy_kernel<<<static_cast<unsigned>((Q + 127) / 128), 128, 0, as_stream(stream)>>>(
      ids, pred, query_labels, Q, k_ids, labels, n_total, reinterpret_cast<unsigned long long*>(counts));
  EN_LAUNCHED("knn_accuracy_kernel");
  return EN_OK;
}

}  // extern "C"
