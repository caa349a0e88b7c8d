/* embeddingnet_b200 -- C ABI of the B200-native distance / mining / loss / bank-kNN hot path.
 *
 * This is the drop-in boundary: plain C, device pointers and sizes only, no torch types.
 * The reference (RocketFlash/EmbeddingNet) has no FFI of its own -- its "API" for this path is a handful of
 * Python callables -- so every entry point below cites the reference call site (file:line under /root/reference)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are asynchronous with respect
 *     to the host unless stated; outputs are valid once the stream has been synchronised;
 *   - return value: 0 = OK, negative = argument error detected on the host before any launch (EN_ERR_*),
 *     positive = cudaError_t.  en_last_error() returns a thread-local message for the last failure;
 *   - functions are re-entrant and keep no global mutable state (tensor-map encoder lookup is idempotent);
 *   - workspace: functions taking (ws, ws_bytes) need scratch of at least the matching en_ws_bytes_*() size,
 *     256-byte aligned; the library never allocates device memory;
 *   - all matrices are row-major fp32, labels are int32, ids are int64 unless stated.
 */
#ifndef EMBEDDINGNET_B200_H
#define EMBEDDINGNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EN_OK 0
#define EN_ERR_ARG (-1)        /* bad shape / null pointer / unsupported value */
#define EN_ERR_WORKSPACE (-2)  /* workspace too small or misaligned */
#define EN_ERR_DRIVER (-3)     /* cuTensorMapEncodeTiled unavailable / failed */
#define EN_ERR_ARCH (-4)       /* device is not sm_100 */
#define EN_ERR_COMM (-5)       /* NCCL reported an error (en_comm_*) */

#define EN_MODE_SEMIHARD 0     /* datagenerators.py:196-199 */
#define EN_MODE_HARDEST 1      /* datagenerators.py:188-190 */
#define EN_MODE_RANDOM_HARD 2  /* datagenerators.py:192-194 */

const char* en_version(void);
const char* en_last_error(void);
/* Number of kernel launches issued by this thread through the library since the last reset (bench accounting). */
int64_t en_launch_count(void);
void en_launch_count_reset(void);

/* Measurement hook for bench.py: while enabled (per thread), entry points that contain a distance GEMM or a bank
 * scan record CUDA events immediately around THAT launch on the caller's stream; en_prof_last_ms synchronises on
 * the second event and returns the kernel's duration in milliseconds (host pointer). */
int en_prof_enable(int on);
int en_prof_last_ms(float* ms_host);
/* Stage times of the last multi-kernel entry point that marks its stages (en_batch_hard_fwd[_bwd]: operand split,
 * distance GEMM, fast finalize, slow finalize), in milliseconds; n_out receives how many were written. */
int en_prof_marks_ms(float* ms_host, int capacity, int* n_out);

/* ---------------------------------------------------------------- row-wise kernels (memory bound) */
/* K.l2_normalize(x, axis=1): y = x * rsqrt(max(sum x^2, 1e-12)).  backbones.py:38,77,118 */
int en_l2_normalize_fwd(const float* x, float* y, int64_t rows, int d, void* stream);
int en_l2_normalize_bwd(const float* x, const float* gy, float* gx, int64_t rows, int d, void* stream);

/* triplet_loss(margin)(y_true, y_pred): y_pred is (B, total_len) = [anchor | positive | negative];
 * loss[b] = max(|a-p|^2 - |a-n|^2 + margin, 0).  losses_and_accuracies.py:26-42.  total_len % 3 must be 0. */
int en_triplet_apn_fwd(const float* y_pred, int64_t B, int total_len, float margin, float* loss, void* stream);
/* gy_pred[b, :] = gloss[b] * d loss[b] / d y_pred[b, :] */
int en_triplet_apn_bwd(const float* y_pred, const float* gloss, int64_t B, int total_len, float margin,
                       float* gy_pred, void* stream);

/* contrastive_loss(y_true, y_pred) = mean(y*d^2 + (1-y)*max(1-d,0)^2), margin literal 1.
 * losses_and_accuracies.py:4-11.  n = number of elements; loss is ONE float. */
int en_contrastive_fwd(const float* y_true, const float* y_pred, int64_t n, float* loss, void* stream);
int en_contrastive_bwd(const float* y_true, const float* y_pred, const float* gloss, int64_t n, float* gy_pred,
                       void* stream);
/* accuracy(y_true, y_pred) = mean(y_true == (y_pred < 0.5)).  losses_and_accuracies.py:47-50 */
int en_pair_accuracy(const float* y_true, const float* y_pred, int64_t n, float* acc, void* stream);

/* Siamese heads, models.py:217-228.  l2: dist[b] = sqrt(max(sum((e1-e2)^2), 1e-7));  l1: out = |e1 - e2| */
int en_siamese_l2_fwd(const float* e1, const float* e2, int64_t B, int d, float* dist, void* stream);
int en_siamese_l2_bwd(const float* e1, const float* e2, const float* gdist, int64_t B, int d, float* g1, float* g2,
                      void* stream);
int en_siamese_l1_fwd(const float* e1, const float* e2, int64_t n, float* out, void* stream);
int en_siamese_l1_bwd(const float* e1, const float* e2, const float* gout, int64_t n, float* g1, float* g2,
                      void* stream);

/* calculate_distances(encoding): the method EmbeddingNet.predict calls at models.py:123 (undefined in the snapshot;
 * np.argmin over its result at models.py:124): dist[i] = sqrt(sum((bank[i] - query)^2)), one query row against all
 * n bank rows, no clamp (an exact match reports 0).  Reads the bank once: n * d * 4 bytes. */
int en_query_distances(const float* bank, const float* query, int64_t n, int d, float* dist, void* stream);

/* x[0..n) *= scale[0] in place (scale is a DEVICE scalar): the upstream gradient of a scalar loss applied to the
 * d loss / d emb that the fused forward+backward entry points store (Keras multiplies the loss by sample / loss
 * weights before differentiating, losses_and_accuracies.py callables are used through model.compile, train.py:160). */
int en_scale_inplace(float* x, int64_t n, const float* scale, void* stream);

/* ---------------------------------------------------------------- pairwise distances + in-batch mining */
/* sklearn.metrics.pairwise_distances(x) as called at datagenerators.py:219, with sklearn's float32 semantics
 * (float64 -2xy+|x|^2+|y|^2, cast f32, clamp >= 0, zero diagonal, sqrt unless `squared`).  out is (n, n).
 * exact != 0: float64 CUDA-core path (bit-compatible with sklearn up to float64 summation order);
 * exact == 0: tcgen05 3xTF32 path (diagnostics / large n). */
size_t en_ws_bytes_pairwise(int64_t n, int d, int exact);
int en_pairwise_dist(const float* x, int64_t n, int d, int squared, int exact, float* out, void* ws, size_t ws_bytes,
                     void* stream);

/* Per (anchor, positive) pair scan of the selection predicates of datagenerators.py:188-199 over ALL rows whose
 * label differs from the anchor's, ascending row order, on loss = (D[a,p] - D[a,n]) + margin in float32:
 *   hardest[p]   = row id of the first maximum of loss if that maximum is > 0, else -1
 *   n_hard[p]    = #{n : loss > 0}
 *   n_semi[p]    = #{n : 0 < loss < margin}
 * D is the (n, n) matrix from en_pairwise_dist; pairs is (n_pairs, 2) int32 (anchor, positive). */
int en_mine_batch_scan(const float* D, const int32_t* labels, int64_t n, const int32_t* pairs, int64_t n_pairs,
                       float margin, int32_t* hardest, int32_t* n_hard, int32_t* n_semi, void* stream);
/* The rank[p]-th candidate (0-based, ascending row id) of pair p under `mode` (EN_MODE_SEMIHARD or
 * EN_MODE_RANDOM_HARD); -1 when rank[p] < 0 or out of range.  The host draws rank[p] from the legacy NumPy RNG
 * in the reference's pair order, which reproduces np.random.choice(candidates) at datagenerators.py:194,199. */
int en_mine_batch_select(const float* D, const int32_t* labels, int64_t n, const int32_t* pairs, int64_t n_pairs,
                         float margin, int mode, const int32_t* rank, int32_t* selected, void* stream);

/* The three selection callables of datagenerators.py:188-199 applied to ONE float32 loss vector (n values):
 * out3 = { index of the first maximum if that maximum is > 0 else -1, #{loss > 0}, #{0 < loss < margin} };
 * en_loss_select returns the rank-th (0-based, ascending index) candidate of `mode` in out1[0] (-1 if none). */
int en_loss_scan(const float* loss_values, int64_t n, float margin, int32_t* out3, void* stream);
int en_loss_select(const float* loss_values, int64_t n, float margin, int mode, int rank, int32_t* out1,
                   void* stream);

/* Gather of the mined triplets (datagenerators.py:241-243,252-256): a/p/n (T, row_len) <- src[triplets[t, 0|1|2]].
 * src is the (n_rows, row_len) sampled set (flattened images or embeddings) already resident on the device, so the
 * batch never returns to the host between embedding, mining and the loss.  triplets (T, 3) int64. */
int en_gather_triplet_rows(const float* src, int64_t n_rows, int64_t row_len, const int64_t* triplets, int64_t T,
                           float* a, float* p, float* n, void* stream);

/* ---------------------------------------------------------------- fused in-batch losses (tcgen05 distance GEMM) */
/* Batch-hard triplet loss (BASELINE.json north_star; Hermans et al. / Moindrot, cited at README.md:112,116 --
 * not implemented by the reference).  emb (B, d), labels (B,).
 *   loss (1 float) = mean_i hinge_i, hinge_i = max(hp_i - hn_i + margin, 0)   (soft: softplus(hp_i - hn_i))
 * Saved for backward: hp_idx/hn_idx (B,) int32 (-1 = none), coef (B,) = d mean / d(hp_i - hn_i),
 * hp/hn (B,) distances (squared or not).  The B x B matrix is never written to memory. */
size_t en_ws_bytes_batch_hard(int64_t B, int d);
int en_batch_hard_fwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared, int soft,
                      float* loss, int32_t* hp_idx, int32_t* hn_idx, float* hp, float* hn, float* coef, void* ws,
                      size_t ws_bytes, void* stream);
/* Loss AND gradient in one call (what a training step needs): gemb (B, d) = gloss[0] * d loss / d emb, with
 * gloss == NULL meaning 1.  The gradient is accumulated by the same kernel that picks the hardest pairs, so the
 * step is prep + memset + distance GEMM + finalize.  Same saved outputs as en_batch_hard_fwd. */
int en_batch_hard_fwd_bwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                          int soft, float* loss, int32_t* hp_idx, int32_t* hn_idx, float* hp, float* hn, float* coef,
                          const float* gloss, float* gemb, void* ws, size_t ws_bytes, void* stream);
/* gemb (B, d) = gloss[0] * d loss / d emb; gemb is fully overwritten. */
int en_batch_hard_bwd(const float* emb, int64_t B, int d, int squared, const int32_t* hp_idx, const int32_t* hn_idx,
                      const float* hp, const float* hn, const float* coef, const float* gloss, float* gemb,
                      void* stream);

/* ---------------------------------------------------------------- host-buffer batch-hard step (pipelined)
 * The loss callable of embedding_net/losses_and_accuracies.py:14-44 for a caller whose batch lives in HOST memory
 * (NumPy arrays, a data loader's pinned staging buffers): embeddings (B, d) + labels (B,) in, loss (1 float) and
 * d loss / d embeddings (B, d) out, every pointer below a HOST pointer.  A pipe owns `depth` (1..8) slots carved
 * from the caller's device block (en_bh_host_pipe_device_bytes, 256-byte aligned), three internal streams
 * (upload | compute | download) and one CUDA graph of the en_batch_hard_fwd_bwd kernels per slot, so consecutive
 * steps overlap: the PCIe link, which bounds this call, stays busy in both directions.
 *   submit: enqueues one step's upload and returns its ticket.  The calling thread also makes the stage hand-offs
 *           of the steps in flight (no stream waits on another stream's event: measured 15 % faster): before it
 *           returns it launches the PREVIOUS step's kernels (waiting for that step's upload) and issues the download
 *           of the step before that (waiting for its kernels), so in a steady loop it blocks for about one upload
 *           time; it also waits for the oldest step when all `depth` slots are in flight.
 *           hp_idx_host / hn_idx_host (B,) int32 are optional (NULL = not wanted).
 *           Host buffers should be page-locked (cudaHostAlloc / cudaHostRegister); pageable memory is accepted
 *           but serialises the copies.  They must stay valid and untouched until wait(ticket) returns.
 *   wait:   carries that step through its remaining stages and returns once its outputs are in host memory
 *           (waiting for the NEWEST ticket drains the pipe; a loop that reads results 2+ steps behind never stalls).
 * One pipe is driven by one thread at a time, with the device that was current at creation current. */
size_t en_bh_host_pipe_device_bytes(int64_t B, int d, int depth);
int en_bh_host_pipe_create(int64_t B, int d, float margin, int squared, int soft, int depth, void* device_mem,
                           size_t device_bytes, void** pipe_out);
int en_bh_host_pipe_submit(void* pipe, const float* emb_host, const int32_t* labels_host, float* loss_host,
                           float* grad_host, int32_t* hp_idx_host, int32_t* hn_idx_host, int64_t* ticket_out);
int en_bh_host_pipe_wait(void* pipe, int64_t ticket);
int en_bh_host_pipe_destroy(void* pipe);

/* Batch-all triplet loss: sum over valid (i,j,k) of max(D_ij - D_ik + margin, 0) / #{terms > 1e-16}.
 * out[0] = loss, out[1] = fraction of positive triplets.  max_positives >= largest class size - 1 (<= 63).
 * stats (3 doubles, device): hinge sum, #positive terms, #valid triplets -- kept for the backward. */
size_t en_ws_bytes_batch_all(int64_t B, int d, int max_positives);
int en_batch_all_fwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                     int max_positives, float* out, double* stats, void* ws, size_t ws_bytes, void* stream);
int en_batch_all_bwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                     int max_positives, const double* stats, const float* gloss, float* gemb, void* ws,
                     size_t ws_bytes, void* stream);
/* Loss AND gradient from ONE pass over the distance tiles (what a training step needs; Keras differentiates the loss
 * callable inside the same train step, train.py:160-162): out / stats as en_batch_all_fwd, gemb (B, d) =
 * gloss[0] * d loss / d emb (gloss == NULL means 1).  Asynchronous (no host read-back): if a class turns out to have
 * more positives per anchor than max_positives, out / stats / gemb are filled with NaN and `overflow` (device int32,
 * optional) receives that count (0 when fine).  Classes with more than 8 positives per anchor (max_positives > 8)
 * stay on the same fused tensor-core kernel: their positives lists are sorted, padded to 64 slots and searched per
 * element; the lists then hold up to 64 entries, so overflow is only reported past that (B <= 65535 rows). */
int en_batch_all_fwd_bwd(const float* emb, const int32_t* labels, int64_t B, int d, float margin, int squared,
                         int max_positives, float* out, double* stats, const float* gloss, float* gemb,
                         int32_t* overflow, void* ws, size_t ws_bytes, void* stream);

/* All-pairs contrastive loss: losses_and_accuracies.py:4-11 over every ordered pair i != j with
 * y_ij = [label_i == label_j] and d_ij = sqrt(max(|e_i - e_j|^2, 1e-7)) (models.py:225).  loss is 1 float. */
size_t en_ws_bytes_contrastive_allpairs(int64_t B, int d);
int en_contrastive_allpairs_fwd(const float* emb, const int32_t* labels, int64_t B, int d, float* loss, void* ws,
                                size_t ws_bytes, void* stream);
int en_contrastive_allpairs_bwd(const float* emb, const int32_t* labels, int64_t B, int d, const float* gloss,
                                float* gemb, void* ws, size_t ws_bytes, void* stream);
/* Loss AND gradient in one pass (see en_batch_all_fwd_bwd); gloss == NULL means 1. */
int en_contrastive_allpairs_fwd_bwd(const float* emb, const int32_t* labels, int64_t B, int d, float* loss,
                                    const float* gloss, float* gemb, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- encoding bank: nearest neighbours */
/* Operand format of the tensor-core scan.  Both split every fp32 value into two planes (hi, lo) and issue three
 * MMAs per k-step (hi*hi + hi*lo + lo*hi):
 *   EN_PREC_TF32X3  TF32 planes (4 B / element each), kind::tf32, ~1e-6 relative: fp32-faithful;
 *   EN_PREC_BF16X3  BF16 planes (2 B / element each), kind::f16 at twice the rate, ~4e-6 relative: enough to SELECT
 *                   candidates, which the exact float64 re-rank then orders. */
#define EN_PREC_TF32X3 0
#define EN_PREC_BF16X3 1
/* Bank preparation (the `fit` of the KNeighborsClassifier-shaped object models.py:58 expects): splits the fp32
 * bank shard into the two planes the tensor-core scan streams and computes squared row norms.
 * dpad = en_bank_dpad(d, precision); hi/lo are (n, dpad) planes of en_bank_plane_bytes() each, norms (n,). */
int en_bank_dpad(int d, int precision);
size_t en_bank_plane_bytes(int64_t n, int d, int precision);
int en_bank_prepare(const float* bank, int64_t n, int d, int precision, void* hi, void* lo, float* norms,
                    void* stream);

/* k nearest bank rows of every query, ordered by (exact distance, global id) -- kneighbors() at models.py:138,
 * np.argmin at models.py:124 for k = 1.  Two stages inside one call: a tcgen05 scan (3 MMAs per k-step on the
 * split planes) keeps the best k + EN_KNN_SLACK candidates per query in registers, then those few are re-evaluated
 * exactly in float64 (sum (q-b)^2) and re-ranked, which makes ids independent of tensor-core rounding and of how
 * the bank is sharded.
 * bank / bank_hi / bank_lo / bank_norms describe THIS shard's n_bank rows whose global ids start at id_offset.
 * exclude_label (may be NULL): candidates whose bank label equals query_labels[q] are skipped -- offline
 * hard-negative mining over a bank (BASELINE config 4).  Outputs: d2 (Q, k) float64 squared distances,
 * ids (Q, k) int64 (-1 = fewer than k candidates), ascending.
 * uncertified (Q,) int32, may be NULL: 0 when the call PROVED (rigorous bound on the scan's rounding error against
 * the gap between the k-th result and the best rejected candidate) that the k ids are the exact nearest rows,
 * 1 when it could not (near-ties at the cut-off, duplicated rows): redo those queries with en_knn_exact_topk. */
#define EN_KNN_SLACK 3
#define EN_KNN_MAX_K 29
size_t en_ws_bytes_knn(int64_t Q, int64_t n_bank, int d, int k);
int en_knn_shard_topk(const float* queries, int64_t Q, int d, const float* bank, const void* bank_hi,
                      const void* bank_lo, const float* bank_norms, int64_t n_bank, int64_t id_offset, int k,
                      int precision, const int32_t* query_labels, const int32_t* bank_labels, double* d2,
                      int64_t* ids, int32_t* uncertified, void* ws, size_t ws_bytes, void* stream);
/* Small query sets (5 <= Q <= EN_KNN_SMALLQ_MAX_Q; between the one-image-per-call pattern of models.py:122,135 and
 * the batched accuracy loop of models.py:144-161) on the tensor cores with the BANK as the 128-row operand and the
 * queries resident in shared memory: the kernel is bounded by the bank stream (n_bank * d * 4 bytes of BF16 planes
 * per call), not by 128-row MMAs on a mostly empty query tile.  Needs the EN_PREC_BF16X3 planes of en_bank_prepare
 * and k <= 5.  Same outputs, exact re-rank and certificate as en_knn_shard_topk.  en_ws_bytes_knn_smallq returns 0
 * when d is too large for the resident query tile (use en_knn_shard_topk then). */
#define EN_KNN_SMALLQ_MAX_Q 64
size_t en_ws_bytes_knn_smallq(int64_t Q, int64_t n_bank, int d, int k);
int en_knn_smallq_topk(const float* queries, int64_t Q, int d, const float* bank, const void* bank_hi,
                       const void* bank_lo, const float* bank_norms, int64_t n_bank, int64_t id_offset, int k,
                       double* d2, int64_t* ids, int32_t* uncertified, void* ws, size_t ws_bytes, void* stream);

/* Float64 brute force over the shard (sum (q-b)^2 for every admissible row, ordered by (d2, id)): the reference
 * semantics with no filter in front.  For the queries en_knn_shard_topk / en_knn_stream_topk flag as uncertified;
 * Q <= EN_KNN_EXACT_MAX_Q per call.  Same outputs. */
#define EN_KNN_EXACT_MAX_Q 64
size_t en_ws_bytes_knn_exact(int64_t Q, int64_t n_bank, int d, int k);
int en_knn_exact_topk(const float* queries, int64_t Q, int d, const float* bank, int64_t n_bank, int64_t id_offset,
                      int k, const int32_t* query_labels, const int32_t* bank_labels, double* d2, int64_t* ids,
                      void* ws, size_t ws_bytes, void* stream);
/* The same brute force decided on the DEVICE: only queries whose flag (the `uncertified` output of a scan) is non-zero
 * are recomputed and overwritten in d2 / ids; with no flag set the launches return at once.  A one-image predict call
 * (models.py:122,135) then never reads the certificate back to the host. */
int en_knn_exact_redo(const float* queries, int64_t Q, int d, const float* bank, int64_t n_bank, int64_t id_offset,
                      int k, const int32_t* query_labels, const int32_t* bank_labels, const int32_t* flags, double* d2,
                      int64_t* ids, void* ws, size_t ws_bytes, void* stream);
/* Small-batch variant for the reference's actual call pattern (one query per predict(), models.py:122,135):
 * a CUDA-core fp32 streaming scan bounded by HBM bandwidth; Q <= EN_KNN_STREAM_MAX_Q.  Same outputs.
 * bank_norms (n_bank squared row norms from en_bank_prepare) may be NULL: the scan then evaluates
 * sum (q-b)^2 directly instead of |b|^2 - 2 q.b (twice the FP32 work per element). */
#define EN_KNN_STREAM_MAX_Q 8
size_t en_ws_bytes_knn_stream(int64_t Q, int64_t n_bank, int d, int k);
int en_knn_stream_topk(const float* queries, int64_t Q, int d, const float* bank, const float* bank_norms,
                       int64_t n_bank, int64_t id_offset, int k, double* d2, int64_t* ids, int32_t* uncertified,
                       void* ws, size_t ws_bytes, void* stream);
/* Merge P per-shard lists (P, Q, k) (as gathered by an NCCL all-gather) into the global top-k by (d2, id). */
int en_knn_merge(const double* d2_parts, const int64_t* id_parts, int n_parts, int64_t Q, int k, double* d2,
                 int64_t* ids, void* stream);
/* Merge over the packed records of ONE all-gather: parts is (n_parts, 2, Q, k) 8-byte words -- per shard the (Q, k)
 * float64 squared distances followed by the (Q, k) int64 ids, i.e. what a rank holds when d2 and ids are the two halves
 * of one buffer. */
int en_knn_merge_packed(const void* parts, int n_parts, int64_t Q, int k, double* d2, int64_t* ids, void* stream);
/* distances (Q, k) float32 = sqrt(d2). */
int en_knn_finalize_dist(const double* d2, int64_t n, float* dist, void* stream);
/* Majority vote of KNeighborsClassifier.predict (models.py:136): labels of the k neighbours are looked up by
 * global id in `labels` (n_total,), ties resolve to the smallest label id.  pred is (Q,) int32. */
int en_knn_vote(const int64_t* ids, int64_t Q, int k, const int32_t* labels, int64_t n_total, int32_t* pred,
                void* stream);
/* calculate_prediction_accuracy, models.py:144-161: counts[0] += #(pred == label), counts[1] += #(label among the
 * labels of the first 5 neighbours).  counts is 2 x int64 and must be zeroed by the caller. */
int en_knn_accuracy(const int64_t* ids, const int32_t* pred, const int32_t* query_labels, int64_t Q, int k_ids,
                    const int32_t* labels, int64_t n_total, int64_t* counts, void* stream);

/* ---------------------------------------------------------------- bank-scale mining with the generator's strategies */
/* datagenerators.py:188-199 over an encoding bank (BASELINE config 4).  For anchor a and its
 * positive slot s with distance pos_d[a, s] (< 0 = unused slot; at most EN_MINE_MAX_SLOTS per anchor):
 *   loss_n = (pos_d[a, s] - d(a, n)) + margin over bank rows n with bank_labels[n] != anchor_labels[a]   (dg:235)
 *   random_hard candidates: loss_n > 0;  semihard candidates: 0 < loss_n < margin                        (dg:193,197)
 * d(a, n) = sqrtf((float) sum_f64 (a - n)^2); elements whose scan value lies within the scan's error bound of a
 * predicate boundary are re-evaluated exactly on the spot, so counts and selected ids are those of the float64
 * definition for either operand format.
 * en_mine_bank_count : counts (A, EN_MINE_MAX_SLOTS, 2) int32 = [random_hard, semihard] candidates in THIS shard.
 * en_mine_bank_select: rank (A, EN_MINE_MAX_SLOTS) int32 = 0-based rank among this shard's candidates of `mode` in
 *   ascending row id (< 0: nothing to select here) -> selected (A, EN_MINE_MAX_SLOTS) int64 global id or -1.
 * n_slots = number of leading slots in use (1..EN_MINE_MAX_SLOTS; the arrays keep the EN_MINE_MAX_SLOTS stride).
 * The host draws the rank from the legacy NumPy RNG (np.random.choice(c) == c[randint(len(c))], dg:194,199); with a
 * sharded bank the per-shard counts are all-gathered and the owning shard resolves the rank.
 * en_pair_dist_exact: dist[i] = d(a_i, b_i) with the definition above (the d_ap inputs).
 * The hardest strategy is the label-excluded nearest neighbour: en_knn_shard_topk with query_labels. */
#define EN_MINE_MAX_SLOTS 8
int en_pair_dist_exact(const float* a, const float* b, int64_t n, int d, float* dist, void* stream);
size_t en_ws_bytes_mine_bank(int64_t A, int d);
int en_mine_bank_count(const float* anchors, const int32_t* anchor_labels, const float* pos_d, int64_t A, int d,
                       int n_slots, float margin, const float* bank, const void* bank_hi, const void* bank_lo,
                       const float* bank_norms, const int32_t* bank_labels, int64_t n_bank, int precision,
                       int32_t* counts, void* ws, size_t ws_bytes, void* stream);
int en_mine_bank_select(const float* anchors, const int32_t* anchor_labels, const float* pos_d, int64_t A, int d,
                        int n_slots, float margin, int mode, const int32_t* rank, const float* bank, const void* bank_hi,
                        const void* bank_lo, const float* bank_norms, const int32_t* bank_labels, int64_t n_bank,
                        int64_t id_offset, int precision, int64_t* selected, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- embedding head */
/* Dense(n_out, activation="relu") [+ K.l2_normalize(axis=1)]: the last layers of every backbone the reference builds
 * (backbones.py:114-119, :36-38, :75-77) -- the producer of the embeddings this library consumes.  One tcgen05
 * 3xTF32 GEMM with bias + ReLU + row normalisation in its epilogue.
 * en_dense_prepare: w is the Keras kernel, (n_in, n_out) row-major; w_hi / w_lo receive its transposed TF32 planes,
 * en_dense_plane_bytes() each (once per set of weights).
 * en_dense_relu_fwd: x (B, n_in) -> out (B, n_out) = relu(x . w + bias), each row scaled by
 * rsqrt(max(sum y^2, 1e-12)) when normalize != 0.  bias may be NULL.  inv_norm (optional, (B,), normalize only)
 * receives that row scale for the backward pass (negated for rows below the 1e-12 clamp).
 * en_dense_relu_bwd: the training direction (the reference trains through these layers, backbones.py:114-119 under
 * model.fit, train.py:172): gy = dL/d out -> gx (B, n_in), gw (n_in, n_out; the Keras kernel layout), gb (n_out,);
 * any of the three may be NULL.  w is the fp32 Keras kernel; y / inv_norm are the forward's outputs.  Two tcgen05
 * 3xTF32 GEMMs (gpre . w^T and x^T . gpre) after a row-wise kernel that undoes the normalisation and the ReLU. */
size_t en_dense_plane_bytes(int n_in, int n_out);
int en_dense_prepare(const float* w, int n_in, int n_out, float* w_hi, float* w_lo, void* stream);
size_t en_ws_bytes_dense(int64_t B, int n_in);
int en_dense_relu_fwd(const float* x, int64_t B, int n_in, const float* w_hi, const float* w_lo, const float* bias,
                      int n_out, int normalize, float* out, float* inv_norm, void* ws, size_t ws_bytes, void* stream);
size_t en_ws_bytes_dense_bwd(int64_t B, int n_in, int n_out);
int en_dense_relu_bwd(const float* x, int64_t B, int n_in, const float* w, int n_out, int normalize, const float* y,
                      const float* inv_norm, const float* gy, float* gx, float* gw, float* gb, void* ws,
                      size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- synthetic data (bench / tests) */
/* x[r, c] = u(r, c) in [-1, 1) from a splitmix64 counter hash; optional class structure:
 * x = relu?(centre[label(r)] + noise * u) with label(r) = (r + row_offset) / rows_per_class (class-major) when
 * rows_per_class > 0, else (r + row_offset) % n_classes.  Bit-identical to embeddingnet_b200.synth (NumPy). */
int en_synth_fill(float* x, int64_t rows, int d, int64_t row_offset, uint64_t seed_centre, uint64_t seed_noise,
                  int64_t n_classes, int64_t rows_per_class, float noise, int relu, int32_t* labels_out,
                  void* stream);

/* ---------------------------------------------------------------- NCCL plumbing of the sharded bank paths */
/* The only partitioned paths: bank kNN (models.py:128-161 at bank scale) and bank mining
 * (datagenerators.py:188-199 at bank scale), bank rows sharded over one process per GPU.  Their exchange step is one
 * all-gather of small records (per-shard top-k lists / candidate counts) and, for mining, one all-reduce(max) of the
 * selected ids.  The Python host issues them through torch.distributed; these entry points let any other host do the
 * same through this library (NCCL is bound with dlopen at first use; the reference itself has no collective).
 *   rank 0:     en_comm_unique_id(id)            -> ship the EN_COMM_ID_BYTES bytes to the other ranks (file, socket..)
 *   every rank: cudaSetDevice(rank's GPU); en_comm_init(nranks, rank, id, &comm)
 *   per search: en_knn_shard_topk(..., d2, ids) into the two halves of one (2, Q, k) buffer,
 *               en_comm_allgather(comm, buffer, all, 2*Q*k*8, stream), en_knn_merge_packed(all, nranks, Q, k, ...)
 * Collectives on one communicator must be issued in the same order on every rank and not concurrently. */
#define EN_COMM_ID_BYTES 128
int en_comm_unique_id(void* id_bytes_host);
int en_comm_init(int nranks, int rank, const void* id_bytes_host, void** comm_out);
int en_comm_allgather(void* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream);
int en_comm_allreduce_max_i64(void* comm, const int64_t* send, int64_t* recv, size_t count, void* stream);
int en_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* EMBEDDINGNET_B200_H */
