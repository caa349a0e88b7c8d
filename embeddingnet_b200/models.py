"""Drop-in for the encoding-bank half of ``embedding_net/models.py`` (RocketFlash/EmbeddingNet).

Kept verbatim from the reference (/root/reference/embedding_net/models.py): ``EmbeddingNet.generate_encodings``
(61-84), ``save_encodings`` (86-90), ``predict`` (115-126), ``predict_knn`` (128-142),
``calculate_prediction_accuracy`` (144-161), ``train_embeddings_classifier`` (52-59) -- names, arguments, return
shapes.  What the snapshot leaves undefined or unwritten is defined here from its call sites:
``calculate_distances`` (models.py:123) and a ``KNeighborsClassifier``-shaped object (``fit`` / ``predict`` /
``kneighbors``, models.py:58,136,138) stored under ``encoded_training_data['knn_classifier']``.

The bank scan is a streaming tcgen05 distance GEMM (split-BF16 operands by default) with an in-register per-query
top-k, followed by an exact float64 re-rank that also PROVES, per query, that no rejected row could have made the
top-k; the rare queries without such a certificate are redone by float64 brute force (``csrc/knn.cu``).  With a ``torch.distributed`` process group the bank is sharded row-wise, one
shard per GPU, and the per-shard top-k lists are merged after one NCCL all-gather; results do not depend on the
number of shards.  Keras model construction, h5 / ONNX export and image IO are out of scope: ``base_model`` is any
object with ``predict(images) -> (n, d) float32``.
"""
from __future__ import annotations

import ctypes
import os
import random

import numpy as np
import torch

from . import _lib
from ._runtime import as_cuda_f32, ptr, require_cuda, stream_ptr, workspace


def draw_candidate_ranks(shard_counts, n_slots, my_rank):
    """Host half of the sharded mining protocol.  shard_counts (P, A, slots): candidates of every
    (anchor, slot) pair in each shard.  For every pair with candidates, in the reference's (anchor, slot) order, draw
    ``np.random.randint(total)`` -- the stream ``np.random.choice(candidates)`` consumes (dg:194,199) -- and hand the
    draw to the shard that owns that position of the ascending-id candidate list (prefix over the shard ranges).
    Returns (A, slots) int32: the rank inside ``my_rank``'s own candidates, -1 where another shard (or none) owns it.
    Every rank calls this with the same counts and the same RNG state, so all agree without further communication."""
    counts = np.asarray(shard_counts).astype(np.int64)
    world, A, slots = counts.shape
    total = counts.sum(axis=0)
    local = np.full((A, slots), -1, dtype=np.int32)
    ii, ss = np.nonzero(total[:, :n_slots] > 0)          # row-major = the reference's (anchor, slot) order
    if ii.size == 0:
        return local
    # ONE vectorised call: RandomState.randint with an array of upper bounds draws element by element from the same
    # bit stream as the per-pair scalar calls (pinned by tests/test_host_logic_cpu.py), without a Python loop per pair
    r = np.random.randint(0, total[ii, ss])
    cum = np.cumsum(counts[:, ii, ss], axis=0)           # (world, pairs): candidates in shards 0..q
    owner = (r[None, :] >= cum).sum(axis=0)              # first shard whose prefix exceeds the draw
    before = np.where(owner > 0, cum[np.maximum(owner - 1, 0), np.arange(ii.size)], 0)
    mine = owner == my_rank
    local[ii[mine], ss[mine]] = (r - before)[mine].astype(np.int32)
    return local


class BankKNNClassifier:
    """``sklearn.neighbors.KNeighborsClassifier``-shaped brute-force classifier over an encoding bank on B200.

    fit(X, y) uploads (this rank's shard of) the bank and prepares the operand planes (``precision``: "bf16x3",
    the default -- two BF16 planes, the scan runs at twice the TF32 rate -- or "tf32x3"); ``kneighbors`` returns
    ``(dist (Q, k) float32, idx (Q, k) int64)`` ascending by (distance, index) -- lowest index wins ties;
    ``predict`` is the uniform majority vote with ties resolved to the smallest class (sklearn's rule).

    process_group: optional torch.distributed group; rank r keeps rows [r*ceil(N/P), (r+1)*ceil(N/P)).
    certify: check the per-query exactness certificate after every scan (one scalar device->host read) and redo
    uncertified queries with the float64 brute-force kernel; ``last_uncertified`` counts them.  ``certify=False``
    skips the read (e.g. inside a CUDA graph); results are then exact only up to the scan's candidate slack.
    """

    PRECISIONS = {"tf32x3": _lib.EN_PREC_TF32X3, "bf16x3": _lib.EN_PREC_BF16X3}

    def __init__(self, n_neighbors=5, process_group=None, device=None, precision="bf16x3", certify=True):
        if precision not in self.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(self.PRECISIONS))
        self.n_neighbors = int(n_neighbors)
        self.process_group = process_group
        self.device = device
        self.precision = precision
        self.certify = bool(certify)
        self._last_flags = None
        self._last_uncertified = 0
        # queries per call up to which the CUDA-core streaming scan is used instead of the tensor-core scan
        # measured on B200 (tools/time_knn_smallq.py, 4M x 512 bank, ms per pass): Q = 1, 2: stream 1.42 / 1.32, small-Q
        # tensor scan 1.38; Q = 4: 1.98 vs 1.37; Q = 32: 1.53; Q = 64: 2.77 (engine: 2.2)
        self.stream_max_q = 2
        self.smallq_max_q = 32
        self._fitted = False

    @property
    def last_uncertified(self):
        """Queries of the last search whose exactness certificate failed (they were redone by float64 brute force).
        Small batches decide the redo on the device, so the count is only read back when someone asks for it."""
        if self._last_flags is not None:
            self._last_uncertified = int(self._last_flags.sum().item())
            self._last_flags = None
        return self._last_uncertified

    # -- sharding helpers
    def _world(self):
        if self.process_group is None:
            return 1, 0
        import torch.distributed as dist

        return dist.get_world_size(self.process_group), dist.get_rank(self.process_group)

    @staticmethod
    def shard_bounds(n_total, world, rank):
        per = (n_total + world - 1) // world
        lo = min(rank * per, n_total)
        return lo, min(lo + per, n_total)

    def fit(self, X, y):
        """X: (N, d) float32 bank (numpy or torch; every rank passes the full bank or see ``fit_shard``), y: N labels
        (any hashable, as in the reference where they are class-name strings, models.py:77)."""
        X = np.asarray(X) if not isinstance(X, torch.Tensor) else X
        n_total = X.shape[0]
        world, rank = self._world()
        lo, hi = self.shard_bounds(n_total, world, rank)
        self.classes_, y_ids = np.unique(np.asarray(y), return_inverse=True)
        return self.fit_shard(X[lo:hi], y_ids.astype(np.int32), lo, n_total, classes=self.classes_)

    def fit_shard(self, X_shard, label_ids_all, id_offset, n_total, classes=None):
        """Fit from this rank's rows only.  label_ids_all: int32 class ids of ALL N rows (4 bytes per row)."""
        dev = self.device or require_cuda()
        self.device = dev
        lib = _lib.load()
        self._bank = as_cuda_f32(X_shard, dev)
        if self._bank.dim() != 2:
            raise ValueError("BankKNNClassifier.fit: X must be (N, d)")
        n, d = self._bank.shape
        self._n_total, self._offset, self._d = int(n_total), int(id_offset), d
        self._prec = self.PRECISIONS[self.precision]
        dpad = lib.en_bank_dpad(d, self._prec)
        plane_dtype = torch.bfloat16 if self._prec == _lib.EN_PREC_BF16X3 else torch.float32
        self._hi = torch.empty((n, dpad), dtype=plane_dtype, device=dev)
        self._lo = torch.empty((n, dpad), dtype=plane_dtype, device=dev)
        assert self._hi.numel() * self._hi.element_size() == lib.en_bank_plane_bytes(n, d, self._prec) or n == 0
        self._norms = torch.empty(n, dtype=torch.float32, device=dev)
        if n > 0:
            _lib.call("en_bank_prepare", ptr(self._bank), n, d, self._prec, ptr(self._hi), ptr(self._lo),
                      ptr(self._norms), stream_ptr())
        ids = label_ids_all if isinstance(label_ids_all, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(label_ids_all, dtype=np.int32)))
        self._labels = ids.to(dev, torch.int32).contiguous()
        if self._labels.numel() != n_total:
            raise ValueError("BankKNNClassifier: need one label id per bank row (%d != %d)" %
                             (self._labels.numel(), n_total))
        if classes is not None:
            self.classes_ = np.asarray(classes)
        elif not hasattr(self, "classes_"):
            self.classes_ = np.arange(int(self._labels.max().item()) + 1)
        self._fitted = True
        return self

    # -- core search on device tensors
    def _search(self, q, k, exclude_labels=None):
        """q: (Q, d) CUDA float32.  Returns (d2 (Q, k) float64, ids (Q, k) int64) global, merged over shards."""
        if not self._fitted:
            raise RuntimeError("BankKNNClassifier: call fit() first")
        lib = _lib.load()
        dev = self.device
        Q, d = q.shape
        if d != self._d:
            raise ValueError("query dimension %d != bank dimension %d" % (d, self._d))
        if not (1 <= k <= _lib.EN_KNN_MAX_K):
            raise ValueError("n_neighbors must be in [1, %d]" % _lib.EN_KNN_MAX_K)
        if k > self._n_total and exclude_labels is None:
            # scikit-learn's message and behaviour (neighbors/_base.py): never pad with id -1, which Python indexing
            # (labels[-1], models.py:139) would silently turn into the LAST bank label
            raise ValueError("Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %d, "
                             "n_samples = %d" % (k, self._n_total, Q))
        n = self._bank.shape[0]
        # one buffer, two halves: the sharded path all-gathers (d2, ids) as ONE packed record per rank
        rec = torch.empty((2, Q, k), dtype=torch.int64, device=dev)
        d2 = rec[0].view(torch.float64)
        ids = rec[1]
        d2.fill_(float("inf"))
        ids.fill_(-1)
        self._last_flags = None
        self._last_uncertified = 0
        if n > 0 and Q > 0:
            ql = bl = None
            if exclude_labels is not None:
                ql = exclude_labels.to(dev, torch.int32).contiguous()
                bl = self._labels[self._offset:self._offset + n]
            flags = torch.empty(Q, dtype=torch.int32, device=dev) if self.certify else None
            smallq_ws = 0
            if (ql is None and self.stream_max_q < Q <= min(self.smallq_max_q, _lib.EN_KNN_SMALLQ_MAX_Q) and
                    self._prec == _lib.EN_PREC_BF16X3):
                smallq_ws = lib.en_ws_bytes_knn_smallq(Q, n, d, k)   # 0: k or d outside what that kernel holds
            if Q <= min(self.stream_max_q, _lib.EN_KNN_STREAM_MAX_Q) and ql is None:
                # the reference's own call pattern: one image per predict() -> HBM-bound streaming scan
                ws = workspace(lib.en_ws_bytes_knn_stream(Q, n, d, k), dev, "knn")
                _lib.call("en_knn_stream_topk", ptr(q), Q, d, ptr(self._bank), ptr(self._norms), n, self._offset, k,
                          ptr(d2), ptr(ids), ptr(flags), ptr(ws), ws.numel(), stream_ptr())
            elif smallq_ws:
                # a handful of queries: bank-stationary tensor-core scan, bounded by the bank stream
                ws = workspace(smallq_ws, dev, "knn")
                _lib.call("en_knn_smallq_topk", ptr(q), Q, d, ptr(self._bank), ptr(self._hi), ptr(self._lo),
                          ptr(self._norms), n, self._offset, k, ptr(d2), ptr(ids), ptr(flags), ptr(ws), ws.numel(),
                          stream_ptr())
            else:
                ws = workspace(lib.en_ws_bytes_knn(Q, n, d, k), dev, "knn")
                _lib.call("en_knn_shard_topk", ptr(q), Q, d, ptr(self._bank), ptr(self._hi), ptr(self._lo),
                          ptr(self._norms), n, self._offset, k, self._prec, ptr(ql), ptr(bl), ptr(d2), ptr(ids),
                          ptr(flags), ptr(ws), ws.numel(), stream_ptr())
            if flags is not None:
                self._redo_uncertified(q, k, ql, bl, flags, d2, ids)
        world, _ = self._world()
        if world > 1:
            import torch.distributed as dist

            rec_all = torch.empty((world, 2, Q, k), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(rec_all, rec, group=self.process_group)   # ONE NCCL all-gather over NVLink
            d2m = torch.empty((Q, k), dtype=torch.float64, device=dev)
            idm = torch.empty((Q, k), dtype=torch.int64, device=dev)
            _lib.call("en_knn_merge_packed", ptr(rec_all), world, Q, k, ptr(d2m), ptr(idm), stream_ptr())
            d2, ids = d2m, idm
        return d2, ids

    def _redo_uncertified(self, q, k, ql, bl, flags, d2, ids):
        """Float64 brute force for the queries whose certificate failed (near-ties at the candidate cut-off)."""
        lib = _lib.load()
        n, d = self._bank.shape
        Q = q.shape[0]
        if Q <= _lib.EN_KNN_EXACT_MAX_Q:
            # small batches (the reference's per-image predict): the redo is decided on the DEVICE -- flagged queries
            # are recomputed in place, no flag means two launches that return at once; no host read-back
            ws = workspace(lib.en_ws_bytes_knn_exact(Q, n, d, k), q.device, "knn_exact")
            _lib.call("en_knn_exact_redo", ptr(q), Q, d, ptr(self._bank), n, self._offset, k, ptr(ql), ptr(bl),
                      ptr(flags), ptr(d2), ptr(ids), ptr(ws), ws.numel(), stream_ptr())
            self._last_flags = flags
            return
        self._last_uncertified = int(flags.sum().item())     # the one host read of the certified path
        if self._last_uncertified == 0:
            return
        todo = flags.nonzero().reshape(-1)
        step = _lib.EN_KNN_EXACT_MAX_Q
        for s in range(0, todo.numel(), step):
            sel = todo[s:s + step]
            qq = q.index_select(0, sel).contiguous()
            qq_l = ql.index_select(0, sel).contiguous() if ql is not None else None
            m = qq.shape[0]
            e_d2 = torch.empty((m, k), dtype=torch.float64, device=q.device)
            e_id = torch.empty((m, k), dtype=torch.int64, device=q.device)
            ws = workspace(lib.en_ws_bytes_knn_exact(m, n, d, k), q.device, "knn_exact")
            _lib.call("en_knn_exact_topk", ptr(qq), m, d, ptr(self._bank), n, self._offset, k, ptr(qq_l), ptr(bl),
                      ptr(e_d2), ptr(e_id), ptr(ws), ws.numel(), stream_ptr())
            d2.index_copy_(0, sel, e_d2)
            ids.index_copy_(0, sel, e_id)

    def kneighbors_device(self, X, n_neighbors=None, exclude_labels=None):
        """Device-resident variant: returns (dist float32, idx int64) CUDA tensors, no host copy."""
        k = self.n_neighbors if n_neighbors is None else int(n_neighbors)
        q = as_cuda_f32(X, self.device)
        if q.dim() == 1:
            q = q.reshape(1, -1)
        d2, ids = self._search(q, k, exclude_labels)
        dist_f = torch.empty(d2.shape, dtype=torch.float32, device=d2.device)
        _lib.call("en_knn_finalize_dist", ptr(d2), d2.numel(), ptr(dist_f), stream_ptr())
        return dist_f, ids

    def kneighbors(self, X, n_neighbors=None, return_distance=True):
        dist_f, ids = self.kneighbors_device(X, n_neighbors)
        if return_distance:
            return dist_f.cpu().numpy(), ids.cpu().numpy()
        return ids.cpu().numpy()

    # -- offline hard-negative mining over the bank with the generator's strategies (dg:188-199; BASELINE config 4)
    def mine_negatives(self, anchors, anchor_labels, positives=None, pos_dist=None, margin=0.5, mode="semihard"):
        """For every anchor a and each of its (up to 8) positives p: one negative bank row chosen by ``mode``
        ('hardest' | 'random_hard' | 'semihard', datagenerators.py:188-199) among ALL bank rows of another class.

        anchors (A, d); anchor_labels (A,) class ids in the numbering of ``fit_shard`` / ``classes_``; either
        ``positives`` (A, S, d) embeddings or ``pos_dist`` (A, S) distances d_ap (negative = unused slot).
        Returns (A, S) int64 global bank ids, -1 where the strategy finds no candidate (the reference's None).
        The random strategies consume the global legacy NumPy RNG exactly as the reference does: one
        ``randint(len(candidates))`` per pair with candidates, pairs in (anchor, slot) order.  With a process group,
        every rank must call this with the same arguments and the same RNG state."""
        from .datagenerators import MODES

        if mode not in MODES:
            raise KeyError(mode)
        if not self._fitted:
            raise RuntimeError("BankKNNClassifier: call fit() first")
        lib = _lib.load()
        dev = self.device
        a = as_cuda_f32(anchors, dev)
        if a.dim() != 2:
            raise ValueError("anchors must be (A, d)")
        A, d = a.shape
        if d != self._d:
            raise ValueError("anchor dimension %d != bank dimension %d" % (d, self._d))
        if A == 0:
            S0 = np.asarray(pos_dist).shape[1] if pos_dist is not None else np.asarray(positives).shape[1]
            return np.zeros((0, S0), dtype=np.int64)
        al = (anchor_labels if isinstance(anchor_labels, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(anchor_labels, dtype=np.int32)))).to(dev, torch.int32).contiguous()
        MS = _lib.EN_MINE_MAX_SLOTS
        if pos_dist is None:
            p = as_cuda_f32(positives, dev)
            if p.dim() != 3 or p.shape[0] != A or p.shape[2] != d:
                raise ValueError("positives must be (A, S, d)")
            S = p.shape[1]
            rep = a.unsqueeze(1).expand(A, S, d).contiguous()
            pd = torch.empty((A, S), dtype=torch.float32, device=dev)
            _lib.call("en_pair_dist_exact", ptr(rep), ptr(p.contiguous()), A * S, d, ptr(pd), stream_ptr())
        else:
            pd = as_cuda_f32(pos_dist, dev)
            S = pd.shape[1]
        if S > MS:
            raise ValueError("at most %d positives per anchor (got %d)" % (MS, S))
        pos_d = torch.full((A, MS), -1.0, dtype=torch.float32, device=dev)
        pos_d[:, :S] = pd
        n = self._bank.shape[0]
        bl = self._labels[self._offset:self._offset + n].contiguous()
        world, rank = self._world()
        if mode == "hardest":
            # The reference takes np.argmax of the FLOAT32 loss (d_ap - d_an) + margin over all negatives and keeps it
            # if positive (dg:188-190, 235): float32 rounding can merge several nearest rows into one maximum, which
            # argmax resolves to the lowest index.  The four nearest other-class rows (exact float64 d2, ascending
            # (d2, id)) cover that: same float32 arithmetic per slot, lowest id among the equal maxima.
            kk = 4
            d2, ids = self._search(a, kk, exclude_labels=al)
            dist = torch.sqrt(d2.to(torch.float32))                                  # (A, kk) float32; inf = no such row
            loss = (pos_d[:, :, None] - dist[:, None, :]) + float(margin)            # (A, MS, kk), float32 left to right
            idx = ids[:, None, :].expand(A, MS, kk)
            tie = (loss == loss[:, :, :1]) & (idx >= 0)
            pick = torch.where(tie, idx, torch.full_like(idx, torch.iinfo(torch.int64).max)).min(dim=2).values
            ok = (loss[:, :, 0] > 0) & (pos_d >= 0) & (ids[:, :1] >= 0)
            out = torch.where(ok, pick, torch.full_like(pick, -1))
            return out[:, :S].cpu().numpy()
        ws = workspace(lib.en_ws_bytes_mine_bank(A, d), dev, "mine_bank")
        counts = torch.zeros((A, MS, 2), dtype=torch.int32, device=dev)
        if n > 0:
            _lib.call("en_mine_bank_count", ptr(a), ptr(al), ptr(pos_d), A, d, S, ctypes.c_float(margin), ptr(self._bank),
                      ptr(self._hi), ptr(self._lo), ptr(self._norms), ptr(bl), n, self._prec, ptr(counts), ptr(ws),
                      ws.numel(), stream_ptr())
        col = 0 if mode == "random_hard" else 1
        mine = counts[:, :, col].contiguous()
        if world > 1:
            import torch.distributed as dist_

            allc = torch.empty((world, A, MS), dtype=torch.int32, device=dev)
            dist_.all_gather_into_tensor(allc, mine, group=self.process_group)   # (pairs, P) counts, SURVEY 8(e)
        else:
            allc = mine.unsqueeze(0)
        local_rank = draw_candidate_ranks(allc.cpu().numpy(), S, rank)
        sel = torch.full((A, MS), -1, dtype=torch.int64, device=dev)
        # The select pass only has work for anchors with a draw that landed in THIS shard: compact them (the ranks are
        # on the host anyway), so a bank where few pairs have a candidate -- or a shard that owns few of the draws --
        # is walked by a fraction of the anchor tiles instead of all of them.
        need = np.nonzero((local_rank >= 0).any(axis=1))[0]
        if n > 0 and need.size > 0:
            if need.size == A:
                a_s, al_s, pd_s, rk = a, al, pos_d, torch.from_numpy(local_rank).to(dev)
                sel_s = sel
            else:
                idx = torch.from_numpy(need).to(dev)
                a_s = a.index_select(0, idx).contiguous()
                al_s = al.index_select(0, idx).contiguous()
                pd_s = pos_d.index_select(0, idx).contiguous()
                rk = torch.from_numpy(np.ascontiguousarray(local_rank[need])).to(dev)
                sel_s = torch.full((need.size, MS), -1, dtype=torch.int64, device=dev)
            ws_s = workspace(lib.en_ws_bytes_mine_bank(a_s.shape[0], d), dev, "mine_bank")
            _lib.call("en_mine_bank_select", ptr(a_s), ptr(al_s), ptr(pd_s), a_s.shape[0], d, S, ctypes.c_float(margin),
                      MODES[mode], ptr(rk), ptr(self._bank), ptr(self._hi), ptr(self._lo), ptr(self._norms), ptr(bl),
                      n, self._offset, self._prec, ptr(sel_s), ptr(ws_s), ws_s.numel(), stream_ptr())
            if need.size != A:
                sel.index_copy_(0, idx, sel_s)
        if world > 1:
            import torch.distributed as dist_

            dist_.all_reduce(sel, op=dist_.ReduceOp.MAX, group=self.process_group)  # one owner per pair, others -1
        return sel[:, :S].cpu().numpy()

    def predict_device(self, X):
        """(Q,) int32 class ids on the device."""
        k = self.n_neighbors
        q = as_cuda_f32(X, self.device)
        if q.dim() == 1:
            q = q.reshape(1, -1)
        _, ids = self._search(q, k)
        pred = torch.empty(q.shape[0], dtype=torch.int32, device=self.device)
        if q.shape[0] > 0:
            _lib.call("en_knn_vote", ptr(ids), q.shape[0], k, ptr(self._labels), self._n_total, ptr(pred),
                      stream_ptr())
        return pred, ids

    def predict(self, X):
        pred, _ = self.predict_device(X)
        return self.classes_[pred.cpu().numpy()]

    def score_topk(self, X, y):
        """Batched ``calculate_prediction_accuracy`` (models.py:144-161): one scan, on-device top-1 / top-5 tally."""
        if self.n_neighbors > self._n_total:
            raise ValueError("Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %d" %
                             (self.n_neighbors, self._n_total))
        k = min(max(self.n_neighbors, 5), self._n_total)  # a bank of fewer than 5 rows: top-"5" = all of them
        q = as_cuda_f32(X, self.device)
        Q = q.shape[0]
        _, ids = self._search(q, k)
        ids_vote = ids[:, :self.n_neighbors].contiguous()
        pred = torch.empty(Q, dtype=torch.int32, device=self.device)
        _lib.call("en_knn_vote", ptr(ids_vote), Q, self.n_neighbors, ptr(self._labels), self._n_total, ptr(pred),
                  stream_ptr())
        lut = {c: i for i, c in enumerate(self.classes_.tolist())}
        want = torch.tensor([lut.get(v, -1) for v in np.asarray(y).tolist()], dtype=torch.int32, device=self.device)
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        _lib.call("en_knn_accuracy", ptr(ids), ptr(pred), ptr(want), Q, k, ptr(self._labels), self._n_total,
                  ptr(counts), stream_ptr())
        c = counts.cpu().numpy()
        return {"top1": float(c[0]) / Q, "top5": float(c[1]) / Q}


class EmbeddingNet:
    """Bank / nearest-neighbour part of the reference class (models.py:22-161)."""

    def __init__(self, params, base_model=None):
        self.params_model = params.get("model", {})
        self.params_dataloader = params.get("dataloader", {})
        self.params_generator = params.get("generator", {})
        self.params_general = params.get("general", {})
        self.params_train = params.get("train", {})
        self.params_encodings = params.get("encodings", {})
        if "softmax" in params:
            self.params_softmax = params["softmax"]
        self.base_model = base_model
        self.backbone_model = None
        self.model = None
        self.input_shape = self.params_model.get("input_shape")
        if "work_dir" in self.params_general and "project_name" in self.params_general:
            self.workdir_path = os.path.join(self.params_general["work_dir"], self.params_general["project_name"])
        self.encoded_training_data = {}

    # -- image hooks (out of scope; overridable)
    def _load_images(self, paths):
        import cv2

        shape = self.params_model.get("input_shape")
        imgs = []
        for p in paths:
            img = cv2.imread(p)
            if img is not None and shape:
                img = cv2.resize(img, (shape[0], shape[1]))
            imgs.append(img)
        return np.array(imgs)

    def _load_image(self, image):
        import cv2

        img = cv2.imread(image) if isinstance(image, str) else image
        shape = self.input_shape or self.params_model.get("input_shape")
        return cv2.resize(img, (shape[0], shape[1]))

    def _generate_encodings(self, imgs):
        return self.base_model.predict(imgs)                                             # models.py:47-49

    def generate_encodings(self, data_loader, max_n_samples=10, shuffle=True):
        """models.py:61-84: at most ``max_n_samples`` per class, rows appended class by class."""
        data_paths, data_labels, data_encodings = [], [], []
        encoded_training_data = {}
        for class_name in data_loader.class_names:
            data_list = data_loader.train_data[class_name]
            if len(data_list) > max_n_samples:
                if shuffle:
                    random.shuffle(data_list)
                data_list = data_list[:max_n_samples]
            data_paths += data_list
            imgs = self._load_images(data_list)
            encods = self._generate_encodings(imgs)
            for encod in encods:
                data_encodings.append(encod)
                data_labels.append(class_name)
        encoded_training_data["paths"] = data_paths
        encoded_training_data["labels"] = data_labels
        encoded_training_data["encodings"] = np.squeeze(np.array(data_encodings))
        self.encoded_training_data = encoded_training_data
        return encoded_training_data

    def save_encodings(self, encoded_training_data, save_folder="./", save_file_name="encodings.pkl"):
        """models.py:86-90 (the fitted GPU classifier is not picklable and is dropped)."""
        from .utils import save_encodings

        save_encodings(encoded_training_data, save_folder, save_file_name)  # keeps {paths, labels, encodings} only

    def load_encodings(self, path_to_encodings, fit_knn=True):
        """What tools/test.py:22 calls (absent in the snapshot): ``utils.load_encodings`` + classifier fit."""
        from .utils import load_encodings

        self.encoded_training_data = load_encodings(path_to_encodings)
        if fit_knn:
            self.fit_knn()
        return self.encoded_training_data

    def fit_knn(self, n_neighbors=None, process_group=None):
        k = n_neighbors or self.params_encodings.get("knn_k", 5) or 5
        clf = BankKNNClassifier(n_neighbors=k, process_group=process_group)
        clf.fit(self._bank_matrix(), self.encoded_training_data["labels"])
        self.encoded_training_data["knn_classifier"] = clf
        return clf

    def train_embeddings_classifier(self, data_loader, classification_model, max_n_samples=10, shuffle=True):
        """models.py:52-59."""
        encodings = self.generate_encodings(data_loader, max_n_samples=max_n_samples, shuffle=shuffle)
        classification_model.fit(encodings["encodings"], encodings["labels"])
        if isinstance(classification_model, BankKNNClassifier):
            self.encoded_training_data["knn_classifier"] = classification_model

    def _bank_matrix(self):
        enc = np.asarray(self.encoded_training_data["encodings"], dtype=np.float32)
        return enc.reshape(1, -1) if enc.ndim == 1 else enc   # np.squeeze quirk for a one-row bank (models.py:82)

    def _nn1(self):
        """1-NN searcher over the current bank.  Cached on the object (never inside the user-visible bank dict, which
        ``save_encodings`` pickles) and rebuilt when 'encodings' / 'labels' are replaced."""
        data = self.encoded_training_data
        enc, labels = data["encodings"], data["labels"]
        key = (id(enc), getattr(enc, "shape", None), id(labels), len(labels))
        cache = getattr(self, "_nn1_cache", None)
        if cache is None or cache[0] != key:
            clf = BankKNNClassifier(n_neighbors=1)
            clf.fit(self._bank_matrix(), labels)
            cache = (key, clf, enc, labels)   # holds the arrays: their ids cannot be recycled while cached
            self._nn1_cache = cache
        return cache[1]

    def calculate_distances(self, encoding):
        """The method ``predict`` calls but the snapshot never defines (models.py:123): Euclidean distance from one
        encoding to every bank row, shape (N,)."""
        clf = self._nn1()                      # the fitted device bank: no re-upload, no (N, d) broadcast of the query
        bank = clf._bank
        n, d = bank.shape
        q = as_cuda_f32(np.asarray(encoding, np.float32).reshape(1, -1), bank.device)
        if q.shape[1] != d:
            raise ValueError("encoding dimension %d != bank dimension %d" % (q.shape[1], d))
        dist = torch.empty(n, dtype=torch.float32, device=bank.device)
        _lib.call("en_query_distances", ptr(bank), ptr(q), n, d, ptr(dist), stream_ptr())  # one pass over the bank
        return dist.cpu().numpy()

    def predict(self, image):
        """models.py:115-126: nearest bank row (np.argmin -> lowest index on ties) -> its label."""
        img = self._load_image(image)
        encoding = self.base_model.predict(np.expand_dims(img, axis=0))
        return self.predict_encoding(encoding)

    def predict_encoding(self, encoding):
        _, idx = self._nn1().kneighbors(np.asarray(encoding, np.float32).reshape(1, -1), n_neighbors=1)
        return self.encoded_training_data["labels"][int(idx[0, 0])]

    def predict_knn(self, image, with_top5=False):
        """models.py:128-142."""
        img = self._load_image(image)
        encoding = self.base_model.predict(np.expand_dims(img, axis=0))
        return self.predict_knn_encoding(encoding, with_top5=with_top5)

    def predict_knn_encoding(self, encoding, with_top5=False):
        clf = self.encoded_training_data["knn_classifier"]
        encoding = np.asarray(encoding, np.float32).reshape(1, -1)
        predicted_label = clf.predict(encoding)                                          # models.py:136, shape (1,)
        if with_top5:
            prediction_top5_idx = clf.kneighbors(encoding, n_neighbors=5)                # models.py:138
            prediction_top5 = [self.encoded_training_data["labels"][prediction_top5_idx[1][0][i]] for i in range(5)]
            return predicted_label, prediction_top5
        return predicted_label

    def calculate_prediction_accuracy(self, data_loader, batched=True):
        """models.py:144-161.  ``batched`` embeds all validation images, then ONE bank scan + on-device tally
        (SURVEY F3); ``batched=False`` keeps the reference's one-image-per-call loop."""
        paths = data_loader.images_paths["val"]
        labels = data_loader.images_labels["val"]
        if batched:
            imgs = self._load_images(paths)
            enc = np.asarray(self.base_model.predict(imgs), np.float32)
            return self.encoded_training_data["knn_classifier"].score_topk(enc, labels)
        correct_top1 = correct_top5 = 0
        for img_path, img_label in zip(paths, labels):
            prediction, prediction_top5 = self.predict_knn(img_path, with_top5=True)
            if prediction[0] == img_label:
                correct_top1 += 1
            if img_label in prediction_top5:
                correct_top5 += 1
        n = len(paths)
        return {"top1": correct_top1 / n, "top5": correct_top5 / n}
