"""Drop-in for the hard-negative mining half of ``embedding_net/datagenerators.py`` (RocketFlash/EmbeddingNet).

``TripletsDataGenerator`` keeps the reference constructor and method signatures
(/root/reference/embedding_net/datagenerators.py:159-261):

    TripletsDataGenerator(embedding_model, class_files_paths, class_names, n_batches=10, input_shape=None,
                          batch_size=32, augmentations=None, k_classes=5, k_samples=5, margin=0.5,
                          negatives_selection_mode='semihard')
    .hardest_negative(loss_values, margin=0.5) / .random_hard_negative(...) / .semihard_negative(...) -> int | None
    .get_batch_triplets_mining() -> ([A, P, N], targets)
    .__getitem__(index)

The arithmetic (pairwise distance matrix, per-pair candidate predicates, arg-max / r-th candidate selection) runs in
CUDA kernels through the C ABI; the host keeps only what must stay on the host to be index-for-index identical to
the reference: the pair enumeration order (dg:225-234) and the draws from the *global legacy NumPy RNG*
(dg:194,199,202,205).  ``np.random.choice(candidates)`` consumes the stream exactly like
``candidates[np.random.randint(0, len(candidates))]``, so the GPU returns candidate *counts*, the host draws the
rank, and the GPU returns the rank-th candidate in ascending row order.

Image loading / augmentation (cv2, albumentations) is outside the hot path; ``_get_images_set`` is kept as the
overridable hook it is in the reference.
"""
from __future__ import annotations

import ctypes
from itertools import combinations

import numpy as np
import torch

from . import _lib
from ._runtime import ptr, require_cuda, stream_ptr, workspace

MODES = {"semihard": _lib.EN_MODE_SEMIHARD, "hardest": _lib.EN_MODE_HARDEST, "random_hard": _lib.EN_MODE_RANDOM_HARD}


# ------------------------------------------------------------------------------------------------ numeric core
def pairwise_distances(x, squared=False, exact=True, return_device=False):
    """``sklearn.metrics.pairwise_distances(x)`` as called at dg:219 (Euclidean, float32 out, zero diagonal).

    exact=True uses the float64 CUDA-core kernel (matches sklearn's float32 results bit for bit up to float64
    summation order); exact=False uses the tcgen05 3xTF32 GEMM."""
    dev = require_cuda()
    xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float32)))
    xt = xt.to(dev, torch.float32).contiguous()
    if xt.dim() != 2:
        raise ValueError("pairwise_distances: expected a (n, d) array")
    n, d = xt.shape
    out = torch.empty((n, n), dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws = workspace(lib.en_ws_bytes_pairwise(n, d, int(exact)), dev, "pairwise")
    _lib.call("en_pairwise_dist", ptr(xt), n, d, int(squared), int(exact), ptr(out), ptr(ws), ws.numel(),
              stream_ptr())
    return out if return_device else out.cpu().numpy()


def enumerate_pairs(labels):
    """(anchor, positive) pairs in the reference's order: classes in first-appearance order (class-major batches:
    dg:225), ``combinations(positive_indices, 2)`` inside a class (dg:231) -- i < j, anchor = i."""
    labels = np.asarray(labels).reshape(-1)
    _, first = np.unique(labels, return_index=True)
    order = labels[np.sort(first)]
    pairs = []
    for c in order:
        idx = np.where(labels == c)[0]
        pairs.extend(combinations(idx.tolist(), 2))
    return np.asarray(pairs, dtype=np.int32).reshape(-1, 2)


def mine_batch_triplets(all_embeddings, labels, margin=0.5, mode="semihard", exact=True):
    """Numeric core of ``get_batch_triplets_mining`` (dg:217-250) on row ids.

    Returns ``(triplets (T, 3) int64, used_fallback)`` in the reference's emission order; consumes the global
    ``np.random`` stream exactly as dg:194,199 do."""
    if mode not in MODES:
        raise KeyError(mode)
    dev = require_cuda()
    labels = np.asarray(labels).reshape(-1)
    D = pairwise_distances(all_embeddings, squared=False, exact=exact, return_device=True)   # dg:219
    n = D.shape[0]
    if labels.shape[0] != n:
        raise ValueError("mine_batch_triplets: one label per embedding row expected")
    pairs_h = enumerate_pairs(labels)
    n_pairs = pairs_h.shape[0]
    if n_pairs == 0:
        raise ValueError("mine_batch_triplets: no (anchor, positive) pair in the batch")
    lab_d = torch.from_numpy(labels.astype(np.int32)).to(dev)
    pairs_d = torch.from_numpy(pairs_h).to(dev)
    scan = torch.empty((3, n_pairs), dtype=torch.int32, device=dev)
    _lib.call("en_mine_batch_scan", ptr(D), ptr(lab_d), n, ptr(pairs_d), n_pairs, ctypes.c_float(margin),
              ptr(scan[0]), ptr(scan[1]), ptr(scan[2]), stream_ptr())
    scan_h = scan.cpu().numpy()
    if mode == "hardest":
        negs = scan_h[0].astype(np.int64)                                   # dg:188-190
    else:
        counts = scan_h[1] if mode == "random_hard" else scan_h[2]          # dg:192-199
        rank = np.full(n_pairs, -1, dtype=np.int32)
        has = counts > 0
        if has.any():
            # one vectorised draw in the reference's pair order: same bit stream as the per-pair
            # np.random.randint(len(candidates)) calls (pinned by tests/test_host_logic_cpu.py)
            rank[has] = np.random.randint(0, counts[has].astype(np.int64))
        rank_d = torch.from_numpy(rank).to(dev)
        sel = torch.empty(n_pairs, dtype=torch.int32, device=dev)
        _lib.call("en_mine_batch_select", ptr(D), ptr(lab_d), n, ptr(pairs_d), n_pairs, ctypes.c_float(margin),
                  MODES[mode], ptr(rank_d), ptr(sel), stream_ptr())
        negs = sel.cpu().numpy().astype(np.int64)
    keep = negs >= 0
    trip = np.stack([pairs_h[keep, 0].astype(np.int64), pairs_h[keep, 1].astype(np.int64), negs[keep]], axis=1)
    fallback = False
    if trip.shape[0] == 0:                                                  # dg:246-250
        a, p = pairs_h[-1]
        other = np.where(labels != labels[a])[0]
        # reference: negative_indices of the LAST class iterated, first entry
        last_cls_mask = labels != labels[pairs_h[-1, 0]]
        first_neg = int(np.where(last_cls_mask)[0][0]) if other.size else int(a)
        trip = np.array([[a, p, first_neg]], dtype=np.int64)
        fallback = True
    return trip, fallback


def _select_on_vector(loss_values, margin, mode):
    """One of dg:188-199 on a single loss vector, evaluated on the GPU."""
    dev = require_cuda()
    lv = torch.from_numpy(np.ascontiguousarray(np.asarray(loss_values, np.float32).reshape(-1))).to(dev)
    n = lv.numel()
    if n == 0:
        return None
    out = torch.empty(3, dtype=torch.int32, device=dev)
    _lib.call("en_loss_scan", ptr(lv), n, ctypes.c_float(margin), ptr(out), stream_ptr())
    arg, n_hard, n_semi = (int(v) for v in out.cpu().numpy())
    if mode == "hardest":
        return arg if arg >= 0 else None
    count = n_hard if mode == "random_hard" else n_semi
    if count == 0:
        return None
    r = int(np.random.randint(0, count))  # == np.random.choice(candidates) stream-wise
    sel = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("en_loss_select", ptr(lv), n, ctypes.c_float(margin), MODES[mode], r, ptr(sel), stream_ptr())
    return int(sel.item())


def gather_triplets(rows, triplets):
    """A, P, N = rows[trip[:, 0]], rows[trip[:, 1]], rows[trip[:, 2]] on the device (dg:241-243,252-256).

    rows: CUDA float32 tensor (n, ...) -- the sampled images or embeddings; triplets: (T, 3) integer array.
    Returns three CUDA tensors of shape (T, ...)."""
    if not (isinstance(rows, torch.Tensor) and rows.is_cuda and rows.dtype == torch.float32):
        raise ValueError("gather_triplets: rows must be a CUDA float32 tensor")
    rows = rows.contiguous()
    n = rows.shape[0]
    row_len = int(rows[0].numel()) if n else 0
    trip = torch.as_tensor(np.ascontiguousarray(np.asarray(triplets, dtype=np.int64))).to(rows.device)
    T = trip.shape[0]
    out = [torch.empty((T,) + tuple(rows.shape[1:]), dtype=torch.float32, device=rows.device) for _ in range(3)]
    if T and row_len:
        _lib.call("en_gather_triplet_rows", ptr(rows), n, row_len, ptr(trip), T, ptr(out[0]), ptr(out[1]), ptr(out[2]),
                  stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ generators
class ENDataGenerator:
    """Bookkeeping of the reference base class (dg:114-156); image IO is delegated to ``image_loader``."""

    def __init__(self, class_files_paths, class_names, val_gen=False, input_shape=None, batch_size=32, n_batches=10,
                 n_batches_val=10, augmentations=None):
        self.input_shape = input_shape
        self.augmentations = augmentations
        self.batch_size = batch_size
        self.n_batches = n_batches
        self.n_batches_val = n_batches_val
        self.val_gen = val_gen
        self.class_files_paths = class_files_paths
        self.class_names = class_names
        self.n_classes = len(self.class_names)
        self.n_samples = {k: len(v) for k, v in self.class_files_paths.items()}

    def __len__(self):
        return self.n_batches_val if self.val_gen else self.n_batches

    def __getitem__(self, index):
        pass

    def _get_images_set(self, clsss, idxs, with_aug=True):
        """dg:145-156.  Needs OpenCV; override for other sources (the tests feed row ids)."""
        import cv2  # image IO is out of scope of the hot path; imported lazily

        if type(clsss) is list:
            img_paths = [self.class_files_paths[cl][idx] for cl, idx in zip(clsss, idxs)]
        else:
            img_paths = [self.class_files_paths[clsss][idx] for idx in idxs]
        imgs = []
        for p in img_paths:
            img = cv2.imread(p)
            if img is not None and self.input_shape:
                img = cv2.resize(img, (self.input_shape[0], self.input_shape[1]))
            imgs.append(img)
        if with_aug:
            imgs = [self.augmentations(image=img)["image"] for img in imgs]
        return np.array(imgs) / 255.0


class TripletsDataGenerator(ENDataGenerator):
    """dg:159-261 with the distance / selection arithmetic on the GPU."""

    def __init__(self, embedding_model, class_files_paths, class_names, n_batches=10, input_shape=None,
                 batch_size=32, augmentations=None, k_classes=5, k_samples=5, margin=0.5,
                 negatives_selection_mode="semihard"):
        super().__init__(class_files_paths=class_files_paths, class_names=class_names, input_shape=input_shape,
                         batch_size=batch_size, n_batches=n_batches, augmentations=augmentations)
        modes = {"semihard": self.semihard_negative, "hardest": self.hardest_negative,
                 "random_hard": self.random_hard_negative}
        self.embedding_model = embedding_model
        self.k_classes = k_classes
        self.k_samples = k_samples
        self.margin = margin
        self.negatives_selection_mode = negatives_selection_mode
        self.negative_selection_fn = modes[negatives_selection_mode]  # KeyError on unknown mode, as in the reference

    def hardest_negative(self, loss_values, margin=0.5):
        return _select_on_vector(loss_values, margin, "hardest")

    def random_hard_negative(self, loss_values, margin=0.5):
        return _select_on_vector(loss_values, margin, "random_hard")

    def semihard_negative(self, loss_values, margin=0.5):
        return _select_on_vector(loss_values, margin, "semihard")

    def get_batch_triplets_mining(self):
        selected_classes_idxs = np.random.choice(self.n_classes, size=self.k_classes, replace=False)       # dg:202
        selected_classes = [self.class_names[cl] for cl in selected_classes_idxs]
        selected_classes_n_elements = [self.n_samples[cl] for cl in selected_classes]
        selected_images = [np.random.choice(cl_n, size=self.k_samples, replace=True)
                           for cl_n in selected_classes_n_elements]                                       # dg:205
        all_embeddings_list, all_images_list = [], []
        for idx, cl_img_idxs in enumerate(selected_images):
            images = self._get_images_set(selected_classes[idx], cl_img_idxs, with_aug=self.augmentations)
            all_images_list.append(images)
            all_embeddings_list.append(self.embedding_model.predict(images))                               # dg:214
        labels = np.repeat(np.arange(self.k_classes), self.k_samples)                                      # dg:226-227
        # Device-resident path: when the image hook and the embedding model hand back CUDA tensors,
        # embeddings go straight into the distance kernel and the mined A / P / N batches are gathered on the device;
        # only the (T, 3) indices and the candidate counts (which the host RNG needs) ever reach the host.
        on_device = all(isinstance(t, torch.Tensor) and t.is_cuda for t in all_embeddings_list + all_images_list)
        if on_device:
            all_embeddings = torch.cat([e.to(torch.float32) for e in all_embeddings_list], dim=0)          # dg:217
            all_images = torch.cat([i.to(torch.float32) for i in all_images_list], dim=0)
        else:
            all_embeddings = np.vstack([np.asarray(e.cpu() if isinstance(e, torch.Tensor) else e)
                                        for e in all_embeddings_list])
            all_images = np.vstack([np.asarray(i.cpu() if isinstance(i, torch.Tensor) else i)
                                    for i in all_images_list])
        trip, _ = mine_batch_triplets(all_embeddings, labels, margin=self.margin,
                                      mode=self.negatives_selection_mode)
        if on_device:
            triplets = gather_triplets(all_images, trip)
        else:
            triplets = [all_images[trip[:, 0]], all_images[trip[:, 1]], all_images[trip[:, 2]]]            # dg:252-256
        targets = np.ones(trip.shape[0], dtype=np.int64)                                                   # dg:244,255
        return triplets, targets

    def __getitem__(self, index):
        return self.get_batch_triplets_mining()
