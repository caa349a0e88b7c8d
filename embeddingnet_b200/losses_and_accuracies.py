"""Drop-in for ``embedding_net/losses_and_accuracies.py`` of RocketFlash/EmbeddingNet, on B200 CUDA kernels.

Same callables, same argument meaning (reference: /root/reference/embedding_net/losses_and_accuracies.py):

    contrastive_loss(y_true, y_pred)                      lac:4-11   -> scalar
    triplet_loss(margin=0.5) -> loss_function(y_true, y_pred)   lac:14-44  -> per-sample (B,) vector
    accuracy(y_true, y_pred)                              lac:47-50  -> scalar

plus the in-batch losses BASELINE.json asks for behind the same ``factory(...) -> fn(y_true, y_pred)`` shape
(``y_true`` = integer labels (B,), ``y_pred`` = embeddings (B, d)); the reference only cites their papers:

    batch_hard_triplet_loss(margin=0.5, squared=False, soft=False)
    batch_all_triplet_loss(margin=0.5, squared=False, max_positives=None)
    contrastive_loss_all_pairs()

and the two Keras Lambdas that feed the losses: ``l2_normalize`` (backbones.py:38) and the Siamese distance heads
``siamese_l2_distance`` / ``siamese_l1_distance`` (models.py:217-228).

TensorFlow is not involved: tensors are ``torch.Tensor`` on a CUDA device (NumPy inputs are uploaded), and every
callable is a ``torch.autograd.Function`` so ``loss.backward()`` runs the hand-written backward kernels.  There is
no CPU fallback and no framework dispatch -- each call goes straight to ``libembeddingnet_b200.so``.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._runtime import as_cuda_f32, as_cuda_i32, ptr, stream_ptr, workspace


def _labels(y_true, device):
    return as_cuda_i32(y_true, device)


# ------------------------------------------------------------------------------------------------ lac:4-11
class _Contrastive(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_true, y_pred):
        n = y_pred.numel()
        if y_true.numel() != n:
            raise ValueError("contrastive_loss: y_true and y_pred must have the same number of elements")
        loss = torch.empty((), dtype=torch.float32, device=y_pred.device)
        _lib.call("en_contrastive_fwd", ptr(y_true), ptr(y_pred), n, ptr(loss), stream_ptr())
        ctx.save_for_backward(y_true, y_pred)
        return loss

    @staticmethod
    def backward(ctx, g):
        y_true, y_pred = ctx.saved_tensors
        g = g.contiguous().to(torch.float32)
        out = torch.empty_like(y_pred)
        _lib.call("en_contrastive_bwd", ptr(y_true), ptr(y_pred), ptr(g), y_pred.numel(), ptr(out), stream_ptr())
        return None, out


def contrastive_loss(y_true, y_pred):
    """Contrastive loss from Hadsell-et-al.'06 (margin fixed at 1, label 1 = similar).  lac:4-11."""
    y_pred = as_cuda_f32(y_pred)
    y_true = as_cuda_f32(y_true, y_pred.device)
    return _Contrastive.apply(y_true, y_pred)


# ------------------------------------------------------------------------------------------------ lac:14-44
class _TripletAPN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, margin):
        B, total = y_pred.shape
        loss = torch.empty(B, dtype=torch.float32, device=y_pred.device)
        _lib.call("en_triplet_apn_fwd", ptr(y_pred), B, total, ctypes.c_float(margin), ptr(loss), stream_ptr())
        ctx.save_for_backward(y_pred)
        ctx.margin = margin
        return loss

    @staticmethod
    def backward(ctx, g):
        (y_pred,) = ctx.saved_tensors
        B, total = y_pred.shape
        g = g.contiguous().to(torch.float32)
        out = torch.empty_like(y_pred)
        _lib.call("en_triplet_apn_bwd", ptr(y_pred), ptr(g), B, total, ctypes.c_float(ctx.margin), ptr(out),
                  stream_ptr())
        return out, None


def triplet_loss(margin=0.5):
    """Triplet hinge on squared L2 over a pre-mined ``(B, 3d)`` concatenation ``[anchor | positive | negative]``.

    Returns ``loss_function(y_true, y_pred)`` giving the per-sample ``(B,)`` vector (Keras applies the batch mean);
    ``y_true`` is ignored, as in the reference.  lac:14-44."""

    def loss_function(y_true, y_pred):
        y_pred = as_cuda_f32(y_pred)
        if y_pred.dim() != 2:
            raise ValueError("triplet_loss: y_pred must be (B, 3*d)")
        return _TripletAPN.apply(y_pred, float(margin))

    return loss_function


# ------------------------------------------------------------------------------------------------ lac:47-50
def accuracy(y_true, y_pred):
    """Classification accuracy with a fixed 0.5 threshold on distances.  lac:47-50."""
    y_pred = as_cuda_f32(y_pred)
    y_true = as_cuda_f32(y_true, y_pred.device)
    if y_true.numel() != y_pred.numel():
        raise ValueError("accuracy: y_true and y_pred must have the same number of elements")
    out = torch.empty((), dtype=torch.float32, device=y_pred.device)
    _lib.call("en_pair_accuracy", ptr(y_true), ptr(y_pred), y_pred.numel(), ptr(out), stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ bb:38 / models:217-228
class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        rows, d = x.shape
        y = torch.empty_like(x)
        _lib.call("en_l2_normalize_fwd", ptr(x), ptr(y), rows, d, stream_ptr())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        rows, d = x.shape
        g = g.contiguous().to(torch.float32)
        gx = torch.empty_like(x)
        _lib.call("en_l2_normalize_bwd", ptr(x), ptr(g), ptr(gx), rows, d, stream_ptr())
        return gx


def l2_normalize(x):
    """``K.l2_normalize(x, axis=1)`` of the backbone head (backbones.py:38,77,118)."""
    x = as_cuda_f32(x)
    if x.dim() != 2:
        raise ValueError("l2_normalize: expected (rows, d)")
    return _L2Normalize.apply(x)


class _SiameseL2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e1, e2):
        B, d = e1.shape
        dist = torch.empty((B, 1), dtype=torch.float32, device=e1.device)
        _lib.call("en_siamese_l2_fwd", ptr(e1), ptr(e2), B, d, ptr(dist), stream_ptr())
        ctx.save_for_backward(e1, e2)
        return dist

    @staticmethod
    def backward(ctx, g):
        e1, e2 = ctx.saved_tensors
        B, d = e1.shape
        g = g.contiguous().to(torch.float32)
        g1, g2 = torch.empty_like(e1), torch.empty_like(e2)
        _lib.call("en_siamese_l2_bwd", ptr(e1), ptr(e2), ptr(g), B, d, ptr(g1), ptr(g2), stream_ptr())
        return g1, g2


def siamese_l2_distance(e1, e2):
    """``sqrt(max(sum((e1-e2)^2, axis=1, keepdims=True), K.epsilon()))`` -> (B, 1).  models.py:225."""
    e1 = as_cuda_f32(e1)
    e2 = as_cuda_f32(e2, e1.device)
    if e1.shape != e2.shape or e1.dim() != 2:
        raise ValueError("siamese_l2_distance: expected two (B, d) tensors")
    return _SiameseL2.apply(e1, e2)


class _SiameseL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e1, e2):
        out = torch.empty_like(e1)
        _lib.call("en_siamese_l1_fwd", ptr(e1), ptr(e2), e1.numel(), ptr(out), stream_ptr())
        ctx.save_for_backward(e1, e2)
        return out

    @staticmethod
    def backward(ctx, g):
        e1, e2 = ctx.saved_tensors
        g = g.contiguous().to(torch.float32)
        g1, g2 = torch.empty_like(e1), torch.empty_like(e2)
        _lib.call("en_siamese_l1_bwd", ptr(e1), ptr(e2), ptr(g), e1.numel(), ptr(g1), ptr(g2), stream_ptr())
        return g1, g2


def siamese_l1_distance(e1, e2):
    """``K.abs(e1 - e2)`` -> (B, d).  models.py:218."""
    e1 = as_cuda_f32(e1)
    e2 = as_cuda_f32(e2, e1.device)
    if e1.shape != e2.shape:
        raise ValueError("siamese_l1_distance: shapes differ")
    return _SiameseL1.apply(e1, e2)


# ------------------------------------------------------------------------------------------------ batch-hard
def _consume_stored_gradient(ctx, g, who):
    """Backward of the fused losses: forward already stored d loss / d emb; hand that buffer to autograd, scaled in
    place by the upstream gradient (the kernel returns at once when it is exactly 1, i.e. plain loss.backward())."""
    gemb = getattr(ctx, "gemb", None)
    if gemb is None:
        raise RuntimeError("%s: backward was already run for this forward (the fused step stores d loss / d emb "
                           "once and hands the buffer over); call the loss again" % who)
    ctx.gemb = None
    g = g.contiguous().to(torch.float32).reshape(1)
    _lib.call("en_scale_inplace", ptr(gemb), gemb.numel(), ptr(g), stream_ptr())
    return gemb


class _BatchHard(torch.autograd.Function):
    """Forward computes the loss AND d loss / d emb in the same pass (en_batch_hard_fwd_bwd) whenever a gradient can
    be asked for; backward is then a scalar multiply."""

    @staticmethod
    def forward(ctx, emb, labels, margin, squared, soft):
        B, d = emb.shape
        dev = emb.device
        lib = _lib.load()
        ws = workspace(lib.en_ws_bytes_batch_hard(B, d), dev, "batch_hard")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        saved_i = torch.empty((2, B), dtype=torch.int32, device=dev)
        saved_f = torch.empty((3, B), dtype=torch.float32, device=dev)
        if ctx.needs_input_grad[0]:
            gemb = torch.empty_like(emb)
            _lib.call("en_batch_hard_fwd_bwd", ptr(emb), ptr(labels), B, d, ctypes.c_float(margin), int(squared),
                      int(soft), ptr(loss), ptr(saved_i[0]), ptr(saved_i[1]), ptr(saved_f[0]), ptr(saved_f[1]),
                      ptr(saved_f[2]), None, ptr(gemb), ptr(ws), ws.numel(), stream_ptr())
            ctx.gemb = gemb   # not save_for_backward: backward hands this very buffer to autograd (no copy)
        else:
            _lib.call("en_batch_hard_fwd", ptr(emb), ptr(labels), B, d, ctypes.c_float(margin), int(squared),
                      int(soft), ptr(loss), ptr(saved_i[0]), ptr(saved_i[1]), ptr(saved_f[0]), ptr(saved_f[1]),
                      ptr(saved_f[2]), ptr(ws), ws.numel(), stream_ptr())
        return loss

    @staticmethod
    def backward(ctx, g):
        return _consume_stored_gradient(ctx, g, "batch_hard_triplet_loss"), None, None, None, None


def batch_hard_triplet_loss(margin=0.5, squared=False, soft=False):
    """Batch-hard triplet loss (Hermans et al. 2017 / Moindrot): for every anchor the hardest positive and hardest
    negative inside the batch, ``mean(max(hp - hn + margin, 0))`` (``soft`` -> ``softplus(hp - hn)``).

    ``fn(y_true=labels (B,), y_pred=embeddings (B, d)) -> scalar``.  One tcgen05 distance GEMM whose epilogue does
    the label masking and per-anchor arg-max / arg-min; the B x B matrix is never stored."""

    def loss_function(y_true, y_pred):
        emb = as_cuda_f32(y_pred)
        if emb.dim() != 2:
            raise ValueError("batch_hard_triplet_loss: y_pred must be (B, d)")
        labels = _labels(y_true, emb.device)
        if labels.numel() != emb.shape[0]:
            raise ValueError("batch_hard_triplet_loss: one label per embedding row expected")
        return _BatchHard.apply(emb, labels, float(margin), bool(squared), bool(soft))

    return loss_function


# ------------------------------------------------------------------------------------------------ batch-all
class _DeferredOverflow:
    """The fused batch-all step never stalls the stream, so "a class has more positives per anchor than
    max_positives" cannot raise at once: the kernel poisons loss and gradient with NaN (never silently wrong) and
    leaves the offending count in a device word, which is copied to pinned host memory behind an event.  The flags of
    earlier calls are looked at whenever that costs nothing: at the next call (non-blocking ``event.query()``) and,
    blocking, by ``check()``."""

    SLOTS = 8

    def __init__(self):
        self._tls = __import__("threading").local()

    def _state(self, device):
        st = getattr(self._tls, "st", None)
        if st is None or st["device"] != device:
            st = {"device": device, "dev": torch.zeros(self.SLOTS, dtype=torch.int32, device=device),
                  "host": torch.zeros(self.SLOTS, dtype=torch.int32).pin_memory(), "events": [None] * self.SLOTS,
                  "limits": [0] * self.SLOTS, "next": 0}
            self._tls.st = st
        return st

    def slot(self, device, limit):
        """Device int32 (1,) view for the next call; its previous use, if any, is checked first."""
        st = self._state(device)
        self.poll(block=False)
        k = st["next"]
        if st["events"][k] is not None:  # ring wrapped around: settle the old use of this slot
            st["events"][k].synchronize()
            self._raise_if_set(st, k)
        st["next"] = (k + 1) % self.SLOTS
        st["limits"][k] = int(limit)
        return k, st["dev"][k:k + 1]

    def arm(self, k):
        st = self._tls.st
        st["host"][k:k + 1].copy_(st["dev"][k:k + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        st["events"][k] = ev

    def _raise_if_set(self, st, k):
        st["events"][k] = None
        n = int(st["host"][k])
        if n > 0:
            raise ValueError("batch_all_triplet_loss: a class has %d positives per anchor but max_positives = %d "
                             "(that call's loss and gradient were filled with NaN)" % (n, st["limits"][k]))

    def poll(self, block=False):
        st = getattr(self._tls, "st", None)
        if st is None:
            return
        for k in range(self.SLOTS):
            ev = st["events"][k]
            if ev is not None and (block or ev.query()):
                if block:
                    ev.synchronize()
                self._raise_if_set(st, k)


_overflow = _DeferredOverflow()


class _BatchAll(torch.autograd.Function):
    """When a gradient can be asked for, forward runs the fused pass (en_batch_all_fwd_bwd: distance tiles computed
    once for loss and gradient) and backward only applies the upstream gradient."""

    @staticmethod
    def forward(ctx, emb, labels, margin, squared, max_pos):
        B, d = emb.shape
        dev = emb.device
        lib = _lib.load()
        ws = workspace(lib.en_ws_bytes_batch_all(B, d, max_pos), dev, "batch_all")
        out = torch.empty(2, dtype=torch.float32, device=dev)
        stats = torch.empty(3, dtype=torch.float64, device=dev)
        if ctx.needs_input_grad[0]:
            gemb = torch.empty_like(emb)
            ones = torch.ones(1, dtype=torch.float32, device=dev)
            k, flag = _overflow.slot(dev, max_pos)
            _lib.call("en_batch_all_fwd_bwd", ptr(emb), ptr(labels), B, d, ctypes.c_float(margin), int(squared),
                      max_pos, ptr(out), ptr(stats), ptr(ones), ptr(gemb), ptr(flag), ptr(ws), ws.numel(),
                      stream_ptr())
            _overflow.arm(k)
            ctx.gemb = gemb
        else:
            _lib.call("en_batch_all_fwd", ptr(emb), ptr(labels), B, d, ctypes.c_float(margin), int(squared), max_pos,
                      ptr(out), ptr(stats), ptr(ws), ws.numel(), stream_ptr())
        loss, frac = out[0].clone(), out[1].clone()
        ctx.mark_non_differentiable(frac)
        return loss, frac

    @staticmethod
    def backward(ctx, g, _gfrac):
        return _consume_stored_gradient(ctx, g, "batch_all_triplet_loss"), None, None, None, None


def batch_all_triplet_loss(margin=0.5, squared=False, max_positives=None, return_fraction=False):
    """Batch-all triplet loss (Moindrot): mean of the positive hinge terms over all valid (anchor, positive,
    negative) triplets.  ``max_positives`` = largest class size minus one (at most 64); when None it is read from
    the labels (one small device->host sync per call)."""

    def loss_function(y_true, y_pred):
        emb = as_cuda_f32(y_pred)
        labels = _labels(y_true, emb.device)
        if emb.dim() != 2 or labels.numel() != emb.shape[0]:
            raise ValueError("batch_all_triplet_loss: y_pred must be (B, d) with one label per row")
        mp = max_positives
        if mp is None:
            mp = int(torch.unique(labels, return_counts=True)[1].max().item()) - 1
        mp = max(int(mp), 1)
        loss, frac = _BatchAll.apply(emb, labels, float(margin), bool(squared), mp)
        loss_function.last_fraction = frac
        return (loss, frac) if return_fraction else loss

    loss_function.last_fraction = None
    loss_function.check = lambda: _overflow.poll(block=True)  # settles the deferred max_positives checks
    return loss_function


# ------------------------------------------------------------------------------------------------ all-pairs contrastive
class _ContrastiveAllPairs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, labels):
        B, d = emb.shape
        lib = _lib.load()
        ws = workspace(lib.en_ws_bytes_contrastive_allpairs(B, d), emb.device, "contrastive_all")
        loss = torch.empty((), dtype=torch.float32, device=emb.device)
        if ctx.needs_input_grad[0]:
            gemb = torch.empty_like(emb)
            _lib.call("en_contrastive_allpairs_fwd_bwd", ptr(emb), ptr(labels), B, d, ptr(loss), None, ptr(gemb),
                      ptr(ws), ws.numel(), stream_ptr())
            ctx.gemb = gemb
        else:
            _lib.call("en_contrastive_allpairs_fwd", ptr(emb), ptr(labels), B, d, ptr(loss), ptr(ws), ws.numel(),
                      stream_ptr())
        return loss

    @staticmethod
    def backward(ctx, g):
        return _consume_stored_gradient(ctx, g, "contrastive_loss_all_pairs"), None


def contrastive_loss_all_pairs():
    """``contrastive_loss`` (lac:4-11) over every ordered pair of the batch with ``y = [label_i == label_j]`` and
    the Siamese head's distance ``sqrt(max(|e_i - e_j|^2, 1e-7))`` (models.py:225), fused into the distance GEMM."""

    def loss_function(y_true, y_pred):
        emb = as_cuda_f32(y_pred)
        labels = _labels(y_true, emb.device)
        if emb.dim() != 2 or labels.numel() != emb.shape[0]:
            raise ValueError("contrastive_loss_all_pairs: y_pred must be (B, d) with one label per row")
        return _ContrastiveAllPairs.apply(emb, labels)

    return loss_function
