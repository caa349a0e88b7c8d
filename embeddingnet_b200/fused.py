"""Autograd-free fused step objects (CUDA-graph capturable): loss AND gradient from one pair of C-ABI calls.

``losses_and_accuracies.py`` is the reference-shaped API (``fn(y_true, y_pred)`` + ``loss.backward()``); training
loops that own their buffers can use these to skip the autograd bookkeeping and to replay the step as a CUDA graph.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._runtime import ptr, stream_ptr


class BatchHardStep:
    """Batch-hard triplet loss + gradient for fixed (B, d); all buffers preallocated, no host sync.

    step(emb, labels) -> (loss 0-dim tensor, grad (B, d) tensor), both owned by this object."""

    def __init__(self, B, d, margin=0.5, squared=False, soft=False, device=None):
        dev = device or torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        self.B, self.d = int(B), int(d)
        self.margin, self.squared, self.soft = float(margin), int(bool(squared)), int(bool(soft))
        self.ws = torch.empty(max(lib.en_ws_bytes_batch_hard(B, d), 256), dtype=torch.uint8, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.saved_i = torch.empty((2, B), dtype=torch.int32, device=dev)
        self.saved_f = torch.empty((3, B), dtype=torch.float32, device=dev)
        self.grad = torch.empty((B, d), dtype=torch.float32, device=dev)
        self.gloss = torch.ones(1, dtype=torch.float32, device=dev)

    def step(self, emb, labels):
        B, d = self.B, self.d
        assert emb.shape == (B, d) and emb.dtype == torch.float32 and emb.is_contiguous() and emb.is_cuda
        assert labels.dtype == torch.int32 and labels.numel() == B and labels.is_cuda
        s = stream_ptr()
        si, sf = self.saved_i, self.saved_f
        _lib.call("en_batch_hard_fwd_bwd", ptr(emb), ptr(labels), B, d, ctypes.c_float(self.margin), self.squared,
                  self.soft, ptr(self.loss), ptr(si[0]), ptr(si[1]), ptr(sf[0]), ptr(sf[1]), ptr(sf[2]),
                  ptr(self.gloss), ptr(self.grad), ptr(self.ws), self.ws.numel(), s)
        return self.loss, self.grad


class BatchHardHostPipeline:
    """Batch-hard triplet loss + gradient for batches that live in HOST memory (pinned torch tensors or NumPy arrays),
    pipelined inside the library (``en_bh_host_pipe_*``): upload, kernels and download of consecutive steps overlap.

    The reference computes its loss on whatever ``y_pred`` Keras hands it
    (embedding_net/losses_and_accuracies.py:26-42); this is the same call for a loop that owns host buffers::

        pipe = BatchHardHostPipeline(B, d, margin=0.5, depth=5)
        t = pipe.submit(emb_h, labels_h, loss_h, grad_h)     # queues the upload, advances the steps in flight
        ...                                                  # submit more steps (up to `depth` in flight)
        pipe.wait(t)                                         # loss_h / grad_h are now filled

    Read results two or more steps behind the newest submit and the loop never stalls (``bench.py``: four behind).

    ``emb_h`` (B, d) float32, ``labels_h`` (B,) int32, ``loss_h`` 1 float32, ``grad_h`` (B, d) float32; buffers must
    stay untouched until ``wait`` returns.  Use ``pinned(...)`` for page-locked buffers (pageable memory works but
    serialises the copies)."""

    def __init__(self, B, d, margin=0.5, squared=False, soft=False, depth=3, device=None):
        dev = device or torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        self.B, self.d, self.depth = int(B), int(d), int(depth)
        nbytes = lib.en_bh_host_pipe_device_bytes(self.B, self.d, self.depth)
        if nbytes == 0:
            raise ValueError("BatchHardHostPipeline: bad shape or depth (B=%d d=%d depth=%d)" % (B, d, depth))
        self._mem = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self._pipe = ctypes.c_void_p(0)
        with torch.cuda.device(dev):
            _lib.call("en_bh_host_pipe_create", self.B, self.d, ctypes.c_float(float(margin)), int(bool(squared)),
                      int(bool(soft)), self.depth, ptr(self._mem), nbytes, ctypes.byref(self._pipe))
        self._submit = lib.en_bh_host_pipe_submit
        self._wait = lib.en_bh_host_pipe_wait
        self._ticket = ctypes.c_int64(0)
        self._alive = {}  # ticket -> the caller's buffers (kept referenced while the copies are in flight)

    @staticmethod
    def pinned(shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype).pin_memory()

    @staticmethod
    def _addr(x, dtype, n):
        if isinstance(x, torch.Tensor):
            assert x.device.type == "cpu" and x.dtype == dtype and x.is_contiguous() and x.numel() == n, \
                "host buffer: contiguous CPU tensor of %s with %d elements expected" % (dtype, n)
            return x.data_ptr()
        want = {torch.float32: "float32", torch.int32: "int32"}[dtype]
        assert x.dtype == want and x.flags["C_CONTIGUOUS"] and x.size == n, \
            "host buffer: C-contiguous NumPy array of %s with %d elements expected" % (want, n)
        return x.ctypes.data

    def submit(self, emb, labels, loss_out, grad_out, hp_idx_out=None, hn_idx_out=None):
        n = self.B * self.d
        a = self._addr
        rc = self._submit(self._pipe, a(emb, torch.float32, n), a(labels, torch.int32, self.B),
                          a(loss_out, torch.float32, 1), a(grad_out, torch.float32, n),
                          a(hp_idx_out, torch.int32, self.B) if hp_idx_out is not None else None,
                          a(hn_idx_out, torch.int32, self.B) if hn_idx_out is not None else None,
                          ctypes.byref(self._ticket))
        if rc:
            _lib.check(rc, "en_bh_host_pipe_submit")
        t = self._ticket.value
        self._alive[t] = (emb, labels, loss_out, grad_out, hp_idx_out, hn_idx_out)
        self._alive.pop(t - self.depth, None)  # that step was waited for inside submit
        return t

    def wait(self, ticket):
        rc = self._wait(self._pipe, ticket)
        if rc:
            _lib.check(rc, "en_bh_host_pipe_wait")
        self._alive.pop(ticket, None)

    def close(self):
        if self._pipe:
            pipe, self._pipe = self._pipe, ctypes.c_void_p(0)
            _lib.call("en_bh_host_pipe_destroy", pipe)
            self._alive.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
