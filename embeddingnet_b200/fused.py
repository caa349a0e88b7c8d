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
