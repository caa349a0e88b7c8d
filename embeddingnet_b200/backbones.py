"""Drop-in for the *head* of ``embedding_net/backbones.py`` (RocketFlash/EmbeddingNet).

The reference ends every backbone with ``Dense(encodings_len // 2, relu) -> Dense(encodings_len, relu) ->
Lambda(K.l2_normalize(axis=1))`` (/root/reference/embedding_net/backbones.py:114-119; the simple backbones end the same
way, :36-38 and :75-77).  The convolutional trunk is cuDNN's job and out of scope; this module is the step between the
pooled trunk features and the distance / mining / kNN path: each Dense is one tcgen05 GEMM whose epilogue applies
bias + ReLU and, for the last layer, the row normalisation (``csrc/head.cu``), so embeddings are born normalised on
the device and can go straight into ``TripletsDataGenerator`` / ``BankKNNClassifier`` without touching the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._runtime import as_cuda_f32, ptr, require_cuda, stream_ptr, workspace


class _DenseReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, bias, normalize):
        lib = _lib.load()
        dev = x.device
        B, n_in = x.shape
        units = kernel.shape[1]
        nbytes = lib.en_dense_plane_bytes(n_in, units)
        w_hi = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        w_lo = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        _lib.call("en_dense_prepare", ptr(kernel), n_in, units, ptr(w_hi), ptr(w_lo), stream_ptr())
        out = torch.empty((B, units), dtype=torch.float32, device=dev)
        inv = torch.empty(B, dtype=torch.float32, device=dev) if normalize else None
        ws = workspace(lib.en_ws_bytes_dense(B, n_in), dev, "dense")
        _lib.call("en_dense_relu_fwd", ptr(x), B, n_in, ptr(w_hi), ptr(w_lo), ptr(bias), units, int(normalize),
                  ptr(out), ptr(inv), ptr(ws), ws.numel(), stream_ptr())
        ctx.save_for_backward(x, kernel, out, inv if inv is not None else torch.empty(0, device=dev))
        ctx.normalize = bool(normalize)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, gy):
        x, kernel, out, inv = ctx.saved_tensors
        lib = _lib.load()
        dev = x.device
        B, n_in = x.shape
        units = kernel.shape[1]
        gy = gy.contiguous().to(torch.float32)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(kernel) if ctx.needs_input_grad[1] else None
        gb = torch.empty(units, dtype=torch.float32, device=dev) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        ws = workspace(lib.en_ws_bytes_dense_bwd(B, n_in, units), dev, "dense_bwd")
        _lib.call("en_dense_relu_bwd", ptr(x), B, n_in, ptr(kernel), units, int(ctx.normalize), ptr(out),
                  ptr(inv) if ctx.normalize else None, ptr(gy), ptr(gx), ptr(gw), ptr(gb), ptr(ws), ws.numel(),
                  stream_ptr())
        return gx, gw, gb, None


class DenseReLU:
    """``Dense(units, activation="relu")`` with an optional fused ``K.l2_normalize``.

    kernel: (n_in, units) -- the Keras layout; bias: (units,) or None.  ``trainable=True`` keeps kernel / bias as
    leaf tensors with ``requires_grad`` (see ``parameters()``) and routes calls through autograd: the backward pass
    is two more tcgen05 GEMMs (``en_dense_relu_bwd``), so the reference's training direction (train.py:172) runs
    through the same kernels as bank building."""

    def __init__(self, kernel, bias=None, normalize=False, device=None, trainable=False):
        dev = device or require_cuda()
        lib = _lib.load()
        k = as_cuda_f32(kernel, dev)
        if k.dim() != 2:
            raise ValueError("DenseReLU: kernel must be (n_in, units)")
        self.n_in, self.units = int(k.shape[0]), int(k.shape[1])
        self.normalize = bool(normalize)
        self.device = dev
        self.trainable = bool(trainable)
        self.kernel = k.detach().clone().requires_grad_(self.trainable)
        self.bias = as_cuda_f32(bias, dev).reshape(-1).detach().clone() if bias is not None else None
        if self.bias is not None and self.bias.numel() != self.units:
            raise ValueError("DenseReLU: bias must have %d entries" % self.units)
        if self.bias is not None:
            self.bias.requires_grad_(self.trainable)
        nbytes = lib.en_dense_plane_bytes(self.n_in, self.units)
        self._hi = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        self._lo = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        _lib.call("en_dense_prepare", ptr(k), self.n_in, self.units, ptr(self._hi), ptr(self._lo), stream_ptr())

    def parameters(self):
        return [t for t in (self.kernel, self.bias) if t is not None]

    def __call__(self, x):
        lib = _lib.load()
        xt = as_cuda_f32(x, self.device)
        if xt.dim() != 2 or xt.shape[1] != self.n_in:
            raise ValueError("DenseReLU: expected (B, %d) input, got %s" % (self.n_in, tuple(xt.shape)))
        B = xt.shape[0]
        if B > 0 and torch.is_grad_enabled() and (self.trainable or xt.requires_grad):
            return _DenseReLUFn.apply(xt, self.kernel, self.bias, self.normalize)
        out = torch.empty((B, self.units), dtype=torch.float32, device=self.device)
        if B == 0:
            return out
        if self.trainable:  # weights may have been updated since the planes were built
            _lib.call("en_dense_prepare", ptr(self.kernel), self.n_in, self.units, ptr(self._hi), ptr(self._lo),
                      stream_ptr())
        ws = workspace(lib.en_ws_bytes_dense(B, self.n_in), self.device, "dense")
        _lib.call("en_dense_relu_fwd", ptr(xt), B, self.n_in, ptr(self._hi), ptr(self._lo), ptr(self.bias), self.units,
                  int(self.normalize), ptr(out), None, ptr(ws), ws.numel(), stream_ptr())
        return out


class EmbeddingHead:
    """``Dense(d // 2, relu) -> Dense(d, relu) [-> l2_normalize]`` on pooled trunk features (backbones.py:114-119).

    ``predict(features)`` mirrors ``base_model.predict``: NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out
    (which keeps ``TripletsDataGenerator``'s device-resident path on the device)."""

    def __init__(self, kernel1, bias1, kernel2, bias2, embeddings_normalization=True, device=None, trainable=False):
        self.fc1 = DenseReLU(kernel1, bias1, normalize=False, device=device, trainable=trainable)
        self.fc2 = DenseReLU(kernel2, bias2, normalize=embeddings_normalization, device=device, trainable=trainable)

    def parameters(self):
        return self.fc1.parameters() + self.fc2.parameters()

    def __call__(self, features):
        """Differentiable forward on CUDA tensors (training: features -> embeddings -> loss -> ``backward()``)."""
        return self.fc2(self.fc1(features))

    def predict(self, features):
        on_device = isinstance(features, torch.Tensor) and features.is_cuda
        y = self.fc2(self.fc1(features))
        return y if on_device else y.cpu().numpy()
