"""Host-side plumbing shared by the drop-in modules: device selection, pointers, streams, scratch buffers.

PyTorch is used for device memory, streams and (in ``distributed.py``) the NCCL process group -- nothing else.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np
import torch

from . import _lib

_tls = threading.local()


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.EmbeddingNetB200Error(
            "embeddingnet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback."
        )
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes: int, device: torch.device, slot: str = "default") -> torch.Tensor:
    """Per-thread, per-device, per-slot scratch that only grows (mining runs on a Keras enqueuer thread beside the
    training thread, SURVEY 3.1, so scratch must not be shared across threads)."""
    if not hasattr(_tls, "ws"):
        _tls.ws = {}
    key = (device.index, slot)
    buf = _tls.ws.get(key)
    nbytes = max(int(nbytes), 256)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _tls.ws[key] = buf
    return buf


def as_cuda_f32(x, device=None) -> torch.Tensor:
    """numpy / torch (any device) -> contiguous float32 CUDA tensor."""
    device = device or require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def as_cuda_i32(x, device=None) -> torch.Tensor:
    device = device or require_cuda()
    if isinstance(x, torch.Tensor):
        t = x.reshape(-1)
        if t.dtype != torch.int32:
            t = t.to(torch.int32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x).reshape(-1).astype(np.int32)))
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def launch_count() -> int:
    return int(_lib.load().en_launch_count())


def launch_count_reset() -> None:
    _lib.load().en_launch_count_reset()
