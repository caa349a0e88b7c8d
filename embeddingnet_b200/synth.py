"""Deterministic synthetic embeddings, bit-identical on CPU (NumPy) and GPU (``en_synth_fill``), SURVEY.md 8(d).

u(r, c, seed) = ((splitmix64(seed ^ (r * 2654435761 + c)) >> 40) * 2^-23) - 1  in [-1, 1), exactly representable in
float32; x = relu?(centre[label(r)] + noise * u) with one float32 multiply and one add (no transcendentals).
"""
from __future__ import annotations

import ctypes

import numpy as np

SEED_CENTRE, SEED_NOISE, SEED_QUERY = 1234, 5678, 91011
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def hash_u(rows, cols, seed):
    """rows (R,), cols (C,) uint64 -> (R, C) float32 in [-1, 1)."""
    with np.errstate(over="ignore"):
        key = np.uint64(seed) ^ (rows[:, None].astype(np.uint64) * np.uint64(2654435761) + cols[None, :].astype(np.uint64))
        h = _splitmix64(key)
    return ((h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)).astype(np.float32)


def labels_for(rows, row_offset=0, n_classes=0, rows_per_class=0):
    R = np.arange(rows, dtype=np.int64) + row_offset
    if n_classes <= 0:
        return np.zeros(rows, np.int32)
    return ((R // rows_per_class) % n_classes if rows_per_class > 0 else R % n_classes).astype(np.int32)


def make_numpy(rows, d, row_offset=0, seed_centre=SEED_CENTRE, seed_noise=SEED_NOISE, n_classes=0, rows_per_class=0,
               noise=0.5, relu=False):
    """Host generator; returns (x (rows, d) float32, labels (rows,) int32)."""
    R = (np.arange(rows, dtype=np.int64) + row_offset).astype(np.uint64)
    C = np.arange(d, dtype=np.uint64)
    labels = labels_for(rows, row_offset, n_classes, rows_per_class)
    if n_classes > 0:
        centre = hash_u(labels.astype(np.uint64), C, seed_centre)
        x = centre + np.float32(noise) * hash_u(R, C, seed_noise)
    else:
        x = hash_u(R, C, seed_noise)
    if relu:
        x = np.maximum(x, np.float32(0))
    return x.astype(np.float32), labels


def make_device(rows, d, row_offset=0, seed_centre=SEED_CENTRE, seed_noise=SEED_NOISE, n_classes=0,
                rows_per_class=0, noise=0.5, relu=False, device=None, out=None):
    """Device generator (same bits as ``make_numpy``); returns (x CUDA float32, labels CUDA int32)."""
    import torch

    from . import _lib
    from ._runtime import ptr, require_cuda, stream_ptr

    dev = device or require_cuda()
    x = out if out is not None else torch.empty((rows, d), dtype=torch.float32, device=dev)
    labels = torch.zeros(rows, dtype=torch.int32, device=dev)
    _lib.call("en_synth_fill", ptr(x), rows, d, row_offset, ctypes.c_uint64(seed_centre), ctypes.c_uint64(seed_noise),
              n_classes, rows_per_class, ctypes.c_float(noise), int(relu), ptr(labels), stream_ptr())
    return x, labels
