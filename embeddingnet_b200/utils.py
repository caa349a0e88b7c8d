"""Bank (de)serialisation: the reference pickle layout plus a sharded raw format for direct per-GPU loading.

Reference: ``utils.load_encodings`` (/root/reference/embedding_net/utils.py:29-33) and
``EmbeddingNet.save_encodings`` (embedding_net/models.py:86-90) pickle a dict
``{'paths': [str]*N, 'labels': [obj]*N, 'encodings': ndarray (N, d) float32}``.
"""
from __future__ import annotations

import json
import os
import pickle

import numpy as np


def load_encodings(path_to_encodings):
    """utils.py:29-33."""
    with open(path_to_encodings, "rb") as f:
        encodings = pickle.load(f)
    return encodings


def save_encodings(encoded_training_data, save_folder="./", save_file_name="encodings.pkl"):
    """models.py:86-90 as a free function."""
    data = {k: v for k, v in encoded_training_data.items() if k not in ("knn_classifier", "_nn1")}
    with open(os.path.join(save_folder, save_file_name), "wb") as f:
        pickle.dump(data, f)


def save_encodings_sharded(encoded_training_data, folder, n_shards):
    """A 20 GB bank is impractical as one pickle: raw float32 row shards + int32 label ids + a small
    JSON index, so that rank r maps only ``shard_r.f32``.  Row order and the label list are the reference's."""
    os.makedirs(folder, exist_ok=True)
    enc = np.ascontiguousarray(np.asarray(encoded_training_data["encodings"], np.float32))
    if enc.ndim == 1:
        enc = enc.reshape(1, -1)
    n, d = enc.shape
    classes, ids = np.unique(np.asarray(encoded_training_data["labels"]), return_inverse=True)
    per = (n + n_shards - 1) // n_shards
    shards = []
    for r in range(n_shards):
        lo, hi = min(r * per, n), min((r + 1) * per, n)
        name = "shard_%03d.f32" % r
        enc[lo:hi].tofile(os.path.join(folder, name))
        shards.append({"file": name, "row_begin": int(lo), "row_end": int(hi)})
    ids.astype(np.int32).tofile(os.path.join(folder, "label_ids.i32"))
    with open(os.path.join(folder, "index.json"), "w") as f:
        json.dump({"n": int(n), "d": int(d), "shards": shards, "classes": [str(c) for c in classes.tolist()],
                   "paths": list(encoded_training_data.get("paths", []))}, f)


def load_encodings_shard(folder, rank):
    """Returns (rows (n_r, d) float32 memmap, label_ids_all (N,) int32, row_begin, N, classes)."""
    with open(os.path.join(folder, "index.json")) as f:
        idx = json.load(f)
    sh = idx["shards"][rank]
    rows = np.memmap(os.path.join(folder, sh["file"]), dtype=np.float32, mode="r",
                     shape=(sh["row_end"] - sh["row_begin"], idx["d"]))
    ids = np.fromfile(os.path.join(folder, "label_ids.i32"), dtype=np.int32)
    return rows, ids, sh["row_begin"], idx["n"], np.asarray(idx["classes"])
