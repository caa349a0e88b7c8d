"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Loads the reference's own source files, UNMODIFIED, from ``/root/reference`` under a ``sys.modules`` shim so that
``embedding_net.losses_and_accuracies`` and ``embedding_net.datagenerators`` execute without TensorFlow 2.2 /
matplotlib / albumentations (none of which can be installed here: Python 3.12, no network).

The shim provides ``tensorflow.keras.backend`` as a torch-CPU (float32) restatement of the handful of backend ops
the reference calls (``lac:9-11,34-41,50``; ``models:218,225``; ``bb:38``).  scikit-learn is the real library.

This module only works where ``/root/reference`` exists (the authoring container).  It is used by
``tests/golden/make_golden.py`` to generate the committed golden fixtures and by the CPU tests (skipped when the
reference tree is absent, e.g. on the GPU box) to validate ``oracle/np_oracle.py``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "embedding_net", "losses_and_accuracies.py"))


class _Shape(tuple):
    """``y_pred.shape.as_list()`` is called at lac:27."""

    def as_list(self):
        return list(self)


def _make_backend():
    import torch

    K = types.ModuleType("tensorflow.keras.backend")

    class KTensor(torch.Tensor):
        @property
        def shape(self):  # type: ignore[override]
            return _Shape(super().shape)

    def wrap(t):
        return t.as_subclass(KTensor) if isinstance(t, torch.Tensor) else t

    def T(x):
        if isinstance(x, torch.Tensor):
            return x
        return torch.as_tensor(x)

    K.KTensor = KTensor
    K.wrap = wrap
    K.square = lambda x: wrap(T(x) * T(x))
    # TF's maximum(x, 0) routes the gradient at x == 0 to x (x >= y mask); clamp(min=) matches that.
    K.maximum = lambda x, y: wrap(torch.clamp(T(x), min=float(y)) if not isinstance(y, torch.Tensor) else torch.maximum(T(x), y))

    def _mean(x, axis=None, keepdims=False):
        x = T(x)
        if x.dtype == torch.bool:  # Keras casts bool to floatx before the mean
            x = x.to(torch.float32)
        return wrap(x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims))

    K.mean = _mean
    K.sum = lambda x, axis=None, keepdims=False: wrap(T(x).sum() if axis is None else T(x).sum(dim=axis, keepdim=keepdims))
    K.equal = lambda a, b: wrap(T(a) == T(b))
    K.cast = lambda x, dtype: wrap(T(x).to(dtype if isinstance(dtype, torch.dtype) else getattr(torch, str(dtype))))
    K.abs = lambda x: wrap(T(x).abs())
    K.sqrt = lambda x: wrap(torch.sqrt(T(x)))
    K.epsilon = lambda: 1e-7

    def _l2_normalize(x, axis=None):
        x = T(x)
        ss = (x * x).sum(dim=axis, keepdim=True)
        return wrap(x * torch.rsqrt(torch.clamp(ss, min=1e-12)))

    K.l2_normalize = _l2_normalize
    return K


_LOADED = {}


def load_reference():
    """Returns (losses_and_accuracies, datagenerators, K) loaded from the reference tree."""
    if _LOADED:
        return _LOADED["lac"], _LOADED["dg"], _LOADED["K"]
    if not reference_available():
        raise RuntimeError("reference tree not available at %s" % REFERENCE_ROOT)
    K = _make_backend()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, n):
            return _Anything()

    stubs = {
        "tensorflow": mod("tensorflow"),
        "tensorflow.keras": mod("tensorflow.keras", backend=K),
        "tensorflow.keras.backend": K,
        "tensorflow.keras.utils": mod("tensorflow.keras.utils", Sequence=object),
        "tensorflow.keras.optimizers": mod("tensorflow.keras.optimizers"),
        "matplotlib": mod("matplotlib"),
        "matplotlib.pyplot": mod("matplotlib.pyplot"),
        "albumentations": mod("albumentations"),
        "cv2": sys.modules.get("cv2") or mod("cv2"),
        "tqdm": sys.modules.get("tqdm") or mod("tqdm"),
        "plotly": mod("plotly"),
        "plotly.express": mod("plotly.express"),
        "keras_radam": mod("keras_radam", RAdam=_Anything),
    }
    stubs["tensorflow"].keras = stubs["tensorflow.keras"]
    stubs["tensorflow.keras"].utils = stubs["tensorflow.keras.utils"]
    stubs["tensorflow.keras"].optimizers = stubs["tensorflow.keras.optimizers"]
    saved = {k: sys.modules.get(k) for k in stubs}
    saved_path = list(sys.path)
    saved_en = {k: v for k, v in sys.modules.items() if k == "embedding_net" or k.startswith("embedding_net.")}
    try:
        for k in saved_en:
            del sys.modules[k]
        sys.modules.update(stubs)
        try:
            import cv2  # noqa: F401  (real one if present)
        except Exception:
            pass
        # Build a package object by hand so the reference's embedding_net/__init__.py (which imports the
        # TF-heavy model code) is not executed; only the two hot-path modules are loaded from source.
        pkg = types.ModuleType("embedding_net")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "embedding_net")]
        sys.modules["embedding_net"] = pkg
        lac = importlib.import_module("embedding_net.losses_and_accuracies")
        dg = importlib.import_module("embedding_net.datagenerators")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k == "embedding_net" or k.startswith("embedding_net.")]:
            del sys.modules[k]
        sys.modules.update(saved_en)
        sys.path[:] = saved_path
    _LOADED.update(lac=lac, dg=dg, K=K)
    return lac, dg, K


class FakeEmbeddingModel:
    """Stands in for the Keras ``base_model``: ``predict(images)`` returns precomputed embeddings.

    The 'images' handed out by ``_get_images_set`` (overridden below) are (k, 1) arrays of global row ids."""

    def __init__(self, table):
        self.table = table

    def predict(self, images):
        import numpy as np

        return self.table[np.asarray(images).reshape(-1).astype(np.int64)]


def make_reference_generator(embeddings_by_class, k_classes, k_samples, margin, mode):
    """Instantiate the reference ``TripletsDataGenerator`` (dg:159-261) over a table of precomputed embeddings.

    embeddings_by_class: list of (n_c, d) float32 arrays.  'Images' are row ids into the stacked table, so the
    triplets returned by ``get_batch_triplets_mining`` are row-id triples."""
    import numpy as np

    _, dg, _ = load_reference()
    class_names = ["c%04d" % i for i in range(len(embeddings_by_class))]
    offsets = np.cumsum([0] + [len(e) for e in embeddings_by_class])
    table = np.vstack(embeddings_by_class).astype(np.float32)
    class_files = {n: ["%d" % (offsets[i] + j) for j in range(len(embeddings_by_class[i]))] for i, n in enumerate(class_names)}

    cls = dg.TripletsDataGenerator

    class Gen(cls):  # only the image loader is replaced; mining code is the reference's
        def _get_images_set(self, clss, idxs, with_aug=True):
            return np.array([[int(self.class_files_paths[clss][i])] for i in idxs], dtype=np.int64)

    g = Gen(
        embedding_model=FakeEmbeddingModel(table),
        class_files_paths=class_files,
        class_names=class_names,
        n_batches=1,
        input_shape=None,
        batch_size=1,
        augmentations=None,
        k_classes=k_classes,
        k_samples=k_samples,
        margin=margin,
        negatives_selection_mode=mode,
    )
    return g, table
