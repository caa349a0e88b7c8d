"""embeddingnet_b200 -- B200-native (sm_100a) implementation of EmbeddingNet's distance / mining / loss / bank-kNN
hot path behind the reference's own Python API.  See DESIGN.md and include/embeddingnet_b200.h.

Importing the package does not load CUDA; the first call into any kernel loads ``libembeddingnet_b200.so`` and
raises if it is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
