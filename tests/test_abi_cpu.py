"""CPU tests of the C-ABI boundary: the library builds for sm_100a without a GPU, loads, and exports exactly the
symbols include/embeddingnet_b200.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "embeddingnet_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(en_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = header_symbols()
    assert len(syms) >= 35
    for must in ("en_batch_hard_fwd", "en_batch_hard_bwd", "en_knn_shard_topk", "en_knn_merge", "en_pairwise_dist",
                 "en_mine_batch_scan", "en_mine_batch_select", "en_l2_normalize_fwd", "en_triplet_apn_fwd"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for name in header_symbols():
        assert hasattr(lib, name), "declared in the header but not exported: " + name


def test_ctypes_signatures_cover_header(lib_built):
    from embeddingnet_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert b"sm_100a" in lib.en_version()


def test_argument_errors_are_reported_without_a_gpu(lib_built):
    """Host-side validation happens before any launch, so it is observable on a CPU-only box."""
    from embeddingnet_b200 import _lib

    lib = _lib.load()
    rc = lib.en_triplet_apn_fwd(ctypes.c_void_p(16), 4, 10, ctypes.c_float(0.5), ctypes.c_void_p(16), None)
    assert rc == -1 and b"multiple of 3" in lib.en_last_error()
    rc = lib.en_l2_normalize_fwd(None, None, 4, 8, None)
    assert rc == -1
    assert lib.en_ws_bytes_batch_hard(4096, 512) > 2 * 4096 * 512 * 4
    assert lib.en_ws_bytes_knn(100, 1000, 64, 40) == 0  # k above EN_KNN_MAX_K
    assert lib.en_bank_dpad(100, _lib.EN_PREC_TF32X3) == 128
    assert lib.en_bank_dpad(100, _lib.EN_PREC_BF16X3) == 128 and lib.en_bank_dpad(33, _lib.EN_PREC_BF16X3) == 64
    assert lib.en_bank_plane_bytes(10, 100, _lib.EN_PREC_BF16X3) == 10 * 128 * 2
    assert lib.en_ws_bytes_knn_exact(65, 1000, 64, 5) == 0  # Q above EN_KNN_EXACT_MAX_Q
    rc2 = lib.en_knn_exact_topk(ctypes.c_void_p(256), 65, 64, ctypes.c_void_p(256), 1000, 0, 5, None, None,
                                ctypes.c_void_p(256), ctypes.c_void_p(256), ctypes.c_void_p(256), 1 << 20, None)
    assert rc2 == -1 and b"at most 64" in lib.en_last_error()
    # host-buffer pipeline: sizes and argument checks happen before anything touches a device
    assert lib.en_bh_host_pipe_device_bytes(4096, 512, 3) >= 3 * (2 * 4096 * 512 * 4 + lib.en_ws_bytes_batch_hard(4096, 512))
    assert lib.en_bh_host_pipe_device_bytes(4096, 512, 9) == 0 and lib.en_bh_host_pipe_device_bytes(0, 512, 3) == 0
    pipe = ctypes.c_void_p(0)
    rc3 = lib.en_bh_host_pipe_create(4096, 512, ctypes.c_float(0.5), 0, 0, 0, ctypes.c_void_p(256), 1 << 30,
                                     ctypes.byref(pipe))
    assert rc3 == -1 and b"depth" in lib.en_last_error() and not pipe.value
    assert lib.en_bh_host_pipe_submit(None, None, None, None, None, None, None, None) == -1
    assert lib.en_bh_host_pipe_destroy(None) == 0
    with pytest.raises(ValueError):
        _lib.check(rc, "en_l2_normalize_fwd")


def test_sass_contains_blackwell_tensor_and_tma_instructions(lib_built):
    """The distance GEMM must really be tcgen05 + TMA (SASS: UTC*MMA, UTMALDG, LDTM), not a legacy mma path."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_built], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass or re.search(r"UTC\w*MMA", sass)
    assert "UTMALDG" in sass
    assert "LDTM" in sass
    assert "HMMA.16" not in sass  # no mma.sync fallback
