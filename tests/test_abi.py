"""The C-ABI library builds, loads and exports every symbol include/latentknn.h declares.
No compute calls: this runs on a box without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from latent_rag_b200 import build, _native

    build.build(verbose=False)
    return _native.load()


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "latentknn.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 19
    for name in names:
        assert hasattr(lib, name), f"{name} declared in latentknn.h but not exported"


def test_python_binding_covers_the_header(lib):
    from latent_rag_b200 import _native

    assert sorted(n for n, _, _ in _native.SYMBOLS) == _declared_symbols()


def test_python_constants_match_the_header():
    from latent_rag_b200 import _native

    with open(os.path.join(ROOT, "include", "latentknn.h")) as f:
        defines = dict(re.findall(r"#define\s+(LK_[A-Z_]+)\s+(\d+)", f.read()))
    for name in ("LK_MAX_K", "LK_MAX_K_FUSED", "LK_MAX_WORLD", "LK_IPC_HANDLE_BYTES"):
        assert int(defines[name]) == getattr(_native, name), name
    assert _native.load().lk_abi_version() == int(defines["LK_ABI_VERSION"])


def test_abi_version_and_error_string(lib):
    from latent_rag_b200 import _native

    assert lib.lk_abi_version() == _native.LK_ABI_VERSION == 2
    assert isinstance(lib.lk_last_error(), bytes)


def test_bad_arguments_are_rejected_without_a_device(lib):
    h = ctypes.c_void_p()
    assert lib.lk_index_create(ctypes.byref(h), 0, 0, 16, 0, 1, None) == -1  # capacity 0
    assert lib.lk_index_create(ctypes.byref(h), 0, 10, 16, 7, 1, None) == -1  # bad metric
    assert b"Unsupported metric" in lib.lk_last_error()
    assert lib.lk_index_create(ctypes.byref(h), 0, 10, 16, 2, 1, None) == -1  # mahalanobis without L
    assert lib.lk_index_search(None, None, 0, 0, 1, 5, None, None, 0, 0, 0, None) == -1
    assert lib.lk_merge_topk(0, None, None, 1, 0, 1, 1, None, None, 0, None) == -1
    assert lib.lk_index_search(None, None, 0, 0, 1, 5000, None, None, 0, 0, 0, None) == -1  # k past LK_MAX_K
    # a truncated persistence image is rejected by its length, never read: argument checks come first
    assert lib.lk_index_import(None, None, 0, None, 0, 5, None) == -1
    assert lib.lk_index_reserve(None, 10, None) == -1
    assert lib.lk_index_export(None, None, None, None) == -1
    assert lib.lk_bert_create(ctypes.byref(h), 0, 100, 64, 128, 4, 256, 2, 1e-12, None) == -1  # no weights
    assert lib.lk_bert_encode(None, None, None, 0, 1, 8, 1, None, 0, None) == -1
    assert lib.lk_linear_forward(0, None, 4, 64, None, 128, None, None, 0, 0, None) == -1


def test_sass_contains_blackwell_instructions():
    """The tensor-core search kernel must really be tcgen05 + TMA code."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from latent_rag_b200 import _native

    sass = subprocess.run([cuobjdump, "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "no tcgen05.mma in the library"
    assert "LDTM" in sass, "no tcgen05.ld in the library"
    assert "UBLKCP" in sass, "no bulk async copy (TMA) in the library"
    assert "sm_100a" in sass or "sm_100" in sass
    # the sentence encoder's linear layers are tcgen05 kernels too
    gemm = sass[sass.find("gemm_umma_kernel"):]
    gemm = gemm[: gemm.find("Function :", 100)] if gemm.find("Function :", 100) > 0 else gemm
    assert "UTCHMMA" in gemm and "UBLKCP" in gemm and "LDTM" in gemm
