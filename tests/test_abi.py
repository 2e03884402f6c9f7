"""CPU tests of the C-ABI boundary: the library builds for sm_100a, loads without a GPU, and
exports exactly the symbols include/lr_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "lr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lr_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound(native_lib):
    from lipreading_b200 import native
    import lipreading_b200.conv_frontend  # noqa: F401  (registers the conv entry points)
    names = _declared()
    assert len(names) >= 20
    raw = ctypes.CDLL(native.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "declared in lr_b200.h but not exported: " + n
        assert n in native.SIGNATURES, "exported but not bound in native.py: " + n
    assert native_lib.lr_abi_version() == 1
    assert native_lib.lr_last_error() is not None


def test_product_library_has_no_diagnostics_or_global_switches(native_lib):
    """kernel variants are per-call arguments; hooks and micro-benchmarks live in
    liblr_b200_diag.so (include/lr_b200_diag.h), which exports everything the product library does plus the hooks."""
    from lipreading_b200 import native
    raw = ctypes.CDLL(native.LIB_PATH)
    for n in ("lr_conv3d_set_debug", "lr_conv3d_set_debug_skip", "lr_umma_microbench", "lr_umma_pattern_bench",
              "lr_umma_issue_bench", "lr_ctc_select_kernel", "lr_proj_select_kernel", "lr_conv3d_set_seam"):
        assert not hasattr(raw, n), n
    diag = ctypes.CDLL(native.DIAG_LIB_PATH)
    for n in list(native.DIAG_SIGNATURES) + _declared():
        assert hasattr(diag, n), n


def test_workspace_queries_need_no_gpu(native_lib):
    assert native_lib.lr_ctc_workspace(32, 75, 65, 30, 0) == 16             # CTA-per-clip: lattices fit shared memory
    assert native_lib.lr_ctc_workspace(256, 75, 65, 30, 0) == 256 * 75 * 64 * 4 + 256 * 4   # warp kernels: alpha + redo flags
    assert native_lib.lr_ctc_workspace(4, 400, 65, 256, 0) == 4 * 2 * 400 * 513 * 4
    assert native_lib.lr_rnn_workspace(1, 256, 75, 256, 2) == 3 * 2 * 256 * 256 * 4
    assert [native_lib.lr_rnn_saved_per_unit(m) for m in (0, 1, 2)] == [0, 4, 5]


def test_no_cpu_fallback():
    import pytest
    import torch
    from lipreading_b200 import functional as LF, native
    with pytest.raises(native.NativeError):
        LF.ctc_nll(torch.zeros(1, 4, 5).log_softmax(-1), torch.ones(1, 2, dtype=torch.int32),
                   torch.tensor([4]), torch.tensor([2]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lipreading_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
