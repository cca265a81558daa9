"""The C-ABI library loads on a GPU-less box, exports every symbol include/fabgpu_ecdsa.h declares, and fails loudly
(no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re

import pytest

from util import ROOT, have_gpu, pkg


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "fabgpu_ecdsa.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fabgpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    b = pkg().binding
    syms = _header_symbols()
    assert len(syms) >= 15
    assert sorted(b.EXPORTS) == syms
    L = ctypes.CDLL(b.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), s


def test_no_cpu_fallback_without_device():
    if have_gpu():
        pytest.skip("a GPU is present")
    b = pkg().binding
    with pytest.raises(b.FabGpuError) as ei:
        b.Context(max_batch=64)
    assert ei.value.code == b.E_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(b.FabGpuError):
        pkg().bccsp.GPUCSP(max_batch=64)


def test_product_does_not_import_oracle():
    # the product tree must not reference oracle/ or tools/ (only tests, smoke and bench's baseline legs may)
    base = os.path.join(ROOT, "fabric-mod_b200")
    for dirpath, _, files in os.walk(base):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".go")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower() and "tools.workload" not in src, (dirpath, f)


def test_small_table_constants_and_handle_codes():
    """No compute: the build constants of the small key tables are self-consistent (windows cover the 257 bits a signed recoding of a
    256-bit scalar needs) and the binding turns small handles (-2 - (generation << 20 | slot)) into the raw device codes (-2 - slot)."""
    import numpy as np
    b = pkg().binding
    wb, nw, nbytes = b.Context.small_table_info()
    assert 4 <= wb <= 12 and nw * wb >= 257 and (nw - 1) * wb < 257
    assert nbytes == nw * (1 << (wb - 1)) * 64
    handles = np.array([-2, -3, -2 - ((5 << 20) | 7), -2 - ((1023 << 20) | 0xFFFFF), -1, 17], np.int32)
    assert b.Context.small_raw_codes(handles).tolist() == [-2, -3, -9, -2 - 0xFFFFF, -1, -1]
