"""The C-ABI library loads and exports every symbol include/drt.h declares (CPU, no compute)."""
import os
import re

import pytest

from dartray_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "drt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drt_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(drt_lib):
    for name in header_symbols():
        assert hasattr(drt_lib, name), name
    assert drt_lib.drt_version() == 100


def test_no_cpu_fallback_without_device(drt_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    with pytest.raises(capi.DrtError) as ei:
        capi.Context(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_host_only_context_refuses_queries(drt_lib):
    import numpy as np
    from tests.util import random_soup, random_rays
    ctx = capi.Context(capi.DEVICE_NONE)
    P, idx = random_soup(10, 0)
    ctx.set_triangles(P, idx)
    ctx.build_bvh()
    ro, rd = random_rays(4, 0)
    with pytest.raises(capi.DrtError) as ei:
        ctx.trace_closest(ro, rd)
    assert ei.value.code == -4
    with pytest.raises(capi.DrtError):
        ctx.trace_any(ro, rd)


def test_argument_errors(drt_lib):
    import numpy as np
    ctx = capi.Context(capi.DEVICE_NONE)
    P = np.zeros((3, 3), np.float32)
    with pytest.raises(capi.DrtError):
        ctx.set_triangles(P, np.array([[0, 1, 3]], np.uint32))  # index out of range
    ctx.set_triangles(P, np.array([[0, 1, 2]], np.uint32))
    with pytest.raises(capi.DrtError):
        ctx.build_bvh(split=7)
    ctx.set_build_order(np.array([0, 0], np.uint32))
    with pytest.raises(capi.DrtError):
        ctx.build_bvh()
    with pytest.raises(capi.DrtError):
        capi.Context(capi.DEVICE_NONE).bvh_info()  # not built
