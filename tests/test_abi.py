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


def test_argument_errors_of_the_widened_scene_and_render_entries(drt_lib):
    """Host-side validation of the entry points added for SURVEY 8f (no GPU needed: setters only stage data)."""
    import numpy as np
    from dartray_b200 import host
    ctx = capi.Context(capi.DEVICE_NONE)
    eye = np.eye(4, dtype=np.float32).reshape(16)
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    ctx.set_triangles(P, np.array([[0, 1, 2]], np.uint32))
    ctx.set_spheres(np.zeros((0, 16), np.float32), np.zeros((0, 16), np.float32), np.zeros((0, 4)))
    with pytest.raises(capi.DrtError, match="quadric kind"):
        ctx.set_quadrics(7, eye, eye, [[1, 0, 1, 360, 0, 0, 0, 0]])
    ctx.set_quadrics(host.QUADRIC_CONE, eye, eye, [[1, 1, 360, 0, 0, 0, 0, 0]])
    with pytest.raises(capi.DrtError, match="mesh index"):
        ctx.set_mesh_shading(np.zeros((3, 3), np.float32), None, None, [3], eye, eye, [1])
    ctx.set_mesh_shading(np.zeros((3, 3), np.float32), None, None, [0], eye, eye, [1])
    ctx.build_bvh()
    assert ctx.bvh_info()["n_prims"] == 2
    # lights: kinds, and the per-kind companions
    with pytest.raises(capi.DrtError, match="light kind"):
        ctx.set_lights([9], [[1, 1, 1]], [[0, 0, 0]], [1], [0, 0], [])
    ctx.set_lights([4, 5, 6, 1], np.ones((4, 3), np.float32), np.zeros((4, 3), np.float32), [1, 1, 1, 1], [0, 0, 0, 0, 0], [])
    with pytest.raises(capi.DrtError, match="power-of-two"):
        ctx.set_infinite_light(0, np.ones((3, 4, 3), np.float32), eye, eye)
    with pytest.raises(capi.DrtError, match="kind 4"):
        ctx.set_infinite_light(3, np.ones((1, 1, 3), np.float32), eye, eye)
    ctx.set_infinite_light(0, np.ones((2, 4, 3), np.float32), eye, eye)
    with pytest.raises(capi.DrtError, match="kind 5 or 6"):
        ctx.set_light_map(0, None, eye)
    with pytest.raises(capi.DrtError, match="projection"):
        ctx.set_light_map(1, None, eye)  # a projection light without its projection matrix / screen window
    ctx.set_light_map(1, None, eye, host.perspective(45.0, 1e-3, 1e30).reshape(16), (-1, 1, -1, 1), 1e-3)
    ctx.set_light_map(2, np.ones((2, 2, 3), np.float32), eye)
    # materials: lobe kinds and wrappers
    lobes = host.mix_lobes(host.matte_lobes(0.5), host.substrate_lobes(0.4, 0.1, 0.2, 0.1), 0.5)
    ctx.set_material_lobes([0, len(lobes)], [l["kind"] for l in lobes], [l["rgb"] for l in lobes], [l["fresnel"] for l in lobes],
                           [l["eta"] for l in lobes], [l["k"] for l in lobes], [(l["param"], l["ei"], l["et"]) for l in lobes])
    with pytest.raises(capi.DrtError, match="lobe count"):
        ctx.set_lobe_wrappers([2], [[1, 1, 1]])
    with pytest.raises(capi.DrtError, match="wrapper bits"):
        ctx.set_lobe_wrappers([2, 8], np.ones((2, 3), np.float32))
    ctx.set_lobe_wrappers([l["wrap"] for l in lobes], [l["scale"] for l in lobes])
    with pytest.raises(capi.DrtError, match="BxDF"):
        ctx.set_material_lobes([0, 1], [8], [[1, 1, 1]], [0], [[0, 0, 0]], [[0, 0, 0]], [(0, 1, 1)])
    # MeasuredMaterial tables
    with pytest.raises(capi.DrtError, match="beyond the data"):
        ctx.L.drt_set_measured.restype = int
        ctx._ck(ctx.L.drt_set_measured(ctx.h, 1, np.array([0], np.int32).ctypes.data, np.array([90, 90, 180], np.int32).ctypes.data,
                                       np.array([0], np.uint64).ctypes.data, np.zeros(8, np.float32).ctypes.data, 8))
    with pytest.raises(capi.DrtError, match="table kind"):
        ctx.set_measured([(3, np.zeros((2, 6), np.float32))])
    ctx.set_measured([(0, np.zeros((2, 2, 2, 3), np.float32)), (1, np.zeros((5, 6), np.float32))])
    # samplers
    with pytest.raises(capi.DrtError, match="sampler kind"):
        ctx.set_sampler(6, 1, 1, 4, 1, 1, 32, 0)
    with pytest.raises(capi.DrtError, match="4096"):
        ctx.set_sample_table(np.zeros((100, 5)))
    ctx.set_sample_table(np.zeros((4096, 5)))
    ctx.set_sampler(host.SAMPLER_BEST_CANDIDATE, 1, 1, 4, 1, 1, 32, 0)
    # and rendering needs a device (no CPU fallback)
    cam = host.PerspectiveCamera(host.look_at((0, 0, -5), (0, 0, 0), (0, 1, 0)), fov=40.0)
    ctx.set_camera(cam.raster_to_camera(8, 8), cam.camera_to_world)
    xw, yw, table = host.Film(8, 8).table()
    ctx.set_film(8, 8, (0.0, 1.0, 0.0, 1.0), xw, yw, table)
    with pytest.raises(capi.DrtError) as ei:
        ctx.render(0, 1)
    assert ei.value.code == -4 and "no CPU fallback" in str(ei.value)
