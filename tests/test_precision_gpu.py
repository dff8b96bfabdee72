"""drt_set_shading_precision(DRT_PRECISION_F32): the path integrator's vertex / resolve kernels in float32 arithmetic.

north_star's bar for the path tracer is the per-pixel mean within 3 sigma of its Monte Carlo variance (the binary64 kernels replay
the reference's samples to ~1e-6, tests/test_baseline_sizes_gpu.py).  The float32 build keeps the samples, the queues and the binary64
traversal, so the two films differ by float32 rounding plus the rare vertex that takes the other branch at an edge; this file bounds
that difference against the oracle with sigma measured from independent renders.
Reference: lib/surface_integrators/path_integrator.dart:44-119, lib/core/integrator.dart:79-185."""
import os

import numpy as np
import pytest

from dartray_b200 import capi, host, scenes
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu

NT = os.cpu_count() or 8


def _lum(rgb):
    rgb = rgb.astype(np.float64)
    return 0.212671 * rgb[..., 0] + 0.715160 * rgb[..., 1] + 0.072169 * rgb[..., 2]


def _smooth(v, k=9):
    h, w = v.shape
    pad = np.pad(v, k // 2, mode="edge")
    out = np.zeros_like(v)
    for dy in range(k):
        for dx in range(k):
            out += pad[dy:dy + h, dx:dx + w]
    return out / (k * k)


def _render(ctx, cam, film, smp, integ, precision=None, **kw):
    host.configure_render(ctx, cam, film, smp, integ)
    if precision is not None:
        ctx.set_shading_precision(precision)
    ctx.film_clear()
    ctx.render(**kw)
    return ctx.film_read(), ctx.render_stats()


def test_precision_argument_is_validated():
    g = capi.Context(0)
    with pytest.raises(capi.DrtError):
        g.set_shading_precision(2)
    g.set_shading_precision(capi.PRECISION_F32)
    g.set_shading_precision(capi.PRECISION_F64)


def test_f32_path_matches_the_oracle_within_3_sigma_small_film():
    """cornell_synth 192x108, lowdiscrepancy 64 spp, path maxdepth 5: float32 kernels against the ORACLE (binary64, the reference's
    arithmetic).  sigma per pixel from two independent 32-spp oracle renders (difference variance = 2 sigma_32^2 = 4 sigma_64^2)."""
    sb, cam = scenes.cornell_synth()
    arrays = sb.arrays()
    g, o = capi.Context(0), Oracle()
    for c in (g, o):
        host.upload_scene(c, arrays)
    film = host.Film(192, 108)
    smp = host.Sampler(kind=host.SAMPLER_LD, spp=64)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    f32, s32 = _render(g, cam, film, smp, integ, capi.PRECISION_F32)
    f64, s64 = _render(g, cam, film, smp, integ, capi.PRECISION_F64)
    host.configure_render(o, cam, film, smp, integ)
    o.film_clear()
    o.render(0, 1, NT)
    fo, so = o.film_read(), o.render_stats()
    assert np.isfinite(f32["rgb"]).all()
    assert np.array_equal(f32["weight"], fo["weight"])
    # the binary64 build replays the oracle; the float32 build must not be the same film (or the switch did nothing)
    e64 = np.abs(f64["rgb"].astype(np.float64) - fo["rgb"]) / np.maximum(np.abs(fo["rgb"]), 1e-3)
    assert e64.max() <= 1e-3
    assert not np.array_equal(f32["rgb"], f64["rgb"])
    halves = []
    for seed in (1, 2):
        host.configure_render(o, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=32, seed=seed), integ)
        o.film_clear()
        o.render(0, 1, NT)
        halves.append(_lum(o.film_read()["rgb"]))
    sigma = np.sqrt(_smooth(((halves[0] - halves[1]) ** 2) / 4.0))
    a, b = _lum(f32["rgb"]), _lum(fo["rgb"])
    lit = sigma > 1e-4 * b.mean()
    z = np.abs(a - b)[lit] / sigma[lit]
    rel = np.abs(f32["rgb"].astype(np.float64) - fo["rgb"]) / np.maximum(np.abs(fo["rgb"]), 1e-3)
    print(f"f32 path vs oracle, 192x108 x 64 spp: max |d| / sigma {z.max():.3e}, max rel err {rel.max():.3e}, "
          f"99.9 % rel err {np.quantile(rel, 0.999):.3e}, mean f32 {a.mean():.6f} oracle {b.mean():.6f}")
    assert z.max() <= 3.0
    assert abs(a.mean() - b.mean()) <= 5e-3 * b.mean()
    assert s32["camera_samples"] == so["camera_samples"]
    for k in ("closest_rays", "shadow_rays"):  # an edge vertex may take the other branch: counts agree to 1e-4, not exactly
        assert abs(int(s32[k]) - int(so[k])) <= 1e-4 * so[k], (k, s32[k], so[k])


def test_f32_path_config4_1080p_256spp_within_3_sigma_of_the_binary64_film_and_of_the_oracle_window():
    """BASELINE config 4 at its full size.  Whole film: float32 against the binary64 GPU film (itself within 1e-3 of the oracle per
    pixel, tests/test_baseline_sizes_gpu.py), sigma from two independent 128-spp binary64 renders.  A 240x135 window: against the
    oracle directly."""
    sb, cam = scenes.cornell_synth()
    arrays = sb.arrays()
    g, o = capi.Context(0), Oracle()
    for c in (g, o):
        host.upload_scene(c, arrays)
    film = host.Film(1920, 1080)
    smp = host.Sampler(kind=host.SAMPLER_LD, spp=256)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    f32, s32 = _render(g, cam, film, smp, integ, capi.PRECISION_F32)
    f64, s64 = _render(g, cam, film, smp, integ, capi.PRECISION_F64)
    halves = []
    for seed in (1, 2):
        h_, _ = _render(g, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=128, seed=seed), integ, capi.PRECISION_F64)
        halves.append(_lum(h_["rgb"]))
    sigma = np.sqrt(_smooth(((halves[0] - halves[1]) ** 2) / 4.0))
    a, b = _lum(f32["rgb"]), _lum(f64["rgb"])
    lit = sigma > 1e-4 * b.mean()
    z = np.abs(a - b)[lit] / sigma[lit]
    rel = np.abs(f32["rgb"].astype(np.float64) - f64["rgb"]) / np.maximum(np.abs(f64["rgb"]), 1e-3)
    print(f"f32 vs f64 path, 1920x1080 x 256 spp: max |d| / sigma {z.max():.3e}, pixels above 1 sigma {int((z > 1).sum())}, "
          f"max rel err {rel.max():.3e}, 99.99 % {np.quantile(rel, 0.9999):.3e}, median {np.median(rel):.3e}; "
          f"mean f32 {a.mean():.7f} f64 {b.mean():.7f}")
    assert np.isfinite(f32["rgb"]).all()
    assert np.array_equal(f32["weight"], f64["weight"])
    assert z.max() <= 3.0
    assert abs(a.mean() - b.mean()) <= 1e-4 * b.mean()
    assert s32["camera_samples"] == s64["camera_samples"]
    for k in ("closest_rays", "shadow_rays"):
        assert abs(int(s32[k]) - int(s64[k])) <= 1e-4 * s64[k], (k, s32[k], s64[k])
    # the oracle on a window of the same film (streams are keyed by absolute pixel)
    x0, y0, w, h = 840, 472, 240, 135
    crop = ((x0 - 0.5) / 1920, (x0 + w - 0.5) / 1920, (y0 - 0.5) / 1080, (y0 + h - 0.5) / 1080)
    fw = host.Film(1920, 1080, crop=crop)
    assert fw.extent() == (x0, y0, w, h)
    host.configure_render(o, cam, fw, smp, integ)
    o.film_clear()
    o.render(0, 1, NT)
    bo = _lum(o.film_read()["rgb"])
    aw, sw = a[y0:y0 + h, x0:x0 + w], sigma[y0:y0 + h, x0:x0 + w]
    litw = sw > 1e-4 * bo.mean()
    zw = np.abs(aw - bo)[litw] / sw[litw]
    print(f"f32 path vs oracle on a 240x135 window of the 1080p film: max |d| / sigma {zw.max():.3e}")
    assert zw.max() <= 3.0


def test_f32_path_on_bxdf_list_materials_within_3_sigma_of_the_oracle():
    """cornell_materials (plastic, OrenNayar matte, metal, uber, glass: BxDF lists with specular bounces), 160x90, lowdiscrepancy 64 spp:
    the float32 kernels against the oracle.  A specular choice that flips at a float32 rounding replaces one sample's radiance by
    another valid sample's, so the bound is the Monte Carlo one, with sigma from two independent 32-spp oracle renders."""
    sb, cam = scenes.cornell_materials()
    arrays = sb.arrays()
    g, o = capi.Context(0), Oracle()
    for c in (g, o):
        host.upload_scene(c, arrays)
    film = host.Film(160, 90)
    smp = host.Sampler(kind=host.SAMPLER_LD, spp=64)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    f32, s32 = _render(g, cam, film, smp, integ, capi.PRECISION_F32)
    f64, s64 = _render(g, cam, film, smp, integ, capi.PRECISION_F64)
    host.configure_render(o, cam, film, smp, integ)
    o.film_clear()
    o.render(0, 1, NT)
    fo, so = o.film_read(), o.render_stats()
    assert np.isfinite(f32["rgb"]).all()
    assert not np.array_equal(f32["rgb"], f64["rgb"])
    halves = []
    for seed in (1, 2):
        host.configure_render(o, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=32, seed=seed), integ)
        o.film_clear()
        o.render(0, 1, NT)
        halves.append(_lum(o.film_read()["rgb"]))
    sigma = np.sqrt(_smooth(((halves[0] - halves[1]) ** 2) / 4.0))
    a, b = _lum(f32["rgb"]), _lum(fo["rgb"])
    lit = sigma > 1e-4 * b.mean()
    z = np.abs(a - b)[lit] / sigma[lit]
    rel = np.abs(f32["rgb"].astype(np.float64) - fo["rgb"]) / np.maximum(np.abs(fo["rgb"]), 1e-3)
    print(f"f32 path vs oracle, cornell_materials 160x90 x 64 spp: max |d| / sigma {z.max():.3e}, pixels above 1 sigma {int((z > 1).sum())}, "
          f"max rel err {rel.max():.3e}, 99.9 % rel err {np.quantile(rel, 0.999):.3e}, mean f32 {a.mean():.6f} oracle {b.mean():.6f}")
    assert z.max() <= 3.0
    assert abs(a.mean() - b.mean()) <= 5e-3 * b.mean()
    assert s32["camera_samples"] == so["camera_samples"]
    for k in ("closest_rays", "shadow_rays"):
        assert abs(int(s32[k]) - int(so[k])) <= 1e-3 * so[k], (k, s32[k], so[k])


def _f32_against_oracle(arrays, cam, film, spp, what, count_tol=1e-3):
    """float32 path render against the oracle: per-pixel |d| <= 3 sigma (sigma from two independent half-size oracle renders), image
    mean within 0.5 %, ray counts within count_tol; the float32 film must differ from the binary64 one (the switch took effect)."""
    g, o = capi.Context(0), Oracle()
    for c in (g, o):
        host.upload_scene(c, arrays)
    smp = host.Sampler(kind=host.SAMPLER_LD, spp=spp)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    f32, s32 = _render(g, cam, film, smp, integ, capi.PRECISION_F32)
    f64, _ = _render(g, cam, film, smp, integ, capi.PRECISION_F64)
    host.configure_render(o, cam, film, smp, integ)
    o.film_clear()
    o.render(0, 1, NT)
    fo, so = o.film_read(), o.render_stats()
    assert np.isfinite(f32["rgb"]).all()
    assert not np.array_equal(f32["rgb"], f64["rgb"])
    e64 = np.abs(f64["rgb"].astype(np.float64) - fo["rgb"]) / np.maximum(np.abs(fo["rgb"]), 1e-3)
    halves = []
    for seed in (1, 2):
        host.configure_render(o, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=spp // 2, seed=seed), integ)
        o.film_clear()
        o.render(0, 1, NT)
        halves.append(_lum(o.film_read()["rgb"]))
    sigma = np.sqrt(_smooth(((halves[0] - halves[1]) ** 2) / 4.0))
    a, b = _lum(f32["rgb"]), _lum(fo["rgb"])
    lit = sigma > 1e-4 * b.mean()
    z = np.abs(a - b)[lit] / sigma[lit]
    rel = np.abs(f32["rgb"].astype(np.float64) - fo["rgb"]) / np.maximum(np.abs(fo["rgb"]), 1e-3)
    print(f"f32 path vs oracle, {what}: max |d| / sigma {z.max():.3e}, pixels above 1 sigma {int((z > 1).sum())}, max rel err {rel.max():.3e}, "
          f"99.9 % {np.quantile(rel, 0.999):.3e} (binary64 kernels: max rel err {e64.max():.1e}), mean f32 {a.mean():.6f} oracle {b.mean():.6f}")
    assert z.max() <= 3.0
    assert abs(a.mean() - b.mean()) <= 5e-3 * b.mean()
    assert s32["camera_samples"] == so["camera_samples"]
    for k in ("closest_rays", "shadow_rays"):
        assert abs(int(s32[k]) - int(so[k])) <= count_tol * so[k], (k, s32[k], so[k])


@pytest.mark.parametrize("which", ["matte", "lobes"])
def test_f32_path_with_per_vertex_mesh_attributes(which):
    """The `extra` build in float32 (render_kernels_f32x.cu): smooth-shaded meshes (per-vertex N / S / uv, triangle.dart:100-160)."""
    from tests.test_render_gpu import _smooth_room
    arrays, cam = _smooth_room(which)
    _f32_against_oracle(arrays, cam, host.Film(96, 72), 64, f"smooth room ({which})")


@pytest.mark.parametrize("which", ["matte", "lobes"])
def test_f32_path_under_an_infinite_light_with_a_cylinder_glass_and_mirror(which):
    """The `extra` build in float32 on the environment-lit scene of the render tests: InfiniteAreaLight importance sampling, a cylinder,
    glass and mirror spheres (specular bounces pick up Le of escaped rays through the binary64 escape kernel)."""
    from tests.test_render_gpu import _sky_scene
    arrays, cam = _sky_scene(which)
    _f32_against_oracle(arrays, cam, host.Film(96, 72), 64, f"sky scene ({which})")


def test_f32_path_on_a_127k_triangle_scene_runs_the_float32_leaf_phase_of_the_quantised_kernel():
    """soup(64) (127 K triangles: the quantised-node kernel) lit by its area light, 96x54 x 64 spp: under DRT_PRECISION_F32 the path
    integrator's queues run traceQKernel with a float32 leaf phase (trace_q_f32.cu)."""
    sb, cam = scenes.soup_render_scene(64)
    arrays = sb.arrays()
    _f32_against_oracle(arrays, cam, host.Film(96, 54), 64, "soup(64) 127 K triangles")


def test_f32_path_on_a_16k_triangle_scene_runs_the_float32_tree_kernel():
    """soup(8) (16 K triangles: below the quantised kernel's threshold, above the leaf-list kernel's): the float32 build of
    traceFastKernel (trace_fast_f32.cu)."""
    sb, cam = scenes.soup_render_scene(8)
    _f32_against_oracle(sb.arrays(), cam, host.Film(96, 54), 64, "soup(8) 16 K triangles")
