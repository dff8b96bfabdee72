"""Participating media: the VolumeRegion plugins and the emission / single-scattering volume integrators.

CPU tests pin the oracle with closed forms written from the reference's algorithm (lib/volume_integrators/*.dart,
lib/volume_regions/*.dart, lib/core/volume/*.dart):
  * Beer-Lambert through a homogeneous slab along a SHADOW ray (analytic tau, homogenous_volume_region.dart:65-73): exact;
  * along the CAMERA ray the emission integrator marches: T = exp(-sigma * (n - 1 + u) * step) for its n steps and scatter sample u
    (emission_integrator.dart:47-76: the last partial step is never taken) — the expectation over the pixel's stratified u values;
  * emission: Lv = Le * sum_i Tr_i * step;
  * exponential / grid densities: tau by the reference's own Riemann sum (density_region.dart:53-77) against numpy;
  * single scattering under a point light in an optically thin medium: sigma_s * phase * I / d^2 integrated along the ray.
GPU tests (marked gpu) replay the oracle's keyed streams per pixel."""
import math

import numpy as np
import pytest

from dartray_b200 import capi, host
from tests.oracle_lib import Oracle


def _emitter_scene(z=10.0, L=(1.0, 1.0, 1.0)):
    sb = host.SceneBuilder()
    sb.mesh([[-50, -50, z], [50, -50, z], [50, 50, z], [-50, 50, z]], [[0, 2, 1], [0, 3, 2]], area_light=L)  # faces -z, towards the camera
    return sb


def _render(ctx, sb, cam, film, smp, integ, nthreads=4):
    host.upload_scene(ctx, sb.arrays())
    host.configure_render(ctx, cam, film, smp, integ)
    if isinstance(ctx, Oracle):
        ctx.render(0, 1, nthreads)
    else:
        ctx.render()
    return ctx.film_read()["rgb"].astype(np.float64)


CAM = host.PerspectiveCamera(host.look_at((0, 0, 0), (0, 0, 1), (0, 1, 0)), fov=0.5)
PATH1 = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=1)


def test_camera_ray_transmittance_is_the_marched_beer_lambert():
    sig = np.array([0.5, 1.0, 2.0])
    sb = _emitter_scene()
    sb.volume("homogeneous", sigma_a=sig, p0=(-5, -5, 3), p1=(5, 5, 5))
    sb.volume_integrator("emission", 0.1)
    spp = 16
    rgb = _render(Oracle(), sb, CAM, host.Film(1, 1), host.Sampler(kind=host.SAMPLER_LD, spp=spp), PATH1)[0, 0]
    # thickness 2 along an (almost) axial ray: n = 20 steps of 0.1; the scatter sample u is a scrambled (0,2)-sequence value:
    # {u_k} = {(k + c) / spp} for some c in [0, 1) -> average of exp(-sigma * (19 + u) * 0.1) over the 16 strata, c unknown:
    # bracket it by c = 0 and c = 1
    n, step = 20, 0.1
    k = np.arange(spp)
    hi = np.mean(np.exp(-sig[None, :] * (n - 1 + (k[:, None] + 0.0) / spp) * step), axis=0)
    lo = np.mean(np.exp(-sig[None, :] * (n - 1 + (k[:, None] + 1.0) / spp) * step), axis=0)
    assert ((rgb >= lo * (1 - 1e-3)) & (rgb <= hi * (1 + 1e-3))).all(), (rgb, lo, hi)
    # no region: T = 1 exactly; a region the ray misses: T = 1
    sb2 = _emitter_scene()
    assert np.allclose(_render(Oracle(), sb2, CAM, host.Film(1, 1), host.Sampler(kind=host.SAMPLER_LD, spp=4), PATH1), 1.0, rtol=1e-6)
    sb3 = _emitter_scene()
    sb3.volume("homogeneous", sigma_a=5.0, p0=(20, 20, 3), p1=(30, 30, 5))
    assert np.allclose(_render(Oracle(), sb3, CAM, host.Film(1, 1), host.Sampler(kind=host.SAMPLER_LD, spp=4), PATH1), 1.0, rtol=1e-6)


def test_emission_adds_le_times_the_marched_transmittance():
    sig, le = 0.5, np.array([2.0, 1.0, 0.5])
    sb = host.SceneBuilder()
    sb.mesh([[-50, -50, 10], [50, -50, 10], [50, 50, 10], [-50, 50, 10]], [[0, 2, 1], [0, 3, 2]])  # a black wall ends the ray
    sb.volume("homogeneous", sigma_a=sig, le=le, p0=(-5, -5, 3), p1=(5, 5, 5))
    sb.volume_integrator("emission", 0.1)
    rgb = _render(Oracle(), sb, CAM, host.Film(1, 1), host.Sampler(kind=host.SAMPLER_LD, spp=64), PATH1)[0, 0]
    # Lv = step * Le * sum_{i < n} exp(-sigma * (i + u) * step): between u = 1 and u = 0
    n, step = 20, 0.1
    i = np.arange(n)
    hi = step * np.exp(-sig * (i + 0.0) * step).sum() * le
    lo = step * np.exp(-sig * (i + 1.0) * step).sum() * le
    assert ((rgb >= lo * (1 - 1e-3)) & (rgb <= hi * (1 + 1e-3))).all(), (rgb, lo, hi)
    # ... and brackets the continuous answer Le (1 - exp(-sigma d)) / sigma
    exact = le * (1 - math.exp(-sig * 2.0)) / sig
    assert np.allclose(rgb, exact, rtol=3e-2)


def _lit_floor(volume_kwargs, integ=None, light_height=8.0, stepsize=0.05):
    """A matte floor under a point light with a slab of medium between the two; the camera sits below the slab and looks at the
    floor point under the light, so only the SHADOW ray crosses the medium.  Returns (radiance, the radiance without a medium)."""
    sb = host.SceneBuilder()
    kd = 0.5
    sb.mesh([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], [[0, 2, 1], [0, 3, 2]], material=sb.material(kd))  # normal +y
    sb.point_light((0.0, light_height, 0.0), (30.0, 30.0, 30.0))
    if volume_kwargs:
        sb.volume(**volume_kwargs)
        sb.volume_integrator("emission", stepsize)
    # the camera looks at the floor point under the light from far to the side, below the slab: its ray never enters the region
    cam = host.PerspectiveCamera(host.look_at((8.0, 2.5, 0.0), (0, 0, 0), (0, 1, 0)), fov=0.05)
    rgb = _render(Oracle(), sb, cam, host.Film(1, 1), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                  integ or host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=1))[0, 0]
    return rgb, kd / math.pi * 30.0 / light_height ** 2


def test_shadow_rays_see_the_analytic_and_the_riemann_optical_depth():
    clear, expect = _lit_floor(None)
    assert np.allclose(clear, expect, rtol=1e-4)
    # homogeneous slab of thickness 2 between light and floor, the camera ray passes below it: exp(-sigma_t * 2) exactly
    sig_a, sig_s = np.array([0.1, 0.3, 0.6]), np.array([0.05, 0.1, 0.2])
    slab, _ = _lit_floor(dict(kind="homogeneous", sigma_a=sig_a, sigma_s=sig_s, p0=(-40, 3, -40), p1=(40, 5, 40)))
    assert np.allclose(slab, expect * np.exp(-(sig_a + sig_s) * 2.0), rtol=1e-4)
    # exponential density a exp(-b h), h measured from the extent's floor: tau = sigma_t * a * (1 - exp(-b H)) / b in the limit;
    # the reference sums sigma_t(p_k) * step at t0 + (u + k) * step with step = 4 x stepsize for a sample-less transmittance call
    a, b, H = 2.0, 0.7, 2.0
    expo, _ = _lit_floor(dict(kind="exponential", sigma_a=sig_a, sigma_s=0.0, p0=(-40, 3, -40), p1=(40, 5, 40), a=a, b=b), stepsize=0.005)
    tau = sig_a * a * (1 - math.exp(-b * H)) / b
    assert np.allclose(expo, expect * np.exp(-tau), rtol=2e-2)
    # volumegrid: trilinear interpolation of a density that is linear in y reproduces it between the voxel centres
    ny = 8
    dens = np.broadcast_to(np.linspace(0.5, 1.5, ny)[None, :, None], (3, ny, 3)).copy()
    grid, _ = _lit_floor(dict(kind="volumegrid", sigma_a=sig_a, sigma_s=0.0, p0=(-40, 3, -40), p1=(40, 5, 40), density=dens), stepsize=0.005)
    # density(y) clamps to the outer voxel centres over the first / last half voxel: mean over the slab
    ys = (np.arange(200000) + 0.5) / 200000
    vox = ys * ny - 0.5
    d = np.interp(vox, np.arange(ny), np.linspace(0.5, 1.5, ny))
    assert np.allclose(grid, expect * np.exp(-sig_a * d.mean() * H), rtol=2e-2)


def test_single_scattering_in_a_thin_medium_matches_the_line_integral():
    # camera ray along +z through a thin isotropic medium lit by a point light off the axis; black wall behind
    sig_s, I, x0 = 1.0e-3, 100.0, 2.0
    sb = host.SceneBuilder()
    sb.mesh([[-50, -50, 10], [50, -50, 10], [50, 50, 10], [-50, 50, 10]], [[0, 2, 1], [0, 3, 2]], material=sb.material(0.0))  # black
    sb.point_light((x0, 0.0, 5.0), (I, I, I))
    sb.volume("homogeneous", sigma_s=sig_s, g=0.0, p0=(-20, -20, 1), p1=(20, 20, 9))
    sb.volume_integrator("single", 0.05)
    rgb = _render(Oracle(), sb, CAM, host.Film(1, 1), host.Sampler(kind=host.SAMPLER_LD, spp=64), PATH1)[0, 0]
    # Lv = int sigma_s * (1 / 4 pi) * I / d(z)^2 dz over z in [1, 9] (attenuation ~ exp(-1e-2): ignored at the tolerance)
    z = np.linspace(1, 9, 200001)
    expect = np.trapezoid(sig_s / (4 * math.pi) * I / (x0 * x0 + (z - 5.0) ** 2), z)
    assert np.allclose(rgb, expect, rtol=2e-2), (rgb, expect)
    # anisotropic phase function: forward-peaked g sends less light sideways; PhaseHG (volume.dart:84-88) by quadrature
    sb.volumes[0]["g"] = 0.6
    rgb_g = _render(Oracle(), sb, CAM, host.Film(1, 1), host.Sampler(kind=host.SAMPLER_LD, spp=64), PATH1)[0, 0]
    g = 0.6
    # w = -ray.d = (0,0,-1); -wo = direction from the light to the point... p(p, w, -wo): cos between -z and (p - light) / |.|
    cosang = -(z - 5.0) / np.sqrt(x0 * x0 + (z - 5.0) ** 2)
    ph = (1 - g * g) / (4 * math.pi * (1 + g * g - 2 * g * cosang) ** 1.5)
    expect_g = np.trapezoid(sig_s * ph * I / (x0 * x0 + (z - 5.0) ** 2), z)
    assert np.allclose(rgb_g, expect_g, rtol=2e-2), (rgb_g, expect_g)


def test_aggregate_of_two_regions_adds_optical_depths():
    one, expect = _lit_floor(dict(kind="homogeneous", sigma_a=0.4, p0=(-40, 3, -40), p1=(40, 5, 40)))
    sb_two = dict(kind="homogeneous", sigma_a=0.4, p0=(-40, 3, -40), p1=(40, 4, 40))
    # build by hand: two stacked slabs of thickness 1
    sb = host.SceneBuilder()
    sb.mesh([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], [[0, 2, 1], [0, 3, 2]], material=sb.material(0.5))
    sb.point_light((0.0, 8.0, 0.0), (30.0, 30.0, 30.0))
    sb.volume(**sb_two)
    sb.volume("homogeneous", sigma_a=0.4, p0=(-40, 4, -40), p1=(40, 5, 40))
    sb.volume_integrator("emission", 0.05)
    cam = host.PerspectiveCamera(host.look_at((8.0, 2.5, 0.0), (0, 0, 0), (0, 1, 0)), fov=0.05)
    two = _render(Oracle(), sb, cam, host.Film(1, 1), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=1))[0, 0]
    assert np.allclose(one, two, rtol=1e-5) and np.allclose(two, expect * math.exp(-0.8), rtol=1e-4)


# ---- GPU parity -----------------------------------------------------------------------------------------------------------------
def _smoke_scene(integrator="single", kind="homogeneous"):
    """A Cornell-like box with a matte sphere, an area light and a medium filling the middle of the box."""
    from dartray_b200 import scenes
    sb, cam = scenes.cornell_synth()
    sb.point_light((3.0, 6.0, -4.0), (60.0, 50.0, 40.0))
    kw = dict(sigma_a=(0.01, 0.02, 0.03), sigma_s=(0.05, 0.04, 0.03), g=0.3, le=(0.02, 0.01, 0.0), p0=(-6, -6, -6), p1=(6, 6, 6))
    if kind == "exponential":
        kw.update(a=1.5, b=0.2, updir=(0.1, 1.0, 0.0))
    if kind == "volumegrid":
        rng = np.random.default_rng(3)
        kw["density"] = rng.uniform(0.2, 2.0, size=(4, 5, 6))
    sb.volume(kind, **kw)
    if kind == "aggregate":
        pass
    sb.volume_integrator(integrator, 1.5)
    return sb, cam


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["homogeneous", "exponential", "volumegrid"])
@pytest.mark.parametrize("vol_integrator", ["emission", "single"])
@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4), host.Integrator(kind=host.INTEGRATOR_DIRECT),
                                   host.Integrator(kind=host.INTEGRATOR_WHITTED), host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=8)],
                         ids=["path", "direct", "whitted", "ao"])
def test_gpu_media_match_the_oracle_per_pixel(kind, vol_integrator, integ):
    sb, cam = _smoke_scene(vol_integrator, kind)
    film, smp = host.Film(48, 36), host.Sampler(kind=host.SAMPLER_LD, spp=4)
    a = _render(capi.Context(0), sb, cam, film, smp, integ)
    b = _render(Oracle(), sb, cam, film, smp, integ, nthreads=8)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
    print(kind, vol_integrator, integ.kind, "max rel err", err.max())
    assert err.max() <= 1e-3


@pytest.mark.gpu
def test_gpu_aggregate_volume_and_stratified_sampler():
    sb, cam = _smoke_scene("single", "homogeneous")
    sb.volume("exponential", sigma_a=0.02, sigma_s=0.03, g=-0.2, p0=(-8, -8, -2), p1=(8, 0, 8), a=1.0, b=0.3)
    film, smp = host.Film(40, 30), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3)
    g = capi.Context(0)
    a = _render(g, sb, cam, film, smp, integ)
    o = Oracle()
    b = _render(o, sb, cam, film, smp, integ, nthreads=8)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
    assert err.max() <= 1e-3
    sg, so = g.render_stats(), o.render_stats()
    assert sg["camera_samples"] == so["camera_samples"] and sg["shadow_rays"] == so["shadow_rays"] and sg["closest_rays"] == so["closest_rays"]


@pytest.mark.gpu
def test_gpu_rejects_media_with_the_specular_recursion():
    from dartray_b200 import scenes
    sb, cam = scenes.cornell_synth()
    sb.mesh([[-1, -9, -1], [1, -9, -1], [1, -9, 1], [-1, -9, 1]], [[0, 2, 1], [0, 3, 2]], material=sb.material_lobes(host.mirror_lobes(0.9)))
    sb.volume("homogeneous", sigma_a=0.01, p0=(-6, -6, -6), p1=(6, 6, 6))
    g = capi.Context(0)
    host.upload_scene(g, sb.arrays())
    host.configure_render(g, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_LD, spp=1), host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=3))
    with pytest.raises(capi.DrtError):
        g.render()
    # the path integrator handles specular bounces itself (no renderer.Li recursion): media are fine there
    host.configure_render(g, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_LD, spp=1), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3))
    g.render()
