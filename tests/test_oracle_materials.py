"""Pins the oracle's BxDF lists (mirror / glass / plastic / metal / uber materials decomposed into
Lambertian, OrenNayar, Microfacet+Blinn, SpecularReflection, SpecularTransmission) with closed forms.  The reference
ships no golden vectors, so each test states the formula it follows with the reference lines it comes from."""
import math

import numpy as np
import pytest

from dartray_b200 import host
from tests.oracle_lib import Oracle


def _oracle(sb, cam, film, sampler, integ):
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, film, sampler, integ)
    return o


def _quad(sb, corners, **kw):
    return sb.mesh(np.asarray(corners, np.float32), [[0, 1, 2], [0, 2, 3]], **kw)


def _floor(sb, y=0.0, half=50.0, **kw):  # normal +y
    return sb.mesh([[-half, y, -half], [half, y, -half], [half, y, half], [-half, y, half]], [[0, 2, 1], [0, 3, 2]], **kw)


def _centre(o, film):
    rgb = o.film_read()["rgb"]
    return rgb[film.yres // 2, film.xres // 2].astype(np.float64)


def _dielectric(cosi, ei, et):  # fresnel_dielectric.dart:24-56
    sint = ei / et * math.sqrt(max(0.0, 1.0 - cosi * cosi))
    if sint >= 1.0:
        return 1.0
    cost = math.sqrt(max(0.0, 1.0 - sint * sint))
    rpar = (et * cosi - ei * cost) / (et * cosi + ei * cost)
    rper = (ei * cosi - et * cost) / (ei * cosi + et * cost)
    return 0.5 * (rpar * rpar + rper * rper)


# ---- host-side flattening (what the Dart shim does when it walks Material objects) ---------------------------------
def test_material_flattening_follows_the_reference_constructors():
    assert host.matte_lobes(0.0) == []                                       # matte_material.dart:57: black Kd adds nothing
    assert host.matte_lobes(0.5)[0]["kind"] == host.LOBE_LAMBERTIAN
    m = host.matte_lobes(0.5, 200.0)[0]
    assert m["kind"] == host.LOBE_OREN_NAYAR and m["param"] == 90.0          # sigma clamp, matte_material.dart:56
    g = host.glass_lobes()
    assert [l["kind"] for l in g] == [host.LOBE_SPECULAR_REFLECTION, host.LOBE_SPECULAR_TRANSMISSION]
    assert (g[0]["ei"], g[0]["et"]) == (1.0, 1.5)                            # glass_material.dart:41-42
    p = host.plastic_lobes(roughness=0.1)
    assert [l["kind"] for l in p] == [host.LOBE_LAMBERTIAN, host.LOBE_MICROFACET_BLINN]
    assert p[1]["param"] == 1.0 / 0.1 and (p[1]["ei"], p[1]["et"]) == (1.5, 1.0)  # plastic_material.dart:45-47
    assert host.plastic_lobes(roughness=0.0)[1]["param"] == 10000.0          # blinn.dart:24-28
    assert host.plastic_lobes(roughness=1e-5)[1]["param"] == 10000.0
    u = host.uber_lobes(kd=0.3, ks=0.2, kr=0.1, kt=0.1, opacity=0.5)
    assert [l["kind"] for l in u] == [host.LOBE_SPECULAR_TRANSMISSION, host.LOBE_LAMBERTIAN, host.LOBE_MICROFACET_BLINN,
                                      host.LOBE_SPECULAR_REFLECTION, host.LOBE_SPECULAR_TRANSMISSION]
    assert np.allclose(u[0]["rgb"], 0.5) and (u[0]["ei"], u[0]["et"]) == (1.0, 1.0)  # uber_material.dart:41-45
    assert np.allclose(u[1]["rgb"], np.float32(0.5) * np.float32(0.3))
    assert [l["kind"] for l in host.uber_lobes()] == [host.LOBE_LAMBERTIAN, host.LOBE_MICROFACET_BLINN]  # defaults: Kr = Kt = 0


# ---- mirror: L = Kr * Le of what the reflected ray sees (specular_reflection.dart:34-41, path_integrator.dart:46-48) --
def test_mirror_shows_the_emitter_scaled_by_kr():
    sb = host.SceneBuilder()
    mirror = sb.material_lobes(host.mirror_lobes((0.9, 0.5, 0.25)))
    _floor(sb, 0.0, material=mirror)
    # emitter: big quad at y = 8 facing down
    _quad(sb, [[-40, 8, -40], [40, 8, -40], [40, 8, 40], [-40, 8, 40]], area_light=(2.0, 2.0, 2.0))
    cam = host.PerspectiveCamera(host.look_at((0, 3, -6), (0, 0, 0), (0, 1, 0)), fov=20.0)
    film = host.Film(9, 9)
    o = _oracle(sb, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5))
    o.render(0, 1, 2)
    # FresnelNoOp = 1; f * |cos| / pdf = Kr exactly in exact arithmetic; float32 storage leaves ~1e-7
    assert np.allclose(_centre(o, film), [1.8, 1.0, 0.5], rtol=1e-5)


def test_mirror_is_black_without_a_specular_path_to_a_light():
    sb = host.SceneBuilder()
    _floor(sb, 0.0, material=sb.material_lobes(host.mirror_lobes(0.9)))
    sb.point_light((0, 5, 0), (10, 10, 10))  # delta lights are never seen by specular bounces; EstimateDirect skips specular lobes
    cam = host.PerspectiveCamera(host.look_at((0, 3, -6), (0, 0, 0), (0, 1, 0)), fov=20.0)
    film = host.Film(5, 5)
    o = _oracle(sb, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=2), host.Integrator(kind=host.INTEGRATOR_PATH))
    o.render(0, 1, 1)
    assert np.all(o.film_read()["rgb"] == 0)


# ---- glass slab at normal incidence: transmission series ------------------------------------------------------------
def test_glass_slab_transmits_the_fresnel_series():
    # camera looks straight down -y through a slab (two interfaces, normals +y / -y) at an emitter below facing up
    sb = host.SceneBuilder()
    glass = sb.material_lobes(host.glass_lobes(1.0, 1.0, 1.5))
    _floor(sb, 2.0, half=20.0, material=glass)                                              # top face, normal +y
    sb.mesh([[-20, 1, -20], [20, 1, -20], [20, 1, 20], [-20, 1, 20]], [[0, 1, 2], [0, 2, 3]], material=glass)  # bottom face, normal -y
    _floor(sb, 0.0, half=20.0, area_light=(1.0, 1.0, 1.0))
    cam = host.PerspectiveCamera(host.look_at((0, 6, 0), (0, 0, 0), (0, 0, 1)), fov=2.0)
    film = host.Film(3, 3)
    spp = 4096
    o = _oracle(sb, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=spp), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=50))
    o.render(0, 1, 8)
    F = _dielectric(1.0, 1.0, 1.5)
    assert abs(F - 0.04) < 1e-12
    # specular_transmission.dart:63-65 carries (1 - F) * Kt per interface and no eta^2 factor; inside the slab the light
    # may bounce an even number of times: T = (1-F)^2 * sum_k F^(2k)
    expect = (1.0 - F) ** 2 / (1.0 - F * F)
    got = _centre(o, film)
    # each path is worth 0 or a product of 2^depth factors: a Monte Carlo estimate, sigma/sqrt(n) ~ 1 / sqrt(spp)
    assert abs(got[0] - expect) < 4.0 * 1.0 / math.sqrt(spp), (got, expect)
    assert np.allclose(got, got[0], rtol=1e-4)  # grey up to the film's float32 XYZ round trip


# ---- plastic under a point light at normal incidence (deterministic) --------------------------------------------------
@pytest.mark.parametrize("roughness", [0.1, 0.02])
def test_plastic_under_a_point_light_matches_the_closed_form(roughness):
    kd, ks, I, h = 0.4, 0.3, 7.0, 3.0
    sb = host.SceneBuilder()
    _floor(sb, 0.0, material=sb.material_lobes(host.plastic_lobes(kd, ks, roughness)))
    sb.point_light((0, h, 0), (I, I, I))
    cam = host.PerspectiveCamera(host.look_at((0, 10, 0), (0, 0, 0), (0, 0, 1)), fov=1.0)  # looks straight down: wo = wi = n
    film = host.Film(1, 1)
    o = _oracle(sb, cam, film, host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False),
                host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=1))
    o.render(0, 1, 1)
    e = 1.0 / roughness
    D = (e + 2.0) / (2.0 * math.pi)                # blinn.dart:31-34 at cos = 1
    Fr = _dielectric(1.0, 1.5, 1.0)                # plastic_material.dart:45: FresnelDielectric(1.5, 1.0) -> 0.04
    f = kd / math.pi + ks * D * 1.0 * Fr / 4.0     # lambertian.dart:35-37 + microfacet.dart:28-46 with G = 1
    expect = f * I / (h * h)                       # point_light.dart:41-47, cos = 1
    assert np.allclose(_centre(o, film), expect, rtol=2e-5), (_centre(o, film), expect)


def test_metal_under_a_point_light_matches_the_conductor_fresnel():
    eta, k, I, h, rough = (0.2, 0.9, 1.1), (3.9, 2.4, 2.2), 5.0, 2.0, 0.05
    sb = host.SceneBuilder()
    _floor(sb, 0.0, material=sb.material_lobes(host.metal_lobes(eta, k, rough)))
    sb.point_light((0, h, 0), (I, I, I))
    cam = host.PerspectiveCamera(host.look_at((0, 10, 0), (0, 0, 0), (0, 0, 1)), fov=1.0)
    film = host.Film(1, 1)
    o = _oracle(sb, cam, film, host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False),
                host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=1))
    o.render(0, 1, 1)
    n, kk = np.asarray(eta), np.asarray(k)
    # fresnel_conductor.dart:29-46 at cosi = 1: Rparl2 == Rperp2 == ((n-1)^2 + k^2) / ((n+1)^2 + k^2)
    F = ((n - 1) ** 2 + kk ** 2) / ((n + 1) ** 2 + kk ** 2)
    e = 1.0 / rough
    expect = (e + 2.0) / (2.0 * math.pi) * F / 4.0 * I / (h * h)
    assert np.allclose(_centre(o, film), expect, rtol=2e-5), (_centre(o, film), expect)


# ---- white furnace (closed emissive sphere seen from inside, as tests/test_oracle_render.py::test_path_white_furnace) ----
def _furnace(lobes, Le=1.0, maxdepth=6, spp=64):
    sb = host.SceneBuilder()
    sb.sphere(host.translate(0, 0, 0), radius=5.0, material=sb.material_lobes(lobes), area_light=(Le, Le, Le), reverse=True)
    cam = host.PerspectiveCamera(host.look_at((0, 0, -1), (0, 0, 0), (0, 1, 0)), fov=60.0)
    o = _oracle(sb, cam, host.Film(16, 16), host.Sampler(kind=host.SAMPLER_LD, spp=spp), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=maxdepth))
    o.render(0, 1, 8)
    return float(o.film_read()["rgb"].mean())


def test_two_lobe_bsdf_sampling_is_consistent_in_a_furnace():
    # two Lambertian lobes of 0.2 + 0.3 are one Lambertian of 0.5: component choice, the averaged pdf and the summed f
    # of BSDF.sample_f (bsdf.dart:53-126) must cancel exactly as for a single lobe: L = Le * sum_{k=0}^{maxdepth+1} rho^k
    expect = sum(0.5 ** k for k in range(6 + 2))
    assert _furnace(host.matte_lobes(0.2) + host.matte_lobes(0.3)) == pytest.approx(expect, rel=3e-3)
    assert _furnace(host.matte_lobes(0.5)) == pytest.approx(expect, rel=3e-3)


def test_plastic_furnace_gains_energy_over_its_diffuse_lobe_only():
    kd, ks = 0.3, 0.25
    diffuse = sum(kd ** k for k in range(6 + 2))
    got = _furnace(host.plastic_lobes(kd, ks, 0.2), spp=256)
    # the Blinn microfacet lobe is not normalised (no closed form), but it only adds energy and F <= 1, G <= 1 bound it
    assert diffuse * 1.01 < got < sum((kd + ks) ** k for k in range(6 + 2)) * 1.2, (diffuse, got)
    # a mirror sphere seen from inside: every bounce is specular, so Le is added at every vertex (path_integrator.dart:46-48);
    # exact up to bounce 3, Russian roulette (:93-99) makes the later terms a Monte Carlo estimate
    kr = 0.7
    assert _furnace(host.mirror_lobes(kr), maxdepth=3) == pytest.approx(sum(kr ** k for k in range(3 + 1)), rel=1e-5)
    assert _furnace(host.mirror_lobes(kr), maxdepth=6) == pytest.approx(sum(kr ** k for k in range(6 + 1)), rel=1e-2)


# ---- directlighting recursion through a mirror (integrator.dart:187-235) ---------------------------------------------
def test_directlighting_sees_a_lit_matte_wall_through_a_mirror():
    kd, I = 0.6, 20.0
    sb = host.SceneBuilder()
    _floor(sb, 0.0, half=4.0, material=sb.material_lobes(host.mirror_lobes(0.8)))
    # matte wall at z = 5 facing -z, lit by a point light in front of it
    _quad(sb, [[-30, -30, 5], [30, -30, 5], [30, 30, 5], [-30, 30, 5]], material=sb.material_lobes(host.matte_lobes(kd)))
    sb.point_light((0, 3, 2), (I, I, I))
    cam = host.PerspectiveCamera(host.look_at((0, 3, -3), (0, 0, 0), (0, 1, 0)), fov=1.0)  # hits the mirror at the origin
    film = host.Film(1, 1)
    smp = host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False)
    o = _oracle(sb, cam, film, smp, host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=5))
    o.render(0, 1, 1)
    # the reflected ray (0,3,-3)->(0,0,0) continues as (0,1,1)/sqrt2 and meets the wall at z = 5: (0, 5, 5)
    p = np.array([0.0, 5.0, 5.0])
    to_l = np.array([0.0, 3.0, 2.0]) - p
    d2 = float(to_l @ to_l)
    cos = abs(to_l[2]) / math.sqrt(d2)
    expect = 0.8 * (kd / math.pi) * I / d2 * cos
    assert np.allclose(_centre(o, film), expect, rtol=1e-4), (_centre(o, film), expect)
    # without recursion depth the mirror shows nothing
    o1 = _oracle(sb, cam, film, smp, host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=1))
    o1.render(0, 1, 1)
    assert np.all(_centre(o1, film) == 0)


# ---- the other delta lights (distant_light.dart:41-48, spot_light.dart:36-70) ---------------------------------------
def test_distant_and_spot_lights_match_their_closed_forms():
    kd = 0.5
    cam = host.PerspectiveCamera(host.look_at((0, 10, 0), (0, 0, 0), (0, 0, 1)), fov=1.0)  # looks straight down at the origin
    film = host.Film(1, 1)
    smp = host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False)
    integ = host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=1)
    # distant light arriving 60 degrees off the normal: L * kd/pi * cos(60)
    sb = host.SceneBuilder()
    _floor(sb, 0.0, material=sb.material((kd, kd, kd)))
    sb.distant_light((0.0, 1.0, math.sqrt(3.0)), (0, 0, 0), (3.0, 2.0, 1.0))
    o = _oracle(sb, cam, film, smp, integ)
    o.render(0, 1, 1)
    assert np.allclose(_centre(o, film), np.array([3.0, 2.0, 1.0]) * kd / math.pi * 0.5, rtol=2e-5)
    # an occluder anywhere along the infinite shadow ray blocks it (visibility_tester.dart:31-33)
    _quad(sb, [[-50, 40, 30], [50, 40, 30], [50, 40, 110], [-50, 40, 110]])
    o = _oracle(sb, cam, film, smp, integ)
    o.render(0, 1, 1)
    assert np.all(_centre(o, film) == 0)
    # spot light above the origin pointing down: full intensity inside the falloff start, smooth step between, zero outside
    I, h = 8.0, 4.0
    for off, cone, delta in ((0.0, 30.0, 5.0), (2.0, 30.0, 5.0), (4.0, 30.0, 5.0)):
        sb = host.SceneBuilder()
        _floor(sb, 0.0, material=sb.material((kd, kd, kd)))
        sb.spot_light((off, h, 0.0), (off, 0.0, 0.0), (I, I, I), cone, delta)
        o = _oracle(sb, cam, film, smp, integ)
        o.render(0, 1, 1)
        d2 = off * off + h * h
        cos_l = h / math.sqrt(d2)  # angle at the light between its axis and the direction to the origin
        ct, cf = math.cos(math.radians(cone)), math.cos(math.radians(cone - delta))
        fall = 0.0 if cos_l < ct else (1.0 if cos_l > cf else ((cos_l - ct) / (cf - ct)) ** 4)
        expect = I * fall / d2 * kd / math.pi * cos_l  # surface cosine equals cos_l here (floor normal parallel to the axis)
        assert np.allclose(_centre(o, film), expect, rtol=1e-4, atol=1e-9), (off, _centre(o, film), expect)


# ---- whitted integrator (whitted_integrator.dart:26-78) ----------------------------------------------------------------
def test_whitted_point_light_and_mirror_recursion_match_closed_forms():
    kd, I = 0.6, 20.0
    smp = host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False)
    film = host.Film(1, 1)
    # matte floor under a point light, seen straight down: L = kd/pi * I/h^2 (one light sample, no MIS)
    sb = host.SceneBuilder()
    _floor(sb, 0.0, material=sb.material((kd, kd, kd)))
    sb.point_light((0, 3, 0), (I, I, I))
    cam = host.PerspectiveCamera(host.look_at((0, 10, 0), (0, 0, 0), (0, 0, 1)), fov=1.0)
    o = _oracle(sb, cam, film, smp, host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=5))
    o.render(0, 1, 1)
    assert np.allclose(_centre(o, film), kd / math.pi * I / 9.0, rtol=2e-5)
    # the same lit wall seen through a mirror (the scene of test_directlighting_sees_a_lit_matte_wall_through_a_mirror)
    sb = host.SceneBuilder()
    _floor(sb, 0.0, half=4.0, material=sb.material_lobes(host.mirror_lobes(0.8)))
    _quad(sb, [[-30, -30, 5], [30, -30, 5], [30, 30, 5], [-30, 30, 5]], material=sb.material_lobes(host.matte_lobes(kd)))
    sb.point_light((0, 3, 2), (I, I, I))
    cam = host.PerspectiveCamera(host.look_at((0, 3, -3), (0, 0, 0), (0, 1, 0)), fov=1.0)
    o = _oracle(sb, cam, film, smp, host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=5))
    o.render(0, 1, 1)
    to_l = np.array([0.0, 3.0, 2.0]) - np.array([0.0, 5.0, 5.0])
    d2 = float(to_l @ to_l)
    expect = 0.8 * (kd / math.pi) * I / d2 * abs(to_l[2]) / math.sqrt(d2)
    assert np.allclose(_centre(o, film), expect, rtol=1e-4), (_centre(o, film), expect)
    o1 = _oracle(sb, cam, film, smp, host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=1))
    o1.render(0, 1, 1)
    assert np.all(_centre(o1, film) == 0)


# ---- translucent / mix / shinymetal (BRDFToBTDF, ScaledBxDF, FresnelApproxEta) ---------------------------------------------
def _sheet_radiance(lobes, light_y, integ=None, sky=None):
    """Radiance leaving the centre of a horizontal sheet (normal +y) towards a camera ABOVE it, lit by a point light at height
    light_y on the axis (negative: behind the sheet) or by a constant sky."""
    sb = host.SceneBuilder()
    m = sb.material_lobes(lobes)
    sb.mesh([[-20, 0, -20], [20, 0, -20], [20, 0, 20], [-20, 0, 20]], [[0, 2, 1], [0, 3, 2]], material=m)
    if sky is not None:
        sb.infinite_light((sky, sky, sky), nsamples=4)
    else:
        sb.point_light((0.0, light_y, 0.0), (30.0, 30.0, 30.0))
    cam = host.PerspectiveCamera(host.look_at((0.0, 6.0, 0.01), (0, 0, 0), (0, 0, 1)), fov=0.5)
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, host.Film(2, 2), host.Sampler(kind=host.SAMPLER_LD, spp=16),
                          integ or host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=3))
    o.render()
    return o.film_read()["rgb"].mean(axis=(0, 1))


def test_translucent_transmits_with_its_btdf_lobes():
    # Lambertian(r * kd) towards the light's side, BRDFToBTDF(Lambertian(t * kd)) through the sheet (translucent_material.dart:62-70)
    kd, r, t, I, h = 0.8, 0.6, 0.3, 30.0, 3.0
    lobes = host.translucent_lobes(kd=kd, ks=0.0, reflect=r, transmit=t)
    assert [l["wrap"] for l in lobes] == [0, host.WRAP_BTDF]
    front = _sheet_radiance(lobes, +h)
    back = _sheet_radiance(lobes, -h)
    assert front == pytest.approx(r * kd / math.pi * I / (h * h), rel=1e-4)
    assert back == pytest.approx(t * kd / math.pi * I / (h * h), rel=1e-4)
    # the glossy pair: reflection only from the front, transmission only from behind, equal lobes up to reflect / transmit
    g = host.translucent_lobes(kd=0.0, ks=0.5, reflect=0.5, transmit=0.25, roughness=0.3)
    assert [l["wrap"] for l in g] == [0, host.WRAP_BTDF]
    gf, gb = _sheet_radiance(g, +h), _sheet_radiance(g, -h)
    assert gf[0] > 0 and gb[0] == pytest.approx(0.5 * gf[0], rel=1e-4)


def test_mix_material_scales_both_bsdfs():
    a, b, s, I, h = 0.9, 0.2, 0.3, 30.0, 3.0
    lobes = host.mix_lobes(host.matte_lobes(a), host.matte_lobes(b), amount=s)
    assert all(l["wrap"] == host.WRAP_SCALED for l in lobes)
    L = _sheet_radiance(lobes, +h)
    assert L == pytest.approx((s * a + (1 - s) * b) / math.pi * I / (h * h), rel=1e-4)
    # amount 1 / 0 reduce to the single materials; a mix of mixes is refused
    one = _sheet_radiance(host.mix_lobes(host.matte_lobes(a), host.matte_lobes(b), amount=1.0), +h)
    assert one == pytest.approx(_sheet_radiance(host.matte_lobes(a), +h), rel=1e-6)
    with pytest.raises(ValueError):
        host.mix_lobes(lobes, host.matte_lobes(b))
    # ScaledBxDF keeps BxDF.pdf's cosine density (scaled_bxdf.dart has no pdf override): a scaled glossy lobe under a sky is still
    # an unbiased estimate as long as the light-sampling half carries it — compare a path-traced furnace against the unscaled lobe
    gl = host.plastic_lobes(0.0, 0.6, 0.2)
    half = host.mix_lobes(gl, gl, amount=0.5)
    full = _sheet_radiance(gl, 0, host.Integrator(kind=host.INTEGRATOR_DIRECT), sky=1.0)
    mixed = _sheet_radiance(half, 0, host.Integrator(kind=host.INTEGRATOR_DIRECT), sky=1.0)
    assert mixed == pytest.approx(full, rel=0.1)


def test_shinymetal_reflectance_at_normal_incidence():
    # FresnelApproxEta (shiny_metal_material.dart:66-70) is built so that the conductor's normal-incidence reflectance is Kr:
    # a mirror-like sheet seen head-on under a constant sky returns Kr * L (clamped at 0.999)
    for kr in (0.2, 0.7, 1.0):
        L = _sheet_radiance(host.shinymetal_lobes(ks=0.0, kr=kr), 0, host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=2), sky=2.0)
        assert L == pytest.approx(min(kr, 0.999) * 2.0, rel=2e-3)


def test_substrate_fresnel_blend_albedo_matches_numeric_quadrature():
    # FresnelBlend over an Anisotropic distribution (fresnel_blend.dart:30-58, anisotropic.dart:39-50): the radiance of a substrate
    # sheet under a constant unit sky is its directional albedo, which a brute-force quadrature of f * cos reproduces; the render
    # gets there through sample_f / pdf (cosine + anisotropic half-vector sampling) and the sky's light sampling with MIS
    Rd, Rs, ex, ey = np.array([0.5, 0.3, 0.2]), np.array([0.05, 0.1, 0.3]), 1.0 / 0.2, 1.0 / 0.05
    lobes = host.substrate_lobes(kd=Rd, ks=Rs, uroughness=0.2, vroughness=0.05)
    assert len(lobes) == 1 and lobes[0]["kind"] == host.LOBE_FRESNEL_BLEND
    L = _sheet_radiance(lobes, 0, host.Integrator(kind=host.INTEGRATOR_DIRECT), sky=1.0)

    # camera of _sheet_radiance: at (0, 6, 0.01) looking at the origin -> wo in the sheet's shading frame
    wo_w = np.array([0.0, 6.0, 0.01])
    wo_w /= np.linalg.norm(wo_w)
    # shading frame of the quad [[-20,0,-20],[20,0,-20],[20,0,20],[-20,0,20]] with indices (0,2,1),(0,3,2) and default uvs:
    # sn = normalize(dpdu), nn = +y; the albedo of this lobe depends on the azimuth of wo only through ex != ey, and wo is 0.1
    # degrees off the normal, so any in-plane orientation of the frame gives the same number to the test's tolerance
    wo = np.array([wo_w[2], wo_w[0], wo_w[1]])
    n_th, n_ph = 400, 800
    th = (np.arange(n_th) + 0.5) / n_th * (np.pi / 2)
    ph = (np.arange(n_ph) + 0.5) / n_ph * (2 * np.pi)
    T, P = np.meshgrid(th, ph, indexing="ij")
    wi = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], axis=-1)
    wh = wi + wo
    wh /= np.linalg.norm(wh, axis=-1, keepdims=True)
    cos_h = np.abs(wh[..., 2])
    d = 1 - cos_h ** 2
    e = (ex * wh[..., 0] ** 2 + ey * wh[..., 1] ** 2) / np.maximum(d, 1e-300)
    D = np.where(d == 0, 0.0, np.sqrt((ex + 2) * (ey + 2)) / (2 * np.pi) * cos_h ** e)
    wi_h = np.abs((wi * wh).sum(-1))
    a = D / (4 * wi_h * np.maximum(wi[..., 2], wo[2]))
    schlick = Rs + (1 - Rs) * ((1 - (wi * wh).sum(-1)) ** 5)[..., None]
    diffuse = Rd * (28 / (23 * np.pi)) * (1 - Rs) * ((1 - (1 - 0.5 * wi[..., 2]) ** 5) * (1 - (1 - 0.5 * wo[2]) ** 5))[..., None]
    f = diffuse + schlick * a[..., None]
    rho = (f * (wi[..., 2] * np.sin(T))[..., None]).sum(axis=(0, 1)) * (np.pi / 2 / n_th) * (2 * np.pi / n_ph)
    assert np.allclose(L, rho, rtol=3e-2), (L, rho)
