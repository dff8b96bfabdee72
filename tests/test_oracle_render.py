"""Pins the CPU oracle's render half (oracle/ref_render.cpp) with analytic known answers and sequence
identities (SURVEY §8c "what pins results instead"): the reference ships no golden vectors, so these
closed forms are what stands between the restatement and a silent mistake."""
import math

import numpy as np
import pytest

from dartray_b200 import host, scenes
from tests.oracle_lib import Oracle, lib


def _oracle(sb, cam, film, sampler, integ):
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, film, sampler, integ)
    return o


def _cam(pos=(0, 0, -5), look=(0, 0, 0), fov=40.0):
    return host.PerspectiveCamera(host.look_at(pos, look, (0, 1, 0)), fov=fov)


def _plane(sb, y=0.0, half=50.0, material=0, **kw):
    # normal +y (triangle normal = normalize(cross(dpdu, dpdv)) with the reference's default uvs)
    return sb.mesh([[-half, y, -half], [half, y, -half], [half, y, half], [-half, y, half]], [[0, 2, 1], [0, 3, 2]],
                   material=material, **kw)


# ---- sequences (montecarlo.dart:486-551) ----------------------------------------------------------------
@pytest.mark.parametrize("mode", [host.RNG_KEYED, host.RNG_SERIAL])
def test_ld_image_samples_form_a_02_net(mode):
    sb, cam = scenes.cornell_synth()
    o = _oracle(sb, cam, host.Film(16, 16), host.Sampler(kind=host.SAMPLER_LD, spp=16, rng_mode=mode),
                host.Integrator(kind=host.INTEGRATOR_PATH))
    s = o.pixel_samples(3, 5)
    assert s.shape == (16, 5 + 14 + 18)  # path integrator: 14 one-D + 9 two-D arrays
    assert (s >= 0).all() and (s < 1).all()
    xy = s[:, :2].astype(np.float64)
    for kx in range(5):  # every elementary interval 2^-kx x 2^-(4-kx) holds exactly one of the 16 points
        ky = 4 - kx
        cell = np.floor(xy[:, 0] * (1 << kx)).astype(int) * (1 << ky) + np.floor(xy[:, 1] * (1 << ky)).astype(int)
        assert sorted(cell) == list(range(16)), (kx, ky)
    # each 1D array is a scrambled van der Corput set: one point per 1/16 stratum
    for col in range(5, 5 + 14):
        assert sorted(np.floor(s[:, col] * 16).astype(int)) == list(range(16))


def test_ld_rounds_pixel_samples_up_to_a_power_of_two():
    sb, cam = scenes.cornell_synth()
    o = _oracle(sb, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_LD, spp=5), host.Integrator(kind=host.INTEGRATOR_AO))
    assert o.pixel_samples(0, 0).shape == (8, 5 + 2)  # AO requests nothing; the emission volume integrator asks for 2


def test_stratified_samples_land_in_their_strata():
    sb, cam = scenes.cornell_synth()
    o = _oracle(sb, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=4, ys=3),
                host.Integrator(kind=host.INTEGRATOR_DIRECT))
    s = o.pixel_samples(2, 1)
    assert s.shape[0] == 12
    ix, iy = np.floor(s[:, 0] * 4).astype(int), np.floor(s[:, 1] * 3).astype(int)
    assert sorted(iy * 4 + ix) == list(range(12))  # image samples stay in stratum order (x fastest)
    assert list(iy * 4 + ix) == list(range(12))
    lens = np.floor(s[:, 2] * 4).astype(int) + 4 * np.floor(s[:, 3] * 3).astype(int)
    assert sorted(lens) == list(range(12))         # lens samples are stratified then shuffled
    assert sorted(np.floor(s[:, 4] * 12).astype(int)) == list(range(12))
    assert s.max() <= np.float32(0.9999999403953552)  # ONE_MINUS_EPSILON clamp, montecarlo.dart:23
    nj = _oracle(sb, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2, jitter=False),
                 host.Integrator(kind=host.INTEGRATOR_AO)).pixel_samples(0, 0)
    assert np.allclose(np.sort(nj[:, 0]), [0.25, 0.25, 0.75, 0.75])


def test_dart_random_restatement_is_deterministic_and_in_range():
    L = lib()
    f1, u1 = np.zeros(64), np.zeros(64, np.uint32)
    f2, u2 = np.zeros(64), np.zeros(64, np.uint32)
    L.orc_dart_random(5489, 64, f1.ctypes.data, u1.ctypes.data)
    L.orc_dart_random(5489, 64, f2.ctypes.data, u2.ctypes.data)
    assert np.array_equal(f1, f2) and np.array_equal(u1, u2)
    assert (f1 >= 0).all() and (f1 < 1).all() and (u1 < 0xffffffff).all()
    L.orc_dart_random(5490, 64, f2.ctypes.data, u2.ctypes.data)
    assert not np.array_equal(f1, f2)
    assert len(set(u1.tolist())) > 60


# ---- integrators: closed forms ------------------------------------------------------------------------------
def test_ambient_occlusion_open_plane_is_one_and_closed_box_is_zero():
    sb = host.SceneBuilder()
    _plane(sb, y=0.0)
    cam = _cam(pos=(0, 3, -6), look=(0, 0, 0))
    o = _oracle(sb, cam, host.Film(16, 12), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False),
                host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=32))
    o.render()
    f = o.film_read()
    hit = f["rgb"][8:, :, 0]  # lower half of the image looks at the plane
    assert np.allclose(hit, 1.0, rtol=0, atol=2e-6)  # RGB -> XYZ -> RGB through the film is not an exact identity
    # inside a closed sphere every hemisphere ray is blocked
    sb = host.SceneBuilder()
    sb.sphere(host.translate(0, 0, 0), radius=4.0)
    o = _oracle(sb, _cam(pos=(0, 0, -1)), host.Film(8, 8), host.Sampler(kind=host.SAMPLER_LD, spp=1),
                host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=16))
    o.render()
    assert (o.film_read()["rgb"] == 0.0).all()
    assert o.render_stats()["shadow_rays"] == 81 * 16  # 9 x 9 sample extent, every camera ray hits


def test_direct_lighting_point_light_on_a_matte_plane_closed_form():
    kd, inten, h = 0.6, 50.0, 4.0
    sb = host.SceneBuilder()
    _plane(sb, y=0.0, material=sb.material((kd, kd, kd)))
    sb.point_light((0.0, h, 0.0), (inten, inten, inten))
    cam = _cam(pos=(0, 6, -8), look=(0, 0, 0), fov=30.0)
    film = host.Film(33, 33)
    o = _oracle(sb, cam, film, host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False),
                host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    rgb = o.film_read()["rgb"]
    # the centre pixel looks at the origin: L = Kd/pi * I / h^2 * cos(0)
    assert rgb[16, 16, 0] == pytest.approx(kd / math.pi * inten / (h * h), rel=2e-3)
    # everywhere on the plane: L(x) = Kd/pi * I * h / (h^2 + r^2)^(3/2); check the image's maximum is at the centre
    assert np.unravel_index(np.argmax(rgb[:, :, 0]), rgb.shape[:2]) == (16, 16)


def test_direct_lighting_quad_light_matches_form_factor():
    # unoccluded matte point under a parallel square light: L = Kd/pi * Le * integral(cos cos / r^2 dA)
    kd, Le, h, a = 0.5, 10.0, 2.0, 1.0
    sb = host.SceneBuilder()
    _plane(sb, y=0.0, material=sb.material((kd, kd, kd)))
    sb.mesh([[-a, h, -a], [a, h, -a], [a, h, a], [-a, h, a]], [[0, 1, 2], [0, 2, 3]], area_light=(Le, Le, Le), nsamples=16)
    cam = _cam(pos=(0.2, 1.5, -0.2), look=(0, 0, 0), fov=1.0)  # narrow, steep view of the origin from under the light
    o = _oracle(sb, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=64), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    xs = (np.arange(400) + 0.5) / 400 * 2 * a - a
    X, Z = np.meshgrid(xs, xs)
    r2 = X * X + Z * Z + h * h
    E = Le * np.sum((h * h) / (r2 * r2)) * (2 * a / 400) ** 2
    assert o.film_read()["rgb"].mean() == pytest.approx(kd / math.pi * E, rel=1e-2)


@pytest.mark.parametrize("rho,Le,maxdepth", [(0.5, 1.0, 5), (0.8, 2.0, 3)])
def test_path_white_furnace(rho, Le, maxdepth):
    # closed emissive matte sphere seen from inside: L = Le * sum_{k=0}^{maxdepth+1} rho^k
    # (emission at the first vertex + one direct-lighting estimate per vertex, path_integrator.dart:44-119)
    sb = host.SceneBuilder()
    sb.sphere(host.translate(0, 0, 0), radius=5.0, material=sb.material((rho, rho, rho)), area_light=(Le, Le, Le), reverse=True)
    o = _oracle(sb, _cam(pos=(0, 0, -1), fov=60.0), host.Film(24, 24), host.Sampler(kind=host.SAMPLER_LD, spp=32),
                host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=maxdepth))
    o.render()
    expect = Le * sum(rho ** k for k in range(maxdepth + 2))
    assert o.film_read()["rgb"].mean() == pytest.approx(expect, rel=3e-3)


# ---- film (image_film.dart:99-185,247-299) -----------------------------------------------------------------
def test_box_filter_weight_equals_sample_count_and_sample_extent():
    sb, cam = scenes.cornell_synth()
    o = _oracle(sb, cam, host.Film(20, 14), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    f = o.film_read()
    assert (f["weight"] == 4.0).all()
    assert o.render_stats()["camera_samples"] == 21 * 15 * 4  # sample extent is one pixel wider and taller
    o2 = _oracle(sb, cam, host.Film(20, 14, filter="gaussian"), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                 host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o2.render()
    assert o2.render_stats()["camera_samples"] == 25 * 19 * 4  # width-2 filter: floor(.5-2) .. ceil(.5+w+2)
    g = o2.film_read()
    assert g["rgb"].shape == (14, 20, 3) and np.isfinite(g["rgb"]).all()
    assert abs(g["rgb"].mean() - f["rgb"].mean()) < 0.1 * f["rgb"].mean()  # blur at a 20 x 14 image border


def test_filter_tables_match_their_closed_forms():
    for name, centre in (("box", 1.0), ("triangle", (2 - 1 / 16) ** 2), ("gaussian", None), ("mitchell", None), ("sinc", None)):
        xw, yw, t = host.filter_table(name)
        assert t.shape == (256,) and np.isfinite(t).all()
        assert np.allclose(t.reshape(16, 16), t.reshape(16, 16).T)  # separable and symmetric in x/y
        if centre is not None:
            assert t[0] == pytest.approx(centre, rel=1e-6)
    assert host.filter_table("gaussian")[2][255] >= 0.0


def test_tasks_cover_the_sample_extent_once():
    sb, cam = scenes.cornell_synth()
    film, sampler, integ = host.Film(30, 22), host.Sampler(kind=host.SAMPLER_LD, spp=2), host.Integrator(kind=host.INTEGRATOR_DIRECT)
    whole = _oracle(sb, cam, film, sampler, integ)
    whole.render()
    parts = _oracle(sb, cam, film, sampler, integ)
    for t in range(4):
        parts.render(t, 4, 2)
    a, b = whole.film_read(), parts.film_read()
    assert np.array_equal(a["weight"], b["weight"])
    assert np.allclose(a["xyz"], b["xyz"], rtol=1e-5, atol=1e-6)


def test_keyed_and_serial_streams_agree_statistically():
    # The GPU replays KEYED streams; the reference runs one SERIAL stream.  Same estimator, different
    # random numbers: image means agree within Monte Carlo noise.
    sb, cam = scenes.cornell_synth()
    means = []
    for mode in (host.RNG_KEYED, host.RNG_SERIAL):
        o = _oracle(sb, cam, host.Film(48, 36), host.Sampler(kind=host.SAMPLER_LD, spp=16, rng_mode=mode),
                    host.Integrator(kind=host.INTEGRATOR_PATH))
        o.render(0, 1, 8)
        means.append(o.film_read()["rgb"].mean(axis=(0, 1)))
    assert np.allclose(means[0], means[1], rtol=0.02)


# ---- per-vertex shading attributes: Triangle.getShadingGeometry (triangle.dart:271-364) -------------------------------
def _lit_quad(N=None, uv=None, S=None, reverse=False, idx=((0, 2, 1), (0, 3, 2))):
    """A matte quad in the plane y = 0 under a point light straight above the point the camera looks at."""
    kd, I, hgt = 0.6, 50.0, 5.0
    sb = host.SceneBuilder()
    sb.mesh([[-4, 0, -4], [4, 0, -4], [4, 0, 4], [-4, 0, 4]], idx, material=sb.material((kd, kd, kd)), N=N, uv=uv, S=S, reverse=reverse)
    sb.point_light((0.0, hgt, 0.0), (I, I, I))
    cam = host.PerspectiveCamera(host.look_at((0.3, 2.0, -0.2), (0, 0, 0), (0, 1, 0)), fov=0.2)
    o = _oracle(sb, cam, host.Film(2, 2), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False),
                host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    return o.film_read()["rgb"].mean(), kd / math.pi * I / (hgt * hgt)


def test_vertex_normals_tilt_the_shading_normal_closed_form():
    flat, base = _lit_quad()
    assert flat == pytest.approx(base, rel=1e-4)  # |cos| = 1 with the geometric normal
    t = 1.0 / math.sqrt(2.0)
    tilted, _ = _lit_quad(N=[[t, t, 0]] * 4)
    assert tilted == pytest.approx(base * t, rel=1e-4)  # L = f * Li * |wi . ns| with ns = normalize(1, 1, 0)
    # unnormalised N are normalised after the transform (triangle.dart:301-303); S alone leaves ns = dg.nn (:304-306)
    assert _lit_quad(N=[[3, 3, 0]] * 4)[0] == pytest.approx(tilted, rel=1e-6)
    assert _lit_quad(S=[[1, 0, 0]] * 4)[0] == pytest.approx(base, rel=1e-4)
    # interpolation: N = +y at x = -4 and normalize(1, 1, 0) at x = +4 -> at x = 0 the blend (t/2, (1 + t)/2, 0), normalised
    mixed, _ = _lit_quad(N=[[0, 1, 0], [t, t, 0], [t, t, 0], [0, 1, 0]])
    nx, ny = t / 2, (1 + t) / 2
    assert mixed == pytest.approx(base * ny / math.hypot(nx, ny), rel=2e-3)


def test_uv_orientation_decides_the_side_an_area_light_emits_to():
    # dg.nn = normalize(cross(dpdu, dpdv)) follows the mesh's uvs (triangle.dart:104-131); DiffuseAreaLight emits where
    # dot(nn, w) > 0 (diffuse_area_light.dart): mirroring the parameterisation turns the light round
    def floor_radiance(uv):
        sb = host.SceneBuilder()
        _plane(sb, material=sb.material((0.5, 0.5, 0.5)))
        sb.mesh([[-1, 3, -1], [1, 3, -1], [1, 3, 1], [-1, 3, 1]], [[0, 1, 2], [0, 2, 3]], area_light=(10.0, 10.0, 10.0), nsamples=4, uv=uv)
        cam = host.PerspectiveCamera(host.look_at((0.5, 1.0, -3.0), (0, 0, 0), (0, 1, 0)), fov=5.0)
        o = _oracle(sb, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=16), host.Integrator(kind=host.INTEGRATOR_DIRECT))
        o.render()
        return o.film_read()["rgb"].mean()
    default = floor_radiance(None)
    same = floor_radiance([[0, 0], [1, 0], [1, 1], [0, 1]])
    mirrored = floor_radiance([[1, 0], [0, 0], [0, 1], [1, 1]])
    assert (default > 0.1) != (mirrored > 0.1) or (same > 0.1) != (mirrored > 0.1)
    assert (same > 0.1) != (mirrored > 0.1)
    assert min(same, mirrored) == 0.0


# ---- InfiniteAreaLight (infinite_area_light.dart) ----------------------------------------------------------------------
def test_constant_infinite_light_closed_forms():
    # an open matte plane under a constant sky: Lo = kd * L for direct lighting, every camera ray that escapes sees L, and a
    # white furnace closed by the sky converges to L / (1 - rho) ... here with rho = kd and one bounce of sky light per vertex
    kd, Lsky = 0.4, 2.0
    sb = host.SceneBuilder()
    _plane(sb, material=sb.material((kd, kd, kd)))
    sb.infinite_light((Lsky, Lsky, Lsky), nsamples=4)
    cam = host.PerspectiveCamera(host.look_at((0, 2, -4), (0, 0, 0), (0, 1, 0)), fov=10.0)
    o = _oracle(sb, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=64), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    assert o.film_read()["rgb"].mean() == pytest.approx(kd * Lsky, rel=1.5e-2)  # irradiance pi * L, BRDF kd / pi
    # looking up at the sky: Li = sum of lights.Le(ray) (sampler_renderer.dart:86-92), exact up to the bilinear weights' rounding
    cam = host.PerspectiveCamera(host.look_at((0, 2, -4), (0, 9, 0), (0, 0, 1)), fov=10.0)
    o = _oracle(sb, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_PATH))
    o.render()
    assert np.allclose(o.film_read()["rgb"], Lsky, rtol=1e-6)
    # path tracing on the open plane: only direct light reaches it (the plane cannot see itself)
    cam = host.PerspectiveCamera(host.look_at((0, 2, -4), (0, 0, 0), (0, 1, 0)), fov=10.0)
    o = _oracle(sb, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=256), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3))
    o.render()
    assert o.film_read()["rgb"].mean() == pytest.approx(kd * Lsky, rel=1.5e-2)


def test_infinite_light_map_is_importance_sampled_and_oriented():
    # a lat-long map that is black except for one bright texel block near the zenith of the LIGHT's frame (+z, v = 0): with the
    # light-to-world rotation that takes +z to +y the plane is lit from above; direct lighting (importance sampled through
    # Distribution2D) and the path tracer's BSDF sampling must agree on the answer
    w, h = 16, 8
    tex = np.zeros((h, w, 3), np.float32)
    tex[0:2, :, :] = 5.0  # theta in [0, pi / 4]: a polar cap
    kd = 0.5
    sb = host.SceneBuilder()
    _plane(sb, material=sb.material((kd, kd, kd)))
    sb.infinite_light((1.0, 1.0, 1.0), nsamples=8, light_to_world=host.rotate(-90, (1, 0, 0)), texels=tex)
    cam = host.PerspectiveCamera(host.look_at((0, 2, -4), (0, 0, 0), (0, 1, 0)), fov=10.0)
    res = {}
    for name, integ, spp in (("direct", host.Integrator(kind=host.INTEGRATOR_DIRECT), 64),
                             ("path", host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=2), 512)):
        o = _oracle(sb, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=spp), integ)
        o.render()
        res[name] = o.film_read()["rgb"].mean()
    # cap of half-angle pi / 4 with (bilinearly blurred) radiance 5: E ~ 5 * pi * sin^2(pi / 4) = 2.5 pi -> Lo ~ kd * 2.5; the
    # blur across the cap's edge moves that by a few per cent, so pin the two estimators against each other tightly and the
    # closed form loosely
    assert res["direct"] == pytest.approx(res["path"], rel=3e-2)
    assert res["direct"] == pytest.approx(kd * 2.5, rel=0.15)
    # rotated the other way the cap is below the horizon: the plane is dark
    sb2 = host.SceneBuilder()
    _plane(sb2, material=sb2.material((kd, kd, kd)))
    sb2.infinite_light((1.0, 1.0, 1.0), nsamples=8, light_to_world=host.rotate(90, (1, 0, 0)), texels=tex)
    o = _oracle(sb2, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=16), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    assert o.film_read()["rgb"].mean() < 1e-3 * res["direct"]


# ---- HaltonSampler (halton_sampler.dart) --------------------------------------------------------------------------------
def test_halton_sampler_radical_inverses_and_window():
    # a white furnace-like flat field: every accepted sample returns the same radiance, so the box-filtered image is constant
    # wherever samples land and the film weights count the samples: spp * delta^2 * (accepted area / delta^2)
    sb = host.SceneBuilder()
    sb.infinite_light((1.0, 1.0, 1.0))
    cam = host.PerspectiveCamera(host.look_at((0, 0, -5), (0, 0, 0), (0, 1, 0)), fov=40.0)
    spp, W, H = 4, 12, 8
    o = _oracle(sb, cam, host.Film(W, H, filter="box", xwidth=0.5, ywidth=0.5), host.Sampler(kind=host.SAMPLER_HALTON, spp=spp),
                host.Integrator(kind=host.INTEGRATOR_PATH))
    o.render()
    f = o.film_read()
    st = o.render_stats()
    # the sample window is [0, W] x [0, H] (box filter 0.5: image_film.dart:247-252) -> delta = W + 1 = 13, wanted = 4 * 169 indices;
    # accepted: imageX <= W and imageY <= H - ... the INCLUSIVE right / bottom quirk (see the oracle): x <= 12, y <= 8
    delta = W + 1
    n = np.arange(spp * delta * delta)
    def radical_inverse(n, base):
        out = np.zeros(n.shape)
        inv = 1.0 / base
        f_ = inv
        n = n.copy()
        while n.any():
            out += (n % base) * f_
            n = (n * inv).astype(np.int64)  # the reference truncates a double product
            f_ *= inv
        return out
    ix, iy = radical_inverse(n, 3) * delta, radical_inverse(n, 2) * delta
    acc = (ix <= W) & (iy <= H)
    assert st["camera_samples"] == int(acc.sum())
    assert np.allclose(f["rgb"][f["weight"] > 0], 1.0, rtol=1e-6)
    # per-pixel sample counts: box filter of half-width 0.5 -> the pixel whose centre is within 0.5 of the sample
    cnt = np.zeros((H, W))
    for x_, y_ in zip(ix[acc], iy[acc]):
        dx, dy = x_ - 0.5, y_ - 0.5
        for py in range(max(int(math.ceil(dy - 0.5)), 0), min(int(math.floor(dy + 0.5)), H - 1) + 1):
            for px in range(max(int(math.ceil(dx - 0.5)), 0), min(int(math.floor(dx + 0.5)), W - 1) + 1):
                cnt[py, px] += 1
    assert np.array_equal(f["weight"], cnt.astype(np.float32))


# ---- AdaptiveSampler (adaptive_sampler.dart) -----------------------------------------------------------------------------
def _adaptive_scene():
    # a bright emitter quad in front of a black void: pixels inside the quad and pixels of the void are uniform, the pixels its
    # outline crosses see both
    sb = host.SceneBuilder()
    sb.mesh([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], [[0, 2, 1], [0, 3, 2]], material=sb.material((0.5, 0.5, 0.5)),
            area_light=(4.0, 4.0, 4.0))
    cam = host.PerspectiveCamera(host.look_at((0.03, 0.02, -5), (0.03, 0.02, 0), (0, 1, 0)), fov=40.0)
    return sb, cam


@pytest.mark.parametrize("method", [host.ADAPTIVE_CONTRAST, host.ADAPTIVE_SHAPE_ID])
def test_adaptive_sampler_supersamples_only_where_the_samples_disagree(method):
    sb, cam = _adaptive_scene()
    W = H = 24
    smp = host.Sampler(kind=host.SAMPLER_ADAPTIVE, xs=3, ys=16, jitter=method)  # minsamples 3 -> 4, maxsamples 16
    o = _oracle(sb, cam, host.Film(W, H, filter="box", xwidth=0.5, ywidth=0.5), smp, host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=1))
    o.render()
    f = o.film_read()
    wt = f["weight"]
    assert set(np.unique(wt)) == {4.0, 16.0}  # box filter of half-width 0.5: the weight is the pixel's own sample count
    lit = f["rgb"][..., 0] > 0
    inside = lit & np.roll(lit, 1, 0) & np.roll(lit, -1, 0) & np.roll(lit, 1, 1) & np.roll(lit, -1, 1)
    outside = ~lit & ~np.roll(lit, 1, 0) & ~np.roll(lit, -1, 0) & ~np.roll(lit, 1, 1) & ~np.roll(lit, -1, 1)
    assert inside.sum() > 20 and outside.sum() > 100
    if method == host.ADAPTIVE_CONTRAST:
        assert (wt[inside] == 4.0).all()        # uniform radiance 4: no contrast
    else:
        assert (wt[inside] == 4.0).mean() > 0.5  # shape ids: the quad is two triangles, its diagonal is supersampled too
    assert (wt[outside] == 4.0).all()            # all misses: Lavg = 0 -> 0 / 0 = NaN > 0.5 is false; equal (absent) ids
    edge = lit & ~inside
    assert (wt[edge] == 16.0).mean() > 0.6       # the outline: radiance 4 and 0 in one pixel
    # the sample window is one pixel wider and taller than the film (image_film.dart:247-252): 2 * 24 + 1 all-miss pixels more
    assert o.render_stats()["camera_samples"] == int((wt == 4).sum() * 4 + (wt == 16).sum() * (4 + 16)) + (2 * W + 1) * 4
    assert np.allclose(f["rgb"][inside], 4.0, rtol=1e-6)


# ---- ProjectionLight / GoniometricLight (projection_light.dart, goniometric_light.dart) -----------------------------------
def _floor_under(light_fn, look=(0.0, 0.0, 0.0)):
    kd = 0.5
    sb = host.SceneBuilder()
    _plane(sb, material=sb.material((kd, kd, kd)))
    light_fn(sb)
    cam = host.PerspectiveCamera(host.look_at((look[0] + 0.01, 3.0, look[2] - 0.01), look, (0, 0, 1)), fov=0.5)
    o = _oracle(sb, cam, host.Film(2, 2), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False),
                host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    return o.film_read()["rgb"].mean(axis=(0, 1)), kd


def test_projection_light_window_and_map():
    I, h = 40.0, 4.0
    down = host.mat_mul(host.translate(0, h, 0), host.rotate(90, (1, 0, 0)))  # the light's +z points at the floor
    # no map: a point light inside the projected square (fov 60 -> half-width h * tan(30 deg) on the floor), nothing outside
    L, kd = _floor_under(lambda sb: sb.projection_light((I, I, I), fov=60.0, light_to_world=down))
    assert L[0] == pytest.approx(kd / math.pi * I / (h * h), rel=1e-4)
    edge = h * math.tan(math.radians(30.0))
    x_in, x_out = 0.9 * edge, 1.1 * edge
    L_in, _ = _floor_under(lambda sb: sb.projection_light((I, I, I), fov=60.0, light_to_world=down), look=(x_in, 0.0, 0.0))
    L_out, _ = _floor_under(lambda sb: sb.projection_light((I, I, I), fov=60.0, light_to_world=down), look=(x_out, 0.0, 0.0))
    d2 = h * h + x_in * x_in
    assert L_in[0] == pytest.approx(kd / math.pi * I * (h / math.sqrt(d2)) / d2, rel=1e-3) and L_out[0] == 0.0
    # a uniform map scales it; a two-colour map (left half red, right half blue) colours the two sides of the axis
    L_half, _ = _floor_under(lambda sb: sb.projection_light((I, I, I), fov=60.0, light_to_world=down, texels=np.full((4, 4, 3), 0.5, np.float32)))
    assert L_half[0] == pytest.approx(0.5 * L[0], rel=1e-5)
    tex = np.zeros((8, 8, 3), np.float32)
    tex[:, :4, 0] = 1.0
    tex[:, 4:, 2] = 1.0
    a, _ = _floor_under(lambda sb: sb.projection_light((I, I, I), fov=60.0, light_to_world=down, texels=tex), look=(0.5 * edge, 0.0, 0.0))
    b, _ = _floor_under(lambda sb: sb.projection_light((I, I, I), fov=60.0, light_to_world=down, texels=tex), look=(-0.5 * edge, 0.0, 0.0))
    e = 1e-3  # the bilinear lookup leaves a trace of the other half
    assert (a[0] > e) != (b[0] > e) and (a[2] > e) != (b[2] > e) and (a[0] > e) == (b[2] > e)


def test_goniometric_light_scales_a_point_light_by_direction():
    I, h = 40.0, 4.0
    at = host.translate(0, h, 0)
    plain, kd = _floor_under(lambda sb: sb.goniometric_light((I, I, I), light_to_world=at))
    assert plain[0] == pytest.approx(kd / math.pi * I / (h * h), rel=1e-4)  # no map: a point light
    quarter, _ = _floor_under(lambda sb: sb.goniometric_light((I, I, I), light_to_world=at, texels=np.full((2, 4, 3), 0.25, np.float32)))
    assert quarter[0] == pytest.approx(0.25 * plain[0], rel=1e-5)
    # the map's pole is the light's +y (the y / z swap of :74-76): rows near t = 1 face the floor below the light
    tex = np.zeros((16, 8, 3), np.float32)
    tex[8:] = 1.0
    tex2 = np.zeros((16, 8, 3), np.float32)
    tex2[:8] = 1.0
    off = (2.0, 0.0, 0.0)  # 27 degrees off the nadir: t = 0.85, rows 13 / 14
    ref, _ = _floor_under(lambda sb: sb.goniometric_light((I, I, I), light_to_world=at), look=off)
    below, _ = _floor_under(lambda sb: sb.goniometric_light((I, I, I), light_to_world=at, texels=tex), look=off)
    dark, _ = _floor_under(lambda sb: sb.goniometric_light((I, I, I), light_to_world=at, texels=tex2), look=off)
    assert below[0] == pytest.approx(ref[0], rel=1e-5) and dark[0] == 0.0
    # exactly at the nadir t = 1: the bilinear lookup blends the last row with row 0 (TEXTURE_REPEAT, mipmap.dart:183-204,341-355)
    nadir, _ = _floor_under(lambda sb: sb.goniometric_light((I, I, I), light_to_world=at, texels=tex))
    assert nadir[0] == pytest.approx(0.5 * plain[0], rel=5e-2)  # the camera looks a hair off the axis


# ---- BestCandidateSampler (best_candidate_sampler.dart) -------------------------------------------------------------------
@pytest.mark.parametrize("W,H", [(24, 12), (12, 24)])
def test_best_candidate_sampler_places_the_pattern_and_rejects_as_written(W, H):
    from tests.util import synthetic_sample_table
    table = synthetic_sample_table()
    sb = host.SceneBuilder()
    sb.infinite_light((1.0, 1.0, 1.0))
    cam = host.PerspectiveCamera(host.look_at((0, 0, -5), (0, 0, 0), (0, 1, 0)), fov=40.0)
    spp = 4
    smp = host.Sampler(kind=host.SAMPLER_BEST_CANDIDATE, spp=spp, sample_table=table)
    o = _oracle(sb, cam, host.Film(W, H, filter="box", xwidth=0.5, ywidth=0.5), smp, host.Integrator(kind=host.INTEGRATOR_PATH))
    o.render()
    f, st = o.film_read(), o.render_stats()
    # window [0, W] x [0, H] (box filter 0.5), right = W inclusive; tiles of 64 / sqrt(4) = 32 pixels
    tw = 64 / math.sqrt(spp)
    nx, ny = int(W // tw) + 1, int(H // tw) + 1
    pts = []
    for ty in range(ny):
        for tx in range(nx):
            pts.append(np.stack([(tx + table[:, 0]) * tw, (ty + table[:, 1]) * tw], axis=1))
    pts = np.concatenate(pts)
    # :117-118 compares BOTH coordinates with left and right: a wide window keeps samples below its bottom edge (they miss the
    # film), a tall one loses everything below y = right
    acc = (pts[:, 0] >= 0) & (pts[:, 0] <= W) & (pts[:, 1] >= 0) & (pts[:, 1] <= W)
    assert st["camera_samples"] == int(acc.sum())
    cnt = np.zeros((H, W))
    for x_, y_ in pts[acc]:
        dx, dy = x_ - 0.5, y_ - 0.5
        for py in range(max(int(math.ceil(dy - 0.5)), 0), min(int(math.floor(dy + 0.5)), H - 1) + 1):
            for px in range(max(int(math.ceil(dx - 0.5)), 0), min(int(math.floor(dx + 0.5)), W - 1) + 1):
                cnt[py, px] += 1
    assert np.array_equal(f["weight"], cnt.astype(np.float32))
    if H > W:
        assert (f["weight"][W + 1:] == 0).all() and (f["weight"][:W - 1] > 0).any()
