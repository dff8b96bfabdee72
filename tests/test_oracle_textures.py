"""Pins the oracle's hit-point textures, bump mapping and ray differentials with numpy restatements written from the reference's
formulas (lib/core/mipmap.dart, lib/core/texture/*.dart, lib/textures/*.dart, lib/core/material.dart:35-88,
lib/core/differential_geometry.dart:122-205, lib/cameras/perspective_camera.dart:50-56,122-128) and with closed forms — nothing
here reads oracle/*.cpp.  CPU only."""
import math

import numpy as np
import pytest

from dartray_b200 import host
from tests.oracle_lib import Oracle

RNG = np.random.default_rng(20261017)


def _oracle_with(textures, spectrum):
    """An oracle holding the given texture objects as nodes; returns (oracle, node ids)."""
    table = host.TextureTable()
    ids = [table.add(t, sp) for t, sp in zip(textures, spectrum)]
    o = Oracle()
    o.set_textures(*table.arrays())
    return o, ids


def _dg(n, u=None, v=None, p=None, dudx=0.0, dvdx=0.0, dudy=0.0, dvdy=0.0, dpdx=None, dpdy=None):
    d = np.zeros((n, 15))
    if p is not None:
        d[:, 0:3] = p
    d[:, 3] = 0.0 if u is None else u
    d[:, 4] = 0.0 if v is None else v
    d[:, 5], d[:, 6], d[:, 7], d[:, 8] = dudx, dvdx, dudy, dvdy
    if dpdx is not None:
        d[:, 9:12] = dpdx
    if dpdy is not None:
        d[:, 12:15] = dpdy
    return d


# ---- numpy MIPMap (mipmap.dart), float images: Dart doubles on float32 texels ------------------------------------------------
def _wrap(i, n, mode):
    if mode == host.WRAP_REPEAT:
        return i % n, True
    if mode == host.WRAP_CLAMP:
        return min(max(i, 0), n - 1), True
    return i, 0 <= i < n


def _texel(level, s, t, mode):
    h, w = level.shape[:2]
    s2, oks = _wrap(s, w, mode)
    t2, okt = _wrap(t, h, mode)
    if not (oks and okt):
        return np.zeros(level.shape[2:], np.float64)
    return level[t2, s2].astype(np.float64)


def _triangle(level, s, t, mode):
    h, w = level.shape[:2]
    s, t = s * w - 0.5, t * h - 0.5
    s0, t0 = math.floor(s), math.floor(t)
    ds, dt = s - s0, t - t0
    return (_texel(level, s0, t0, mode) * ((1 - ds) * (1 - dt)) + _texel(level, s0, t0 + 1, mode) * ((1 - ds) * dt) +
            _texel(level, s0 + 1, t0, mode) * (ds * (1 - dt)) + _texel(level, s0 + 1, t0 + 1, mode) * (ds * dt))


def _n_levels(h, w):
    """mipmap.dart:143 with common.dart:98-103: 1 + (log(max) * (1 / log 2)).toInt() — the product is 2.9999999999999996 for 8 (and
    falls short for 64, 128, 4096 ...), so those resolutions get one level less and no 1 x 1 top."""
    return 1 + int(math.log(max(h, w)) * (1.0 / math.log(2.0)))


def _pyramid_float(img):
    levels = [img.astype(np.float32)]
    while len(levels) < _n_levels(*img.shape[:2]):
        p = levels[-1].astype(np.float64)
        h, w = p.shape
        sh, sw = max(1, h // 2), max(1, w // 2)
        nxt = np.zeros((sh, sw), np.float32)
        for t in range(sh):
            for s in range(sw):  # texel() with TEXTURE_REPEAT addressing (a 1-wide level repeats its only column)
                nxt[t, s] = np.float32((p[(2 * t) % h, (2 * s) % w] + p[(2 * t) % h, (2 * s + 1) % w] + p[(2 * t + 1) % h, (2 * s) % w] +
                                        p[(2 * t + 1) % h, (2 * s + 1) % w]) * 0.25)
        levels.append(nxt)
    return levels


_LUT = np.asarray([math.exp(-2.0 * (i / 127)) - math.exp(-2.0) for i in range(128)], np.float32)


def _ewa(level, s, t, ds0, dt0, ds1, dt1, mode):  # mipmap.dart:270-339
    h, w = level.shape[:2]
    s, t = s * w - 0.5, t * h - 0.5
    ds0, dt0, ds1, dt1 = ds0 * w, dt0 * h, ds1 * w, dt1 * h
    A, B, Cc = dt0 * dt0 + dt1 * dt1 + 1, -2.0 * (ds0 * dt0 + ds1 * dt1), ds0 * ds0 + ds1 * ds1 + 1
    invF = 1.0 / (A * Cc - B * B * 0.25)
    A, B, Cc = A * invF, B * invF, Cc * invF
    det = -B * B + 4.0 * A * Cc
    inv = 1.0 / det
    us, vs = math.sqrt(det * Cc), math.sqrt(A * det)
    s0, s1 = math.ceil(s - 2.0 * inv * us), math.floor(s + 2.0 * inv * us)
    t0, t1 = math.ceil(t - 2.0 * inv * vs), math.floor(t + 2.0 * inv * vs)
    acc, wsum = 0.0, 0.0
    for it in range(t0, t1 + 1):
        tt = it - t
        for si in range(s0, s1 + 1):
            ss = si - s
            r2 = A * ss * ss + B * ss * tt + Cc * tt * tt
            if r2 < 1.0:
                wgt = float(_LUT[int(min(r2 * 128, 127))])
                acc = acc + _texel(level, si, it, mode) * wgt
                wsum += wgt
    return acc / wsum


def _lookup2(levels, s, t, ds0, dt0, ds1, dt1, mode, trilinear, max_aniso):  # mipmap.dart:224-268 and :206-222
    n = len(levels)
    if trilinear:
        width = 2.0 * max(abs(ds0), abs(dt0), abs(ds1), abs(dt1))
        level = n - 1 + math.log(max(width, 1e-8)) / math.log(2.0)
        if level < 0:
            return _triangle(levels[0], s, t, mode)
        if level >= n - 1:
            return _texel(levels[-1], 0, 0, mode)
        il = math.floor(level)
        d = level - il
        return _triangle(levels[il], s, t, mode) * (1 - d) + _triangle(levels[il + 1], s, t, mode) * d
    if ds0 * ds0 + dt0 * dt0 < ds1 * ds1 + dt1 * dt1:
        ds0, ds1, dt0, dt1 = ds1, ds0, dt1, dt0
    major, minor = math.hypot(ds0, dt0), math.hypot(ds1, dt1)
    if minor * max_aniso < major and minor > 0:
        sc = major / (minor * max_aniso)
        ds1, dt1, minor = ds1 * sc, dt1 * sc, minor * sc
    if minor == 0:
        return _triangle(levels[0], s, t, mode)
    lod = max(0.0, n - 1.0 + math.log(minor) / math.log(2.0))
    il = math.floor(lod)
    d = lod - il

    def e(lv):
        return _texel(levels[-1], 0, 0, mode) if lv >= n else _ewa(levels[lv], s, t, ds0, dt0, ds1, dt1, mode)
    return e(il) * (1 - d) + e(il + 1) * d


def test_float_pyramid_is_the_box_average_and_the_spectrum_pyramid_loses_its_first_texel_as_written():
    """mipmap.dart:152-166.  A float image averages four doubles.  A spectrum image goes through SpectrumImage.operator[], which
    hands out ONE shared RGBColor (spectrum_image.dart:103-112,133-135): in `texel(a) + texel(b)` both operands are the same
    object when operator+ runs, so a level holds (2 b + c + d) / 4 of the finer one (b = texel(2s + 1, 2t))."""
    img = RNG.random((8, 16)).astype(np.float32)
    rgb = RNG.random((8, 8, 3)).astype(np.float32)
    o, ids = _oracle_with([host.ImageTexture(img), host.ImageTexture(rgb)], [False, True])
    got = o.image_levels(ids[0], 1)
    want = _pyramid_float(img)
    assert len(got) == len(want) == 5
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    got = o.image_levels(ids[1], 3)
    assert len(got) == _n_levels(8, 8) == 3 and np.array_equal(got[0], rgb)
    fine = rgb
    for lv in range(1, 3):
        h, w = fine.shape[:2]
        b, c, d = fine[0::2, 1::2], fine[1::2, 0::2], fine[1::2, 1::2]
        f32 = lambda x: x.astype(np.float32)  # noqa: E731
        acc = f32(b.astype(np.float64) + b)
        acc = f32(acc.astype(np.float64) + c)
        acc = f32(acc.astype(np.float64) + d)
        want = f32(acc.astype(np.float64) * 0.25)
        assert np.array_equal(got[lv], want), lv
        fine = want
    # and NOT the plain average
    assert not np.allclose(got[1], (rgb[0::2, 0::2] + rgb[0::2, 1::2] + rgb[1::2, 0::2] + rgb[1::2, 1::2]) / 4)


@pytest.mark.parametrize("wrap", [host.WRAP_REPEAT, host.WRAP_CLAMP, host.WRAP_BLACK])
def test_bilinear_lookup_without_differentials(wrap):
    """No ray differentials -> minorLength == 0 -> triangle(0, s, t) (mipmap.dart:252-255,341-355), for the three wrap modes;
    su / sv / du / dv of UVMapping2D (uv_mapping_2d.dart:25-36) move (u, v) outside [0, 1)."""
    img = RNG.random((8, 4)).astype(np.float32)
    rgb = RNG.random((4, 8, 3)).astype(np.float32)
    mp = host.UVMapping(su=2.5, sv=-1.5, du=0.25, dv=0.1)
    if wrap == host.WRAP_BLACK:
        mp = host.UVMapping(su=0.8, sv=0.7, du=0.1, dv=0.15)  # a float image cannot return TEXTURE_BLACK's Spectrum: stay inside
    o, ids = _oracle_with([host.ImageTexture(img, mp, wrap=wrap), host.ImageTexture(rgb, host.UVMapping(su=2.5, sv=-1.5, du=0.25, dv=0.1), wrap=wrap)],
                          [False, True])
    u, v = RNG.random(200), RNG.random(200)
    gf = o.texture_eval(ids[0], _dg(200, u, v))[:, 0]
    gs = o.texture_eval(ids[1], _dg(200, u, v))
    for i in range(200):
        s, t = mp.su * u[i] + mp.du, mp.sv * v[i] + mp.dv
        assert gf[i] == pytest.approx(float(_triangle(img, s, t, wrap)), rel=1e-12, abs=1e-15)
        s, t = 2.5 * u[i] + 0.25, -1.5 * v[i] + 0.1
        assert np.allclose(gs[i], _triangle(rgb, s, t, wrap), rtol=3e-7, atol=1e-7)  # float32 rounding per RGBColor operation


@pytest.mark.parametrize("trilinear", [True, False])
def test_filtered_lookup_matches_the_numpy_mipmap(trilinear):
    """MIPMap.lookup2 (mipmap.dart:224-268): the trilinear branch (:206-222) or the EWA filter (:270-339) with the anisotropy
    clamp, on a float image (Dart doubles: agreement to rounding of the sums)."""
    img = RNG.random((32, 32)).astype(np.float32)
    o, ids = _oracle_with([host.ImageTexture(img, trilinear=trilinear, max_anisotropy=4.0)], [False])
    levels = _pyramid_float(img)
    n = 120
    u, v = RNG.random(n), RNG.random(n)
    scale = 10.0 ** RNG.uniform(-3.0, -0.3, n)
    d = RNG.normal(size=(n, 4)) * scale[:, None]
    d[:10, 2:] *= 0.01  # very anisotropic footprints: the clamp at maxanisotropy
    got = o.texture_eval(ids[0], _dg(n, u, v, dudx=d[:, 0], dvdx=d[:, 1], dudy=d[:, 2], dvdy=d[:, 3]))[:, 0]
    for i in range(n):
        want = float(_lookup2(levels, u[i], v[i], d[i, 0], d[i, 1], d[i, 2], d[i, 3], host.WRAP_REPEAT, trilinear, 4.0))
        assert got[i] == pytest.approx(want, rel=1e-10), i


def _pyramid_spectrum(rgb):
    """The spectrum pyramid AS WRITTEN (see the first test): (2 b + c + d) * 0.25 in float32 steps."""
    f32 = lambda x: x.astype(np.float32)  # noqa: E731
    levels = [f32(rgb)]
    while len(levels) < _n_levels(*rgb.shape[:2]):
        fine = levels[-1]
        b, c, d = fine[0::2, 1::2], fine[1::2, 0::2], fine[1::2, 1::2]
        acc = f32(b.astype(np.float64) + b)
        acc = f32(acc.astype(np.float64) + c)
        acc = f32(acc.astype(np.float64) + d)
        levels.append(f32(acc.astype(np.float64) * 0.25))
    return levels


@pytest.mark.parametrize("trilinear", [True, False])
def test_spectrum_lookup_matches_the_numpy_mipmap_on_the_pyramid_as_written(trilinear):
    """The same filters through RGBColor arithmetic (float32 per operation) against the double-precision numpy filter over the
    spectrum pyramid as the reference builds it: agreement to float32 accuracy."""
    rgb = RNG.random((16, 16, 3)).astype(np.float32)
    o, ids = _oracle_with([host.ImageTexture(rgb, trilinear=trilinear, wrap=host.WRAP_CLAMP)], [True])
    levels = _pyramid_spectrum(rgb)
    n = 80
    u, v = RNG.random(n), RNG.random(n)
    d = RNG.normal(size=(n, 4)) * (10.0 ** RNG.uniform(-2.5, -0.5, n))[:, None]
    gs = o.texture_eval(ids[0], _dg(n, u, v, dudx=d[:, 0], dvdx=d[:, 1], dudy=d[:, 2], dvdy=d[:, 3]))
    for i in range(n):
        want = _lookup2(levels, u[i], v[i], d[i, 0], d[i, 1], d[i, 2], d[i, 3], host.WRAP_CLAMP, trilinear, 8.0)
        assert np.allclose(gs[i], want, rtol=2e-5, atol=1e-6), i


def test_mappings_scale_mix_checkerboard_bilerp():
    """The 2D mappings through UVTexture (uv_texture.dart:26-37 returns (frac s, frac t, 0)); ScaleTexture / MixTexture over
    hit-point textures; the checkerboard's closed-form box filter (checkerboard_texture.dart:50-74); BilerpTexture."""
    w2t = host.mat_mul(host.rotate(30.0, (1, 2, 3)), host.translate(0.1, -0.2, 0.3))
    sph, cyl = host.UVTexture(host.SphericalMapping(w2t)), host.UVTexture(host.CylindricalMapping(w2t))
    pla = host.UVTexture(host.PlanarMapping((0.5, 0.25, 0.0), (0.0, 0.3, 0.7), 0.2, 0.4))
    uvt = host.UVTexture(host.UVMapping(3.0, 2.0, 0.5, 0.25))
    chk = host.CheckerboardTexture(0.9, 0.2, host.UVMapping(4.0, 4.0))
    chk_none = host.CheckerboardTexture(0.9, 0.2, host.UVMapping(4.0, 4.0), aa="none")
    bil = host.BilerpTexture(0.1, 0.4, 0.7, 1.5)
    sc = host.ScaleTexture(bil, chk_none)
    mix = host.MixTexture(uvt, (0.2, 0.4, 0.6), bil)
    o, ids = _oracle_with([sph, cyl, pla, uvt, chk, chk_none, bil, sc, mix], [True, True, True, True, False, False, False, False, True])
    n = 300
    p = RNG.normal(size=(n, 3)).astype(np.float32).astype(np.float64)
    u, v = RNG.random(n), RNG.random(n)
    dd = RNG.normal(size=(n, 4)) * 0.05
    dg = _dg(n, u, v, p, dd[:, 0], dd[:, 1], dd[:, 2], dd[:, 3])
    frac = lambda x: x - np.floor(x)  # noqa: E731
    # spherical / cylindrical: Normalize(worldToTexture(p)) -> (theta / pi, phi / 2 pi), ((pi + atan2(y, x)) / 2 pi, z)
    q = (host._m(w2t).astype(np.float64) @ np.concatenate([p, np.ones((n, 1))], 1).T).T[:, :3]
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    phi = np.arctan2(q[:, 1], q[:, 0])
    phi = np.where(phi < 0, phi + 2 * np.pi, phi)
    g = o.texture_eval(ids[0], dg)
    assert np.allclose(g[:, 0], frac(np.arccos(np.clip(q[:, 2], -1, 1)) / np.pi), atol=2e-6) and np.allclose(g[:, 1], frac(phi / (2 * np.pi)), atol=2e-6)
    g = o.texture_eval(ids[1], dg)
    assert np.allclose(g[:, 0], frac((np.pi + np.arctan2(q[:, 1], q[:, 0])) / (2 * np.pi)), atol=2e-6) and np.allclose(g[:, 1], frac(q[:, 2]), atol=2e-6)
    g = o.texture_eval(ids[2], dg)
    assert np.allclose(g[:, 0], frac(0.2 + p @ np.float32([0.5, 0.25, 0.0]).astype(np.float64)), atol=1e-6)
    assert np.allclose(g[:, 1], frac(0.4 + p @ np.float32([0.0, 0.3, 0.7]).astype(np.float64)), atol=1e-6)
    g = o.texture_eval(ids[3], dg)
    assert np.allclose(g[:, 0], frac(3 * u + 0.5), atol=1e-6) and np.allclose(g[:, 1], frac(2 * v + 0.25), atol=1e-6) and (g[:, 2] == 0).all()
    # checkerboard: point sampled, and the box filter's closed form
    s, t = 4 * u, 4 * v
    point = np.where((np.floor(s) + np.floor(t)) % 2 == 0, 0.9, 0.2)
    assert np.array_equal(o.texture_eval(ids[5], dg)[:, 0], point)
    ds, dt = np.maximum(abs(4 * dd[:, 0]), abs(4 * dd[:, 2])), np.maximum(abs(4 * dd[:, 1]), abs(4 * dd[:, 3]))
    bump = lambda x: np.floor(x / 2) + 2 * np.maximum(x / 2 - np.floor(x / 2) - 0.5, 0)  # noqa: E731
    sint, tint = (bump(s + ds) - bump(s - ds)) / (2 * ds), (bump(t + dt) - bump(t - dt)) / (2 * dt)
    area2 = np.where((ds > 1) | (dt > 1), 0.5, sint + tint - 2 * sint * tint)
    inside = (np.floor(s - ds) == np.floor(s + ds)) & (np.floor(t - dt) == np.floor(t + dt))
    want = np.where(inside, point, 0.9 * (1 - area2) + 0.2 * area2)
    assert np.allclose(o.texture_eval(ids[4], dg)[:, 0], want, rtol=1e-12)
    # the integral the closed form stands for: the fraction of the filter box covered by tex2's checks (brute-force quadrature)
    i = int(np.argmax(~inside & (ds < 1) & (dt < 1)))
    xs, ys = np.meshgrid(np.linspace(s[i] - ds[i], s[i] + ds[i], 801), np.linspace(t[i] - dt[i], t[i] + dt[i], 801))
    frac2 = ((np.floor(xs) + np.floor(ys)) % 2 == 1).mean()
    assert area2[i] == pytest.approx(frac2, abs=5e-3)
    bl = 0.1 * ((1 - u) * (1 - v)) + 0.4 * (1 - u) * v + 0.7 * u * (1 - v) + 1.5 * u * v
    assert np.allclose(o.texture_eval(ids[6], dg)[:, 0], bl, rtol=1e-13)
    assert np.allclose(o.texture_eval(ids[7], dg)[:, 0], point * bl, rtol=1e-13)
    g = o.texture_eval(ids[8], dg)
    uvv = np.stack([frac(3 * u + 0.5), frac(2 * v + 0.25), np.zeros(n)], 1)
    assert np.allclose(g, uvv * (1 - bl)[:, None] + np.float32([0.2, 0.4, 0.6]) * bl[:, None], rtol=1e-6, atol=1e-7)


# ---- through the renderer ----------------------------------------------------------------------------------------------
QUAD_P = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
QUAD_I = np.array([[0, 1, 2], [2, 3, 0]], np.uint32)
QUAD_UV = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)


def _render(sb, cam, film, sampler, integ, nthreads=4):
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, film, sampler, integ)
    o.render(0, 1, nthreads)
    return o.film_read()["rgb"]


PROGRAM_CASES = {
    "matte": (host.matte_lobes, dict(kd=(0.6, 0.5, 0.4), sigma=20.0)),
    "mirror": (host.mirror_lobes, dict(kr=(0.9, 0.8, 0.7))),
    "glass": (host.glass_lobes, dict(kr=0.9, kt=(0.7, 0.8, 0.9), index=1.4)),
    "plastic": (host.plastic_lobes, dict(kd=(0.3, 0.4, 0.5), ks=0.3, roughness=0.05)),
    "metal": (host.metal_lobes, dict(eta=(0.2, 0.9, 1.1), k=(3.9, 2.4, 2.2), roughness=0.02)),
    "shinymetal": (host.shinymetal_lobes, dict(ks=(0.8, 0.7, 0.6), kr=0.5, roughness=0.08)),
    "substrate": (host.substrate_lobes, dict(kd=(0.5, 0.3, 0.2), ks=0.4, uroughness=0.05, vroughness=0.2)),
    "translucent": (host.translucent_lobes, dict(kd=0.4, ks=0.3, reflect=(0.6, 0.5, 0.4), transmit=0.3, roughness=0.1)),
    "uber": (host.uber_lobes, dict(kd=(0.4, 0.3, 0.2), ks=0.2, kr=0.1, kt=0.15, roughness=0.07, index=1.3, opacity=0.8)),
    "subsurface": (host.subsurface_lobes, dict(kr=(0.9, 0.8, 0.7), index=1.4)),
    "kdsubsurface": (host.subsurface_lobes, dict(kr=0.8, index=1.25)),
}


@pytest.mark.parametrize("plugin", sorted(PROGRAM_CASES))
def test_a_program_over_constant_textures_builds_the_flattened_bsdf(plugin):
    """The per-hit getBSDF of every material plugin against the host-side flattening of the same plugin with the same constants
    (host.*_lobes, pinned in tests/test_oracle_materials.py): the two films must be bit-identical."""
    fn, params = PROGRAM_CASES[plugin]
    films = []
    for program in (False, True):
        sb = host.SceneBuilder()
        m = sb.material_program(plugin, **params) if program else sb.material_lobes(fn(**params))
        back = sb.material(0.7)
        sb.mesh(QUAD_P, QUAD_I, material=m, uv=QUAD_UV)
        sb.mesh(QUAD_P * 3 + np.float32([0, 0, 1.5]), QUAD_I, material=back)
        sb.point_light((0.7, 0.9, -2.5), (12, 11, 10))
        sb.point_light((-0.5, 0.3, 0.8), (3, 3, 3))
        cam = host.PerspectiveCamera(host.look_at((0.3, 0.2, -3.5), (0, 0, 0), (0, 1, 0)), fov=38.0)
        films.append(_render(sb, cam, host.Film(24, 18), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                             host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4)))
    assert films[0].max() > 0 and np.array_equal(films[0], films[1])


def test_mix_program_matches_the_flattened_mix():
    films = []
    for program in (False, True):
        sb = host.SceneBuilder()
        if program:
            a = sb.material_program("plastic", kd=(0.3, 0.4, 0.5), ks=0.3, roughness=0.05)
            b = sb.material_lobes(host.matte_lobes((0.7, 0.2, 0.2), 0.0))
            m = sb.material_program("mix", m1=a, m2=b, amount=(0.3, 0.5, 0.7))
        else:
            m = sb.material_lobes(host.mix_lobes(host.plastic_lobes((0.3, 0.4, 0.5), 0.3, 0.05), host.matte_lobes((0.7, 0.2, 0.2), 0.0), (0.3, 0.5, 0.7)))
        sb.mesh(QUAD_P, QUAD_I, material=m, uv=QUAD_UV)
        sb.point_light((0.7, 0.9, -2.5), (12, 11, 10))
        cam = host.PerspectiveCamera(host.look_at((0.3, 0.2, -3.5), (0, 0, 0), (0, 1, 0)), fov=38.0)
        films.append(_render(sb, cam, host.Film(24, 18), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_DIRECT)))
    assert films[0].max() > 0 and np.array_equal(films[0], films[1])


def test_bump_map_tilts_the_shading_normal_by_the_displacement_gradient():
    """Material.Bump (material.dart:35-88) on a flat quad (dndu = dndv = 0) with the displacement d(u, v) = a u + b v (a BilerpTexture):
    dpdu' = dpdu + nn a, dpdv' = dpdv + nn b exactly, for any step du / dv.  Under a distant light a matte surface then returns
    Kd / pi * L * |n' . wi| with n' = normalize(cross(dpdu', dpdv')) flipped towards the geometric normal."""
    a, b = 0.35, -0.2
    disp = host.BilerpTexture(0.0, b, a, a + b)  # v00 + (v10 - v00) s + (v01 - v00) t: no s t term
    sb = host.SceneBuilder()
    m = sb.material_program("matte", kd=(0.8, 0.6, 0.4), bumpmap=disp)
    sb.mesh(QUAD_P, QUAD_I, material=m, uv=QUAD_UV)
    wi = np.array([0.3, 0.2, -1.0]) / np.linalg.norm([0.3, 0.2, -1.0])
    sb.distant_light(tuple(wi), (0, 0, 0), (2.0, 2.0, 2.0))  # from -> to: light travels along -wi
    cam = host.PerspectiveCamera(host.look_at((0, 0, -4), (0, 0, 0), (0, 1, 0)), fov=20.0)
    img = _render(sb, cam, host.Film(16, 16), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False),
                  host.Integrator(kind=host.INTEGRATOR_DIRECT))
    # uv (0,0)-(1,1) over x, y in [-1, 1]: dpdu = (2, 0, 0), dpdv = (0, 2, 0), nn = +z; the camera sits on the -z side
    nn = np.array([0.0, 0.0, 1.0])
    n2 = np.cross(np.array([2.0, 0, 0]) + nn * a, np.array([0, 2.0, 0]) + nn * b)
    n2 /= np.linalg.norm(n2)
    want = np.float32([0.8, 0.6, 0.4]).astype(np.float64) / np.pi * 2.0 * abs(n2 @ wi)
    centre = img[4:12, 4:12]
    assert np.allclose(centre, want, rtol=2e-6), (centre[0, 0], want)
    flat = np.float32([0.8, 0.6, 0.4]) / np.pi * 2.0 * abs(nn @ wi)
    assert not np.allclose(centre, flat, rtol=1e-3)  # the bump map did something


def test_camera_ray_differentials_size_the_texture_filter():
    """perspective_camera.dart:50-56,122-128 (rx / ry through the neighbouring raster positions), sampler_renderer.dart:166
    (scaled by 1 / sqrt(samplesPerPixel)), differential_geometry.dart:122-205 ((u, v) offsets of the auxiliary plane hits),
    uv_mapping_2d.dart and the checkerboard's box filter: a matte quad whose Kd is a closed-form-filtered checkerboard(1, 0),
    seen head on under a distant light, returns L / pi * cos * (1 - area2) with area2 from the pixel's footprint — computed here
    from pinhole geometry alone."""
    xres = yres = 24
    fov, dist, freq, spp = 30.0, 5.0, 9.0, 4
    sb = host.SceneBuilder()
    m = sb.material_program("matte", kd=host.CheckerboardTexture(1.0, 0.0, host.UVMapping(freq, freq)))
    sb.mesh(QUAD_P, QUAD_I, material=m, uv=QUAD_UV)
    sb.distant_light((0, 0, -1), (0, 0, 0), (np.pi, np.pi, np.pi))
    cam = host.PerspectiveCamera(host.look_at((0, 0, -dist), (0, 0, 0), (0, 1, 0)), fov=fov)
    img = _render(sb, cam, host.Film(xres, yres), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2, jitter=False),
                  host.Integrator(kind=host.INTEGRATOR_DIRECT))[:, :, 0]
    # pinhole: raster (x, y) -> screen [-1, 1]^2 (square film), tan(fov / 2) at the screen edge; look_at(0, 0, -d) with up +y
    # maps camera +x to world -x?  Only |du/dx| matters: the footprint is symmetric.  Plane z = 0 at distance `dist`.
    tanh = math.tan(math.radians(fov) / 2)
    pix = 2 * tanh / xres * dist                      # world-space distance between neighbouring pixels' hits on the plane
    d_uv = pix / 2.0 * freq / math.sqrt(spp)          # u = (x + 1) / 2, times the mapping's scale, times scaleDifferentials
    bump = lambda x: np.floor(x / 2) + 2 * np.maximum(x / 2 - np.floor(x / 2) - 0.5, 0)  # noqa: E731
    want = np.zeros((yres, xres))
    for py in range(yres):
        for px in range(xres):
            acc = 0.0
            for sx, sy in ((0.25, 0.25), (0.75, 0.25), (0.25, 0.75), (0.75, 0.75)):
                X, Y = (px + sx) / xres * 2 - 1, 1 - (py + sy) / yres * 2   # screen coordinates of the sample
                wx, wy = -X * tanh * dist, Y * tanh * dist                   # world hit (camera x is world -x for this look_at)
                s, t = (wx + 1) / 2 * freq, (wy + 1) / 2 * freq
                if np.floor(s - d_uv) == np.floor(s + d_uv) and np.floor(t - d_uv) == np.floor(t + d_uv):
                    val = 1.0 if (np.floor(s) + np.floor(t)) % 2 == 0 else 0.0
                else:
                    sint = (bump(s + d_uv) - bump(s - d_uv)) / (2 * d_uv)
                    tint = (bump(t + d_uv) - bump(t - d_uv)) / (2 * d_uv)
                    val = 1.0 - (sint + tint - 2 * sint * tint)
                acc += val
            want[py, px] = acc / 4
    inner = (slice(6, 18), slice(6, 18))  # the quad fills the middle of the frame
    # the footprint at the frame centre is exact; off centre the perspective foreshortening of the differentials is second order
    assert np.abs(img[inner] - want[inner]).max() < 0.02, np.abs(img[inner] - want[inner]).max()
    assert img[inner].std() > 0.05  # the checks are resolved, not averaged away


# ---- the noise textures (lib/core/texture.dart:40-140 and lib/textures/{fbm,wrinkled,windy,marble,dots}_texture.dart) ------------
# Ken Perlin's published permutation table (his 2002 reference implementation), typed here independently of the C++ sides
PERLIN = [151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6,
          148, 247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171,
          168, 68, 175, 74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46,
          245, 40, 244, 102, 143, 54, 65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159,
          86, 164, 100, 109, 198, 173, 186, 3, 64, 52, 217, 226, 250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227,
          47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213, 119, 248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129,
          22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104, 218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12, 191,
          179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157, 184, 84, 204, 176, 115, 121, 50, 45, 127,
          4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180] * 2


def _noise(x, y=0.5, z=0.5):
    ix, iy, iz = math.floor(x), math.floor(y), math.floor(z)
    dx, dy, dz = x - ix, y - iy, z - iz
    ix, iy, iz = ix & 255, iy & 255, iz & 255

    def grad(X, Y, Z, a, b, c):
        h = PERLIN[PERLIN[PERLIN[X] + Y] + Z] & 15
        u = a if (h < 8 or h in (12, 13)) else b
        v = b if (h < 4 or h in (12, 13)) else c
        return (-u if h & 1 else u) + (-v if h & 2 else v)

    def wgt(t):
        return 6 * t ** 5 - 15 * t ** 4 + 10 * t ** 3
    lerp = lambda t, a, b: (1 - t) * a + t * b  # noqa: E731
    wx, wy, wz = wgt(dx), wgt(dy), wgt(dz)
    x00 = lerp(wx, grad(ix, iy, iz, dx, dy, dz), grad(ix + 1, iy, iz, dx - 1, dy, dz))
    x10 = lerp(wx, grad(ix, iy + 1, iz, dx, dy - 1, dz), grad(ix + 1, iy + 1, iz, dx - 1, dy - 1, dz))
    x01 = lerp(wx, grad(ix, iy, iz + 1, dx, dy, dz - 1), grad(ix + 1, iy, iz + 1, dx - 1, dy, dz - 1))
    x11 = lerp(wx, grad(ix, iy + 1, iz + 1, dx, dy - 1, dz - 1), grad(ix + 1, iy + 1, iz + 1, dx - 1, dy - 1, dz - 1))
    return lerp(wz, lerp(wy, x00, x10), lerp(wy, x01, x11))


def _f32v(v):
    return np.asarray(v, np.float64).astype(np.float32).astype(np.float64)


def _fbm(P, dpdx, dpdy, omega, max_oct, turbulence=False):
    s2 = max(float(dpdx @ dpdx), float(dpdy @ dpdy))
    l2 = math.log(s2) / math.log(2.0) if s2 > 0 else -math.inf
    foct = min(float(max_oct), max(0.0, -1.0 - 0.5 * l2))
    octv = math.floor(foct)
    total, lam, o = 0.0, 1.0, 1.0
    f = abs if turbulence else (lambda q: q)
    for _ in range(octv):
        total += o * f(_noise(*_f32v(P * lam)))  # P * lambda is a new float32 Point
        lam *= 1.99
        o *= omega
    v = min(max((foct - octv - 0.3) / (0.7 - 0.3), 0.0), 1.0)
    total += o * (v * v * (-2.0 * v + 3.0)) * f(_noise(*_f32v(P * lam)))
    if turbulence:
        total += (max_oct - foct) * 0.2
    return total


def test_perlin_noise_has_the_properties_of_the_published_function():
    """Independent of any table copy: gradient noise vanishes on the integer lattice, is bounded, and has period 256."""
    for x, y, z in RNG.integers(-300, 300, (50, 3)):
        assert _noise(float(x), float(y), float(z)) == 0.0
    pts = RNG.uniform(-40, 40, (300, 3))
    vals = np.array([_noise(*q) for q in pts])
    assert np.abs(vals).max() <= 1.5 and vals.std() > 0.1
    for q in pts[:40]:
        assert _noise(q[0] + 256.0, q[1], q[2] - 256.0) == pytest.approx(_noise(*q), abs=1e-9)


def test_noise_textures_match_the_numpy_restatement():
    """fbm / wrinkled / windy / marble over IdentityMapping3D, dots over a 2D mapping, the 3D checkerboard."""
    xf = host.mat_mul(host.scale(1.7, 0.9, 1.3), host.rotate(20.0, (0, 1, 1)))
    fbm, wr, wi = host.FBmTexture(6, 0.6, xf), host.WrinkledTexture(5, 0.45, xf), host.WindyTexture(xf)
    mar = host.MarbleTexture(7, 0.55, 2.5, 0.3, xf)
    dots = host.DotsTexture(0.9, 0.1, host.UVMapping(7.0, 5.0, 0.3, 0.1))
    ch3 = host.Checkerboard3DTexture(0.8, 0.25, xf)
    o, ids = _oracle_with([fbm, fbm, wr, wi, mar, dots, ch3], [False, True, False, False, True, False, False])
    n = 60
    p = _f32v(RNG.uniform(-3, 3, (n, 3)))
    dpx, dpy = _f32v(RNG.normal(size=(n, 3)) * 10.0 ** RNG.uniform(-4, -0.5, (n, 1))), _f32v(RNG.normal(size=(n, 3)) * 0.01)
    dpx[:5] = 0.0
    dpy[:5] = 0.0  # no differentials: every octave
    u, v = RNG.random(n), RNG.random(n)
    dg = _dg(n, u, v, p, dpdx=dpx, dpdy=dpy)
    M = host._m(xf).astype(np.float64)
    xp = lambda q: _f32v(M[:3, :3] @ q + M[:3, 3])  # noqa: E731
    xv = lambda q: _f32v(M[:3, :3] @ q)  # noqa: E731
    g_fbm, g_fbms, g_wr, g_wi = (o.texture_eval(i, dg) for i in ids[:4])
    g_mar, g_dots, g_ch3 = (o.texture_eval(i, dg) for i in ids[4:])
    spline = np.array([[0.58, 0.58, 0.6], [0.58, 0.58, 0.6], [0.58, 0.58, 0.6], [0.5, 0.5, 0.5], [0.6, 0.59, 0.58], [0.58, 0.58, 0.6],
                       [0.58, 0.58, 0.6], [0.2, 0.2, 0.33], [0.58, 0.58, 0.6]])
    for i in range(n):
        P, dx, dy = xp(p[i]), xv(dpx[i]), xv(dpy[i])
        want = _fbm(P, dx, dy, 0.6, 6)
        assert g_fbm[i, 0] == pytest.approx(want, rel=1e-9, abs=1e-12)
        assert np.allclose(g_fbms[i], np.float32(want), rtol=1e-7)  # new Spectrum(n)
        assert g_wr[i, 0] == pytest.approx(_fbm(P, dx, dy, 0.45, 5, True), rel=1e-9, abs=1e-12)
        wind = abs(_fbm(_f32v(P * 0.1), _f32v(dx * 0.1), _f32v(dy * 0.1), 0.5, 3)) * _fbm(P, dx, dy, 0.5, 6)
        assert g_wi[i, 0] == pytest.approx(wind, rel=1e-9, abs=1e-12)
        Ps = _f32v(P * 2.5)
        t = 0.5 + 0.5 * math.sin(Ps[1] + 0.3 * _fbm(Ps, _f32v(dx * 2.5), _f32v(dy * 2.5), 0.55, 7))
        first = math.floor(t * 6)
        t = t * 6 - first
        c0, c1, c2, c3 = spline[first:first + 4]
        s0, s1, s2 = c0 * (1 - t) + c1 * t, c1 * (1 - t) + c2 * t, c2 * (1 - t) + c3 * t
        s0, s1 = s0 * (1 - t) + s1 * t, s1 * (1 - t) + s2 * t
        assert np.allclose(g_mar[i], (s0 * (1 - t) + s1 * t) * 1.5, rtol=2e-6)
        s, tt = 7.0 * u[i] + 0.3, 5.0 * v[i] + 0.1
        sc, tc = math.floor(s + 0.5), math.floor(tt + 0.5)
        inside = False
        if _noise(sc + 0.5, tc + 0.5) > 0:
            cs, ct = sc + 0.15 * _noise(sc + 1.5, tc + 2.8), tc + 0.15 * _noise(sc + 4.5, tc + 9.8)
            inside = (s - cs) ** 2 + (tt - ct) ** 2 < 0.35 * 0.35
        assert g_dots[i, 0] == (0.9 if inside else 0.1)
        assert g_ch3[i, 0] == (0.8 if (math.floor(P[0]) + math.floor(P[1]) + math.floor(P[2])) % 2 == 0 else 0.25)
    assert 0 < (g_dots[:, 0] == 0.9).sum() < n  # both branches were taken
