"""GPU vs oracle on materials whose parameters are textures that read the hit point, and on bump maps:
camera ray differentials, DifferentialGeometry.computeDifferentials, the texture mappings, the MIPMap trilinear / EWA filters,
Material.Bump and the per-hit getBSDF of the material plugins — through the C ABI (drt_set_textures, drt_set_material_programs),
per pixel at the north_star tolerance (1e-3 relative; the replayed streams agree to ~1e-6) and with equal ray counts.  The
oracle side of every piece is pinned by tests/test_oracle_textures.py."""
import numpy as np
import pytest

from dartray_b200 import capi, host
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(77)
IMG_RGB = RNG.random((32, 16, 3)).astype(np.float32)
IMG_F = RNG.random((16, 16)).astype(np.float32)
QUAD_I = np.array([[0, 1, 2], [2, 3, 0]], np.uint32)
QUAD_UV = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)


def _quad(z, size=1.0, y=None):
    if y is not None:  # a floor at height y
        return np.array([[-size, y, -size], [size, y, -size], [size, y, size], [-size, y, size]], np.float32)
    return np.array([[-size, -size, z], [size, -size, z], [size, size, z], [-size, size, z]], np.float32)


def _rel_err(a, b, floor=1e-3):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def _both(sb, cam, film, sampler, integ):
    arrays = sb.arrays()
    g, o = capi.Context(0), Oracle()
    for c in (g, o):
        host.upload_scene(c, arrays)
        host.configure_render(c, cam, film, sampler, integ)
    g.render(0, 1)
    o.render(0, 1, 8)
    return g, o, g.film_read()["rgb"], o.film_read()["rgb"]


def _check(g, o, fg, fo, what):
    err = _rel_err(fg, fo)
    print(what, "max rel err", err.max(), "mean", float(fo.mean()))
    assert np.isfinite(fg).all() and fo.max() > 0
    assert err.max() <= 1e-3
    sg, so = g.render_stats(), o.render_stats()
    assert sg["camera_samples"] == so["camera_samples"]
    assert abs(sg["closest_rays"] - so["closest_rays"]) <= 1e-4 * so["closest_rays"]
    assert abs(sg["shadow_rays"] - so["shadow_rays"]) <= 1e-4 * max(so["shadow_rays"], 1)


def _scene(material_of, lights="point"):
    """A textured back wall (mesh with uv), a floor with per-vertex normals, a sphere, a disk and the four other quadrics, each
    wearing material_of(shape name)."""
    sb = host.SceneBuilder()
    sb.mesh(_quad(1.0, 2.0), QUAD_I, material=material_of(sb, "wall"), uv=QUAD_UV * 2.0)
    n = np.array([[0.1, 1, 0], [-0.1, 1, 0.1], [0, 1, -0.1], [0.05, 1, 0.05]], np.float32)
    sb.mesh(_quad(0, 2.0, y=-1.0), QUAD_I, material=material_of(sb, "floor"), uv=QUAD_UV, N=n)
    sb.sphere(host.translate(-0.8, -0.4, 0.0), radius=0.5, material=material_of(sb, "sphere"))
    sb.disk(host.mat_mul(host.translate(0.9, 0.6, 0.2), host.rotate(60.0, (1, 0, 0))), radius=0.5, innerradius=0.1, material=material_of(sb, "disk"))
    sb.cylinder(host.mat_mul(host.translate(0.9, -0.6, 0.0), host.rotate(90.0, (1, 0, 0))), radius=0.25, zmin=-0.4, zmax=0.4,
                material=material_of(sb, "cylinder"))
    sb.cone(host.mat_mul(host.translate(0.0, -1.0, -0.3), host.rotate(-90.0, (1, 0, 0))), radius=0.3, height=0.7, material=material_of(sb, "cone"))
    sb.paraboloid(host.mat_mul(host.translate(-0.2, 0.8, 0.2), host.rotate(200.0, (1, 0.3, 0))), radius=0.35, zmin=0.05, zmax=0.5,
                  material=material_of(sb, "paraboloid"))
    sb.hyperboloid(host.mat_mul(host.translate(-1.4, 0.7, 0.3), host.rotate(70.0, (1, 0, 0))), p1=(0.3, 0.0, -0.3), p2=(0.2, 0.25, 0.3),
                   material=material_of(sb, "hyperboloid"))
    if lights == "point":
        sb.point_light((0.5, 1.5, -3.0), (14, 13, 12))
        sb.point_light((-1.5, 0.5, -1.0), (3, 3, 4))
    else:
        sb.mesh(_quad(0, 0.4, y=1.9), QUAD_I, material=sb.material(0.0), area_light=(9, 9, 8), nsamples=2)
    return sb


CAM = host.PerspectiveCamera(host.look_at((0.2, 0.3, -4.0), (0, 0, 0), (0, 1, 0)), fov=45.0)


def _uber_everywhere(sb, name):
    cache = getattr(sb, "_mats", None)
    if cache is None:
        cache = sb._mats = {}
        kd = host.ScaleTexture(host.ImageTexture(IMG_RGB, host.UVMapping(3.0, 2.0, 0.1, 0.2)), (0.9, 0.8, 0.7))
        bump = host.ScaleTexture(host.ImageTexture(IMG_F, host.UVMapping(2.0, 2.0)), -0.08)
        cache["m"] = sb.material_program("uber", kd=kd, ks=0.05, roughness=0.05, bumpmap=bump)
    return cache["m"]


@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4), host.Integrator(kind=host.INTEGRATOR_DIRECT),
                                   host.Integrator(kind=host.INTEGRATOR_DIRECT, strategy=1), host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=1)],
                         ids=["path", "direct_all", "direct_one", "whitted"])
def test_image_textured_bumped_uber_on_every_shape(integ):
    """bump-sphere.pbrt's material (uber, Kd = scale(imagemap, colour), bumpmap = scale(imagemap, -0.1)) on a mesh with uv, a mesh
    with per-vertex normals and all six quadrics: every shape's (u, v), dndu / dndv and the EWA filter under real footprints."""
    sb = _scene(_uber_everywhere)
    g, o, fg, fo = _both(sb, CAM, host.Film(96, 72), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
    _check(g, o, fg, fo, f"uber everywhere, integrator {integ.kind}")


def test_area_light_and_stratified_sampler():
    sb = _scene(_uber_everywhere, lights="area")
    g, o, fg, fo = _both(sb, CAM, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2),
                         host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3))
    _check(g, o, fg, fo, "area light")


PLUGINS = {
    "matte": dict(kd=host.ImageTexture(IMG_RGB, trilinear=True), sigma=host.ScaleTexture(host.ImageTexture(IMG_F), 40.0)),
    "mirror": dict(kr=host.CheckerboardTexture((0.9, 0.9, 0.9), (0.2, 0.3, 0.4), host.UVMapping(6.0, 6.0))),
    "glass": dict(kr=host.ImageTexture(IMG_RGB), kt=host.UVTexture(host.UVMapping(2.0, 2.0)), index=host.BilerpTexture(1.3, 1.4, 1.5, 1.6)),
    "plastic": dict(kd=host.ImageTexture(IMG_RGB, wrap=host.WRAP_CLAMP), ks=host.ImageTexture(IMG_RGB, wrap=host.WRAP_BLACK, mapping=host.UVMapping(0.9, 0.9, 0.05, 0.05)),
                    roughness=host.MixTexture(0.02, 0.3, host.ImageTexture(IMG_F))),
    "metal": dict(eta=host.MixTexture((0.2, 0.9, 1.1), (1.5, 1.0, 0.6), host.ImageTexture(IMG_F)), k=(3.9, 2.4, 2.2), roughness=host.BilerpTexture(0.01, 0.2, 0.05, 0.3)),
    "shinymetal": dict(ks=host.ImageTexture(IMG_RGB), kr=host.CheckerboardTexture(0.8, 0.0, host.UVMapping(3.0, 3.0), aa="none"), roughness=0.1),
    "substrate": dict(kd=host.ImageTexture(IMG_RGB), ks=0.3, uroughness=host.BilerpTexture(0.02, 0.02, 0.4, 0.4), vroughness=0.1),
    "translucent": dict(kd=host.ImageTexture(IMG_RGB), ks=0.2, reflect=host.UVTexture(), transmit=0.4, roughness=0.1),
    "uber": dict(kd=host.ImageTexture(IMG_RGB), ks=0.2, kr=host.CheckerboardTexture(0.3, 0.0, host.UVMapping(5.0, 5.0)), kt=0.1, roughness=0.07,
                 index=1.3, opacity=host.MixTexture((1.0, 1.0, 1.0), (0.5, 0.6, 0.7), host.ImageTexture(IMG_F))),
    "subsurface": dict(kr=host.ImageTexture(IMG_RGB), index=host.BilerpTexture(1.2, 1.3, 1.4, 1.5)),
}


@pytest.mark.parametrize("plugin", sorted(PLUGINS))
def test_every_material_plugin_with_textured_parameters(plugin):
    """The per-hit getBSDF of each material (lib/materials/*.dart) with textures on its parameters, path integrator (the specular
    lobes that come and go per hit drive specularBounce and the emitted-light rule)."""
    def mats(sb, name):
        if not hasattr(sb, "_m"):
            sb._m = sb.material_program(plugin, **PLUGINS[plugin])
            sb._plain = sb.material((0.6, 0.6, 0.6))
        return sb._m if name in ("wall", "sphere", "cylinder", "floor") else sb._plain
    sb = _scene(mats, lights="area")
    g, o, fg, fo = _both(sb, CAM, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5))
    _check(g, o, fg, fo, plugin)


def test_mix_material_of_a_textured_and_a_flattened_material():
    def mats(sb, name):
        if not hasattr(sb, "_m"):
            a = sb.material_program("plastic", kd=host.ImageTexture(IMG_RGB), ks=0.3, roughness=0.05, bumpmap=host.ScaleTexture(host.ImageTexture(IMG_F), 0.05))
            b = sb.material_lobes(host.matte_lobes((0.7, 0.2, 0.2), 20.0))
            sb._m = sb.material_program("mix", m1=a, m2=b, amount=host.CheckerboardTexture((0.9, 0.9, 0.9), (0.1, 0.2, 0.3), host.UVMapping(4.0, 4.0)))
        return sb._m
    sb = _scene(mats)
    g, o, fg, fo = _both(sb, CAM, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    _check(g, o, fg, fo, "mix")


@pytest.mark.parametrize("mapping", ["spherical", "cylindrical", "planar"])
def test_mappings_that_read_the_hit_point_and_its_screen_space_derivatives(mapping):
    w2t = host.mat_mul(host.rotate(25.0, (1, 1, 0)), host.translate(0.1, 0.2, -0.3))
    mp = {"spherical": host.SphericalMapping(w2t), "cylindrical": host.CylindricalMapping(w2t), "planar": host.PlanarMapping((0.7, 0.1, 0.0), (0.0, 0.6, 0.3), 0.1, 0.2)}[mapping]

    def mats(sb, name):
        if not hasattr(sb, "_m"):
            sb._m = sb.material_program("matte", kd=host.ImageTexture(IMG_RGB, mp), bumpmap=host.ScaleTexture(host.ImageTexture(IMG_F, mp, trilinear=True), 0.05))
        return sb._m
    sb = _scene(mats)
    g, o, fg, fo = _both(sb, CAM, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=2), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    _check(g, o, fg, fo, mapping)


@pytest.mark.parametrize("camera", ["thinlens", "orthographic", "environment"])
def test_ray_differentials_of_the_other_cameras(camera):
    """perspective with a lens (the offsets ignore the lens, perspective_camera.dart:122-126), orthographic (its offset origins stay in
    camera space, orthographic_camera.dart:111-115 as written), environment (the generic one-pixel shifts, camera.dart:37-62)."""
    c2w = host.look_at((0.2, 0.3, -4.0), (0, 0, 0), (0, 1, 0))
    cam = {"thinlens": host.PerspectiveCamera(c2w, fov=45.0, lens_radius=0.05, focal_distance=4.0),
           "orthographic": host.OrthographicCamera(c2w, screen_window=(-2.2, 2.2, -1.7, 1.7)),
           "environment": host.EnvironmentCamera(host.look_at((0.0, 0.0, -0.5), (0, 0, 1), (0, 1, 0)))}[camera]
    sb = _scene(_uber_everywhere)
    g, o, fg, fo = _both(sb, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    _check(g, o, fg, fo, camera)


def test_constant_programs_reproduce_the_flattened_materials_on_the_gpu():
    """A program over constant textures must give the film of the host-flattened lobe list, bit for bit (same kernels after the
    texture pass, same lobes)."""
    films = []
    for program in (False, True):
        def mats(sb, name):
            if not hasattr(sb, "_m"):
                kw = dict(kd=(0.4, 0.3, 0.2), ks=0.2, kr=0.1, kt=0.15, roughness=0.07, index=1.3, opacity=0.8)
                sb._m = sb.material_program("uber", **kw) if program else sb.material_lobes(host.uber_lobes(**kw))
            return sb._m
        sb = _scene(mats)
        g = capi.Context(0)
        host.upload_scene(g, sb.arrays())
        host.configure_render(g, CAM, host.Film(48, 36), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4))
        g.render(0, 1)
        films.append(g.film_read()["rgb"])
    assert films[0].max() > 0 and np.array_equal(films[0], films[1])


def test_programs_with_the_specular_recursion_are_rejected_not_approximated():
    def mats(sb, name):
        if not hasattr(sb, "_m"):
            sb._m = sb.material_program("glass", kr=host.ImageTexture(IMG_RGB), kt=0.9, index=1.5)
        return sb._m
    sb = _scene(mats)
    g = capi.Context(0)
    host.upload_scene(g, sb.arrays())
    host.configure_render(g, CAM, host.Film(16, 12), host.Sampler(kind=host.SAMPLER_LD, spp=1), host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=5))
    with pytest.raises(RuntimeError, match="specular recursion"):
        g.render(0, 1)


def test_invalid_texture_tables_are_rejected():
    g = capi.Context(0)
    nodes = np.zeros(1, host.TEX_DTYPE)
    nodes["kind"], nodes["spectrum"], nodes["image_width"], nodes["image_height"], nodes["image_channels"] = 3, 1, 12, 8, 3
    nodes["tex1"] = nodes["tex2"] = nodes["amount"] = -1
    with pytest.raises(RuntimeError, match="power-of-two"):
        g.set_textures(nodes, np.zeros(12 * 8 * 3, np.float32))
    nodes["image_width"] = 16
    with pytest.raises(RuntimeError, match="beyond the texel array"):
        g.set_textures(nodes, np.zeros(10, np.float32))
    nodes["kind"], nodes["tex1"] = 1, 0
    with pytest.raises(RuntimeError, match="earlier nodes"):
        g.set_textures(nodes, np.zeros(0, np.float32))


NOISE_XF = host.mat_mul(host.scale(2.5, 2.0, 3.0), host.rotate(20.0, (0, 1, 1)))
NOISE_CASES = {
    "marble_fbm_bump": ("matte", dict(kd=host.MarbleTexture(6, 0.5, 3.0, 0.4, NOISE_XF), bumpmap=host.ScaleTexture(host.FBmTexture(5, 0.6, NOISE_XF), 0.05))),
    "wrinkled_windy": ("plastic", dict(kd=host.ScaleTexture(host.WrinkledTexture(4, 0.5, NOISE_XF), (0.8, 0.6, 0.4)), ks=host.MixTexture((0.05, 0.05, 0.05), (0.4, 0.4, 0.4), host.WindyTexture(NOISE_XF)), roughness=0.1)),
    "dots_checker3d": ("uber", dict(kd=host.DotsTexture((0.9, 0.2, 0.1), (0.2, 0.3, 0.8), host.UVMapping(9.0, 9.0)), ks=0.1,
                                    roughness=host.Checkerboard3DTexture(0.05, 0.3, NOISE_XF))),
}


@pytest.mark.parametrize("case", sorted(NOISE_CASES))
def test_noise_textures(case):
    """fbm / wrinkled / windy / marble (Perlin noise over IdentityMapping3D, octaves cut by the screen-space footprint), dots and the
    3D checkerboard, on every shape, camera vertices with differentials and later bounces without."""
    plugin, params = NOISE_CASES[case]

    def mats(sb, name):
        if not hasattr(sb, "_m"):
            sb._m = sb.material_program(plugin, **params)
        return sb._m
    sb = _scene(mats)
    g, o, fg, fo = _both(sb, CAM, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3))
    _check(g, o, fg, fo, case)
