"""GPU wavefront renderer (drt_render through the C ABI) against the CPU oracle in keyed-stream mode.

north_star tolerances: replayed sample streams; deterministic integrators (ambient occlusion, direct
lighting) per pixel within 1e-3 relative; path tracing per-pixel mean within 3 sigma.  Because the GPU
shading code keeps the reference's arithmetic (float32 storage, float64 expressions), the renders
agree far more tightly than that; the tests assert the stated tolerance on every pixel and report
the observed maximum."""
import math

import numpy as np
import pytest

from dartray_b200 import capi, host, scenes
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _pair(arrays):
    g, o = capi.Context(0), Oracle()
    for c in (g, o):
        host.upload_scene(c, arrays)
    return g, o


def _cornell():
    sb, cam = scenes.cornell_synth()
    return sb.arrays(), cam


def _render_both(arrays, cam, film, sampler, integ, task=(0, 1)):
    g, o = _pair(arrays)
    for c in (g, o):
        host.configure_render(c, cam, film, sampler, integ)
    g.render(*task)
    o.render(task[0], task[1], 8)
    return g, o, g.film_read(), o.film_read()


def _rel_err(a, b, floor=1e-4):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


# ---- sampler sequences: bit-exact replay -----------------------------------------------------------------
@pytest.mark.parametrize("sampler,integ", [
    (host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_PATH)),
    (host.Sampler(kind=host.SAMPLER_LD, spp=64, seed=7), host.Integrator(kind=host.INTEGRATOR_PATH)),
    (host.Sampler(kind=host.SAMPLER_LD, spp=6), host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    (host.Sampler(kind=host.SAMPLER_LD, spp=8), host.Integrator(kind=host.INTEGRATOR_DIRECT, strategy=1)),
    (host.Sampler(kind=host.SAMPLER_LD, spp=2), host.Integrator(kind=host.INTEGRATOR_AO)),
    (host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=3, ys=2), host.Integrator(kind=host.INTEGRATOR_PATH)),
    (host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2, jitter=False), host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    (host.Sampler(kind=host.SAMPLER_RANDOM, spp=5), host.Integrator(kind=host.INTEGRATOR_PATH)),
    # index-shuffle fast path edge cases (one / two samples per pixel) and its fallback above 2048 samples per pixel
    (host.Sampler(kind=host.SAMPLER_LD, spp=1), host.Integrator(kind=host.INTEGRATOR_PATH)),
    (host.Sampler(kind=host.SAMPLER_LD, spp=2), host.Integrator(kind=host.INTEGRATOR_AO)),
    (host.Sampler(kind=host.SAMPLER_LD, spp=1024, seed=11), host.Integrator(kind=host.INTEGRATOR_PATH)),
    (host.Sampler(kind=host.SAMPLER_LD, spp=4096, seed=5), host.Integrator(kind=host.INTEGRATOR_PATH)),
])
def test_sampler_sequences_replay_the_oracle_bit_for_bit(sampler, integ):
    sb, cam = scenes.cornell_synth()
    # a second light with 4 samples exercises nSamples > 1 arrays of the direct-lighting layout
    sb.mesh([[-1, 9.8, -1], [1, 9.8, -1], [1, 9.8, 1], [-1, 9.8, 1]], [[0, 1, 2], [0, 2, 3]], area_light=(1, 2, 3), nsamples=3)
    g, o = _pair(sb.arrays())
    film = host.Film(32, 24)
    for c in (g, o):
        host.configure_render(c, cam, film, sampler, integ)
    for (x, y) in [(0, 0), (5, 7), (31, 23), (-1, 24)]:
        sg, so = g.pixel_samples(x, y), o.pixel_samples(x, y)
        assert sg.shape == so.shape and sg.shape[0] > 0
        assert np.array_equal(sg.view(np.uint32), so.view(np.uint32)), (x, y, np.argwhere(sg != so)[:4])


# ---- deterministic integrators ------------------------------------------------------------------------------
def test_ambient_occlusion_matches_oracle_per_pixel():
    arrays, cam = _cornell()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(96, 72), host.Sampler(kind=host.SAMPLER_LD, spp=2),
                                host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=32))
    assert np.array_equal(fg["weight"], fo["weight"])
    err = _rel_err(fg["rgb"], fo["rgb"])
    print("AO max rel err", err.max(), "pixels differing", int((fg["rgb"] != fo["rgb"]).any(axis=2).sum()))
    assert err.max() <= 1e-3
    sg, so = g.render_stats(), o.render_stats()
    assert sg["camera_samples"] == so["camera_samples"]
    assert sg["shadow_rays"] == so["shadow_rays"]


@pytest.mark.parametrize("ns", [1, 2, 4, 13, 64, 128])
def test_ambient_occlusion_ray_queue_layout_keeps_every_sample(ns):
    """The AO rays are queued by (0,2)-net cell and hit block, not in the reference's order (render_kernels.cu: aoRayPos): every sample
    of every hit must still be traced exactly once — the film equals the oracle's BIT FOR BIT (an occlusion count over 2^k rays) for
    sample counts that exercise every split of the cell grid (13 rounds up to 16), odd film sizes (a partial last block of hits)."""
    arrays, cam = _cornell()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(61, 37), host.Sampler(kind=host.SAMPLER_LD, spp=1),
                                host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=ns))
    assert np.array_equal(fg["rgb"], fo["rgb"]), int((fg["rgb"] != fo["rgb"]).any(axis=2).sum())
    sg, so = g.render_stats(), o.render_stats()
    assert sg["shadow_rays"] == so["shadow_rays"]


@pytest.mark.parametrize("strategy", [0, 1])
def test_direct_lighting_matches_oracle_per_pixel(strategy):
    sb, cam = scenes.cornell_synth()
    sb.point_light((0.0, 5.0, -5.0), (40.0, 30.0, 20.0))
    sb.sphere(host.translate(5, 6, 2), radius=0.8, area_light=(10.0, 10.0, 10.0), nsamples=2)
    arrays = sb.arrays()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(80, 60), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                                host.Integrator(kind=host.INTEGRATOR_DIRECT, strategy=strategy))
    err = _rel_err(fg["rgb"], fo["rgb"])
    print("direct max rel err", err.max(), "pixels differing", int((fg["rgb"] != fo["rgb"]).any(axis=2).sum()))
    assert err.max() <= 1e-3
    sg, so = g.render_stats(), o.render_stats()
    assert sg["shadow_rays"] == so["shadow_rays"]
    assert sg["closest_rays"] == so["closest_rays"]


def test_direct_lighting_stratified_sampler_and_oren_nayar():
    sb, cam = scenes.cornell_synth()
    rough = sb.material((0.6, 0.5, 0.4), sigma=25.0)
    sb.mesh([[-3, -9.9, -3], [3, -9.9, -3], [3, -9.9, 3], [-3, -9.9, 3]], [[0, 2, 1], [0, 3, 2]], material=rough)
    g, o, fg, fo = _render_both(sb.arrays(), cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2),
                                host.Integrator(kind=host.INTEGRATOR_DIRECT))
    assert _rel_err(fg["rgb"], fo["rgb"]).max() <= 1e-3


def _quadric_room():
    """cornell_synth plus one of each remaining quadric and an outward-emitting cylinder
    light (the only one of them with Shape.sample, cylinder.dart:230-240)."""
    sb, cam = scenes.cornell_synth()
    blue = sb.material((0.2, 0.3, 0.7))
    sb.cylinder(host.mat_mul(host.translate(5, 2, 3), host.rotate(70, (1, 0.2, 0))), radius=1.2, zmin=-2.0, zmax=2.5, phimax=300.0, material=blue)
    sb.cone(host.mat_mul(host.translate(-5, 3, 4), host.rotate(-90, (1, 0, 0))), radius=2.0, height=4.0, material=blue)
    sb.paraboloid(host.mat_mul(host.translate(0, -9, -2), host.rotate(-90, (1, 0, 0))), radius=2.5, zmin=0.5, zmax=4.0, material=blue)
    sb.hyperboloid(host.mat_mul(host.translate(5, -4, -3), host.rotate(-80, (1, 0, 0.1))), p1=(1.5, 0.0, -2.0), p2=(0.5, 1.2, 2.0),
                   material=blue, reverse=True)
    sb.cylinder(host.mat_mul(host.translate(-6, 7, 0), host.rotate(90, (0, 1, 0))), radius=0.4, zmin=-1.5, zmax=1.5,
                area_light=(20.0, 18.0, 12.0), nsamples=2)
    return sb.arrays(), cam


@pytest.mark.parametrize("strategy", [0, 1])
def test_quadrics_direct_lighting_with_a_cylinder_light_matches_oracle(strategy):
    arrays, cam = _quadric_room()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(96, 72), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                                host.Integrator(kind=host.INTEGRATOR_DIRECT, strategy=strategy))
    err = _rel_err(fg["rgb"], fo["rgb"])
    print("quadric room direct max rel err", err.max())
    assert err.max() <= 1e-3
    sg, so = g.render_stats(), o.render_stats()
    assert sg["shadow_rays"] == so["shadow_rays"] and sg["closest_rays"] == so["closest_rays"]


def test_quadrics_path_and_ao_match_oracle():
    arrays, cam = _quadric_room()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=8),
                                host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4))
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("quadric room path max rel err", err.max())
    assert err.max() <= 1e-3
    assert abs(fg["rgb"].mean() - fo["rgb"].mean()) <= 5e-3 * fo["rgb"].mean()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=1),
                                host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=16, ao_maxdist=6.0))
    assert (_rel_err(fg["rgb"], fo["rgb"]) > 1e-3).mean() <= 1e-3


def test_cone_light_is_rejected():
    sb, cam = scenes.cornell_synth()
    sb.cone(host.translate(0, 0, 0), material=0)
    a = sb.arrays()
    # hand the cone to the first light's ShapeSet: Shape.sample is unimplemented for it (shape.dart:83-86)
    a["light_shape_prims"] = np.asarray([a["idx"].shape[0] + a["sph_params"].shape[0]], np.uint32)
    a["light_shape_offsets"] = np.asarray([0, 1], np.uint32)
    g = capi.Context(0)
    host.upload_scene(g, a)
    host.configure_render(g, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_LD, spp=1), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    # a shape listed under a light has to carry that light's index itself (the MIS test reads it, integrator.dart:170)
    with pytest.raises(RuntimeError, match="whose light index is another light"):
        g.render(0, 1)
    a["quad_light"] = np.asarray([0], np.int32)
    g = capi.Context(0)
    host.upload_scene(g, a)
    host.configure_render(g, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_LD, spp=1), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    with pytest.raises(RuntimeError, match="area lights"):  # the scene tables are validated when the render starts
        g.render(0, 1)


def _smooth_room(which):
    """cornell_synth plus smooth-shaded meshes: per-vertex N / S / uv through Triangle.getShadingGeometry (triangle.dart:271-364)."""
    from tests.util import uv_sphere_mesh
    sb, cam = (scenes.cornell_materials() if which == "lobes" else scenes.cornell_synth())
    P, idx, N, S, UV = uv_sphere_mesh(10, 14, 1.0)
    m = sb.material_lobes(host.plastic_lobes((0.3, 0.5, 0.7), 0.4, 0.1)) if which == "lobes" else sb.material((0.3, 0.5, 0.7))
    # non-uniform scale + rotation: normals go through the inverse transpose, tangents through the matrix itself
    sb.mesh(P, idx, material=m, o2w=host.mat_mul(host.mat_mul(host.translate(4, 2, -2), host.rotate(35, (1, 0.4, 0.2))), host.scale(3.0, 2.0, 2.5)), N=N, S=S, uv=UV)
    sb.mesh(P, idx, material=m, o2w=host.mat_mul(host.translate(-5, 5, 3), host.scale(2.5, 2.5, 2.5)), N=N)          # normals only
    sb.mesh(P, idx, material=m, o2w=host.mat_mul(host.translate(0, -7, -5), host.scale(2.0, 2.0, 2.0)), uv=UV)        # uvs only
    sb.mesh(P, idx, material=m, o2w=host.mat_mul(host.translate(6, 7, 4), host.scale(1.5, 1.5, 1.5)), S=S, reverse=True)  # tangents only
    return sb.arrays(), cam


@pytest.mark.parametrize("which,integ", [
    ("matte", host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    ("matte", host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4)),
    ("lobes", host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4)),
    ("lobes", host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=3)),
    ("matte", host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=2)),
])
def test_smooth_shaded_meshes_match_oracle(which, integ):
    arrays, cam = _smooth_room(which)
    g, o, fg, fo = _render_both(arrays, cam, host.Film(80, 60), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("smooth room", which, integ.kind, "max rel err", err.max(), "q999", np.quantile(err, 0.999))
    assert err.max() <= 1e-3
    assert abs(fg["rgb"].mean() - fo["rgb"].mean()) <= 2e-3 * fo["rgb"].mean()
    if integ.kind == host.INTEGRATOR_DIRECT and which == "matte":
        assert err.max() <= 1e-3
        sg, so = g.render_stats(), o.render_stats()
        assert sg["shadow_rays"] == so["shadow_rays"] and sg["closest_rays"] == so["closest_rays"]
    # the attributes are in use: the same room without them renders differently
    a2 = dict(arrays)
    a2["mesh_shading"] = False
    g2 = capi.Context(0)
    host.upload_scene(g2, a2)
    host.configure_render(g2, cam, host.Film(80, 60), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
    g2.render(0, 1)
    assert np.abs(g2.film_read()["rgb"] - fg["rgb"]).max() > 1e-2


def _sky_scene(which, constant=False):
    """Objects on a ground plane under an InfiniteAreaLight (infinite_area_light.dart): a procedural lat-long map with a
    gradient and a bright patch (or the 1 x 1 constant map), rotated so that the map's pole is +y, plus a small quad light."""
    sb = host.SceneBuilder()
    lobes = which == "lobes"
    mat = (lambda kd: sb.material_lobes(host.matte_lobes(kd, 0.0))) if lobes else (lambda kd: sb.material(kd))
    ground = mat((0.5, 0.5, 0.45))
    sb.mesh([[-30, 0, -30], [30, 0, -30], [30, 0, 30], [-30, 0, 30]], [[0, 2, 1], [0, 3, 2]], material=ground)
    sb.sphere(host.translate(-2.5, 1.5, 0), radius=1.5, material=sb.material_lobes(host.glass_lobes(1.0, 1.0, 1.5)) if lobes else mat((0.7, 0.3, 0.3)))
    sb.sphere(host.translate(2.5, 1.5, 1), radius=1.5, material=sb.material_lobes(host.mirror_lobes((0.9, 0.9, 0.9))) if lobes else mat((0.3, 0.7, 0.3)))
    sb.cylinder(host.mat_mul(host.translate(0, 1.0, 3), host.rotate(90, (1, 0, 0))), radius=0.8, zmin=-1.0, zmax=1.0, material=mat((0.3, 0.3, 0.8)))
    sb.mesh([[-1, 6, -1], [1, 6, -1], [1, 6, 1], [-1, 6, 1]], [[0, 1, 2], [0, 2, 3]], material=ground, area_light=(6.0, 6.0, 6.0), nsamples=2)
    tex = None
    if not constant:
        w, h = 32, 16
        v, u = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing="ij")
        tex = np.stack([0.3 + 0.7 * (1 - v), 0.4 + 0.5 * (1 - v), 0.6 + 0.4 * u], axis=2).astype(np.float32)
        tex[2:5, 5:9, :] = (30.0, 28.0, 20.0)  # a sun
        tex[h // 2:, :, :] *= 0.2               # dim below the horizon
    sb.infinite_light((0.8, 0.9, 1.0), nsamples=4, light_to_world=host.mat_mul(host.rotate(-90, (1, 0, 0)), host.rotate(40, (0, 0, 1))),
                      texels=tex)
    cam = host.PerspectiveCamera(host.look_at((0, 4, -12), (0, 1.5, 0), (0, 1, 0)), fov=40.0)
    return sb.arrays(), cam


@pytest.mark.parametrize("which,constant,integ", [
    ("matte", False, host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    ("matte", True, host.Integrator(kind=host.INTEGRATOR_DIRECT, strategy=1)),
    ("matte", False, host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3)),
    ("lobes", False, host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)),
    ("lobes", False, host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=3)),
    ("lobes", True, host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=3)),
    ("matte", False, host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=8)),
])
def test_infinite_light_matches_oracle(which, constant, integ):
    arrays, cam = _sky_scene(which, constant)
    g, o, fg, fo = _render_both(arrays, cam, host.Film(80, 60), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("sky scene", which, constant, integ.kind, "max rel err", err.max(), "q999", np.quantile(err, 0.999))
    assert err.max() <= 1e-3
    assert abs(fg["rgb"].mean() - fo["rgb"].mean()) <= 2e-3 * fo["rgb"].mean()
    sky = fo["rgb"][:10].mean()
    assert sky > 0.05  # the escaped camera rays at the top of the frame see the map
    if integ.kind in (host.INTEGRATOR_DIRECT, host.INTEGRATOR_AO) and which == "matte":
        assert err.max() <= 1e-3
        sg, so = g.render_stats(), o.render_stats()
        assert sg["shadow_rays"] == so["shadow_rays"] and sg["closest_rays"] == so["closest_rays"]


def test_infinite_light_needs_its_map():
    arrays, cam = _sky_scene("matte")
    arrays["light_infinite"] = []
    g = capi.Context(0)
    host.upload_scene(g, arrays)
    host.configure_render(g, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_LD, spp=1), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    with pytest.raises(RuntimeError, match="drt_set_infinite_light"):
        g.render(0, 1)
    with pytest.raises(RuntimeError, match="power-of-two"):
        g.set_infinite_light(1, np.ones((3, 5, 3), np.float32), np.eye(4), np.eye(4))


@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4), host.Integrator(kind=host.INTEGRATOR_DIRECT),
                                   host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=8)])
def test_halton_sampler_matches_oracle(integ):
    """halton_sampler.dart: a global sequence scattered over the window with rejection, LatinHypercube integrator samples."""
    arrays, cam = _cornell()
    smp = host.Sampler(kind=host.SAMPLER_HALTON, spp=6, seed=3)
    g, o, fg, fo = _render_both(arrays, cam, host.Film(72, 40), smp, integ)
    sg, so = g.render_stats(), o.render_stats()
    assert sg["camera_samples"] == so["camera_samples"] > 0  # the same indices of the sequence are accepted
    assert np.array_equal(fg["weight"], fo["weight"])         # ... and land in the same pixels (box filter: sample counts)
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("halton", integ.kind, "max rel err", err.max())
    assert err.max() <= 1e-3
    if integ.kind != host.INTEGRATOR_PATH:
        assert err.max() <= 1e-3 and sg["shadow_rays"] == so["shadow_rays"] and sg["closest_rays"] == so["closest_rays"]
    # shards of the sequence tile it: the union of two shards is the whole render
    films = []
    for sh in range(2):
        gs = capi.Context(0)
        host.upload_scene(gs, arrays)
        host.configure_render(gs, cam, host.Film(72, 40), smp, integ)
        gs.render_shard(sh, 2)
        films.append(gs.film_read())
    assert np.array_equal(films[0]["weight"] + films[1]["weight"], fg["weight"])
    assert np.allclose(films[0]["xyz"] + films[1]["xyz"], fg["xyz"], rtol=1e-5, atol=1e-6)  # xyz: the unnormalised sums


@pytest.mark.parametrize("method,integ", [
    (host.ADAPTIVE_CONTRAST, host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3)),
    (host.ADAPTIVE_CONTRAST, host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    (host.ADAPTIVE_SHAPE_ID, host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    (host.ADAPTIVE_SHAPE_ID, host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=8)),
])
def test_adaptive_sampler_matches_oracle(method, integ):
    """adaptive_sampler.dart: minSamples everywhere, maxSamples where reportResults asks (contrast or shape ids), the first
    visit's samples dropped there."""
    arrays, cam = _cornell()
    smp = host.Sampler(kind=host.SAMPLER_ADAPTIVE, xs=2, ys=8, jitter=method, seed=9)
    g, o, fg, fo = _render_both(arrays, cam, host.Film(64, 48), smp, integ)
    # box filter: the weight is each pixel's own sample count, 2 or 8 — the same pixels are supersampled
    same = fg["weight"] == fo["weight"]
    print("adaptive", method, integ.kind, "supersampled", int((fo["weight"] == 8).sum()), "weights differing", int((~same).sum()))
    assert set(np.unique(fo["weight"])) == {2.0, 8.0}
    if method == host.ADAPTIVE_SHAPE_ID or integ.kind != host.INTEGRATOR_PATH:
        assert same.all()
        assert g.render_stats()["camera_samples"] == o.render_stats()["camera_samples"]
    else:
        assert (~same).mean() <= 2e-3  # a contrast ratio within rounding of 0.5 may fall either way
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)[same]
    assert err.max() <= 1e-3
    if integ.kind != host.INTEGRATOR_PATH:
        assert err.max() <= 1e-3


def _wrapped_materials_room():
    """cornell_synth with TranslucentMaterial, MixMaterial, ShinyMetalMaterial and SubstrateMaterial BSDFs (BRDFToBTDF / ScaledBxDF
    wrappers, FresnelBlend over an Anisotropic distribution)."""
    return scenes.cornell_synth({
        "grey": host.mix_lobes(host.matte_lobes((0.7, 0.7, 0.65)), host.plastic_lobes((0.2, 0.3, 0.6), 0.4, 0.1), amount=(0.3, 0.5, 0.7)),
        "red": host.shinymetal_lobes(ks=(0.8, 0.5, 0.3), kr=(0.2, 0.1, 0.1), roughness=0.15),
        "green": host.substrate_lobes(kd=(0.2, 0.5, 0.2), ks=(0.1, 0.15, 0.1), uroughness=0.3, vroughness=0.05),
        "box": host.translucent_lobes(kd=(0.6, 0.7, 0.5), ks=0.3, reflect=0.4, transmit=0.6, roughness=0.2),
        "sphere": host.mix_lobes(host.glass_lobes(1.0, 1.0, 1.5), host.matte_lobes((0.8, 0.4, 0.2)), amount=0.6),
    })


@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5), host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=3),
                                   host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=3)])
def test_translucent_mix_and_shinymetal_match_oracle(integ):
    sb, cam = _wrapped_materials_room()
    arrays = sb.arrays()
    assert (arrays["lobe_wrap"] & 1).any() and (arrays["lobe_wrap"] & 2).any()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(80, 60), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("wrapped materials", integ.kind, "max rel err", err.max(), "q999", np.quantile(err, 0.999))
    assert err.max() <= 1e-3
    assert abs(fg["rgb"].mean() - fo["rgb"].mean()) <= 2e-3 * fo["rgb"].mean()
    sg, so = g.render_stats(), o.render_stats()
    assert abs(sg["closest_rays"] - so["closest_rays"]) <= 1e-3 * so["closest_rays"]
    # the wrappers are in use: the same lobes without them render differently
    a2 = dict(arrays)
    a2["lobe_wrap"] = np.zeros_like(arrays["lobe_wrap"])
    g2 = capi.Context(0)
    host.upload_scene(g2, a2)
    host.configure_render(g2, cam, host.Film(80, 60), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
    g2.render(0, 1)
    assert np.abs(g2.film_read()["rgb"] - fg["rgb"]).max() > 1e-2


def _mapped_lights_room():
    """cornell_synth lit by projection and goniometric lights, with and without maps (drt_set_light_map)."""
    sb, cam = scenes.cornell_synth()
    w, h = 16, 8
    v, u = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing="ij")
    stripes = np.stack([0.5 + 0.5 * np.sin(12 * u), 0.5 + 0.5 * np.cos(9 * v), 0.3 + 0.7 * u * v], axis=2).astype(np.float32)
    sb.projection_light((300.0, 280.0, 250.0), fov=50.0, texels=stripes,
                        light_to_world=host.mat_mul(host.translate(-3, 6, -6), host.rotate(60, (1, 0.2, 0))))
    sb.projection_light((150.0, 150.0, 150.0), fov=30.0, light_to_world=host.mat_mul(host.translate(5, 8, -2), host.rotate(80, (1, 0, 0.3))))
    sb.goniometric_light((120.0, 140.0, 160.0), texels=stripes[:, ::-1].copy(), light_to_world=host.mat_mul(host.translate(0, 2, -4), host.rotate(30, (0, 0, 1))))
    sb.goniometric_light((40.0, 40.0, 40.0), light_to_world=host.translate(-6, -6, -3))
    return sb.arrays(), cam


def _wrapped_room_arrays():
    sb, cam = _wrapped_materials_room()
    return sb.arrays(), cam


@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_DIRECT), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3),
                                   host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=2)])
def test_projection_and_goniometric_lights_match_oracle(integ):
    """projection_light.dart / goniometric_light.dart: point lights scaled by a map lookup (drt_set_light_map)."""
    arrays, cam = _mapped_lights_room()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(80, 60), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("mapped lights", integ.kind, "max rel err", err.max())
    assert err.max() <= 1e-3
    if integ.kind == host.INTEGRATOR_DIRECT:
        assert err.max() <= 1e-3
        sg, so = g.render_stats(), o.render_stats()
        assert sg["shadow_rays"] == so["shadow_rays"] and sg["closest_rays"] == so["closest_rays"]
    # a projection / goniometric light without its drt_set_light_map is refused
    arrays["light_mapped"] = []
    g2 = capi.Context(0)
    host.upload_scene(g2, arrays)
    host.configure_render(g2, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_LD, spp=1), integ)
    with pytest.raises(RuntimeError, match="drt_set_light_map"):
        g2.render(0, 1)


@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4), host.Integrator(kind=host.INTEGRATOR_DIRECT)])
def test_best_candidate_sampler_matches_oracle(integ):
    """best_candidate_sampler.dart: the pattern table through drt_set_sample_table, per-tile shifts from dart:math Random."""
    from tests.util import synthetic_sample_table
    arrays, cam = _cornell()
    sb2 = None
    smp = host.Sampler(kind=host.SAMPLER_BEST_CANDIDATE, spp=5, seed=2, sample_table=synthetic_sample_table())
    cam.lens_radius, cam.focal_distance = 0.3, 30.0  # lens samples come from the pattern + the tile's shifts
    g, o, fg, fo = _render_both(arrays, cam, host.Film(80, 60), smp, integ)
    sg, so = g.render_stats(), o.render_stats()
    assert sg["camera_samples"] == so["camera_samples"] > 0
    assert np.array_equal(fg["weight"], fo["weight"])
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("bestcandidate", integ.kind, "max rel err", err.max())
    assert err.max() <= 1e-3
    if integ.kind == host.INTEGRATOR_DIRECT:
        assert err.max() <= 1e-3 and sg["shadow_rays"] == so["shadow_rays"]
    # without its table the sampler is refused
    g2 = capi.Context(0)
    host.upload_scene(g2, arrays)
    host.configure_render(g2, cam, host.Film(8, 8), host.Sampler(kind=host.SAMPLER_BEST_CANDIDATE, spp=4), integ)
    with pytest.raises(RuntimeError, match="drt_set_sample_table"):
        g2.render(0, 1)


@pytest.mark.parametrize("smp", [
    host.Sampler(kind=host.SAMPLER_ADAPTIVE, xs=2, ys=8, jitter=host.ADAPTIVE_CONTRAST, seed=4),
    host.Sampler(kind=host.SAMPLER_BEST_CANDIDATE, spp=4, seed=4),
    host.Sampler(kind=host.SAMPLER_HALTON, spp=3, seed=4),
])
def test_shards_and_tasks_of_the_sequence_samplers_tile_the_render(smp):
    """drt_render_shard splits the pixel list / the sample sequence in blocks; drt_render(task, count) gives every task its own
    sampler over its sub-window, as the reference does (dartray.dart:1009-1023)."""
    from tests.util import synthetic_sample_table
    if smp.kind == host.SAMPLER_BEST_CANDIDATE:
        smp.sample_table = synthetic_sample_table()
    arrays, cam = _cornell()
    integ = host.Integrator(kind=host.INTEGRATOR_DIRECT)
    film = host.Film(96, 64)

    def render(fn):
        c = capi.Context(0)
        host.upload_scene(c, arrays)
        host.configure_render(c, cam, film, smp, integ)
        fn(c)
        return c.film_read(), c.render_stats()
    whole, st = render(lambda c: c.render(0, 1))
    parts = [render(lambda c, k=k: c.render_shard(k, 3)) for k in range(3)]
    assert np.array_equal(sum(p[0]["weight"] for p in parts), whole["weight"])
    assert np.allclose(sum(p[0]["xyz"] for p in parts), whole["xyz"], rtol=1e-5, atol=1e-6)
    assert sum(p[1]["camera_samples"] for p in parts) == st["camera_samples"]
    # tasks: GPU == oracle task by task (each task is its own sampler window)
    for task in ((0, 2), (1, 2)):
        g, o, fg, fo = _render_both(arrays, cam, film, smp, integ, task=task)
        assert np.array_equal(fg["weight"], fo["weight"])
        assert _rel_err(fg["rgb"], fo["rgb"]).max() <= 1e-3


# ---- path tracing ------------------------------------------------------------------------------------------
def test_path_integrator_matches_oracle():
    arrays, cam = _cornell()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=16),
                                host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5))
    # same streams, same arithmetic: the per-pixel means agree far inside 3 sigma of the Monte Carlo noise
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("path max rel err", err.max(), "pixels differing", int((fg["rgb"] != fo["rgb"]).any(axis=2).sum()))
    sigma = fo["rgb"].std() / math.sqrt(16)
    assert np.abs(fg["rgb"] - fo["rgb"]).max() <= 3 * sigma
    assert abs(fg["rgb"].mean() - fo["rgb"].mean()) <= 5e-3 * fo["rgb"].mean()
    assert err.max() <= 1e-3
    sg, so = g.render_stats(), o.render_stats()
    assert abs(sg["closest_rays"] - so["closest_rays"]) <= 1e-3 * so["closest_rays"]
    assert abs(sg["shadow_rays"] - so["shadow_rays"]) <= 1e-3 * so["shadow_rays"]


def test_path_deep_bounces_use_the_integrator_stream():
    arrays, cam = _cornell()
    g, o, fg, fo = _render_both(arrays, cam, host.Film(40, 30), host.Sampler(kind=host.SAMPLER_RANDOM, spp=2),
                                host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=9))
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    assert err.max() <= 1e-3
    assert abs(fg["rgb"].mean() - fo["rgb"].mean()) <= 5e-3 * fo["rgb"].mean()


# ---- materials beyond matte: BxDF lists through drt_set_material_lobes -------------------------------
def _material_cornell(which):
    ov = {
        "specular": {"sphere": host.glass_lobes(1.0, 1.0, 1.5), "box": host.mirror_lobes((0.9, 0.85, 0.7))},
        "glossy": {"grey": host.plastic_lobes((0.6, 0.6, 0.55), 0.3, 0.08), "sphere": host.metal_lobes((0.2, 0.92, 1.1), (3.9, 2.45, 2.14), 0.05),
                   "box": host.matte_lobes((0.5, 0.4, 0.3), 25.0)},
        "uber": {"sphere": host.uber_lobes(kd=(0.3, 0.2, 0.2), ks=0.3, kr=0.2, kt=0.25, roughness=0.15, index=1.33, opacity=(0.8, 0.7, 0.9)),
                 "box": host.uber_lobes(kd=0.4, ks=0.2, roughness=0.3), "red": host.plastic_lobes((0.48, 0.1, 0.07), 0.2, 0.3)},
    }[which]
    sb, cam = scenes.cornell_synth(ov)
    return sb.arrays(), cam


@pytest.mark.parametrize("which", ["specular", "glossy", "uber"])
def test_path_integrator_with_bxdf_lists_matches_oracle(which):
    arrays, cam = _material_cornell(which)
    spp = 16
    g, o, fg, fo = _render_both(arrays, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=spp),
                                host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=7))
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print(which, "max rel err", err.max(), "q999", np.quantile(err, 0.999), "pixels differing", int((fg["rgb"] != fo["rgb"]).any(axis=2).sum()))
    assert np.isfinite(fg["rgb"]).all()
    # north_star: per-pixel mean within 3 sigma of the Monte Carlo noise; replayed streams do far better
    sigma = fo["rgb"].std() / math.sqrt(spp)
    assert np.abs(fg["rgb"] - fo["rgb"]).max() <= 3 * sigma
    assert abs(fg["rgb"].mean() - fo["rgb"].mean()) <= 5e-3 * fo["rgb"].mean()
    # every pixel, not a quantile: the replayed streams agree to ~2e-6 on a B200 (profiles/r02p_pytest_gpu.log)
    assert err.max() <= 1e-3
    sg, so = g.render_stats(), o.render_stats()
    assert abs(sg["closest_rays"] - so["closest_rays"]) <= 2e-3 * so["closest_rays"]
    assert abs(sg["shadow_rays"] - so["shadow_rays"]) <= 2e-3 * so["shadow_rays"]
    # and the materials do change the picture
    plain = capi.Context(0)
    sb0, _ = scenes.cornell_synth()
    host.upload_scene(plain, sb0.arrays())
    host.configure_render(plain, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=spp), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=7))
    plain.render()
    assert np.abs(plain.film_read()["rgb"] - fg["rgb"]).mean() > 1e-2


def test_direct_lighting_with_glossy_bxdf_lists_matches_oracle_per_pixel():
    arrays, cam = _material_cornell("glossy")
    for strategy in (0, 1):
        g, o, fg, fo = _render_both(arrays, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                                    host.Integrator(kind=host.INTEGRATOR_DIRECT, strategy=strategy, maxdepth=5))
        err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
        print("direct glossy strategy", strategy, "max rel err", err.max())
        assert err.max() <= 1e-3  # north_star: deterministic integrators per pixel within 1e-3 relative


def test_bxdf_lists_of_plain_matte_equal_the_matte_entry_point():
    """drt_set_material_lobes with one Lambertian / OrenNayar lobe per material takes the general kernels and must
    reproduce drt_set_materials bit for bit."""
    sb, cam = scenes.cornell_synth()
    sb.materials[3] = (0, (0.48, 0.48, 0.48), 20.0)  # OrenNayar on the box
    a = sb.arrays()
    film, smp, integ = host.Film(48, 36), host.Sampler(kind=host.SAMPLER_LD, spp=8), host.Integrator(kind=host.INTEGRATOR_PATH)
    g1, g2 = capi.Context(0), capi.Context(0)
    host.upload_scene(g1, a)
    b = dict(a)
    b["mat_general"] = True
    host.upload_scene(g2, b)
    for g in (g1, g2):
        host.configure_render(g, cam, film, smp, integ)
        g.render()
    assert np.array_equal(g1.film_read()["rgb"], g2.film_read()["rgb"])


@pytest.mark.parametrize("which,maxdepth,strategy", [("specular", 5, 0), ("specular", 3, 1), ("uber", 4, 0), ("specular", 1, 0)])
def test_directlighting_specular_recursion_matches_oracle(which, maxdepth, strategy):
    """DirectLightingIntegrator.Li with SpecularReflect / SpecularTransmit (integrator.dart:187-290): the GPU evaluates
    the recursion chain by chain; the uber case has two SpecularTransmission BxDFs, so the component draw of every
    branch call must come from the same position of the integrator stream as in the oracle's depth-first recursion."""
    arrays, cam = _material_cornell(which)
    g, o, fg, fo = _render_both(arrays, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                                host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=maxdepth, strategy=strategy))
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("direct + specular", which, maxdepth, "max rel err", err.max(), "q999", np.quantile(err, 0.999))
    assert err.max() <= 1e-3  # north_star: deterministic integrators per pixel within 1e-3 relative
    sg, so = g.render_stats(), o.render_stats()
    assert sg["closest_rays"] == so["closest_rays"] and abs(sg["shadow_rays"] - so["shadow_rays"]) <= 2
    if maxdepth > 1:  # the recursion does show: mirror box and glass sphere are not black
        g1 = capi.Context(0)
        host.upload_scene(g1, arrays)
        host.configure_render(g1, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                              host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=1, strategy=strategy))
        g1.render()
        assert np.abs(g1.film_read()["rgb"] - fg["rgb"]).mean() > 1e-3


@pytest.mark.parametrize("which,maxdepth", [("specular", 5), ("glossy", 5), ("uber", 4), ("specular", 1), (None, 3)])
def test_whitted_integrator_matches_oracle(which, maxdepth):
    """WhittedIntegrator.Li (whitted_integrator.dart:26-78): one LightSample.random(rng) per light at every vertex — so an
    area light's sample position depends on how many draws the recursion has consumed before the vertex is reached."""
    if which is None:
        sb, cam = scenes.cornell_synth()
    else:
        sb, cam = scenes.cornell_synth({
            "specular": {"sphere": host.glass_lobes(1.0, 1.0, 1.5), "box": host.mirror_lobes((0.9, 0.85, 0.7))},
            "glossy": {"grey": host.plastic_lobes((0.6, 0.6, 0.55), 0.3, 0.08), "sphere": host.metal_lobes((0.2, 0.92, 1.1), (3.9, 2.45, 2.14), 0.05)},
            "uber": {"sphere": host.uber_lobes(kd=(0.3, 0.2, 0.2), ks=0.3, kr=0.2, kt=0.25, roughness=0.15, index=1.33, opacity=(0.8, 0.7, 0.9))},
        }[which])
    sb.point_light((0.0, 5.0, -5.0), (40.0, 30.0, 20.0))
    sb.spot_light((-6.0, 8.0, -8.0), (2.0, -8.0, 2.0), (300.0, 300.0, 400.0), 25.0, 8.0)
    g, o, fg, fo = _render_both(sb.arrays(), cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4),
                                host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=maxdepth))
    err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
    print("whitted", which, maxdepth, "max rel err", err.max())
    assert err.max() <= 1e-3 and fo["rgb"].mean() > 0.05
    sg, so = g.render_stats(), o.render_stats()
    assert sg["closest_rays"] == so["closest_rays"] and abs(sg["shadow_rays"] - so["shadow_rays"]) <= 2


def test_distant_and_spot_lights_match_oracle_per_pixel():
    """lib/lights/distant_light.dart:41-48 and spot_light.dart:36-70 through drt_set_lights kinds 2 / 3 + drt_set_spot_params."""
    sb, cam = scenes.cornell_synth()
    sb.distant_light((3.0, 6.0, -10.0), (0, 0, 0), (2.0, 1.5, 1.0))
    sb.spot_light((-6.0, 8.0, -8.0), (2.0, -8.0, 2.0), (300.0, 300.0, 400.0), 25.0, 8.0)
    arrays = sb.arrays()
    for integ in (host.Integrator(kind=host.INTEGRATOR_DIRECT), host.Integrator(kind=host.INTEGRATOR_DIRECT, strategy=1),
                  host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4)):
        g, o, fg, fo = _render_both(arrays, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
        err = _rel_err(fg["rgb"], fo["rgb"], floor=1e-3)
        print("delta lights, integrator", integ.kind, "max rel err", err.max())
        assert err.max() <= 1e-3
    # the spot's cone edge is in the picture: some pixels lit by it, some not
    base, _ = scenes.cornell_synth()
    g0 = capi.Context(0)
    host.upload_scene(g0, base.arrays())
    host.configure_render(g0, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4))
    g0.render()
    assert np.abs(g0.film_read()["rgb"] - fg["rgb"]).max() > 0.05


def test_thin_lens_camera_and_random_sampler():
    """perspective_camera.dart:104-119: lensRadius > 0 moves the ray origin on the lens (ConcentricSampleDisk)."""
    sb, cam = scenes.cornell_synth()
    cam.lens_radius, cam.focal_distance = 0.8, 33.0
    for sampler in (host.Sampler(kind=host.SAMPLER_LD, spp=8), host.Sampler(kind=host.SAMPLER_RANDOM, spp=3),
                    host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2)):
        g, o, fg, fo = _render_both(sb.arrays(), cam, host.Film(48, 36), sampler, host.Integrator(kind=host.INTEGRATOR_DIRECT))
        assert np.array_equal(fg["weight"], fo["weight"])
        assert _rel_err(fg["rgb"], fo["rgb"], floor=1e-3).max() <= 1e-3
    sharp = host.PerspectiveCamera(cam.camera_to_world, fov=cam.fov)
    g2 = capi.Context(0)
    host.upload_scene(g2, sb.arrays())
    host.configure_render(g2, sharp, host.Film(48, 36), host.Sampler(kind=host.SAMPLER_LD, spp=8), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    g2.render()
    assert np.abs(g2.film_read()["rgb"] - fg["rgb"]).max() > 1e-3  # the lens does change the image


def test_orthographic_and_environment_cameras():
    """lib/cameras/orthographic_camera.dart:52-80 and environment_camera.dart:42-52."""
    sb, pcam = scenes.cornell_synth()
    integ = host.Integrator(kind=host.INTEGRATOR_DIRECT)
    ortho = host.OrthographicCamera(pcam.camera_to_world, screen_window=(-10.0, 10.0, -7.5, 7.5), lens_radius=0.3, focal_distance=30.0)
    env = host.EnvironmentCamera(host.look_at((0, 0, 0), (0, 0, 1), (0, 1, 0)))
    images = []
    for cam in (ortho, env):
        g, o, fg, fo = _render_both(sb.arrays(), cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
        assert fo["rgb"].mean() > 0.05
        assert _rel_err(fg["rgb"], fo["rgb"], floor=1e-3).max() <= 1e-3
        assert g.render_stats()["closest_rays"] == o.render_stats()["closest_rays"]
        images.append(fg["rgb"])
    assert np.abs(images[0] - images[1]).max() > 0.05
    with pytest.raises(capi.DrtError):
        capi.Context(0).set_camera_kind(7)


# ---- film --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flt", ["gaussian", "mitchell", "triangle", "sinc", "box"])
def test_filters_and_crop_window(flt):
    arrays, cam = _cornell()
    film = host.Film(48, 36, filter=flt, crop=(0.25, 0.9, 0.1, 0.8))
    g, o, fg, fo = _render_both(arrays, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=4),
                                host.Integrator(kind=host.INTEGRATOR_DIRECT))
    assert g.film_size() == o.film_size()
    assert np.allclose(fg["weight"], fo["weight"], rtol=1e-5, atol=1e-6)
    assert _rel_err(fg["rgb"], fo["rgb"], floor=1e-3).max() <= 1e-3


def test_tasks_and_shards_tile_the_image():
    arrays, cam = _cornell()
    film = host.Film(50, 38)
    sampler, integ = host.Sampler(kind=host.SAMPLER_LD, spp=2), host.Integrator(kind=host.INTEGRATOR_DIRECT)
    full = capi.Context(0)
    host.upload_scene(full, arrays)
    host.configure_render(full, cam, film, sampler, integ)
    full.render()
    ref = full.film_read()
    # reference semantics: task sub-windows of the sample extent (dartray.dart:1009-1023), one film
    tasks = capi.Context(0)
    host.upload_scene(tasks, arrays)
    host.configure_render(tasks, cam, film, sampler, integ)
    o = Oracle()
    host.upload_scene(o, arrays)
    host.configure_render(o, cam, film, sampler, integ)
    for t in range(3):
        tasks.render(t, 3)
        o.render(t, 3, 4)
    ft = tasks.film_read()
    assert np.array_equal(ft["weight"], ref["weight"])
    assert np.allclose(ft["xyz"], ref["xyz"], rtol=1e-6, atol=1e-7)
    assert _rel_err(ft["rgb"], o.film_read()["rgb"], floor=1e-3).max() <= 1e-3
    # load-balanced shards: interleaved pixel blocks
    sh = capi.Context(0)
    host.upload_scene(sh, arrays)
    host.configure_render(sh, cam, film, sampler, integ)
    sh.set_batch_slots(700)  # several batches per shard
    for s in range(4):
        sh.render_shard(s, 4)
    fs = sh.film_read()
    assert np.array_equal(fs["weight"], ref["weight"])
    assert np.allclose(fs["xyz"], ref["xyz"], rtol=1e-6, atol=1e-7)
    assert sh.render_stats()["camera_samples"] == full.render_stats()["camera_samples"]


def test_empty_scene_and_no_lights():
    cam = host.PerspectiveCamera(host.look_at((0, 0, -5), (0, 0, 0), (0, 1, 0)), fov=40.0)
    sb = host.SceneBuilder()
    g = capi.Context(0)
    host.upload_scene(g, sb.arrays())
    host.configure_render(g, cam, host.Film(16, 12), host.Sampler(spp=1), host.Integrator(kind=host.INTEGRATOR_PATH))
    g.render()
    f = g.film_read()
    assert (f["rgb"] == 0).all() and (f["weight"] == 1).all()
    # geometry but no lights: black, and the integrator draws differ (UniformSampleOneLight returns early)
    sb = host.SceneBuilder()
    sb.sphere(host.translate(0, 0, 0), radius=1.0, material=sb.material((0.5, 0.5, 0.5)))
    g, o, fg, fo = _render_both(sb.arrays(), cam, host.Film(16, 12), host.Sampler(spp=2),
                                host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=6))
    assert (fg["rgb"] == 0).all() and (fo["rgb"] == 0).all()
    assert g.render_stats()["closest_rays"] == o.render_stats()["closest_rays"]


def test_render_errors():
    g = capi.Context(0)
    with pytest.raises(capi.DrtError):
        g.render()  # nothing set
    with pytest.raises(ValueError):
        g.set_sampler(0, 1, 1, 4, 1, 1, 32, 0, rng_mode=0)  # serial stream cannot be replayed in parallel
    with pytest.raises(capi.DrtError):
        g.set_integrator(5, 5, 0, 1, 0.0, 1.0)
    with pytest.raises(capi.DrtError):
        g.set_materials(np.array([3], np.int32), np.ones((1, 3), np.float32), np.zeros(1, np.float32))


# ---- drt_render_profile: event spans per kernel class, reference-walk work of every traced queue --------------------------------
def test_render_profile_counts_the_reference_work_of_every_ray():
    arrays, cam = _cornell()
    g, o = _pair(arrays)
    film, smp = host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    for c in (g, o):
        host.configure_render(c, cam, film, smp, integ)
    g.set_render_profiling(capi.PROFILE_TIME | capi.PROFILE_WORK)
    g.film_clear()
    g.render()
    o.render(0, 1, 8)
    prof, sg, so = g.render_profile(), g.render_stats(), o.render_stats()
    # the counting walk saw exactly the rays the render traced ...
    assert prof["closest"]["rays"] == sg["closest_rays"] == so["closest_rays"]
    assert prof["any"]["rays"] == sg["shadow_rays"] == so["shadow_rays"]
    # ... and did the slab / primitive tests the oracle's walk did (bvh_accel.dart:125/131/187/193)
    assert prof["closest"]["nodes_visited"] + prof["any"]["nodes_visited"] == so["nodes_visited"]
    assert prof["closest"]["prims_tested"] + prof["any"]["prims_tested"] == so["prims_tested"]
    # every class that launched took time; the picture is unchanged by profiling
    for k in ("trace_closest", "trace_any", "integrator", "sampler", "resolve", "film"):
        assert prof["launches"][k] > 0 and prof["ms"][k] > 0.0, k
    a = g.film_read()["rgb"]
    g.set_render_profiling(0)
    g.film_clear()
    g.render()
    assert np.array_equal(a, g.film_read()["rgb"])
    assert g.render_profile()["launches"]["film"] == 0  # cleared by drt_film_clear, nothing recorded with profiling off
