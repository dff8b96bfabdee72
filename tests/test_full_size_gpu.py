"""BASELINE.json configs at FULL size (soup_1m: 1,015,810 triangles; 8,388,608 coherent + 8,388,608 incoherent rays;
1920x1080 films), where the CPU oracle cannot check every ray in seconds: size-independent properties instead.

  * the production kernel (traceFastKernel) against the literal f64 walk (traceKernel, DRT_KERNEL_EXACT_WALK): two
    independent GPU implementations of bvh_accel.dart:101-226, bit-identical on every ray;
  * the oracle itself on a strided subsample of the same rays against the same 1M-triangle tree, bit-exact;
  * geometry: o + t d lies on the reported triangle at the reported barycentrics;
  * closest-hit / any-hit coupling: a shadow ray cut just short of the closest hit is unoccluded, one running just
    past it is occluded;
  * ambient occlusion at 1080p x 64: every pixel a multiple of 1/64, weights exactly one, the oracle's image on a
    coarse crop of the same camera;  path tracing at 1080p: image mean against the oracle's low-resolution render."""
import numpy as np
import pytest

from dartray_b200 import capi, host, scenes
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def soup_1m():
    P, idx = scenes.soup(512)
    assert idx.shape[0] == 1_015_810
    c = capi.Context(0)
    c.set_triangles(P, idx)
    c.build_bvh(capi.SPLIT_SAH, 4)
    return P, idx, c


@pytest.fixture(scope="module")
def ray_sets():
    return {"coherent": scenes.coherent_rays(4096, 2048), "incoherent": scenes.incoherent_rays(8_388_608)}


@pytest.mark.parametrize("which", ["coherent", "incoherent"])
def test_production_kernel_equals_the_literal_walk_on_every_ray(soup_1m, ray_sets, which):
    P, idx, c = soup_1m
    ro, rd = ray_sets[which]
    c.set_kernel_variant(capi.KERNEL_FAST)  # soup_1m: the quantised-node kernel (trace_fast2.cu)
    assert c.bvh_info()["device_bytes"] < 90e6  # 64-byte wide nodes: 81 MB; the float32 nodes make it 114 MB
    fast, fast_any = c.trace_closest(ro, rd), c.trace_any(ro, rd)
    try:
        c.set_kernel_variant(capi.KERNEL_FAST_V1)
        v1, v1_any = c.trace_closest(ro, rd), c.trace_any(ro, rd)
        c.set_kernel_variant(capi.KERNEL_EXACT_WALK)
        walk, walk_any = c.trace_closest(ro, rd), c.trace_any(ro, rd)
    finally:
        c.set_kernel_variant(capi.KERNEL_FAST)
    for other, other_any in ((walk, walk_any), (v1, v1_any)):
        assert np.array_equal(fast["prim"], other["prim"])
        for k in ("t", "b1", "b2"):
            assert np.array_equal(fast[k].view(np.uint32), other[k].view(np.uint32)), k
        assert np.array_equal(fast_any, other_any)
    hit = fast["prim"] >= 0
    assert 0.2 < hit.mean() < 1.0
    # shadow rays with the full interval: intersectP's float32 edge vectors (triangle.dart:162-194) may disagree with
    # intersect's f64 ones only for grazing hits
    assert np.mean(fast_any.astype(bool) != hit) < 1e-5

    # o + t d is the point of the reported triangle at the reported barycentrics (triangle.dart:77-95)
    o, d = ro[hit, :3].astype(np.float64), rd[hit, :3].astype(np.float64)
    t, b1, b2 = (fast[k][hit].astype(np.float64) for k in ("t", "b1", "b2"))
    tri = P[idx[fast["prim"][hit]]].astype(np.float64)
    p_bary = (1.0 - b1 - b2)[:, None] * tri[:, 0] + b1[:, None] * tri[:, 1] + b2[:, None] * tri[:, 2]
    p_ray = o + t[:, None] * d
    assert np.abs(p_ray - p_bary).max() <= 1e-5 * max(1.0, np.abs(p_ray).max())
    assert (b1 >= 0).all() and (b2 >= 0).all() and (b1 + b2 <= 1.0 + 1e-7).all()

    # closest-hit / any-hit coupling on the rays that hit
    ro_h, rd_short, rd_long = ro[hit].copy(), rd[hit].copy(), rd[hit].copy()
    rd_short[:, 3] = fast["t"][hit] * np.float32(1.0 - 1e-3)
    rd_long[:, 3] = fast["t"][hit] * np.float32(1.0 + 1e-3)
    # a ray cut short of its closest hit is unoccluded — except where intersectP's float32 edge vectors
    # (triangle.dart:162-194) hit a grazing triangle that intersect's f64 ones miss: a handful of rays in millions
    assert np.mean(c.trace_any(ro_h, rd_short) != 0) < 1e-5, "rays cut short of their closest hit are occluded"
    # ... and one running past it is occluded, except where intersectP's float32 edge vectors miss a grazing triangle
    # that intersect's f64 ones hit (the same handful of rays as above)
    assert np.mean(c.trace_any(ro_h, rd_long) == 0) < 1e-5, "rays running past their closest hit are not occluded"


def test_oracle_agrees_on_a_strided_subsample_of_the_full_workload(soup_1m, ray_sets):
    P, idx, c = soup_1m
    o = Oracle()
    o.set_triangles(P, idx)
    o.build_bvh(2, 4)
    for name, (ro, rd) in ray_sets.items():
        ro_s, rd_s = np.ascontiguousarray(ro[5::64]), np.ascontiguousarray(rd[5::64])
        hg, ho = c.trace_closest(ro_s, rd_s), o.trace_closest(ro_s, rd_s, nthreads=16)
        assert np.array_equal(hg["prim"], ho["prim"]), name
        for k in ("t", "b1", "b2"):
            assert np.array_equal(hg[k].view(np.uint32), ho[k].view(np.uint32)), (name, k)
        assert np.array_equal(c.trace_any(ro_s, rd_s), o.trace_any(ro_s, rd_s, nthreads=16)), name
    # no exact-t ties in either set: the runner-up primitive never shares the winner's t
    # (checked on the subsample through the oracle's brute-force walk being order-independent)
    ro_s, rd_s = np.ascontiguousarray(ray_sets["incoherent"][0][::4096]), np.ascontiguousarray(ray_sets["incoherent"][1][::4096])
    hb, nties, _second = o.trace_closest_brute(ro_s, rd_s, nthreads=16)
    hg = c.trace_closest(ro_s, rd_s)
    assert (nties[hb["prim"] >= 0] <= 1).all()  # the exhaustive walk saw no second primitive at the winner's t
    assert np.array_equal(hb["prim"], hg["prim"]) and np.array_equal(hb["t"].view(np.uint32), hg["t"].view(np.uint32))


def test_ambient_occlusion_at_1080p_has_the_structure_of_the_estimator(soup_1m):
    P, idx, c = soup_1m
    cam = host.PerspectiveCamera(host.look_at((0, 0, -4), (0, 0, 0), (0, 1, 0)), fov=40.0)
    smp = host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False)
    integ = host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=64)
    host.configure_render(c, cam, host.Film(1920, 1080), smp, integ)
    c.film_clear()
    c.render()
    f = c.film_read()
    assert (f["weight"] == 1.0).all()
    v = f["rgb"].astype(np.float64) * 64.0
    assert np.abs(v - np.round(v)).max() < 1e-3  # nClear / 64 through the film's float32 XYZ round trip
    assert 0.0 <= f["rgb"].min() and f["rgb"].max() <= 1.0 + 1e-6
    st = c.render_stats()
    assert st["closest_rays"] == st["camera_samples"] == 1921 * 1081  # one camera ray per sample of the sample extent
    assert st["shadow_rays"] % 64 == 0 and 0 < st["shadow_rays"] <= 64 * st["camera_samples"]  # 64 per camera ray that hit
    # the oracle on a coarse film of the same camera: pixel centres of a 60x34 film are not 1080p pixel centres, so
    # compare the image means (the AO field is smooth at that scale)
    o = Oracle()
    o.set_triangles(P, idx)
    o.build_bvh(2, 4)
    host.configure_render(o, cam, host.Film(120, 68), smp, integ)
    o.render(0, 1, 16)
    assert abs(f["rgb"].mean() - o.film_read()["rgb"].mean()) < 0.02


def test_path_tracing_at_1080p_matches_the_oracle_image_mean():
    """SURVEY 8d config 4 parity: image mean within 0.5 % of the oracle (its render is a reduced film of the same camera)."""
    sb, cam = scenes.cornell_synth()
    g, o = capi.Context(0), Oracle()
    for x in (g, o):
        host.upload_scene(x, sb.arrays())
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    host.configure_render(g, cam, host.Film(1920, 1080), host.Sampler(kind=host.SAMPLER_LD, spp=16), integ)
    host.configure_render(o, cam, host.Film(240, 135), host.Sampler(kind=host.SAMPLER_LD, spp=64), integ)
    g.render()
    o.render(0, 1, 16)
    fg, fo = g.film_read()["rgb"].astype(np.float64), o.film_read()["rgb"].astype(np.float64)
    # box filter of width 0.5: a pixel gets its own 16 samples, plus a neighbour's whenever a sample offset is exactly 0
    # (image_film.dart:108-111: ceil(x - 1) .. floor(x) is two pixels then; probability 2^-24 per sample and axis)
    w = g.film_read()["weight"]
    assert np.isfinite(fg).all() and (np.abs(w - 16.0) <= 1.0).all() and (w == 16.0).mean() > 0.9999
    assert abs(fg.mean() - fo.mean()) <= 5e-3 * fo.mean(), (fg.mean(), fo.mean())
    # 8x8 block means of the 1080p image against the oracle's pixels: the same integral per low-res pixel, two
    # independent sample sets -> within 3 sigma of the oracle's per-pixel Monte Carlo noise almost everywhere
    blocks = fg.reshape(135, 8, 240, 8, 3).mean(axis=(1, 3))
    sigma = max(fo.std() / np.sqrt(64.0), 1e-3)
    assert np.mean(np.abs(blocks - fo) > 3.0 * sigma * 4.0) < 0.02
