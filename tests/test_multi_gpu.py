"""drt_create_multi: one context over several GPUs, through the C ABI only (ctypes, no torch, no NCCL in the test process).

The reference's precedent is lib/dartray_web/render_manager.dart:100-141 (one isolate per image region, each with a private
copy of the scene, regions copied into one image).  Here: one BVH build, the arrays uploaded to every device, the render
split in interleaved pixel blocks, the films summed into the first device's over NVLink.  Keyed sample streams make the union
of the devices' samples exactly the one-device sample set, so the multi-device film must EQUAL the one-device film.

On a one-GPU box DRT_ALLOW_DUPLICATE_DEVICES=1 lets two contexts of the same device stand in for two GPUs: every line of the
multi-device path runs (forwarded setters, shared build, per-device host threads, film sum, delta semantics)."""
import os

import numpy as np
import pytest

from dartray_b200 import capi, host, scenes

pytestmark = pytest.mark.gpu


def _devices():
    """All the GPUs of the box when there are several, else device 0 twice."""
    ids = []
    for d in range(16):
        try:
            capi.Context(d).close()
            ids.append(d)
        except capi.DrtError:
            break
    if len(ids) >= 2:
        return ids
    os.environ["DRT_ALLOW_DUPLICATE_DEVICES"] = "1"
    return [0, 0]


def _render(ctx, arrays, cam, film, smp, integ, passes=1):
    host.upload_scene(ctx, arrays)
    host.configure_render(ctx, cam, film, smp, integ)
    ctx.film_clear()
    for _ in range(passes):
        ctx.render()
    return ctx.film_read(), ctx.render_stats()


def test_multi_device_film_equals_the_one_device_film_config4():
    """BASELINE.json configs[3] geometry and integrator (cornell_synth, path maxdepth 5, lowdiscrepancy), 1920x1080 at 16 spp."""
    ids = _devices()
    sb, cam = scenes.cornell_synth()
    arrays = sb.arrays()
    film, smp = host.Film(1920, 1080), host.Sampler(kind=host.SAMPLER_LD, spp=16)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    one, multi = capi.Context(ids[0]), capi.Context(ids)
    assert one.device_count == 1 and multi.device_count == len(ids)
    f1, s1 = _render(one, arrays, cam, film, smp, integ)
    fm, sm = _render(multi, arrays, cam, film, smp, integ)
    assert s1 == sm  # camera samples and ray counts of all devices together
    assert np.array_equal(f1["weight"], fm["weight"])
    # box filter: a pixel's samples all come from the device that owns its block, the others add zeros -> bit-equal
    assert np.array_equal(f1["xyz"], fm["xyz"]) and np.array_equal(f1["rgb"], fm["rgb"])
    # a second pass into the same film adds ITS samples once (the peers' films hold the delta since the last sum)
    multi.render()
    f2 = multi.film_read()
    assert np.array_equal(f2["weight"], 2.0 * f1["weight"])
    assert np.allclose(f2["rgb"], f1["rgb"], rtol=1e-6, atol=1e-7)  # same samples twice: same mean
    # task windows (taskNum / taskCount, dartray.dart:1009-1023) split over the devices as well
    multi.film_clear()
    for t in range(3):
        multi.render(t, 3)
    f3 = multi.film_read()
    assert np.array_equal(f3["weight"], f1["weight"]) and np.allclose(f3["rgb"], f1["rgb"], rtol=1e-6, atol=1e-7)


def test_multi_device_context_forwards_the_shading_precision():
    """drt_set_shading_precision on a multi-device context reaches every device: the float32 film of the devices together equals the
    float32 film of one device bit for bit (same kernels, keyed streams), and differs from the binary64 film."""
    ids = _devices()
    sb, cam = scenes.cornell_synth()
    arrays = sb.arrays()
    film, smp = host.Film(480, 270), host.Sampler(kind=host.SAMPLER_LD, spp=16)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    one, multi = capi.Context(ids[0]), capi.Context(ids)
    f64, _ = _render(one, arrays, cam, film, smp, integ)
    for c in (one, multi):
        c.set_shading_precision(capi.PRECISION_F32)
    f1, s1 = _render(one, arrays, cam, film, smp, integ)
    fm, sm = _render(multi, arrays, cam, film, smp, integ)
    assert s1 == sm
    assert np.array_equal(f1["rgb"], fm["rgb"]) and np.array_equal(f1["weight"], fm["weight"])
    assert not np.array_equal(f1["rgb"], f64["rgb"])


def test_multi_device_filter_footprints_cross_block_borders():
    """A wide gaussian filter spreads every sample over pixels of other devices' blocks: the SUM of the films (not a copy of
    regions) keeps those contributions; AO on a mesh with per-device BVH copies."""
    ids = _devices()
    P, idx = scenes.soup(8)
    sb = host.SceneBuilder()
    sb.mesh(P, idx)
    arrays = sb.arrays()
    cam = host.PerspectiveCamera(host.look_at((0, 0, -4), (0, 0, 0), (0, 1, 0)), fov=40.0)
    film = host.Film(320, 200, filter="gaussian", xwidth=2.0, ywidth=2.0)
    smp = host.Sampler(kind=host.SAMPLER_LD, spp=4)
    integ = host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=16)
    f1, s1 = _render(capi.Context(ids[0]), arrays, cam, film, smp, integ)
    fm, sm = _render(capi.Context(ids), arrays, cam, film, smp, integ)
    assert s1 == sm
    assert np.allclose(f1["weight"], fm["weight"], rtol=1e-12)
    assert np.abs(f1["rgb"] - fm["rgb"]).max() <= 1e-6  # f64 film sums in a different order


def test_multi_device_context_shares_one_bvh_build_and_answers_queries_on_the_first_device():
    ids = _devices()
    P, idx = scenes.soup(8)
    one, multi = capi.Context(ids[0]), capi.Context(ids)
    for c in (one, multi):
        c.set_triangles(P, idx)
        c.build_bvh()
    assert one.bvh_info()["n_nodes"] == multi.bvh_info()["n_nodes"]
    for k in ("offset", "n_primitives", "axis", "ordered", "bounds"):
        assert np.array_equal(one.bvh_export()[k], multi.bvh_export()[k]), k
    ro, rd = scenes.incoherent_rays(1 << 15)
    a, b = one.trace_closest(ro, rd), multi.trace_closest(ro, rd)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    with pytest.raises(capi.DrtError):
        os.environ.pop("DRT_ALLOW_DUPLICATE_DEVICES", None)
        capi.Context([ids[0], ids[0]])
