"""GPU parity: libdartray_gpu's CUDA traversal vs the CPU oracle, through the C ABI.
Bar (BASELINE.json north_star): closest-hit primitive index bit-exact, t and barycentrics within
1e-5 relative.  The kernels reproduce the reference arithmetic op for op, so the tests below ask
for more: float32(t), b1, b2 BIT-IDENTICAL and the any-hit flag identical."""
import numpy as np
import pytest

from dartray_b200 import capi, host, scenes
from tests.oracle_lib import Oracle
from tests.util import mesh_refine_order, random_rays, random_soup, translate

pytestmark = pytest.mark.gpu


# DRT_KERNEL_FAST_Q (quantised 64-byte nodes, forced whatever the scene size), DRT_KERNEL_FAST_V1 (float32 128-byte nodes),
# DRT_KERNEL_EXACT_WALK (the literal walk); DRT_KERNEL_FAST = 0 picks between the first two by scene size
# by scene size — and runs the leaf-list kernel (traceSmallKernel) on scenes of <= 32 leaves
@pytest.fixture(params=[3, 2, 1, 0], ids=["fast_q", "fast_v1", "exact_walk", "default"])
def variant(request):
    return request.param


def make_pair(P, idx, spheres=None, order=None, split=2, maxprims=4, variant=0, disks=None):
    o = Oracle()
    c = capi.Context(0)
    c.set_kernel_variant(variant)
    for x in (o, c):
        x.set_triangles(P, idx)
        if spheres is not None:
            x.set_spheres(*spheres)
        if disks is not None:
            x.set_disks(*disks)
        x.set_build_order(order)
        x.build_bvh(split, maxprims)
    return o, c


def assert_hits_equal(hg, ho):
    assert (hg["prim"] == ho["prim"]).all(), f"{(hg['prim'] != ho['prim']).sum()} primitive ids differ"
    for k in ("t", "b1", "b2"):
        assert (hg[k].view(np.uint32) == ho[k].view(np.uint32)).all(), k


@pytest.mark.parametrize("split", [0, 1, 2])
def test_random_soup_closest_and_any(drt_lib, split, variant):
    P, idx = random_soup(5000, seed=split)
    o, c = make_pair(P, idx, split=split, variant=variant)
    ro, rd = random_rays(50000, seed=100 + split)
    assert_hits_equal(c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8))
    assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()
    # bounded intervals, as shadow rays use them
    ro, rd = random_rays(50000, seed=200 + split, tmin=0.7, tmax=2.9)
    assert_hits_equal(c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8))
    assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()


def test_mixed_triangles_and_spheres(drt_lib, variant):
    P, idx = random_soup(800, seed=5)
    mats = [translate(0.3, 0.1, -0.2), translate(-0.4, 0.2, 0.5), translate(0.0, -0.5, 0.0)]
    sph = (np.stack([m[0] for m in mats]), np.stack([m[1] for m in mats]),
           [[0.25, -0.25, 0.25, 360.0], [0.4, -0.1, 0.3, 200.0], [0.3, -0.3, 0.1, 360.0]])
    o, c = make_pair(P, idx, sph, mesh_refine_order([300, 500], 3), variant=variant)
    ro, rd = random_rays(40000, seed=6)
    hg, ho = c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8)
    assert (hg["prim"] == ho["prim"]).all()
    assert (hg["t"].view(np.uint32) == ho["t"].view(np.uint32)).all()
    sphere_hit = ho["prim"] >= 800
    assert sphere_hit.sum() > 1000
    # sphere (u, v) go through atan2/acos: libm vs CUDA may differ in the last ulp of the f64
    np.testing.assert_allclose(hg["b1"], ho["b1"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(hg["b2"], ho["b2"], rtol=1e-6, atol=1e-7)
    assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()


def test_disks_with_triangles_and_spheres(drt_lib, variant):
    """Disk (lib/shapes/disk.dart): quadric ids continue after the spheres."""
    P, idx = random_soup(600, seed=9)
    mats = [translate(0.3, 0.1, -0.2), translate(-0.4, 0.2, 0.5)]
    sph = (np.stack([m[0] for m in mats]), np.stack([m[1] for m in mats]), [[0.25, -0.25, 0.25, 360.0], [0.3, -0.1, 0.2, 300.0]])
    rot = host.rotate(35.0, (1.0, 0.3, 0.2))
    dm = [host.mat_mul(host.translate(0.1, -0.3, 0.2), rot), host.translate(-0.5, 0.5, -0.4), host.mat_mul(host.translate(0.6, 0.6, 0.1), host.rotate(80, (0, 1, 0)))]
    dsk = (np.stack([m.reshape(16) for m in dm]), np.stack([host.mat_inv(m).reshape(16) for m in dm]),
           [[0.0, 0.6, 0.0, 360.0], [0.1, 0.5, 0.2, 360.0], [-0.05, 0.45, 0.0, 250.0]])
    o, c = make_pair(P, idx, sph, None, variant=variant, disks=dsk)
    for tmin, tmax in ((0.0, np.inf), (0.5, 2.8)):
        ro, rd = random_rays(60000, seed=10, tmin=tmin, tmax=tmax)
        hg, ho = c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8)
        assert (hg["prim"] == ho["prim"]).all()
        assert (hg["t"].view(np.uint32) == ho["t"].view(np.uint32)).all()
        assert (ho["prim"] >= 602).sum() > 1500  # the disks are hit
        np.testing.assert_allclose(hg["b1"], ho["b1"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(hg["b2"], ho["b2"], rtol=1e-6, atol=1e-7)
        assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()
    for k in ("offset", "n_primitives", "axis", "ordered", "bounds"):
        assert np.array_equal(c.bvh_export()[k], o.bvh_export()[k]), k


def test_cylinders_cones_paraboloids_hyperboloids(drt_lib, variant):
    """cylinder.dart / cone.dart / paraboloid.dart / hyperboloid.dart: kinds 2..5 of drt_set_quadrics, ids after the disks."""
    from tests.test_oracle_trace import quadric_zoo
    P, idx = random_soup(500, seed=31)
    mats = [translate(0.3, 0.6, -0.2)]
    sph = (np.stack([m[0] for m in mats]), np.stack([m[1] for m in mats]), [[0.2, -0.2, 0.2, 360.0]])
    dm = [host.mat_mul(host.translate(0.6, -0.6, 0.1), host.rotate(80, (0, 1, 0)))]
    dsk = (np.stack([m.reshape(16) for m in dm]), np.stack([host.mat_inv(m).reshape(16) for m in dm]), [[0.0, 0.3, 0.0, 360.0]])
    o, c = Oracle(), capi.Context(0)
    c.set_kernel_variant(variant)
    for x in (o, c):
        x.set_triangles(P, idx)
        x.set_spheres(*sph)
        x.set_disks(*dsk)
        for kind, prm, m in quadric_zoo(host):
            x.set_quadrics(kind, m.reshape(16), host.mat_inv(m).reshape(16), [prm])
        x.build_bvh(2, 4)
    for tmin, tmax in ((0.0, np.inf), (0.5, 2.9)):
        ro, rd = random_rays(80000, seed=32, tmin=tmin, tmax=tmax)
        hg, ho = c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8)
        assert (hg["prim"] == ho["prim"]).all()
        assert (hg["t"].view(np.uint32) == ho["t"].view(np.uint32)).all()
        for k in range(4):
            assert (ho["prim"] == 502 + k).sum() > 300, k  # every kind is hit
        np.testing.assert_allclose(hg["b1"], ho["b1"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(hg["b2"], ho["b2"], rtol=1e-6, atol=1e-7)
        assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()
    for k in ("offset", "n_primitives", "axis", "ordered", "bounds"):
        assert np.array_equal(c.bvh_export()[k], o.bvh_export()[k]), k
    # the reference's NaN hit for an axial ray inside a cylinder (tests/test_oracle_trace.py::test_cylinder_known_answers)
    o2, c2 = Oracle(), capi.Context(0)
    c2.set_kernel_variant(variant)
    eye = np.eye(4, dtype=np.float32).reshape(16)
    for x in (o2, c2):
        x.set_triangles(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
        x.set_spheres(np.zeros((0, 16), np.float32), np.zeros((0, 16), np.float32), np.zeros((0, 4)))
        x.set_quadrics(2, eye, eye, [[0.5, -1.0, 2.0, 360.0, 0, 0, 0, 0]])
        x.build_bvh(2, 4)
    ro, rd = scenes.pack_rays(np.array([[0.1, 0.1, -5], [0.6, 0.1, -5], [-3, 0, 0.5]], np.float32),
                              np.array([[0, 0, 1], [0, 0, 1], [1, 0, 0]], np.float32), 0.0, np.inf)
    hg, ho = c2.trace_closest(ro, rd), o2.trace_closest(ro, rd)
    assert (hg["prim"] == ho["prim"]).all() and list(ho["prim"]) == [0, -1, 0]
    assert np.isnan(hg["t"][0]) and np.isnan(ho["t"][0]) and hg["t"][2] == ho["t"][2] == 2.5
    assert (c2.trace_any(ro, rd) == o2.trace_any(ro, rd)).all()


def test_known_answer_edge_cases(drt_lib, variant):
    """Same quirks the oracle test pins: inclusive triangle edges, strict flat-box culling and the
    NaN behaviour of the slab test for axis-parallel rays (bvh_accel.dart:441-471)."""
    P = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 0, 2], [1, 0, 2], [1, 1, 3]], np.float32)
    idx = np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
    o, c = make_pair(P, idx, variant=variant)
    cases = [((0.75, 0.5, 1.0), (0, 0, -1), 0, np.inf), ((0.5, 0.5, 1.0), (0, 0, -1), 0, np.inf),
             ((1.0, 0.5, 1.0), (0, 0, -1), 0, np.inf), ((0.5, 0.0, 1.0), (0, 0, -1), 0, np.inf),
             ((0.2, 0.1, 1.0), (1, 0, 0), 0, np.inf), ((0.5, 0.2, 1.0), (0, 0, -1), 1.0, 5.0),
             ((0.5, 0.2, 1.0), (0, 0, -1), 0.0, 1.0), ((0.5, 0.25, 5.0), (0, 0, -1), 0.0, 2.75),
             ((0.5, 0.25, 5.0), (0, 0, -1), 2.75, 9.0), ((0.5, 0.25, 5.0), (0, 0, -1), 0.0, np.inf),
             ((0.5, 0.2, -1.0), (0, 0, 1), 0.0, np.inf), ((0.5, 0.2, -1.0), (0, 0, 1), 1.5, np.inf)]
    ro = np.array([[*c_[0], c_[2]] for c_ in cases], np.float32)
    rd = np.array([[*c_[1], c_[3]] for c_ in cases], np.float32)
    assert_hits_equal(c.trace_closest(ro, rd), o.trace_closest(ro, rd))
    assert (c.trace_any(ro, rd) == o.trace_any(ro, rd)).all()


def test_big_leaf_and_equal_t_ties(drt_lib, variant):
    """Coincident triangles: one 40-primitive leaf; the LAST tested equal-t hit wins (triangle.dart:96)."""
    tri = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0]], np.float32)
    P = np.tile(tri, (40, 1))
    idx = np.arange(120, dtype=np.uint32).reshape(-1, 3)
    o, c = make_pair(P, idx, variant=variant)
    ro, rd = scenes.pack_rays(np.array([[0.6, 0.3, 1.0]], np.float32), np.array([[0.01, 0.02, -1.0]], np.float32))
    hg, ho = c.trace_closest(ro, rd), o.trace_closest(ro, rd)
    assert ho["prim"][0] == 39
    assert_hits_equal(hg, ho)


def test_empty_scene_and_zero_rays(drt_lib):
    c = capi.Context(0)
    c.set_triangles(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    c.build_bvh()
    ro, rd = random_rays(100, 1)
    h = c.trace_closest(ro, rd)
    assert (h["prim"] == -1).all() and np.isinf(h["t"]).all()
    assert (c.trace_any(ro, rd) == 0).all()
    P, idx = random_soup(10, 0)
    c.set_triangles(P, idx)
    with pytest.raises(capi.DrtError):
        c.trace_closest(ro, rd)  # scene changed, BVH not rebuilt
    c.build_bvh()
    assert c.trace_closest(ro[:0], rd[:0]).shape[0] == 0


def test_counters_match_reference_work(drt_lib):
    """nodes_visited / prims_tested equal the reference traversal's slab and primitive test counts."""
    P, idx = scenes.soup(16)
    o, c = make_pair(P, idx, order=mesh_refine_order([idx.shape[0]]))
    ro, rd = scenes.incoherent_rays(20000)
    c.set_counting(True)
    hg = c.trace_closest(ro, rd)
    cg = c.counters()
    ho = o.trace_closest(ro, rd, nthreads=4)
    co = o.counters()
    assert_hits_equal(hg, ho)
    assert cg["rays"] == co["rays"] == 20000
    assert cg["nodes_visited"] == co["nodes_visited"]
    assert cg["prims_tested"] == co["prims_tested"]
    occ = c.trace_any(ro, rd)
    cg = c.counters()
    assert (occ == o.trace_any(ro, rd, nthreads=4)).all()
    co = o.counters()
    assert cg["nodes_visited"] == co["nodes_visited"] and cg["prims_tested"] == co["prims_tested"]
    c.set_counting(False)


def test_small_scene_with_every_shape_kind_runs_the_leaf_list_kernel(drt_lib):
    """<= 32 leaves: the default variant is traceSmallKernel (trace_fast.cu).  24 triangles, a full and a clipped sphere, two disks
    (one an annulus with a phi cut) and the four other quadrics; random, bounded, axis-parallel, far-away and on-surface rays; every
    split method (the visiting-order tables follow the tree).  Bit-exact against the oracle like the tree kernels."""
    from tests.test_oracle_trace import quadric_zoo
    P, idx = random_soup(24, seed=77)
    mats = [translate(0.3, 0.1, -0.2), translate(-0.4, 0.2, 0.5)]
    sph = (np.stack([m[0] for m in mats]), np.stack([m[1] for m in mats]), [[0.25, -0.25, 0.25, 360.0], [0.4, -0.1, 0.3, 200.0]])
    dm = [host.mat_mul(host.translate(0.1, -0.3, 0.2), host.rotate(35.0, (1.0, 0.3, 0.2))), host.translate(-0.5, 0.5, -0.4)]
    dsk = (np.stack([m.reshape(16) for m in dm]), np.stack([host.mat_inv(m).reshape(16) for m in dm]),
           [[0.0, 0.6, 0.0, 360.0], [0.1, 0.5, 0.2, 250.0]])
    rng = np.random.default_rng(78)
    for split in (0, 1, 2):
        for with_quadrics in (False, True):
            o, c = Oracle(), capi.Context(0)
            for x in (o, c):
                x.set_triangles(P, idx)
                x.set_spheres(*sph)
                x.set_disks(*dsk)
                if with_quadrics:
                    for kind, prm, m in quadric_zoo(host):
                        x.set_quadrics(kind, m.reshape(16), host.mat_inv(m).reshape(16), [prm])
                x.build_bvh(split, 4)
            assert c.bvh_info()["n_leaves"] <= 32
            sets = [random_rays(60000, seed=79 + split), random_rays(60000, seed=90 + split, tmin=0.4, tmax=1.9)]
            ro, rd = random_rays(6000, seed=95)
            rd = rd.copy()
            rd[:2000, 0] = 0.0   # zero direction components: inf / NaN in the slab test, "slow" rays
            rd[2000:4000, :2] = 0.0
            rd[4000:5000, 1] = -0.0
            ro = ro.copy()
            ro[5000:5500, :3] *= 1.0e6      # far away
            ro[5500:5750, :3] = 3.0e38      # at the edge of the float32 filter's preconditions: products overflow, the filter says "undecided"
            ro[5750:6000, :3] = 3.2e38      # beyond them: "slow" rays, every box in binary64
            sets.append((ro, rd))
            n = 20000  # rays that start on a triangle of the soup (secondary-ray style, tmin = 1e-3)
            tri = P[idx[rng.integers(0, idx.shape[0], n)]]
            b = rng.uniform(0, 1, (n, 2)).astype(np.float32)
            flip = b.sum(axis=1) > 1
            b[flip] = 1 - b[flip]
            org = (tri[:, 0] * (1 - b[:, :1] - b[:, 1:]) + tri[:, 1] * b[:, :1] + tri[:, 2] * b[:, 1:]).astype(np.float32)
            d2 = rng.normal(size=(n, 3))
            d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
            sets.append(scenes.pack_rays(org, d2, 1e-3, np.inf))
            for ro, rd in sets:
                hg, ho = c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8)
                assert (hg["prim"] == ho["prim"]).all(), (split, with_quadrics, int((hg["prim"] != ho["prim"]).sum()))
                same_t = (hg["t"].view(np.uint32) == ho["t"].view(np.uint32)) | (np.isnan(hg["t"]) & np.isnan(ho["t"]))
                assert same_t.all()
                np.testing.assert_allclose(hg["b1"], ho["b1"], rtol=1e-6, atol=1e-7)
                np.testing.assert_allclose(hg["b2"], ho["b2"], rtol=1e-6, atol=1e-7)
                assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()
            assert (ho["prim"] >= 0).sum() > 1000
            # the same rays through the forced tree kernel: identical records
            c.set_kernel_variant(2)
            ht = c.trace_closest(ro, rd)
            assert hg.tobytes() == ht.tobytes() or (np.isnan(hg["t"]).any() and (hg["prim"] == ht["prim"]).all())


def test_axis_aligned_walls_and_rays(drt_lib, variant):
    """Cornell-style flat quads: leaf boxes are flat, so after the first triangle of a wall is hit the
    sibling's box entry distance equals ray.maxDistance up to rounding — the float32 filter cannot
    decide and the exact float64 path must give the reference's answer.  Also axis-parallel rays
    (zero direction components: inf / NaN in the slab test)."""
    q = lambda a, b, c_, d: [a, b, c_, d]
    quads = [q((-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)),      # back wall z=1
             q((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1)),  # floor y=-1
             q((-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1)),      # ceiling
             q((-1, -1, -1), (-1, 1, -1), (-1, 1, 1), (-1, -1, 1)),  # left
             q((1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1))]      # right
    P = np.array([v for qd in quads for v in qd], np.float32)
    idx = np.array([[4 * k, 4 * k + 1, 4 * k + 2] for k in range(5)] + [[4 * k, 4 * k + 2, 4 * k + 3] for k in range(5)],
                   np.uint32)
    o, c = make_pair(P, idx, variant=variant)
    rng = np.random.default_rng(3)
    n = 60000
    org = np.zeros((n, 3), np.float32)
    org[:, 2] = -3.0
    d = np.stack([rng.uniform(-0.4, 0.4, n), rng.uniform(-0.4, 0.4, n), np.ones(n)], axis=1)
    d[:2000, 0] = 0.0          # axis-parallel in x
    d[2000:4000, :2] = 0.0     # straight down the z axis
    d[4000:6000, 1] = d[4000:6000, 0]  # exactly on the quad diagonals x == y
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[2000:4000] = (0, 0, 1)
    ro, rd = scenes.pack_rays(org, d)
    assert_hits_equal(c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8))
    assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()
    # rays starting ON a wall (secondary-ray style, tmin = 1e-3)
    org2 = np.stack([rng.uniform(-1, 1, n), np.full(n, -1.0), rng.uniform(-1, 1, n)], axis=1).astype(np.float32)
    d2 = rng.normal(size=(n, 3))
    d2[:, 1] = np.abs(d2[:, 1])
    d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
    ro, rd = scenes.pack_rays(org2, d2, 1e-3, np.inf)
    assert_hits_equal(c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8))
    assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()


def test_soup_scene_coherent_and_incoherent(drt_lib, variant):
    """A 127k-triangle slice of the config-2 workload, both ray sets, bit-exact."""
    P, idx = scenes.soup(64)
    o, c = make_pair(P, idx, variant=variant)
    for ro, rd in (scenes.coherent_rays(512, 256), scenes.incoherent_rays(1 << 17)):
        assert_hits_equal(c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8))
        assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()


def test_concurrent_device_launches_on_two_streams_do_not_share_a_work_counter(drt_lib):
    """drt_trace_*_device is asynchronous on the caller's stream: two launches in flight on different streams of one
    context must each trace every ray (they used to share the persistent kernel's work counter)."""
    import torch
    P, idx = scenes.soup(8)
    c = capi.Context(0)
    c.set_triangles(P, idx)
    c.build_bvh()
    n = 1 << 18
    ro, rd = scenes.incoherent_rays(n)
    ro2, rd2 = scenes.coherent_rays(512, 512)
    href, href2 = c.trace_closest(ro, rd), c.trace_closest(ro2, rd2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    dro, drd, dro2, drd2 = (torch.from_numpy(a).cuda() for a in (ro, rd, ro2, rd2))
    for rep in range(8):
        dh = torch.full((n, 4), -7.0, dtype=torch.float32, device="cuda")
        dh2 = torch.full((n, 4), -7.0, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        c.trace_closest_device(dro.data_ptr(), drd.data_ptr(), n, dh.data_ptr(), s1.cuda_stream)
        c.trace_closest_device(dro2.data_ptr(), drd2.data_ptr(), n, dh2.data_ptr(), s2.cuda_stream)
        torch.cuda.synchronize()
        for d, h in ((dh, href), (dh2, href2)):
            hd = d.cpu().numpy().view(capi.HIT_DTYPE).reshape(-1)
            assert (hd["prim"] == h["prim"]).all() and (hd["t"].view(np.uint32) == h["t"].view(np.uint32)).all(), rep


def test_default_variant_picks_the_kernel_by_scene_size(drt_lib):
    for ns, expect_q in ((8, False), (64, True)):
        P, idx = scenes.soup(ns)
        c = capi.Context(0)
        c.set_triangles(P, idx)
        c.build_bvh()
        per_prim = c.bvh_info()["device_bytes"] / idx.shape[0]
        # 64-byte nodes (~32 B per primitive) + 48-byte records, against 128-byte nodes (~64 B per primitive)
        assert (per_prim < 96) == expect_q, (ns, per_prim)


def test_device_pointer_entry_matches_host_entry(drt_lib):
    import torch
    P, idx = scenes.soup(8)
    c = capi.Context(0)
    c.set_triangles(P, idx)
    c.build_bvh()
    ro, rd = scenes.incoherent_rays(1 << 15)
    href = c.trace_closest(ro, rd)
    dro, drd = torch.from_numpy(ro).cuda(), torch.from_numpy(rd).cuda()
    dh = torch.empty((ro.shape[0], 4), dtype=torch.float32, device="cuda")
    docc = torch.empty(ro.shape[0], dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    c.trace_closest_device(dro.data_ptr(), drd.data_ptr(), ro.shape[0], dh.data_ptr(), st)
    c.trace_any_device(dro.data_ptr(), drd.data_ptr(), ro.shape[0], docc.data_ptr(), st)
    torch.cuda.synchronize()
    hd = dh.cpu().numpy().view(capi.HIT_DTYPE).reshape(-1)
    assert (hd["prim"] == href["prim"]).all() and (hd["t"].view(np.uint32) == href["t"].view(np.uint32)).all()
    assert (docc.cpu().numpy() == c.trace_any(ro, rd)).all()


# ---- the quantised-node kernel at the edges of its float32 preconditions (trace_fast2.cu header) -------------------------------
@pytest.mark.parametrize("scale,offset", [(1.0e-20, 0.0), (1.0e15, 0.0), (1.0, 3.0e6), (1.0e-3, -7.5e4), (3.0e17, 1.0e18)],
                         ids=["tiny", "huge", "far_small", "far_tiny", "near_2^62"])
def test_quantised_nodes_at_extreme_coordinates(drt_lib, variant, scale, offset):
    """Grid steps clamp at 2^-60, coordinates approach the 2^62 bound, boxes shrink to a few ulps of their position: the node
    boxes stay conservative (the builder checks every inequality in binary64) and every leaf decision is the reference's."""
    P, idx = random_soup(3000, seed=11)
    P = (P.astype(np.float64) * scale + offset).astype(np.float32)
    o, c = make_pair(P, idx, variant=variant)
    ro, rd = random_rays(30000, seed=12)
    ro[:, :3] = (ro[:, :3].astype(np.float64) * scale + offset).astype(np.float32)
    assert_hits_equal(c.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8))
    assert (c.trace_any(ro, rd) == o.trace_any(ro, rd, nthreads=8)).all()
    hit = c.trace_closest(ro, rd)["prim"] >= 0
    assert scale * 1e3 < abs(offset) or hit.mean() > 0.05  # far-away flat soups may be missed by most rays; the others are hit


def test_quantised_nodes_with_unnormalised_and_degenerate_directions(drt_lib, variant):
    """Directions scaled by 1e-25 / 1e25 (|invDir| leaves [2^-60, 2^40]: the binary64 path on the decoded boxes), axis-parallel rays
    (infinite invDir), rays from beyond 2^62 and rays starting exactly on box planes."""
    P, idx = random_soup(2000, seed=21)
    o, c = make_pair(P, idx, variant=variant)
    ro, rd = random_rays(20000, seed=22)
    rd2 = rd.copy()
    rd2[0::4, :3] *= np.float32(1e-25)
    rd2[1::4, :3] *= np.float32(1e25)
    rd2[2::8, 0] = 0.0                      # parallel to the yz plane
    rd2[6::8, 1:3] = 0.0                    # along x
    rd2[:, 3] = np.inf
    ro2 = ro.copy()
    ro2[3::16, :3] *= np.float32(1e19)      # beyond the 2^62 origin bound
    # origins exactly on triangle vertices' coordinates: on leaf / node box planes
    k = np.arange(5, ro2.shape[0], 16)
    ro2[k, 0] = P[idx[k % idx.shape[0], 0], 0]
    for a, b in ((ro2, rd2), (ro, rd2)):
        assert_hits_equal(c.trace_closest(a, b), o.trace_closest(a, b, nthreads=8))
        assert (c.trace_any(a, b) == o.trace_any(a, b, nthreads=8)).all()
