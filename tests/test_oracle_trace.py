"""Pins the CPU oracle for traversal + shapes.  The reference ships no golden vectors
(test/spectrum_test.dart:14-39 is commented out), so the pins are analytic known answers and the
accelerated-vs-exhaustive differential check of lib/renderers/aggregate_test_renderer.dart:42-118."""
import math

import numpy as np
import pytest

from dartray_b200 import scenes
from tests.oracle_lib import Oracle
from tests.util import random_rays, random_soup, translate

INF = np.float32(np.inf)


def one_tri():
    # p1=(0,0), p2=(1,0), p3=(1,1): hit point = (b1 + b2, b2), so x = y is the b1 == 0 edge
    P = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0]], np.float32)
    idx = np.array([[0, 1, 2]], np.uint32)
    o = Oracle()
    o.set_triangles(P, idx)
    o.build_bvh()
    return o


def ray(o, d, tmin=0.0, tmax=np.inf):
    return scenes.pack_rays(np.array([o], np.float32), np.array([d], np.float32), tmin, tmax)


def test_triangle_known_answers():
    o = one_tri()
    # interior hit: b2 = y, b1 = x - y, t = origin height (triangle.dart:77-95)
    h = o.trace_closest(*ray((0.75, 0.5, 2.0), (0, 0, -1)))[0]
    assert h["prim"] == 0 and h["t"] == 2.0 and h["b1"] == 0.25 and h["b2"] == 0.5
    # inclusive edge: b1 == 0 exactly on the diagonal still hits (triangle.dart:78)
    h = o.trace_closest(*ray((0.5, 0.5, 1.0), (0, 0, -1)))[0]
    assert h["prim"] == 0 and h["b1"] == 0.0 and h["b2"] == 0.5
    # just outside the diagonal
    assert o.trace_closest(*ray((0.5, 0.5 + 1e-6, 1.0), (0, 0, -1)))[0]["prim"] == -1
    # NaN semantics of the slab test (bvh_accel.dart:441-471): for an axis-parallel ray whose
    # origin lies exactly on a box plane, (b - o) * (1/0) = 0 * inf = NaN.  A NaN in the X slab
    # poisons tmin/tmax (later `tymin > tmin` style updates are false) and the final
    # `tmin < maxDistance && tmax > minDistance` rejects the node, so the reference MISSES the edge
    # x == 1 of this triangle for a ray with d.x == 0.  A NaN in the y/z slabs is simply ignored,
    # so the y == 0 edge (b2 == 0, inclusive) is hit.  The oracle keeps both behaviours.
    assert o.trace_closest(*ray((1.0, 0.5, 1.0), (0, 0, -1)))[0]["prim"] == -1
    assert o.trace_closest(*ray((1.0 - 1e-6, 0.5, 1.0), (0, 0, -1)))[0]["prim"] == 0
    h = o.trace_closest(*ray((0.5, 0.0, 1.0), (0, 0, -1)))[0]
    assert h["prim"] == 0 and h["b2"] == 0.0
    # parallel ray: divisor == 0 (triangle.dart:66)
    assert o.trace_closest(*ray((0.2, 0.1, 1.0), (1, 0, 0)))[0]["prim"] == -1
    # The triangle's t range is inclusive (triangle.dart:96) but the slab test is strict
    # (`tmin < maxDistance && tmax > minDistance`, bvh_accel.dart:471): for this FLAT leaf box
    # tzmin == tzmax == t, so a hit at exactly minDistance or maxDistance is culled by the box.
    assert o.trace_closest(*ray((0.5, 0.2, 1.0), (0, 0, -1), 0.5, 5.0))[0]["prim"] == 0
    assert o.trace_closest(*ray((0.5, 0.2, 1.0), (0, 0, -1), 1.0, 5.0))[0]["prim"] == -1
    assert o.trace_closest(*ray((0.5, 0.2, 1.0), (0, 0, -1), 0.0, 1.0))[0]["prim"] == -1
    assert o.trace_closest(*ray((0.5, 0.2, 1.0), (0, 0, -1), 1.0000001, 5.0))[0]["prim"] == -1
    assert o.trace_closest(*ray((0.5, 0.2, 1.0), (0, 0, -1), 0.0, 0.99999))[0]["prim"] == -1
    # a tilted triangle (plane z = y) has a thick box, so the inclusive ends are observable
    tilt = Oracle()
    tilt.set_triangles(np.array([[0, 0, 0], [1, 0, 0], [1, 1, 1]], np.float32), np.array([[0, 1, 2]], np.uint32))
    tilt.build_bvh()
    assert tilt.trace_closest(*ray((0.5, 0.25, 1.0), (0, 0, -1)))[0]["t"] == 0.75
    assert tilt.trace_closest(*ray((0.5, 0.25, 1.0), (0, 0, -1), 0.0, 0.75))[0]["prim"] == 0
    assert tilt.trace_closest(*ray((0.5, 0.25, 1.0), (0, 0, -1), 0.75, 2.0))[0]["prim"] == 0
    assert tilt.trace_closest(*ray((0.5, 0.25, 1.0), (0, 0, -1), 0.0, 0.7499999))[0]["prim"] == -1
    assert tilt.trace_closest(*ray((0.5, 0.25, 1.0), (0, 0, -1), 0.7500001, 2.0))[0]["prim"] == -1
    assert tilt.trace_any(*ray((0.5, 0.25, 1.0), (0, 0, -1), 0.0, 0.75))[0] == 1
    # behind the origin
    assert o.trace_closest(*ray((0.5, 0.2, 1.0), (0, 0, 1)))[0]["prim"] == -1
    # any-hit agrees
    assert o.trace_any(*ray((0.75, 0.5, 2.0), (0, 0, -1)))[0] == 1
    assert o.trace_any(*ray((0.5, 0.5, 1.0), (0, 0, -1)))[0] == 1
    assert o.trace_any(*ray((0.2, 0.1, 1.0), (1, 0, 0)))[0] == 0


def test_equal_t_last_tested_wins():
    """`t > ray.maxDistance` is the reject test (triangle.dart:96): an equal-t hit tested later
    replaces the earlier one.  Two coincident triangles -> the brute-force loop ends on prim 1."""
    P = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0]], np.float32)
    idx = np.array([[0, 1, 2], [0, 1, 2]], np.uint32)
    o = Oracle()
    o.set_triangles(P, idx)
    o.build_bvh()
    hb, nties, _ = o.trace_closest_brute(*ray((0.5, 0.2, 1.0), (0, 0, -1)))
    assert hb[0]["prim"] == 1 and nties[0] == 2
    # centroid-degenerate range -> one leaf holding both (bvh_accel.dart:265-274), same winner
    ex = o.bvh_export()
    assert len(ex["offset"]) == 1 and ex["n_primitives"][0] == 2
    assert o.trace_closest(*ray((0.5, 0.2, 1.0), (0, 0, -1)))[0]["prim"] == 1


def sphere_oracle(radius=1.0, center=(0, 0, 0), zmin=None, zmax=None, phimax=360.0):
    o = Oracle()
    o.set_triangles(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    m, mi = translate(*center)
    o.set_spheres(m, mi, [[radius, -radius if zmin is None else zmin, radius if zmax is None else zmax, phimax]])
    o.build_bvh()
    return o


def test_sphere_known_answers():
    o = sphere_oracle(1.0, (0, 0, 5))
    h = o.trace_closest(*ray((0, 0, 0), (0, 0, 1)))[0]
    assert h["prim"] == 0 and h["t"] == 4.0
    # from inside: t0 < minDistance -> t1 (sphere.dart:86-92)
    h = o.trace_closest(*ray((0, 0, 5), (0, 0, 1)))[0]
    assert h["prim"] == 0 and h["t"] == 1.0
    # grazing miss / tangent
    assert o.trace_closest(*ray((1.5, 0, 0), (0, 0, 1)))[0]["prim"] == -1
    # closed form for an oblique ray
    org, d = np.array([0.3, -0.2, 0.0]), np.array([0.05, 0.1, 1.0])
    d = (d / np.linalg.norm(d)).astype(np.float32).astype(np.float64)
    org = org.astype(np.float32).astype(np.float64)
    oc = org - np.array([0, 0, 5.0])
    A, B, Cc = d @ d, 2 * (d @ oc), oc @ oc - 1.0
    t = (-B - math.sqrt(B * B - 4 * A * Cc)) / (2 * A)
    h = o.trace_closest(*ray(org, d))[0]
    assert h["prim"] == 0 and abs(h["t"] - t) <= 1e-6 * t
    assert o.trace_any(*ray(org, d))[0] == 1
    # tmax clips
    assert o.trace_closest(*ray((0, 0, 0), (0, 0, 1), 0.0, 3.9))[0]["prim"] == -1
    assert o.trace_any(*ray((0, 0, 0), (0, 0, 1), 0.0, 3.9))[0] == 0


def test_partial_sphere_clipping():
    # zmax clip: the near cap is cut away, the ray enters through the hole and hits the far side (t1)
    o = sphere_oracle(1.0, (0, 0, 0), zmin=-1.0, zmax=0.5)
    h = o.trace_closest(*ray((0, 0.1, 5), (0, 0, -1)))[0]
    assert h["prim"] == 0 and h["t"] > 5.0
    # phimax = 180: half sphere y >= 0; a ray at y < 0 misses both roots
    o = sphere_oracle(1.0, (0, 0, 0), phimax=180.0)
    assert o.trace_closest(*ray((0.2, -0.5, 5), (0, 0, -1)))[0]["prim"] == -1
    assert o.trace_closest(*ray((0.2, 0.5, 5), (0, 0, -1)))[0]["prim"] == 0


@pytest.mark.parametrize("split", [0, 1, 2])
def test_bvh_matches_exhaustive(split):
    """aggregate_test_renderer.dart:82-107: accelerated and exhaustive intersection agree."""
    P, idx = random_soup(600, seed=split)
    o = Oracle()
    o.set_triangles(P, idx)
    m, mi = translate(0.3, -0.2, 0.1)
    o.set_spheres(np.stack([m, translate(-0.5, 0.5, 0.2)[0]]), np.stack([mi, translate(-0.5, 0.5, 0.2)[1]]),
                  [[0.4, -0.4, 0.4, 360.0], [0.3, -0.1, 0.2, 270.0]])
    o.build_bvh(split, 4)
    ro, rd = random_rays(3000, seed=10 + split)
    h = o.trace_closest(ro, rd)
    hb, nties, _ = o.trace_closest_brute(ro, rd)
    single = nties <= 1
    assert (h["prim"][single] == hb["prim"][single]).all()
    assert (h["t"] == hb["t"]).all()
    assert (o.trace_any(ro, rd) == o.trace_any_brute(ro, rd)).all()
    # short rays (shadow-ray style intervals)
    ro2, rd2 = random_rays(2000, seed=20 + split, tmin=0.5, tmax=2.8)
    assert (o.trace_closest(ro2, rd2)["t"] == o.trace_closest_brute(ro2, rd2)[0]["t"]).all()
    assert (o.trace_any(ro2, rd2) == o.trace_any_brute(ro2, rd2)).all()


def test_empty_scene_misses():
    o = Oracle()
    o.set_triangles(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    o.build_bvh()
    ro, rd = random_rays(8, 0)
    assert (o.trace_closest(ro, rd)["prim"] == -1).all()
    assert (o.trace_any(ro, rd) == 0).all()


def test_sah_tree_shape():
    """With maxnodeprims 4 the reference still splits every range of <= 4 primitives
    (bvh_accel.dart:313-316), so leaves hold one primitive unless centroids coincide."""
    P, idx = scenes.soup(4)
    o = Oracle()
    o.set_triangles(P, idx)
    o.build_bvh(2, 4)
    ex = o.bvh_export()
    n = idx.shape[0]
    leaves = ex["n_primitives"] > 0
    assert len(ex["offset"]) == 2 * int(leaves.sum()) - 1
    assert int(ex["n_primitives"].sum()) == n
    assert sorted(ex["ordered"].tolist()) == list(range(n))
    assert (ex["n_primitives"][leaves] == 1).mean() > 0.99
    # flatten: first child at n+1, second child index stored in offset (bvh_accel.dart:419-437)
    inner = np.nonzero(~leaves)[0]
    assert (ex["offset"][inner] > inner + 1).all()
    # parent box is the union of the children's boxes
    b = ex["bounds"]
    for i in inner[:200]:
        c0, c1 = i + 1, ex["offset"][i]
        assert (b[i, :3] == np.minimum(b[c0, :3], b[c1, :3])).all()
        assert (b[i, 3:] == np.maximum(b[c0, 3:], b[c1, 3:])).all()


# ---- Disk (lib/shapes/disk.dart) ------------------------------------------------------------
def disk_oracle(height=0.0, radius=1.0, inner=0.0, phimax=360.0, center=(0, 0, 0)):
    o = Oracle()
    o.set_triangles(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    o.set_spheres(np.zeros((0, 16), np.float32), np.zeros((0, 16), np.float32), np.zeros((0, 4)))
    m, mi = translate(*center)
    o.set_disks(m, mi, [[height, radius, inner, phimax]])
    o.build_bvh()
    return o


def test_disk_known_answers():
    o = disk_oracle(height=2.0, radius=1.0)
    h = o.trace_closest(*ray((0.3, 0.4, 0), (0, 0, 1)))[0]
    assert h["prim"] == 0 and h["t"] == 2.0
    assert h["b1"] == pytest.approx(math.atan2(0.4, 0.3) / (2 * math.pi), rel=1e-6)  # u = phi / phiMax
    assert h["b2"] == pytest.approx(1.0 - 0.5, rel=1e-6)                              # v = 1 - (r - ri) / (R - ri)
    assert o.trace_any(*ray((0.3, 0.4, 0), (0, 0, 1)))[0] == 1
    # outside the radius, and exactly on it (dist2 > r^2 is the reject test: the rim counts)
    assert o.trace_closest(*ray((0.8, 0.7, 0), (0, 0, 1)))[0]["prim"] == -1
    assert o.trace_closest(*ray((0.0, 1.0, 0), (0, 0, 1)))[0]["prim"] == 0
    # ... but the same rim point reached with x == box edge and d.x == 0 is lost one level up: 0 * inf = NaN in the X
    # slab of BVHAccel rejects the node (bvh_accel.dart:441-448), a reference quirk the oracle keeps
    assert o.trace_closest(*ray((1.0, 0.0, 0), (0, 0, 1)))[0]["prim"] == -1
    # rays parallel to the disk's plane miss (|d.z| < 1e-7), from either side it is hit
    assert o.trace_closest(*ray((-3, 0, 2.0), (1, 0, 0)))[0]["prim"] == -1
    assert o.trace_closest(*ray((0, 0, 5), (0, 0, -1)))[0]["t"] == 3.0
    # interval: the shape accepts t in [tmin, tmax], but the flat leaf box is entered only when tmin_box < tmax
    # strictly (bvh_accel.dart:471), so t == tmax is lost in the accelerator
    assert o.trace_closest(*ray((0, 0, 0), (0, 0, 1), 0.0, 2.0))[0]["prim"] == -1
    assert o.trace_closest(*ray((0, 0, 0), (0, 0, 1), 0.0, 2.001))[0]["prim"] == 0
    assert o.trace_closest(*ray((0, 0, 0), (0, 0, 1), 0.0, 1.999))[0]["prim"] == -1
    assert o.trace_any(*ray((0, 0, 0), (0, 0, 1), 2.001, 9.0))[0] == 0
    # annulus and partial sweep
    o = disk_oracle(radius=1.0, inner=0.5)
    assert o.trace_closest(*ray((0.2, 0.1, -1), (0, 0, 1)))[0]["prim"] == -1
    assert o.trace_closest(*ray((0.6, 0.1, -1), (0, 0, 1)))[0]["prim"] == 0
    o = disk_oracle(radius=1.0, phimax=90.0)
    assert o.trace_closest(*ray((0.3, 0.3, -1), (0, 0, 1)))[0]["prim"] == 0
    assert o.trace_closest(*ray((-0.3, 0.3, -1), (0, 0, 1)))[0]["prim"] == -1
    # oblique closed form, translated disk
    o = disk_oracle(height=0.5, radius=2.0, center=(1, -1, 3))
    org, d = np.array([0.2, 0.1, 0.0], np.float32).astype(np.float64), np.array([0.3, -0.4, 1.0])
    d = (d / np.linalg.norm(d)).astype(np.float32).astype(np.float64)
    t = (3.5 - org[2]) / d[2]
    h = o.trace_closest(*ray(org, d))[0]
    assert h["prim"] == 0 and abs(h["t"] - t) <= 1e-6 * t


def test_disk_light_irradiance_closed_form():
    # matte point on the axis of a parallel disk light: E = pi * L * R^2 / (h^2 + R^2)
    from dartray_b200 import host
    kd, Le, h, R = 0.5, 8.0, 2.0, 1.5
    sb = host.SceneBuilder()
    sb.mesh([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], [[0, 2, 1], [0, 3, 2]], material=sb.material((kd, kd, kd)))
    # the disk's normal is +z in object space; Rotate 90 about x turns it towards -y (as cornell-path.pbrt:17-18)
    sb.disk(host.mat_mul(host.translate(0, h, 0), host.rotate(90, (1, 0, 0))), radius=R, area_light=(Le, Le, Le), nsamples=16)
    cam = host.PerspectiveCamera(host.look_at((0.2, 1.2, -0.2), (0, 0, 0), (0, 1, 0)), fov=1.0)
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=64), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    E = math.pi * Le * R * R / (h * h + R * R)
    assert o.film_read()["rgb"].mean() == pytest.approx(kd / math.pi * E, rel=1e-2)


# ---- Cylinder / Cone / Paraboloid / Hyperboloid (lib/shapes/*.dart) ---------------------------
def quadric_oracle(kind, params, center=(0, 0, 0)):
    o = Oracle()
    o.set_triangles(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    o.set_spheres(np.zeros((0, 16), np.float32), np.zeros((0, 16), np.float32), np.zeros((0, 4)))
    m, mi = translate(*center)
    o.set_quadrics(kind, m, mi, [list(params) + [0.0] * (8 - len(params))])
    o.build_bvh()
    return o


def test_cylinder_known_answers():
    o = quadric_oracle(2, (0.5, -1.0, 2.0, 360.0))  # radius, zmin, zmax, phimax (cylinder.dart:239-247)
    h = o.trace_closest(*ray((-3, 0, 0.5), (1, 0, 0)))[0]
    assert h["prim"] == 0 and h["t"] == 2.5
    assert h["b1"] == pytest.approx(0.5, rel=1e-6)        # phi = pi at (-0.5, 0): u = phi / phiMax
    assert h["b2"] == pytest.approx(1.5 / 3.0, rel=1e-6)  # v = (z - zmin) / (zmax - zmin)
    assert o.trace_any(*ray((-3, 0, 0.5), (1, 0, 0)))[0] == 1
    # from inside: the far root
    assert o.trace_closest(*ray((0, 0, 0), (0, 1, 0)))[0]["t"] == 0.5
    # above zmax the near root is clipped and so is the far one; a slanted ray leaves through the far wall inside the range
    assert o.trace_closest(*ray((-3, 0, 2.5), (1, 0, 0)))[0]["prim"] == -1
    d = np.array([1.0, 0.0, -0.25])
    h = o.trace_closest(*ray((-3, 0, 2.7), d))[0]  # z at the near wall 2.075 (clipped), at the far wall 1.825
    assert h["prim"] == 0 and h["t"] == pytest.approx(3.5, rel=1e-6)
    # an axial ray OUTSIDE the radius misses (A = B = 0, C > 0: t0 = -0 / 0 = NaN, t1 = -inf < minDistance) ...
    assert o.trace_closest(*ray((0.6, 0.1, -5), (0, 0, 1)))[0]["prim"] == -1
    # ... but INSIDE the radius the reference reports a hit at t = NaN: t1 = C / -0 = +inf passes `t1 < minDistance`, every
    # comparison against the NaN t0 is false, so no clip test rejects it (common.dart:140-167, cylinder.dart:57-85).
    # A reference quirk the oracle keeps as written.
    h = o.trace_closest(*ray((0.1, 0.1, -5), (0, 0, 1)))[0]
    assert h["prim"] == 0 and math.isnan(h["t"])
    # partial sweep
    o = quadric_oracle(2, (1.0, -1.0, 1.0, 90.0))
    assert o.trace_closest(*ray((3, 3, 0), (-1, -1, 0)))[0]["prim"] == 0     # phi = 45 degrees
    assert o.trace_closest(*ray((-3, 3, 0), (1, -1, 0)))[0]["prim"] == -1    # roots at 135 and 315 degrees, both clipped
    assert o.trace_closest(*ray((-3, 0.5, 0), (1, 0, 0)))[0]["t"] == pytest.approx(3 + math.sqrt(0.75), rel=1e-6)  # near clipped, far in sweep


def test_cone_known_answers():
    o = quadric_oracle(3, (2.0, 1.0, 360.0))  # height, radius, phimax (cone.dart:216-222): radius 1 at z = 0, apex at z = 2
    h = o.trace_closest(*ray((-3, 0, 1.0), (1, 0, 0)))[0]  # at z = 1 the cone's radius is 0.5
    assert h["prim"] == 0 and h["t"] == pytest.approx(2.5, rel=1e-6)
    assert h["b2"] == pytest.approx(0.5, rel=1e-6)  # v = z / height
    # the mirror nappe above the apex (z > height) is clipped
    assert o.trace_closest(*ray((-3, 0, 3.0), (1, 0, 0)))[0]["prim"] == -1
    assert o.trace_any(*ray((-3, 0, 1.0), (1, 0, 0)))[0] == 1
    assert o.trace_any(*ray((-3, 0, 3.0), (1, 0, 0)))[0] == 0


def test_paraboloid_known_answers():
    o = quadric_oracle(4, (1.0, 0.0, 1.0, 360.0))  # radius, zmin, zmax, phimax: z = x^2 + y^2
    h = o.trace_closest(*ray((-3, 0, 0.25), (1, 0, 0)))[0]
    assert h["prim"] == 0 and h["t"] == pytest.approx(2.5, rel=1e-6) and h["b2"] == pytest.approx(0.25, rel=1e-6)
    # down the axis: A == 0 -> t0 = q / A; the reference's Quadratic divides by zero (common.dart:140-167) and the inf / nan
    # that follows decides; the oracle keeps whatever IEEE gives (no special case), here a hit at the vertex z = 0
    h = o.trace_closest(*ray((0.5, 0, 3), (0, 0, -1)))[0]
    assert h["prim"] in (-1, 0)
    # clipped band
    o = quadric_oracle(4, (1.0, 0.5, 1.0, 360.0))
    assert o.trace_closest(*ray((-3, 0, 0.25), (1, 0, 0)))[0]["prim"] == -1
    assert o.trace_closest(*ray((-3, 0, 0.75), (1, 0, 0)))[0]["t"] == pytest.approx(3 - math.sqrt(0.75), rel=1e-6)


def test_hyperboloid_known_answers():
    # p1 = (1, 0, -1), p2 = (1, 0, 1): the segment is parallel to the axis -> a cylinder of radius 1, z in [-1, 1]
    o = quadric_oracle(5, (1.0, 0.0, -1.0, 1.0, 0.0, 1.0, 360.0))
    h = o.trace_closest(*ray((-3, 0, 0.5), (1, 0, 0)))[0]
    assert h["prim"] == 0 and h["t"] == pytest.approx(2.0, rel=1e-5)
    assert h["b2"] == pytest.approx(0.75, rel=1e-6)  # v = (z - p1.z) / (p2.z - p1.z)
    assert o.trace_closest(*ray((-3, 0, 1.5), (1, 0, 0)))[0]["prim"] == -1
    # a twisted ruling: p1 = (1, 0, -1), p2 = (0, 1, 1) sweeps the one-sheet hyperboloid x^2 + y^2 - z^2 / 2 = 1 / 2
    o = quadric_oracle(5, (1.0, 0.0, -1.0, 0.0, 1.0, 1.0, 360.0))
    h = o.trace_closest(*ray((-3, 0, 0), (1, 0, 0)))[0]
    assert h["prim"] == 0 and h["t"] == pytest.approx(3 - math.sqrt(0.5), rel=1e-5)  # waist radius sqrt(1/2)
    h = o.trace_closest(*ray((-3, 0, 1), (1, 0, 0)))[0]
    assert h["t"] == pytest.approx(3 - 1.0, rel=1e-5)  # radius 1 at z = +-1


def test_quadrics_bvh_matches_exhaustive():
    """aggregate_test_renderer.dart:42-118 with every quadric kind in one scene, under rotated transforms."""
    from dartray_b200 import host
    P, idx = random_soup(300, seed=21)
    o = Oracle()
    o.set_triangles(P, idx)
    o.set_spheres(np.zeros((0, 16), np.float32), np.zeros((0, 16), np.float32), np.zeros((0, 4)))
    for kind, prm, m in quadric_zoo(host):
        o.set_quadrics(kind, m.reshape(16), host.mat_inv(m).reshape(16), [prm])
    o.build_bvh()
    ro, rd = random_rays(20000, seed=22)
    h = o.trace_closest(ro, rd, nthreads=8)
    hb, nties, _ = o.trace_closest_brute(ro, rd, nthreads=8)
    single = nties <= 1
    assert (h["prim"][single] == hb["prim"][single]).all() and (h["t"] == hb["t"]).all()
    assert (o.trace_any(ro, rd, nthreads=8) == o.trace_any_brute(ro, rd, nthreads=8)).all()
    for k in range(4):
        assert (h["prim"] == 300 + k).sum() > 100, k  # each kind is hit


def quadric_zoo(host):
    return [
        (2, [0.3, -0.4, 0.5, 300.0, 0, 0, 0, 0], host.mat_mul(host.translate(0.4, 0.1, -0.3), host.rotate(40.0, (1.0, 0.2, 0.1)))),
        (3, [0.8, 0.4, 360.0, 0, 0, 0, 0, 0], host.mat_mul(host.translate(-0.5, -0.3, 0.2), host.rotate(-70.0, (0.1, 1.0, 0.3)))),
        (4, [0.45, 0.1, 0.7, 330.0, 0, 0, 0, 0], host.mat_mul(host.translate(0.0, 0.5, 0.4), host.rotate(120.0, (0.3, 0.2, 1.0)))),
        (5, [0.4, 0.0, -0.4, 0.1, 0.35, 0.4, 360.0, 0], host.mat_mul(host.translate(-0.1, -0.5, -0.5), host.rotate(25.0, (1.0, 1.0, 0.0)))),
    ]


def test_cylinder_light_irradiance_closed_form():
    # a matte element at the centre of an inward-emitting cylinder (radius R, z in [-H, H]), its normal along the axis:
    # the wall fills the polar angles atan(R / H) .. pi / 2, so E = pi * L * H^2 / (H^2 + R^2)
    from dartray_b200 import host
    kd, Le, H, R = 0.5, 3.0, 1.5, 2.0
    sb = host.SceneBuilder()
    e = 0.02
    sb.mesh([[-e, -e, 0], [e, -e, 0], [e, e, 0], [-e, e, 0]], [[0, 1, 2], [0, 2, 3]], material=sb.material((kd, kd, kd)))
    sb.cylinder(host.translate(0, 0, 0), radius=R, zmin=-H, zmax=H, area_light=(Le, Le, Le), nsamples=16, reverse=True)
    cam = host.PerspectiveCamera(host.look_at((0.05, 0.03, 0.6), (0, 0, 0), (0, 1, 0)), fov=0.5)
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, host.Film(4, 4), host.Sampler(kind=host.SAMPLER_LD, spp=64), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render()
    E = math.pi * Le * H * H / (H * H + R * R)
    assert o.film_read()["rgb"].mean() == pytest.approx(kd / math.pi * E, rel=1e-2)
