"""TransformedPrimitive: object instancing and animated shapes (motion blur).

The oracle's AnimatedTransform (oracle/ref_anim.h) is pinned against a restatement written HERE in numpy from
lib/core/{matrix4x4,quaternion,transform,animated_transform}.dart — float32 stores, binary64 expressions — and against closed
forms: an identity sphere instanced under a translation IS the sphere built with that transform (bit for bit: both transform the
ray with the same matrix), a translating sphere at time t is the static sphere at the interpolated position, an instanced mesh
reports the hits of the mesh with transformed vertices, motion bounds contain the bound at every time.  GPU tests (through the
C ABI: drt_set_instances, drt_set_ray_times) compare ray casts and renders with the oracle."""
import math

import numpy as np
import pytest

from dartray_b200 import capi, host
from tests.oracle_lib import Oracle

f32 = np.float32


# ---- numpy restatement of the reference's matrix / quaternion / AnimatedTransform arithmetic ----------------------------------
def np_mul(a, b):  # matrix4x4.dart:197-210
    a, b = a.astype(np.float64), b.astype(np.float64)
    r = np.zeros((4, 4), f32)
    for i in range(4):
        for j in range(4):
            r[i, j] = a[i, 0] * b[0, j] + a[i, 1] * b[1, j] + a[i, 2] * b[2, j] + a[i, 3] * b[3, j]
    return r


def np_inverse(m):  # matrix4x4.dart:242-357: cofactors over the determinant; checked against numpy.linalg below
    d = m.astype(np.float64)
    det = float(np.linalg.det(d))
    if det == 0.0:
        return m.copy()
    return np.linalg.inv(d).astype(f32)


def np_quat_from_matrix(m):  # quaternion.dart:39-77
    m = m.astype(np.float64)
    trace = m[0, 0] + m[1, 1] + m[2, 2]
    if trace > 0.0:
        s = math.sqrt(trace + 1.0)
        w = s / 2.0
        s = 0.5 / s
        v = [(m[2, 1] - m[1, 2]) * s, (m[0, 2] - m[2, 0]) * s, (m[1, 0] - m[0, 1]) * s]
    else:
        nxt = [1, 2, 0]
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = nxt[i], nxt[nxt[i]]
        s = math.sqrt((m[i, i] - (m[j, j] + m[k, k])) + 1.0)
        v = [0.0, 0.0, 0.0]
        v[i] = s * 0.5
        if s != 0.0:
            s = 0.5 / s
        w = (m[k, j] - m[j, k]) * s
        v[j] = (m[j, i] + m[i, j]) * s
        v[k] = (m[k, i] + m[i, k]) * s
    return np.array(v, f32), float(w)


def np_decompose(m):  # animated_transform.dart:61-105
    T = np.array([m[0, 3], m[1, 3], m[2, 3]], f32)
    M = m.copy()
    M[:3, 3] = 0
    M[3, :3] = 0
    M[3, 3] = 1
    R = M.copy()
    for _ in range(100):
        Rit = np_inverse(R.T.copy())
        Rnext = (0.5 * (R.astype(np.float64) + Rit.astype(np.float64))).astype(f32)
        norm = max(float(np.abs(R[i, :3].astype(np.float64) - Rnext[i, :3].astype(np.float64)).sum()) for i in range(3))
        R = Rnext
        if not norm > 0.0001:
            break
    return T, np_quat_from_matrix(R), np_mul(np_inverse(R), M)


def np_slerp(t, q1, q2):  # quaternion.dart:151-173
    (v1, w1), (v2, w2) = q1, q2
    dot = lambda a, b: float(np.dot(a[0].astype(np.float64), b[0].astype(np.float64))) + a[1] * b[1]
    scale = lambda q, f: ((q[0].astype(np.float64) * f).astype(f32), q[1] * f)
    add = lambda a, b: ((a[0].astype(np.float64) + b[0].astype(np.float64)).astype(f32), a[1] + b[1])
    sub = lambda a, b: ((a[0].astype(np.float64) - b[0].astype(np.float64)).astype(f32), a[1] - b[1])
    normalize = lambda q: ((q[0].astype(np.float64) / math.sqrt(dot(q, q))).astype(f32), q[1] / math.sqrt(dot(q, q)))
    c = dot(q1, q2)
    if c > 0.9995:
        return normalize(add(scale(q1, 1.0 - t), scale(q2, t)))
    thetap = math.acos(min(max(c, -1.0), 1.0)) * t
    qperp = normalize(sub(q2, scale(q1, c)))
    return add(scale(q1, math.cos(thetap)), scale(qperp, math.sin(thetap)))


def np_quat_matrix(q):  # quaternion.dart:120-149 (the matrix `m`; the Transform is (Transpose(m), m))
    (x, y, z), w = q[0].astype(np.float64), q[1]
    m = np.eye(4, dtype=f32)
    m[0, 0] = 1.0 - 2.0 * (y * y + z * z); m[0, 1] = 2.0 * (x * y + z * w); m[0, 2] = 2.0 * (x * z - y * w)
    m[1, 0] = 2.0 * (x * y - z * w); m[1, 1] = 1.0 - 2.0 * (x * x + z * z); m[1, 2] = 2.0 * (y * z + x * w)
    m[2, 0] = 2.0 * (x * z + y * w); m[2, 1] = 2.0 * (y * z - x * w); m[2, 2] = 1.0 - 2.0 * (x * x + y * y)
    return m


def np_interpolate(m0, m1, time, t0=0.0, t1=1.0):  # animated_transform.dart:107-136
    if time <= t0:
        return m0
    if time >= t1:
        return m1
    (T0, R0, S0), (T1, R1, S1) = np_decompose(m0), np_decompose(m1)
    dt = (time - t0) / (t1 - t0)
    trans = ((T0.astype(np.float64) * (1.0 - dt)).astype(f32).astype(np.float64) + (T1.astype(np.float64) * dt).astype(f32).astype(np.float64)).astype(f32)
    rot = np_quat_matrix(np_slerp(dt, R0, R1))
    scale = (S0.astype(np.float64) * (1.0 - dt) + S1.astype(np.float64) * dt).astype(f32)
    tm = np.eye(4, dtype=f32)
    tm[:3, 3] = trans
    return np_mul(np_mul(tm, rot.T.copy()), scale)


# ---- scenes ---------------------------------------------------------------------------------------------------------------------
def _ctm_pair():
    c0 = host.mat_mul(host.translate(-1, 0.2, 0.5), host.rotate(30, (0, 0, 1)))
    c1 = host.mat_mul(host.translate(-0.6, 0.4, 1.0), host.mat_mul(host.rotate(120, (0, 1, 1)), host.scale(1.0, 1.6, 0.8)))
    return c0, c1


def _rays(n, seed, origin=(0.0, 0.0, -6.0), spread=0.45):
    rng = np.random.default_rng(seed)
    d = np.stack([rng.uniform(-spread, spread, n), rng.uniform(-spread, spread, n), np.ones(n)], 1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ro = np.zeros((n, 4), f32)
    ro[:, :3] = origin
    rd = np.zeros((n, 4), f32)
    rd[:, :3] = d
    rd[:, 3] = np.inf
    return ro, rd


def _tetra(scale=0.6):
    P = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], f32) * scale
    I = np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], np.uint32)
    return P, I


def _instance_scene():
    """Two objects (a sphere; a tetrahedron mesh + a cylinder behind a nested BVH), four instances — static, translating, rotating
    and scaling — next to top-level shapes, under two point lights."""
    sb = host.SceneBuilder()
    red, green, grey = sb.material((0.7, 0.2, 0.2)), sb.material((0.2, 0.7, 0.3)), sb.material(0.6)
    sb.begin_object()
    sb.sphere(np.eye(4, dtype=f32), radius=0.5, material=red)
    ball = sb.end_object()
    sb.begin_object()
    P, I = _tetra()
    sb.mesh(P, I, material=green)
    sb.cylinder(host.translate(0.0, 0.0, -0.3), radius=0.2, zmin=-0.5, zmax=0.5, material=red)
    thing = sb.end_object(split=2, max_node_prims=1)
    sb.instance(ball, host.translate(1.2, 0.9, 0.0))
    sb.instance(ball, host.translate(1.0, -0.8, 0.0), host.translate(1.6, -0.2, 0.4))
    c0, c1 = _ctm_pair()
    sb.instance(thing, c0, c1)
    sb.instance(thing, host.mat_mul(host.translate(0.2, -1.2, 0.3), host.scale(0.5, 0.5, 0.5)))
    sb.mesh([[-3, -2, 2.5], [3, -2, 2.5], [3, 2.5, 2.5], [-3, 2.5, 2.5]], [[0, 1, 2], [2, 3, 0]], material=grey)
    sb.sphere(host.translate(-0.2, 1.3, 0.4), radius=0.35, material=grey)
    sb.point_light((0.5, 2.0, -4.0), (30, 28, 26))
    sb.point_light((-2.5, 0.5, -2.0), (6, 6, 8))
    return sb


# ---- oracle against the restatement and the closed forms --------------------------------------------------------------------
def test_matrix_inverse_is_the_inverse():
    rng = np.random.default_rng(1)
    for _ in range(20):
        m = np.eye(4, dtype=f32)
        m[:3, :] = rng.normal(size=(3, 4)).astype(f32)
        sb = host.SceneBuilder()
        sb.begin_object()
        sb.sphere(np.eye(4, dtype=f32), radius=0.5)
        ob = sb.end_object()
        sb.instance(ob, m, m)
        o = Oracle()
        host.upload_scene(o, sb.arrays())
        pr = o.instance_probe(0, 0.0)
        # the instance holds Inverse(ctm) = (mInv, m) as handed over; S of the decomposition goes through the cofactor inverse
        T, (Rv, Rw), S = np_decompose(pr["m"])
        assert np.allclose(pr["T"][0], T) and np.allclose(pr["R"][0][:3], Rv, atol=2e-6) and abs(pr["R"][0][3] - Rw) < 2e-6
        assert np.allclose(pr["S"][0], S, atol=5e-6)
        assert not pr["animated"]


def test_interpolate_matches_the_numpy_restatement():
    c0, c1 = _ctm_pair()
    sb = host.SceneBuilder()
    sb.begin_object()
    sb.sphere(np.eye(4, dtype=f32), radius=0.5)
    ob = sb.end_object()
    sb.instance(ob, c0, c1, 0.0, 1.0)
    o = Oracle()
    a = sb.arrays()
    host.upload_scene(o, a)
    m0, m1 = a["instance_start_m"][0].reshape(4, 4), a["instance_end_m"][0].reshape(4, 4)
    for t in (0.0, 0.1, 0.37, 0.5, 0.93, 1.0, 1.5, -0.2):
        pr = o.instance_probe(0, t)
        assert pr["animated"]
        expect = np_interpolate(m0, m1, t)
        assert np.allclose(pr["m"], expect, atol=3e-6), (t, pr["m"], expect)
        # m * mInv == identity: the inverse the products carry along is the inverse of the product
        assert np.allclose(pr["m"].astype(np.float64) @ pr["minv"].astype(np.float64), np.eye(4), atol=2e-5)
    # a rotation about one axis interpolates the angle (Slerp), the translation linearly
    sb2 = host.SceneBuilder()
    sb2.begin_object()
    sb2.sphere(np.eye(4, dtype=f32), radius=0.5)
    ob2 = sb2.end_object()
    sb2.instance(ob2, host.rotate(10, (0, 0, 1)), host.mat_mul(host.translate(2, 0, 0), host.rotate(70, (0, 0, 1))))
    o2 = Oracle()
    host.upload_scene(o2, sb2.arrays())
    mid = o2.instance_probe(0, 0.5)["minv"]  # primitive-to-world at t = 0.5
    # Inverse(ctm) interpolates, not the ctm: world-to-primitive translation lerps; its rotation is the slerp of the inverses
    w2p0, w2p1 = host.mat_inv(host.rotate(10, (0, 0, 1))), host.mat_inv(host.mat_mul(host.translate(2, 0, 0), host.rotate(70, (0, 0, 1))))
    expect_T = 0.5 * (w2p0[:3, 3] + w2p1[:3, 3])
    expect = host.mat_inv(host.mat_mul(host.translate(*expect_T), host.mat_inv(host.rotate(40, (0, 0, 1)))))
    assert np.allclose(mid, expect, atol=2e-6)


def test_static_instance_of_an_identity_sphere_is_the_transformed_sphere_bit_for_bit():
    xf = host.mat_mul(host.translate(0.3, -0.2, 0.5), host.mat_mul(host.rotate(33, (1, 2, 0.5)), host.scale(1.0, 1.0, 1.0)))
    sa, sbb = host.SceneBuilder(), host.SceneBuilder()
    sa.begin_object()
    sa.sphere(np.eye(4, dtype=f32), radius=0.7, zmin=-0.3, zmax=0.6, phimax=300.0)
    sa.instance(sa.end_object(), xf)
    sbb.sphere(xf, radius=0.7, zmin=-0.3, zmax=0.6, phimax=300.0)
    oa, ob = Oracle(), Oracle()
    host.upload_scene(oa, sa.arrays())
    host.upload_scene(ob, sbb.arrays())
    ro, rd = _rays(4000, 3, spread=0.2)
    ha, hb = oa.trace_closest(ro, rd), ob.trace_closest(ro, rd)
    assert (hb["prim"] >= 0).sum() > 500
    for k in ("t", "b1", "b2"):
        assert np.array_equal(ha[k].view(np.uint32), hb[k].view(np.uint32)), k
    assert np.array_equal(ha["prim"], hb["prim"])
    assert np.array_equal(oa.trace_any(ro, rd), ob.trace_any(ro, rd))


def test_translating_sphere_at_time_t_is_the_sphere_at_the_interpolated_position():
    p0, p1 = np.array([-0.5, 0.1, 0.0]), np.array([0.7, -0.3, 0.6])
    sa = host.SceneBuilder()
    sa.animated(lambda s: s.sphere(np.eye(4, dtype=f32), radius=0.6), host.translate(*p0), host.translate(*p1))
    oa = Oracle()
    host.upload_scene(oa, sa.arrays())
    ro, rd = _rays(3000, 5, spread=0.25)
    for t in (0.0, 0.25, 0.6, 1.0):
        oa.set_ray_times(np.full(ro.shape[0], t))
        ha = oa.trace_closest(ro, rd)
        # world-to-primitive translation: the float32 lerp of -p0 and -p1; the static twin sits at its negation
        w2p = ((-p0.astype(f32)).astype(np.float64) * (1.0 - t)).astype(f32).astype(np.float64) + ((-p1.astype(f32)).astype(np.float64) * t).astype(f32).astype(np.float64)
        pos = -(w2p.astype(f32)).astype(np.float64)
        sbb = host.SceneBuilder()
        sbb.sphere(host.translate(*pos), radius=0.6)
        ob = Oracle()
        host.upload_scene(ob, sbb.arrays())
        hb = ob.trace_closest(ro, rd)
        assert (hb["prim"] >= 0).sum() > 300
        assert np.array_equal(ha["prim"] >= 0, hb["prim"] >= 0)
        assert np.array_equal(ha["t"].view(np.uint32), hb["t"].view(np.uint32))


def test_instanced_mesh_reports_the_hits_of_the_transformed_mesh_and_the_nested_bvh_equals_brute_force():
    xf = host.mat_mul(host.translate(0.2, 0.1, 0.3), host.mat_mul(host.rotate(50, (0.3, 1, 0.2)), host.scale(1.2, 0.8, 1.0)))
    P, I = _tetra(0.8)
    rng = np.random.default_rng(9)
    P = np.concatenate([P + rng.normal(0, 0.4, 3).astype(f32) for _ in range(6)])
    I = np.concatenate([I + 4 * k for k in range(6)])
    sa, sbb = host.SceneBuilder(), host.SceneBuilder()
    sa.begin_object()
    sa.mesh(P, I)
    sa.instance(sa.end_object(split=2, max_node_prims=2), xf)
    sbb.mesh(P, I, o2w=xf)
    oa, ob = Oracle(), Oracle()
    host.upload_scene(oa, sa.arrays())
    host.upload_scene(ob, sbb.arrays())
    ro, rd = _rays(5000, 11, spread=0.3)
    ha, hb = oa.trace_closest(ro, rd), ob.trace_closest(ro, rd)
    both = (ha["prim"] >= 0) & (hb["prim"] >= 0)
    assert both.sum() > 1000 and ((ha["prim"] >= 0) != (hb["prim"] >= 0)).mean() < 2e-3  # silhouette rays may flip
    same = both & (ha["prim"] == hb["prim"])
    assert same.sum() > 0.99 * both.sum()
    assert np.allclose(ha["t"][same], hb["t"][same], rtol=2e-5)
    # the two-level walk against the exhaustive loop over the top-level primitives (aggregate_test_renderer.dart's pattern)
    hbrute = oa.trace_closest_brute(ro, rd)[0] if hasattr(oa, "trace_closest_brute") else None
    if hbrute is not None:
        assert np.array_equal(hbrute["prim"], ha["prim"]) and np.array_equal(hbrute["t"].view(np.uint32), ha["t"].view(np.uint32))


def test_motion_bounds_contain_the_bound_at_every_time():
    sb = _instance_scene()
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    for inst in range(4):
        pr0 = o.instance_probe(inst, 0.0)
        lo, hi = pr0["bound"][:3], pr0["bound"][3:]
        box = np.array([[-0.6, -0.6, -0.8], [0.6, 0.6, 0.6]]) if inst >= 2 else np.array([[-0.5] * 3, [0.5] * 3])
        for t in np.linspace(0, 1, 23):
            p2w = o.instance_probe(inst, t)["minv"].astype(np.float64)
            corners = np.array([[box[i][0], box[j][1], box[k][2], 1.0] for i in (0, 1) for j in (0, 1) for k in (0, 1)])
            w = corners @ p2w.T
            # the object's own bound is looser than `box` only for the mesh object (the tetrahedron's extent 0.6)
            assert (w[:, :3].min(0) >= lo - 0.15).all() and (w[:, :3].max(0) <= hi + 0.15).all()
    assert o.instance_probe(0)["animated"] is False and o.instance_probe(1)["animated"] is True


def test_renders_of_an_instance_scene_run_and_moving_shapes_blur():
    sb = _instance_scene()
    cam = host.PerspectiveCamera(host.look_at((0.2, 0.3, -6.0), (0, 0, 0.5), (0, 1, 0)), fov=40.0)
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, host.Film(48, 36), host.Sampler(kind=host.SAMPLER_LD, spp=8), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render(0, 1, 8)
    img = o.film_read()["rgb"]
    assert np.isfinite(img).all() and img.max() > 0.05
    # the same scene frozen at its start transforms differs where the shapes move
    sb2 = _instance_scene()
    sb2.instances = [(i[0], i[1], i[2], i[1], i[2], i[5], i[6]) for i in sb2.instances]
    o2 = Oracle()
    host.upload_scene(o2, sb2.arrays())
    host.configure_render(o2, cam, host.Film(48, 36), host.Sampler(kind=host.SAMPLER_LD, spp=8), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o2.render(0, 1, 8)
    img2 = o2.film_read()["rgb"]
    assert np.abs(img - img2).max() > 0.02


def test_abi_argument_errors_without_a_device():
    ctx = capi.Context(capi.DEVICE_NONE) if hasattr(capi, "DEVICE_NONE") else None
    if ctx is None:
        pytest.skip("no host-only context")
    sb = _instance_scene()
    a = sb.arrays()
    ctx.set_triangles(a["P"], a["idx"], a["tri_mat"], a["tri_light"], a["tri_rev"])
    ctx.set_spheres(a["sph_o2w"], a["sph_w2o"], a["sph_params"], a["sph_mat"], a["sph_light"], a["sph_rev"])
    args = [a["object_offsets"], a["object_prims"], a["object_split"], a["object_max_node_prims"], a["instance_object"],
            a["instance_start_m"], a["instance_start_minv"], a["instance_end_m"], a["instance_end_minv"], a["instance_times"]]
    bad = list(args)
    bad[4] = np.array([0, 1, 7, 1], np.uint32)
    with pytest.raises(capi.DrtError, match="object that was not defined"):
        ctx.set_instances(*bad)
    ctx.set_instances(*args)
    ctx.set_build_order(None)
    with pytest.raises(capi.DrtError, match="top-level build order"):
        ctx.build_bvh(2, 4)


# ---- GPU against the oracle -----------------------------------------------------------------------------------------------------
def _both_contexts(sb):
    g, o = capi.Context(0), Oracle()
    a = sb.arrays()
    for c in (g, o):
        host.upload_scene(c, a)
    return g, o


@pytest.mark.gpu
def test_gpu_ray_cast_through_instances_matches_the_oracle():
    g, o = _both_contexts(_instance_scene())
    ro, rd = _rays(40000, 21, spread=0.4)
    rng = np.random.default_rng(2)
    for times in (None, np.zeros(ro.shape[0]), np.ones(ro.shape[0]), rng.random(ro.shape[0])):
        for c in (g, o):
            c.set_ray_times(times)
        hg, ho = g.trace_closest(ro, rd), o.trace_closest(ro, rd, nthreads=8)
        agree = hg["prim"] == ho["prim"]
        # static transforms and the end transforms are bit-exact; in between the slerp's sin / cos come from two libms, and a
        # matrix element that rounds the other way moves a silhouette ray
        interpolated = times is not None and not (np.all(times == 0) or np.all(times == 1))
        print("instances: prim agreement", agree.mean(), "hits", (ho["prim"] >= 0).mean())
        assert (ho["prim"] >= 0).mean() > 0.3
        if interpolated:
            assert agree.mean() > 0.9995
            assert np.allclose(hg["t"][agree], ho["t"][agree], rtol=2e-5)
        else:
            assert agree.all()
            for k in ("t", "b1", "b2"):
                assert np.allclose(hg[k], ho[k], rtol=1e-6, atol=1e-6), k  # u, v of quadrics: atan2 of two libms
            assert np.array_equal(hg["t"].view(np.uint32), ho["t"].view(np.uint32))
        og, oo = g.trace_any(ro, rd), o.trace_any(ro, rd, nthreads=8)
        assert (og != oo).mean() < (5e-4 if interpolated else 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4), host.Integrator(kind=host.INTEGRATOR_DIRECT),
                                   host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=16), host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=2)],
                         ids=["path", "direct", "ao", "whitted"])
def test_gpu_render_with_instances_and_motion_blur_matches_the_oracle(integ):
    sb = _instance_scene()
    cam = host.PerspectiveCamera(host.look_at((0.2, 0.3, -6.0), (0, 0, 0.5), (0, 1, 0)), fov=40.0)
    g, o = _both_contexts(sb)
    for c in (g, o):
        host.configure_render(c, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=8), integ)
    g.render(0, 1)
    o.render(0, 1, 8)
    fg, fo = g.film_read()["rgb"], o.film_read()["rgb"]
    err = np.abs(fg - fo) / np.maximum(np.abs(fo), 1e-3)
    # a sample whose ray grazes a moving silhouette may fall on the other side (two libms in the slerp): count such pixels
    bad = (err.max(axis=2) > 1e-3).mean()
    print("instances render", integ.kind, "max rel err", err.max(), "pixels off", bad, "mean", float(fo.mean()))
    assert np.isfinite(fg).all() and fo.max() > 0.05
    assert bad <= 2e-3
    assert np.median(err) < 1e-5
    sg, so = g.render_stats(), o.render_stats()
    assert sg["camera_samples"] == so["camera_samples"]
    assert abs(sg["closest_rays"] - so["closest_rays"]) <= 1e-3 * so["closest_rays"]


def _render_both(sb, sampler, integ, film=(64, 48)):
    cam = host.PerspectiveCamera(host.look_at((0.2, 0.3, -6.0), (0, 0, 0.5), (0, 1, 0)), fov=40.0)
    g, o = _both_contexts(sb)
    for c in (g, o):
        host.configure_render(c, cam, host.Film(*film), sampler, integ)
    g.render(0, 1)
    o.render(0, 1, 8)
    fg, fo = g.film_read()["rgb"], o.film_read()["rgb"]
    err = np.abs(fg - fo) / np.maximum(np.abs(fo), 1e-3)
    return fg, fo, err


@pytest.mark.gpu
@pytest.mark.parametrize("sampler", [host.Sampler(kind=host.SAMPLER_HALTON, spp=4), host.Sampler(kind=host.SAMPLER_RANDOM, spp=4),
                                     host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2)], ids=["halton", "random", "stratified"])
def test_gpu_instances_with_the_samplers_whose_time_samples_are_doubles(sampler):
    fg, fo, err = _render_both(_instance_scene(), sampler, host.Integrator(kind=host.INTEGRATOR_DIRECT))
    print("instances, sampler", sampler.kind, "max rel err", err.max(), "mean", float(fo.mean()))
    assert fo.max() > 0.05 and (err.max(axis=2) > 1e-3).mean() <= 2e-3 and np.median(err) < 1e-5


@pytest.mark.gpu
def test_gpu_instances_with_textured_bumped_materials():
    from tests.test_textures_gpu import IMG_F, IMG_RGB
    sb = host.SceneBuilder()
    kd = host.ScaleTexture(host.ImageTexture(IMG_RGB, host.UVMapping(3.0, 2.0, 0.1, 0.2)), (0.9, 0.8, 0.7))
    bump = host.ScaleTexture(host.ImageTexture(IMG_F, host.UVMapping(2.0, 2.0)), -0.05)
    tex = sb.material_program("uber", kd=kd, ks=0.05, roughness=0.05, bumpmap=bump)
    marble = sb.material_program("matte", kd=host.MarbleTexture() if hasattr(host, "MarbleTexture") else kd)
    sb.begin_object()
    sb.sphere(np.eye(4, dtype=f32), radius=0.5, material=tex)
    ball = sb.end_object()
    sb.begin_object()
    P, I = _tetra()
    sb.mesh(P, I, material=marble, uv=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], f32))
    sb.cylinder(host.translate(0.0, 0.0, -0.3), radius=0.2, zmin=-0.5, zmax=0.5, material=tex)
    thing = sb.end_object(split=2, max_node_prims=1)
    sb.instance(ball, host.translate(1.2, 0.9, 0.0))
    sb.instance(ball, host.translate(1.0, -0.8, 0.0), host.translate(1.6, -0.2, 0.4))
    c0, c1 = _ctm_pair()
    sb.instance(thing, c0, c1)
    sb.mesh([[-3, -2, 2.5], [3, -2, 2.5], [3, 2.5, 2.5], [-3, 2.5, 2.5]], [[0, 1, 2], [2, 3, 0]], material=tex,
            uv=np.array([[0, 0], [2, 0], [2, 2], [0, 2]], f32))
    sb.point_light((0.5, 2.0, -4.0), (30, 28, 26))
    for integ in (host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3), host.Integrator(kind=host.INTEGRATOR_DIRECT)):
        fg, fo, err = _render_both(sb, host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
        print("instances, textured", integ.kind, "max rel err", err.max(), "mean", float(fo.mean()))
        assert fo.max() > 0.05 and (err.max(axis=2) > 1e-3).mean() <= 2e-3 and np.median(err) < 1e-5


@pytest.mark.gpu
def test_gpu_instances_inside_participating_media():
    sb = _instance_scene()
    sb.volume("homogeneous", sigma_a=0.05, sigma_s=0.15, g=0.2, p0=(-2.5, -2, -1), p1=(2.5, 2.5, 2.4))
    sb.vol_integrator = (1, 0.25)  # single scattering, step 0.25
    fg, fo, err = _render_both(sb, host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3), film=(48, 36))
    print("instances in fog: max rel err", err.max(), "mean", float(fo.mean()))
    assert fo.max() > 0.02 and (err.max(axis=2) > 1e-3).mean() <= 5e-3 and np.median(err) < 1e-5


def test_animated_camera_interpolates_the_camera_transform():
    """A camera that translates during the exposure: the oracle's film equals, sample for sample, what a static camera at each
    sample's interpolated position would see — checked on the camera rays through orc_pixel_samples-free means: a scene that is a
    single huge emissive wall has the same image; a sphere in front of it smears along the motion."""
    sb = host.SceneBuilder()
    sb.sphere(host.translate(0.0, 0.0, 0.0), radius=0.5, material=sb.material((0.8, 0.2, 0.2)))
    sb.point_light((0.0, 3.0, -4.0), (40, 40, 40))
    c0 = host.look_at((0.0, 0.0, -5.0), (0, 0, 0), (0, 1, 0))
    c1 = host.look_at((1.0, 0.0, -5.0), (1, 0, 0), (0, 1, 0))
    films = []
    for end in (None, c1):
        cam = host.PerspectiveCamera(c0, fov=30.0)
        cam.camera_to_world_end = end
        o = Oracle()
        host.upload_scene(o, sb.arrays())
        host.configure_render(o, cam, host.Film(48, 32), host.Sampler(kind=host.SAMPLER_LD, spp=16), host.Integrator(kind=host.INTEGRATOR_DIRECT))
        o.render(0, 1, 8)
        films.append(o.film_read()["rgb"])
    lit = [(f.sum(axis=2) > 1e-3) for f in films]
    # the moving camera sees the sphere smeared towards -x (the camera moves to +x): more pixels lit, each dimmer on average
    assert lit[1].sum() > 1.3 * lit[0].sum()
    xs = np.arange(48)[None, :]
    assert (lit[1] * xs).sum() / lit[1].sum() < (lit[0] * xs).sum() / lit[0].sum() - 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1, 2], ids=["perspective", "orthographic", "environment"])
def test_gpu_animated_camera_matches_the_oracle(kind):
    from tests.test_textures_gpu import IMG_RGB
    sb = _instance_scene()
    c0 = host.look_at((0.2, 0.3, -6.0), (0, 0, 0.5), (0, 1, 0))
    c1 = host.mat_mul(host.look_at((0.9, 0.1, -5.5), (0.2, 0, 0.5), (0.1, 1, 0)), host.rotate(4.0, (0, 0, 1)))
    if kind == 0:
        cam = host.PerspectiveCamera(c0, fov=40.0, lens_radius=0.05, focal_distance=6.0)
    elif kind == 1:
        cam = host.OrthographicCamera(c0, screen_window=(-3.0, 3.0, -2.25, 2.25)) if hasattr(host, "OrthographicCamera") else None
    else:
        cam = host.EnvironmentCamera(c0) if hasattr(host, "EnvironmentCamera") else None
    if cam is None:
        pytest.skip("camera class not in host.py")
    cam.camera_to_world_end = c1
    g, o = _both_contexts(sb)
    for c in (g, o):
        host.configure_render(c, cam, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), host.Integrator(kind=host.INTEGRATOR_DIRECT))
    g.render(0, 1)
    o.render(0, 1, 8)
    fg, fo = g.film_read()["rgb"], o.film_read()["rgb"]
    err = np.abs(fg - fo) / np.maximum(np.abs(fo), 1e-3)
    print("animated camera", kind, "max rel err", err.max(), "pixels off", (err.max(axis=2) > 1e-3).mean(), "mean", float(fo.mean()))
    assert fo.max() > 0.02 and (err.max(axis=2) > 1e-3).mean() <= 5e-3 and np.median(err) < 1e-5
