"""MeasuredMaterial: the two BxDFs it adds — RegularHalfangleBRDF over a .merl table, IrregularIsotropicBRDF over the
samples of a .brdf file — pinned against restatements written HERE from lib/core/reflection/{regular_halfangle_brdf,
irregular_isotropic_brdf,brdf_remap}.dart and lib/core/kdtree.dart (independent of oracle/ref_render.cpp), the host-side table
loaders, and the GPU path against the oracle (flattened lobes and material program kind 11 with a bump map)."""
import math
import struct

import numpy as np
import pytest

from dartray_b200 import host
from tests.oracle_lib import Oracle

RNG = np.random.default_rng(2024)
# a small regular-halfangle table (the reference fixes 90 x 90 x 180; the BxDF takes any) and an irregular sample cloud
TABLE = (RNG.random((6, 5, 8, 3)) * 0.5).astype(np.float32)
N_SAMPLES = 700


def _cloud(n=N_SAMPLES, seed=5):
    rng = np.random.default_rng(seed)
    th_i, th_o = np.arccos(rng.random(n)), np.arccos(rng.random(n))
    ph_i, ph_o = rng.random(n) * 2 * np.pi, rng.random(n) * 2 * np.pi
    rgb = np.stack([0.3 + 0.2 * np.cos(th_i), 0.2 + 0.3 * np.cos(th_o), 0.1 + 0.1 * np.cos(ph_i - ph_o) ** 2], 1)
    return host.brdf_samples(th_i, ph_i, th_o, ph_o, rgb)


CLOUD = _cloud()


def _dirs(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v[:, 2] = np.abs(v[:, 2])
    return v.astype(np.float32).astype(np.float64)


def _oracle():
    sb = host.SceneBuilder()
    t0 = sb.measured_table(host.MEASURED_REGULAR_HALFANGLE, TABLE)
    t1 = sb.measured_table(host.MEASURED_IRREGULAR_ISOTROPIC, CLOUD)
    ids = [sb.material_lobes(host.measured_lobes(t0, host.MEASURED_REGULAR_HALFANGLE)),
           sb.material_lobes(host.measured_lobes(t1, host.MEASURED_IRREGULAR_ISOTROPIC))]
    sb.mesh([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]], material=ids[0])
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    return o, ids


# ---- restatements ------------------------------------------------------------------------------------------------------
def py_regular_halfangle(table, wo, wi):  # regular_halfangle_brdf.dart:27-75, scalar, as written
    f32 = lambda v: np.asarray(v, np.float32).astype(np.float64)
    n_h, n_d, n_p = table.shape[:3]
    wo, wi = f32(wo), f32(wi)
    wh = f32(wo + wi)
    if wh[2] < 0.0:
        wo, wi, wh = -wo, -wi, -wh
    if not wh.any():
        return np.zeros(3)
    wh = f32(wh / math.sqrt(float(wh @ wh)))
    sin_h = math.sqrt(max(0.0, 1.0 - wh[2] * wh[2]))
    cos_phi = 1.0 if sin_h == 0.0 else min(max(wh[0] / sin_h, -1.0), 1.0)
    sin_phi = 0.0 if sin_h == 0.0 else min(max(wh[1] / sin_h, -1.0), 1.0)
    whx = f32([cos_phi * wh[2], sin_phi * wh[2], -sin_h])
    why = f32([-sin_phi, cos_phi, 0.0])
    wd = f32([wi @ whx, wi @ why, wi @ wh])
    wh_theta = math.acos(min(max(wh[2], -1.0), 1.0))
    wd_theta = math.acos(min(max(wd[2], -1.0), 1.0))
    wd_phi = math.atan2(wd[1], wd[0])
    if wd_phi < 0.0:
        wd_phi += 2.0 * math.pi
    if wd_phi > math.pi:
        wd_phi -= math.pi
    # int REMAP(V, MAX, COUNT) => ((V / MAX).toInt() * COUNT).clamp(0, COUNT - 1): truncation first
    remap = lambda v, mx, count: min(max(int(v / mx) * count, 0), count - 1)
    ih = remap(math.sqrt(max(0.0, wh_theta / (math.pi / 2.0))), 1.0, n_h)
    idt = remap(wd_theta, math.pi / 2.0, n_d)
    ip = remap(wd_phi, math.pi, n_p)
    return table[ih, idt, ip].astype(np.float64)


class PyKdTree:  # kdtree.dart:24-112 (left child = nodeNum + 1; any median partition is a valid nth_element)
    def __init__(self, pts):
        self.pts = np.asarray(pts, np.float64)
        n = len(pts)
        self.split_axis, self.split_pos = np.full(n, 3, np.int32), np.zeros(n)
        self.has_left, self.right = np.zeros(n, bool), np.full(n, n, np.int64)
        self.data = np.zeros(n, np.int64)
        self.next_free = 1
        self._build(0, list(range(n)))

    def _build(self, node, items):
        if len(items) == 1:
            self.data[node] = items[0]
            return
        p = self.pts[items]
        axis = int(np.argmax(p.max(0) - p.min(0)))
        items = sorted(items, key=lambda i: (self.pts[i, axis], i))
        mid = len(items) // 2
        self.split_axis[node], self.split_pos[node], self.data[node] = axis, self.pts[items[mid], axis], items[mid]
        if mid > 0:
            self.has_left[node] = True
            child = self.next_free
            self.next_free += 1
            self._build(child, items[:mid])
        if mid + 1 < len(items):
            self.right[node] = self.next_free
            self.next_free += 1
            self._build(self.right[node], items[mid + 1:])

    def lookup(self, p, max_d2, node=0, out=None):
        out = [] if out is None else out
        n, axis = len(self.pts), self.split_axis[node]
        if axis != 3:
            d2 = (p[axis] - self.split_pos[node]) ** 2
            if p[axis] <= self.split_pos[node]:
                if self.has_left[node]:
                    self.lookup(p, max_d2, node + 1, out)
                if d2 < max_d2 and self.right[node] < n:
                    self.lookup(p, max_d2, self.right[node], out)
            else:
                if self.right[node] < n:
                    self.lookup(p, max_d2, self.right[node], out)
                if d2 < max_d2 and self.has_left[node]:
                    self.lookup(p, max_d2, node + 1, out)
        i = self.data[node]
        d = (self.pts[i].astype(np.float32) - np.asarray(p, np.float32)).astype(np.float64)  # DistanceSquared of float32 Points
        if float(d @ d) < max_d2:
            out.append((int(i), float(d @ d)))
        return out


def py_irregular_isotropic(tree, values, wo, wi):  # irregular_isotropic_brdf.dart:36-62
    m = host.brdf_remap(wo, wi).astype(np.float64)
    last = 0.001
    while True:
        found = tree.lookup(m, last)
        if len(found) > 2 or last > 1.5:
            w = np.array([math.exp(-100.0 * d2) for _, d2 in found])
            v = (values[[i for i, _ in found]].astype(np.float64) * w[:, None]).sum(0) if found else np.zeros(3)
            return np.maximum(v, 0.0) / w.sum(), sorted(i for i, _ in found)
        last *= 2.0


# ---- oracle against the restatements -----------------------------------------------------------------------------------
def test_regular_halfangle_matches_the_restatement_and_reads_cell_0_as_written():
    o, ids = _oracle()
    wo, wi = _dirs(400, 1), _dirs(400, 2)
    f, pdf = o.bsdf_eval(ids[0], wo, wi)
    expect = np.stack([py_regular_halfangle(TABLE, a, b) for a, b in zip(wo, wi)])
    assert np.array_equal(f, expect.astype(np.float32))
    # REMAP truncates before it multiplies: every generic direction pair lands in cell (0, 0, 0)
    assert np.array_equal(f, np.broadcast_to(TABLE[0, 0, 0], f.shape))
    assert np.allclose(pdf, wi[:, 2] / np.pi, rtol=1e-12)  # bxdf.dart:84-88: neither BxDF overrides pdf / sample_f


def test_irregular_isotropic_matches_a_kdtree_transliteration():
    o, ids = _oracle()
    wo, wi = _dirs(300, 3), _dirs(300, 4)
    f, pdf = o.bsdf_eval(ids[1], wo, wi)
    tree = PyKdTree(CLOUD[:, :3])
    counts = []
    for k in range(wo.shape[0]):
        expect, found = py_irregular_isotropic(tree, CLOUD[:, 3:], wo[k], wi[k])
        counts.append(len(found))
        # the kd-tree hands over exactly the samples inside the radius: a brute-force scan finds the same set
        m = host.brdf_remap(wo[k], wi[k])
        d = (CLOUD[:, :3] - m).astype(np.float64)
        d2 = (d * d).sum(1)
        last = 0.001
        while not ((d2 < last).sum() > 2 or last > 1.5):
            last *= 2.0
        assert sorted(np.nonzero(d2 < last)[0].tolist()) == found
        assert np.allclose(f[k], expect, rtol=5e-6, atol=1e-9), (k, f[k], expect)  # float32 Spectrum sums, order not fixed by the reference
    assert min(counts) >= 3 and max(counts) > 10  # the radius search was exercised at several sizes
    assert np.allclose(pdf, wi[:, 2] / np.pi, rtol=1e-12)


def test_sample_f_is_the_cosine_sampling_of_bxdf():
    o, ids = _oracle()
    u = np.random.default_rng(9).random((500, 3))
    wo = _dirs(1, 7)[0]
    for mid in ids:
        wi, f, pdf, ty = o.bsdf_sample(mid, wo, u)
        assert (ty == 1 | 8).all()  # BSDF_REFLECTION | BSDF_GLOSSY
        assert np.allclose(pdf, np.abs(wi[:, 2]) / np.pi, rtol=1e-6)
        f2, pdf2 = o.bsdf_eval(mid, wo, wi)
        assert np.array_equal(f, f2)


def test_brdf_remap_and_samples():
    # mirror directions: dphi = pi -> y = 1; equal directions: dphi = 0
    w = np.array([0.6, 0.0, 0.8])
    assert np.allclose(host.brdf_remap(w, w * [-1, 1, 1]), [0.36, 1.0, 0.64], atol=1e-6)
    assert np.allclose(host.brdf_remap(w, w), [0.36, 0.0, 0.64], atol=1e-6)
    s = host.brdf_samples([0.3], [0.1], [0.7], [2.0], [[0.1, 0.2, 0.3]])
    assert s.shape == (1, 6) and s.dtype == np.float32
    assert np.allclose(s[0, :3], [math.sin(0.3) * math.sin(0.7), 1.9 / math.pi, math.cos(0.3) * math.cos(0.7)], atol=1e-6)
    assert np.allclose(s[0, 3:], [0.1, 0.2, 0.3])


def test_merl_table_loader():
    n = 90 * 90 * 180
    rng = np.random.default_rng(3)
    planes = rng.normal(300.0, 400.0, (3, n))  # some negative: clamped to 0
    blob = struct.pack("<3i", 90, 90, 180) + planes.astype("<f8").tobytes()
    t = host.merl_table(blob)
    assert t.shape == (90, 90, 180, 3) and t.dtype == np.float32
    k = 123457
    cell = t.reshape(-1, 3)[k]
    for c, s in enumerate((1.0 / 1500.0, 1.15 / 1500.0, 1.66 / 1500.0)):
        assert cell[c] == np.float32(max(0.0, float(np.float32(planes[c, k])) * s))
    assert (t >= 0).all() and (t == 0).any()
    with pytest.raises(ValueError):
        host.merl_table(struct.pack("<3i", 90, 90, 90) + b"\0" * 64)


def test_program_without_bump_renders_the_flattened_lobes_film():
    """Material program kind 11 builds the BSDF the flattened lobe list holds (oracle, CPU)."""
    from tests.test_textures_gpu import CAM
    films = []
    for program in (False, True):
        sb = _measured_scene(program, bump=False)
        o = Oracle()
        host.upload_scene(o, sb.arrays())
        host.configure_render(o, CAM, host.Film(24, 18), host.Sampler(kind=host.SAMPLER_LD, spp=2), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3))
        o.render(0, 1, 8)
        films.append(o.film_read()["rgb"])
    assert films[0].max() > 0 and np.array_equal(films[0], films[1])


# ---- GPU against the oracle ----------------------------------------------------------------------------------------------
def _measured_scene(program: bool, bump: bool = True):
    from tests.test_textures_gpu import IMG_F, _scene
    tables = {}

    def material_of(sb, name):
        if not tables:
            tables["merl"] = sb.measured_table(host.MEASURED_REGULAR_HALFANGLE, TABLE)
            tables["brdf"] = sb.measured_table(host.MEASURED_IRREGULAR_ISOTROPIC, CLOUD)
            if program:
                bm = host.ScaleTexture(host.ImageTexture(IMG_F, host.UVMapping(2.0, 2.0)), -0.06) if bump else None
                tables["m_merl"] = sb.material_program("measured", m1=tables["merl"], bumpmap=bm)
                tables["m_brdf"] = sb.material_program("measured", m1=tables["brdf"], bumpmap=bm)
            else:
                tables["m_merl"] = sb.material_lobes(host.measured_lobes(tables["merl"], host.MEASURED_REGULAR_HALFANGLE))
                tables["m_brdf"] = sb.material_lobes(host.measured_lobes(tables["brdf"], host.MEASURED_IRREGULAR_ISOTROPIC))
        return tables["m_merl"] if name in ("wall", "cone", "disk") else tables["m_brdf"]
    return _scene(material_of)


@pytest.mark.gpu
@pytest.mark.parametrize("program", [False, True], ids=["lobes", "program_bump"])
@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=3), host.Integrator(kind=host.INTEGRATOR_DIRECT)],
                         ids=["path", "direct"])
def test_gpu_measured_materials_match_the_oracle(program, integ):
    from tests.test_textures_gpu import CAM, _both, _check
    sb = _measured_scene(program)
    g, o, fg, fo = _both(sb, CAM, host.Film(64, 48), host.Sampler(kind=host.SAMPLER_LD, spp=4), integ)
    _check(g, o, fg, fo, f"measured ({'program + bump' if program else 'lobes'}), integrator {integ.kind}")
