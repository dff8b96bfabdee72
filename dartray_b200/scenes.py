"""Synthetic procedural workloads named by BASELINE.json / SURVEY.md §8d.

The reference ships no procedural scenes; these generators pin the byte-exact inputs that both
the CUDA path and the CPU oracle consume (same arrays, uploaded through the same C ABI).
Everything is derived from PCG32 streams so a scene is reproducible from its name alone.

* ``soup(n_spheres)``  — ``soup_1m`` (512 UV-spheres, 1,015,808 triangles + 2-triangle ground quad)
  and ``soup_10m`` (5,120 UV-spheres).
* ``coherent_rays`` / ``incoherent_rays`` — the config-2 ray sets.
* ``cornell_synth`` — the 22 triangles + sphere of web/scenes/cornell-path.pbrt:23-59 with the disk
  light replaced by a 2-triangle quad light.
"""
from __future__ import annotations

import hashlib
import math

import numpy as np

_MULT = 6364136223846793005
_MASK = (1 << 64) - 1


class PCG32:
    """PCG-XSH-RR 64/32 (O'Neill), vectorised with LCG jump-ahead so draw i is O(1)."""

    _BLOCK = 1 << 20

    def __init__(self, seed: int, stream: int = 1):
        self.inc = ((stream << 1) | 1) & _MASK
        s = 0
        s = (s * _MULT + self.inc) & _MASK
        s = (s + seed) & _MASK
        s = (s * _MULT + self.inc) & _MASK
        self.state = s
        # A[i], C[i] with state_{k+i} = A[i]*state_k + C[i]  (mod 2^64)
        n = self._BLOCK
        A = np.empty(n + 1, dtype=np.uint64)
        C = np.empty(n + 1, dtype=np.uint64)
        A[0], C[0] = 1, 0
        A[1], C[1] = _MULT, self.inc
        have = 2
        with np.errstate(over="ignore"):
            while have < n + 1:
                m = min(have - 1, n + 1 - have)  # extend using jump by (have-1)
                j = have - 1
                A[have:have + m] = A[1:1 + m] * A[j]
                C[have:have + m] = A[1:1 + m] * C[j] + C[1:1 + m]
                have += m
        self._A, self._C = A, C

    def u32(self, n: int) -> np.ndarray:
        """Next n 32-bit outputs."""
        out = np.empty(n, dtype=np.uint32)
        done = 0
        with np.errstate(over="ignore"):
            while done < n:
                m = min(self._BLOCK, n - done)
                old = self._A[:m] * np.uint64(self.state) + self._C[:m]
                xs = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
                rot = (old >> np.uint64(59)).astype(np.uint32)
                out[done:done + m] = (xs >> rot) | (xs << ((np.uint32(32) - rot) & np.uint32(31)))
                self.state = (int(self._A[m]) * self.state + int(self._C[m])) & _MASK
                done += m
        return out

    def uniform(self, n: int) -> np.ndarray:
        """n doubles in [0, 1): u32 * 2^-32."""
        return self.u32(n).astype(np.float64) * (1.0 / 4294967296.0)


def _uv_sphere_template(stacks: int, slices: int):
    """Unit UV-sphere: (dirs f64 [nv,3], tris int64 [nt,3]), outward-facing winding."""
    dirs = [(0.0, 1.0, 0.0)]
    for i in range(1, stacks):
        th = math.pi * i / stacks
        for j in range(slices):
            ph = 2.0 * math.pi * j / slices
            dirs.append((math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)))
    dirs.append((0.0, -1.0, 0.0))
    south = len(dirs) - 1
    ring = lambda i, j: 1 + (i - 1) * slices + (j % slices)
    tris = []
    for j in range(slices):
        tris.append((0, ring(1, j + 1), ring(1, j)))
    for i in range(1, stacks - 1):
        for j in range(slices):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            tris.append((a, b, c))
            tris.append((b, d, c))
    for j in range(slices):
        tris.append((south, ring(stacks - 1, j), ring(stacks - 1, j + 1)))
    return np.asarray(dirs, dtype=np.float64), np.asarray(tris, dtype=np.int64)


def soup(n_spheres: int = 512, stacks: int = 32, slices: int = 32, seed: int = 1):
    """``soup_1m`` (n_spheres=512) / ``soup_10m`` (5120): returns (P f32 [nv,3], idx u32 [nt,3]).

    Centres uniform in [-1,1]^3, radii uniform in [0.03,0.09]; PCG32(seed, stream=1), draws in order
    cx,cy,cz,r per sphere; vertices computed in f64 and rounded to f32; then a 2-triangle ground
    quad at y=-1.1 with half-size 3."""
    rng = PCG32(seed, 1)
    u = rng.uniform(4 * n_spheres).reshape(n_spheres, 4)
    centres = -1.0 + 2.0 * u[:, :3]
    radii = 0.03 + 0.06 * u[:, 3]
    dirs, tris = _uv_sphere_template(stacks, slices)
    nv = dirs.shape[0]
    P = (centres[:, None, :] + radii[:, None, None] * dirs[None, :, :]).astype(np.float32).reshape(-1, 3)
    idx = (tris[None, :, :] + (np.arange(n_spheres, dtype=np.int64) * nv)[:, None, None]).reshape(-1, 3)
    base = P.shape[0]
    quad = np.array([[-3, -1.1, -3], [3, -1.1, -3], [3, -1.1, 3], [-3, -1.1, 3]], dtype=np.float32)
    qidx = np.array([[0, 2, 1], [0, 3, 2]], dtype=np.int64) + base  # normal +y
    P = np.ascontiguousarray(np.concatenate([P, quad], axis=0))
    idx = np.ascontiguousarray(np.concatenate([idx, qidx], axis=0).astype(np.uint32))
    return P, idx


def pack_rays(o, d, tmin=0.0, tmax=np.inf):
    """(origin, direction) f32 [n,3] -> the ABI's two float4 arrays (o|tmin, d|tmax)."""
    n = o.shape[0]
    ro = np.empty((n, 4), dtype=np.float32)
    rd = np.empty((n, 4), dtype=np.float32)
    ro[:, :3] = o
    ro[:, 3] = tmin
    rd[:, :3] = d
    rd[:, 3] = tmax
    return ro, rd


def coherent_rays(width: int = 4096, height: int = 2048, eye=(0.0, 0.0, -4.0), vfov_deg: float = 40.0):
    """Pinhole rays through pixel centres, row-major; eye looks at the origin, up +y (left-handed
    like the reference's LookAt, lib/core/transform.dart:306-331)."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = -eye / np.linalg.norm(eye)
    up = np.array([0.0, 1.0, 0.0])
    left = np.cross(up, fwd)
    left /= np.linalg.norm(left)
    nup = np.cross(fwd, left)
    th = math.tan(math.radians(vfov_deg) / 2.0)
    aspect = width / height
    sx = ((np.arange(width, dtype=np.float64) + 0.5) / width * 2.0 - 1.0) * th * aspect
    sy = (1.0 - (np.arange(height, dtype=np.float64) + 0.5) / height * 2.0) * th
    d = (fwd[None, None, :] + sx[None, :, None] * left[None, None, :] + sy[:, None, None] * nup[None, None, :])
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    d = d.reshape(-1, 3).astype(np.float32)
    o = np.broadcast_to(eye.astype(np.float32), d.shape)
    return pack_rays(o, d)


def incoherent_rays(n: int = 8_388_608, radius: float = 2.5, seed_origin: int = 2, seed_target: int = 3):
    """Origin uniform on the sphere of `radius` (PCG32 seed 2: z=1-2u1, phi=2pi*u2); direction =
    normalise(uniform point in the unit ball (PCG32 seed 3, rejection on x^2+y^2+z^2<1) - origin)."""
    u = PCG32(seed_origin, 1).uniform(2 * n).reshape(n, 2)
    z = 1.0 - 2.0 * u[:, 0]
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = 2.0 * math.pi * u[:, 1]
    o = radius * np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    rng = PCG32(seed_target, 1)
    tgt = np.empty((n, 3), dtype=np.float64)
    got = 0
    while got < n:
        m = max(1024, int((n - got) * 2.0) + 1024)
        c = 2.0 * rng.uniform(3 * m).reshape(m, 3) - 1.0
        c = c[(c * c).sum(axis=1) < 1.0]
        k = min(n - got, c.shape[0])
        tgt[got:got + k] = c[:k]
        got += k
    o32 = o.astype(np.float32)
    d = tgt - o32.astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return pack_rays(o32, d.astype(np.float32))


def cornell_synth(overrides: dict | None = None):
    """the 22 triangles + sphere of web/scenes/cornell-path.pbrt:23-59 (same
    transforms and Kd values) with the disk light (:15-19) replaced by a 2-triangle quad light,
    4.24 x 4.24 at y = 9.9 facing -y, L = (36, 36, 36), nsamples 1.  Returns (SceneBuilder, camera).
    `overrides` maps "grey" / "red" / "green" / "box" / "sphere" to BxDF lists (host.mirror_lobes(...), ...) for the
    material tests; None keeps the file's matte materials."""
    from . import host

    sb = host.SceneBuilder()
    ov = overrides or {}

    def mat(name, kd):
        return sb.material_lobes(ov[name]) if name in ov else sb.material(kd)

    grey = mat("grey", (0.75, 0.75, 0.75))
    red = mat("red", (0.48, 0.1125, 0.075))
    green = mat("green", (0.1125, 0.375, 0.1125))
    box = mat("box", (0.48, 0.48, 0.48))
    ball = mat("sphere", (0.48, 0.48, 0.48)) if "sphere" in ov else box
    h = 2.12
    sb.mesh([[-h, 9.9, -h], [h, 9.9, -h], [h, 9.9, h], [-h, 9.9, h]], [[0, 1, 2], [0, 2, 3]], material=grey,
            area_light=(36.0, 36.0, 36.0), nsamples=1)
    quad = [[0, 1, 2], [0, 2, 3]]
    walls = [
        (grey, [10, -10, -10, -10, -10, -10, -10, -10, 10, 10, -10, 10]),
        (grey, [10, 10, -10, 10, 10, 10, -10, 10, 10, -10, 10, -10]),
        (grey, [10, -10, 10, -10, -10, 10, -10, 10, 10, 10, 10, 10]),
        (red, [-10, -10, 10, -10, -10, -10, -10, 10, -10, -10, 10, 10]),
        (green, [10, -10, -10, 10, -10, 10, 10, 10, 10, 10, 10, -10]),
    ]
    for m, p in walls:
        sb.mesh(np.asarray(p, np.float32).reshape(4, 3), quad, material=m)
    # short box: Translate 4 -7 4, Scale 0.3 0.4 0.3, Rotate 30 0 1 0 (CTM post-multiplies)
    o2w = host.mat_mul(host.mat_mul(host.translate(4, -7, 4), host.scale(0.3, 0.4, 0.3)), host.rotate(30, (0, 1, 0)))
    rquad = [[0, 2, 1], [0, 3, 2]]
    faces = [
        (rquad, [10, -10, -10, -10, -10, -10, -10, -10, 10, 10, -10, 10]),
        (rquad, [10, 10, -10, 10, 10, 10, -10, 10, 10, -10, 10, -10]),
        (rquad, [10, -10, 10, -10, -10, 10, -10, 10, 10, 10, 10, 10]),
        (rquad, [-10, -10, 10, -10, -10, -10, -10, 10, -10, -10, 10, 10]),
        (rquad, [10, -10, -10, 10, -10, 10, 10, 10, 10, 10, 10, -10]),
        (quad, [10, -10, -10, -10, -10, -10, -10, 10, -10, 10, 10, -10]),
    ]
    for q, p in faces:
        sb.mesh(np.asarray(p, np.float32).reshape(4, 3), q, material=box, o2w=o2w)
    sb.sphere(host.translate(-4, -4, 0), radius=3.0, material=ball)
    cam = host.PerspectiveCamera(host.look_at((0, 0, -35), (0, 0, 0), (0, 1, 0)), fov=35.0)
    return sb, cam


def cornell_materials():
    """cornell_synth with one material of each family the BxDF-list path covers: plastic back wall
    and ceiling, glass sphere, mirror-ish uber box, metal floor stays matte grey, OrenNayar red wall."""
    from . import host

    return cornell_synth({
        "grey": host.plastic_lobes((0.6, 0.6, 0.55), 0.3, 0.08),
        "red": host.matte_lobes((0.48, 0.1125, 0.075), 30.0),
        "green": host.metal_lobes((0.2, 0.92, 1.1), (3.9, 2.45, 2.14), 0.1),
        "box": host.uber_lobes(kd=(0.3, 0.2, 0.2), ks=0.3, kr=0.3, kt=0.0, roughness=0.15, index=1.33, opacity=(0.9, 0.8, 0.9)),
        "sphere": host.glass_lobes(1.0, 1.0, 1.5),
    })


def cornell_path():
    """web/scenes/cornell-path.pbrt:1-61 as shipped: disk area light (radius 3 at y = 9.9, rotated 90 degrees about x,
    L = 36, nsamples 1, default matte Kd 0.5), five walls, the short box, the sphere.  Returns (SceneBuilder, camera);
    the file's film is 320x240 with lowdiscrepancy 16 spp, BASELINE.json configs[0] overrides it to 256x256, 4 spp."""
    from . import host

    sb = host.SceneBuilder()
    default = sb.material((0.5, 0.5, 0.5))  # matte_material.dart Create: Kd default 0.5
    grey = sb.material((0.75, 0.75, 0.75))
    red = sb.material((0.48, 0.1125, 0.075))
    green = sb.material((0.1125, 0.375, 0.1125))
    box = sb.material((0.48, 0.48, 0.48))
    sb.disk(host.mat_mul(host.translate(0, 9.9, 0), host.rotate(90, (1, 0, 0))), radius=3.0, material=default,
            area_light=(36.0, 36.0, 36.0), nsamples=1)
    quad = [[0, 1, 2], [0, 2, 3]]
    walls = [
        (grey, [10, -10, -10, -10, -10, -10, -10, -10, 10, 10, -10, 10]),
        (grey, [10, 10, -10, 10, 10, 10, -10, 10, 10, -10, 10, -10]),
        (grey, [10, -10, 10, -10, -10, 10, -10, 10, 10, 10, 10, 10]),
        (red, [-10, -10, 10, -10, -10, -10, -10, 10, -10, -10, 10, 10]),
        (green, [10, -10, -10, 10, -10, 10, 10, 10, 10, 10, 10, -10]),
    ]
    for m, p in walls:
        sb.mesh(np.asarray(p, np.float32).reshape(4, 3), quad, material=m)
    o2w = host.mat_mul(host.mat_mul(host.translate(4, -7, 4), host.scale(0.3, 0.4, 0.3)), host.rotate(30, (0, 1, 0)))
    rquad = [[0, 2, 1], [0, 3, 2]]
    faces = [
        (rquad, [10, -10, -10, -10, -10, -10, -10, -10, 10, 10, -10, 10]),
        (rquad, [10, 10, -10, 10, 10, 10, -10, 10, 10, -10, 10, -10]),
        (rquad, [10, -10, 10, -10, -10, 10, -10, 10, 10, 10, 10, 10]),
        (rquad, [-10, -10, 10, -10, -10, -10, -10, 10, -10, -10, 10, 10]),
        (rquad, [10, -10, -10, 10, -10, 10, 10, 10, 10, 10, 10, -10]),
        (quad, [10, -10, -10, -10, -10, -10, -10, 10, -10, 10, 10, -10]),
    ]
    for q, p in faces:
        sb.mesh(np.asarray(p, np.float32).reshape(4, 3), q, material=box, o2w=o2w)
    sb.sphere(host.translate(-4, -4, 0), radius=3.0, material=box)
    cam = host.PerspectiveCamera(host.look_at((0, 0, -35), (0, 0, 0), (0, 1, 0)), fov=35.0)
    return sb, cam


def soup_render_scene(n_spheres: int = 512):
    """`soup` as a renderable scene (config 3 / 5): grey matte, one quad light above, config-2 camera."""
    from . import host

    P, idx = soup(n_spheres)
    sb = host.SceneBuilder()
    grey = sb.material((0.6, 0.6, 0.6))
    sb.mesh(P, idx, material=grey)
    sb.mesh([[-1.5, 2.5, -1.5], [1.5, 2.5, -1.5], [1.5, 2.5, 1.5], [-1.5, 2.5, 1.5]], [[0, 1, 2], [0, 2, 3]], material=grey,
            area_light=(20.0, 20.0, 20.0), nsamples=1)
    cam = host.PerspectiveCamera(host.look_at((0, 0, -4), (0, 0, 0), (0, 1, 0)), fov=40.0)
    return sb, cam


def rays_hash(ro: np.ndarray, rd: np.ndarray) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(ro).tobytes())
    h.update(np.ascontiguousarray(rd).tobytes())
    return h.hexdigest()[:16]
