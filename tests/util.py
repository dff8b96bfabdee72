"""Small scene builders shared by the tests."""
import numpy as np

from dartray_b200 import scenes


def random_soup(n_tris, seed, extent=1.0, size=0.2):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n_tris, 1, 3))
    P = (c + rng.uniform(-size, size, (n_tris, 3, 3))).astype(np.float32).reshape(-1, 3)
    idx = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    return P, idx


def random_rays(n, seed, radius=3.0, tmin=0.0, tmax=np.inf):
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n, 3))
    o = radius * o / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = rng.uniform(-1, 1, (n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return scenes.pack_rays(o.astype(np.float32), d.astype(np.float32), tmin, tmax)


def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    mi = np.eye(4, dtype=np.float32)
    mi[:3, 3] = (-x, -y, -z)
    return m.reshape(16), mi.reshape(16)


def mesh_refine_order(mesh_tri_counts, n_spheres_after=0):
    """Primitive.fullyRefine is LIFO (primitive.dart:71-84): each mesh's triangles come out reversed."""
    order, base = [], 0
    for n in mesh_tri_counts:
        order.extend(range(base + n - 1, base - 1, -1))
        base += n
    order.extend(range(base, base + n_spheres_after))
    return np.asarray(order, dtype=np.uint32)


def uv_sphere_mesh(stacks=12, slices=16, radius=1.0):
    """A latitude/longitude sphere with per-vertex normals N (object space), tangents S and uvs: what a smooth-shaded
    "trianglemesh" of the shipped scenes carries (triangle_mesh.dart:88-140)."""
    P, N, S, UV, idx = [], [], [], [], []
    for i in range(stacks + 1):
        th = np.pi * i / stacks
        for j in range(slices + 1):
            ph = 2 * np.pi * j / slices
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            P.append(radius * n)
            N.append(n)
            S.append([-np.sin(ph), 0.0, np.cos(ph)])
            UV.append([j / slices, i / stacks])
    w = slices + 1
    for i in range(stacks):
        for j in range(slices):
            a, b, c, d = i * w + j, i * w + j + 1, (i + 1) * w + j, (i + 1) * w + j + 1
            if i > 0:
                idx.append([a, b, c])
            if i < stacks - 1:
                idx.append([b, d, c])
    return (np.asarray(P, np.float32), np.asarray(idx, np.uint32), np.asarray(N, np.float32), np.asarray(S, np.float32),
            np.asarray(UV, np.float32))


def synthetic_sample_table(seed=7):
    """A stand-in for BestCandidateSampler's _SAMPLE_TABLE (best_candidate_sampler.dart:163-4258): 4096 x 5 doubles, image
    positions from a jittered 64 x 64 grid in shuffled order, time / lens uniform.  The reference's own table is data of the
    reference and is handed over by the caller (drt_set_sample_table); the tests only need its shape."""
    rng = np.random.default_rng(seed)
    g = (np.stack(np.meshgrid(np.arange(64), np.arange(64), indexing="ij"), -1).reshape(-1, 2) + rng.uniform(0.05, 0.95, (4096, 2))) / 64.0
    rng.shuffle(g)
    return np.concatenate([g, rng.uniform(0, 1, (4096, 3))], axis=1).astype(np.float64)
