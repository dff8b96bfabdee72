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
