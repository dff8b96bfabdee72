"""Host BVH builder of libdartray_gpu (csrc/bvh_builder.cpp) == the oracle's restatement of
BVHAccel's constructor (bvh_accel.dart:41-91, 228-437): node for node, box for box, leaf order."""
import numpy as np
import pytest

from dartray_b200 import capi, scenes
from tests.oracle_lib import Oracle
from tests.util import mesh_refine_order, random_soup, translate


def both(P, idx, spheres=None, order=None, split=2, maxprims=4):
    o = Oracle()
    c = capi.Context(capi.DEVICE_NONE)
    for x in (o, c):
        x.set_triangles(P, idx)
        if spheres is not None:
            x.set_spheres(*spheres)
        x.set_build_order(order)
        x.build_bvh(split, maxprims)
    return o.bvh_export(), c.bvh_export(), c


def assert_same(a, b):
    for k in ("bounds", "offset", "n_primitives", "axis", "ordered"):
        assert a[k].shape == b[k].shape, k
        assert (a[k] == b[k]).all(), k


@pytest.mark.parametrize("split", [0, 1, 2])
@pytest.mark.parametrize("maxprims", [1, 4, 16])
def test_random_soup_topology(drt_lib, split, maxprims):
    P, idx = random_soup(3000, seed=7)
    a, b, _ = both(P, idx, split=split, maxprims=maxprims)
    assert_same(a, b)


def test_mixed_scene_with_refine_order(drt_lib):
    P, idx = random_soup(500, seed=3)
    m0, i0 = translate(0.3, 0.1, -0.2)
    m1, i1 = translate(-0.4, 0.2, 0.5)
    sph = (np.stack([m0, m1]), np.stack([i0, i1]), [[0.25, -0.25, 0.25, 360.0], [0.4, -0.1, 0.3, 200.0]])
    order = mesh_refine_order([200, 300], n_spheres_after=2)
    a, b, _ = both(P, idx, sph, order)
    assert_same(a, b)
    # the build order matters for the result only through ties; the id set is preserved
    assert sorted(b["ordered"].tolist()) == list(range(502))


def test_degenerate_centroids_make_big_leaves(drt_lib):
    """All centroids equal -> one leaf (bvh_accel.dart:265-274), count beyond the 4-bit inline field."""
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    P = np.tile(tri, (40, 1))
    idx = np.arange(120, dtype=np.uint32).reshape(-1, 3)
    a, b, c = both(P, idx)
    assert_same(a, b)
    assert c.bvh_info()["max_leaf_prims"] == 40 and c.bvh_info()["n_nodes"] == 1


def test_equal_centroid_ties_follow_insertion_sort(drt_lib):
    """Ranges of <= 4 primitives are split by Dart's insertion sort with a never-equal comparator
    (common.dart:289-297): equal keys end up in reverse order."""
    P, idx = [], []
    for k, x in enumerate([0.0, 0.0, 1.0, 1.0]):  # pairs share the x centroid, differ in y extent
        P += [[x - 0.1, k, 0], [x + 0.1, k, 0], [x, k + 0.5, 0.1 * k]]
        idx.append([3 * k, 3 * k + 1, 3 * k + 2])
    a, b, _ = both(np.array(P, np.float32), np.array(idx, np.uint32))
    assert_same(a, b)


def test_soup_scene_topology(drt_lib):
    P, idx = scenes.soup(32)
    order = mesh_refine_order([idx.shape[0]])
    a, b, c = both(P, idx, order=order)
    assert_same(a, b)
    info = c.bvh_info()
    assert info["n_prims"] == idx.shape[0] and info["max_depth"] < 64
