"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")

HIT_DTYPE = np.dtype([("t", np.float32), ("b1", np.float32), ("b2", np.float32), ("prim", np.int32)])


def build_oracle(force: bool = False) -> str:
    src_dir = os.path.join(_ROOT, "oracle")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cpp", ".h"))]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", src_dir], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_SO)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.orc_create.restype = vp
        L.orc_destroy.argtypes = [vp]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [vp]
        L.orc_set_triangles.argtypes = [vp, vp, u32, vp, u32, vp, vp, vp]
        L.orc_set_spheres.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp]
        L.orc_set_build_order.argtypes = [vp, vp, u32]
        L.orc_build_bvh.argtypes = [vp, i32, i32]
        L.orc_build_seconds.restype = C.c_double
        L.orc_build_seconds.argtypes = [vp]
        L.orc_bvh_num_nodes.restype = u32
        L.orc_bvh_num_nodes.argtypes = [vp]
        L.orc_num_prims.restype = u32
        L.orc_num_prims.argtypes = [vp]
        L.orc_bvh_export.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_trace_closest.argtypes = [vp, vp, vp, u64, vp, vp, i32]
        L.orc_trace_any.argtypes = [vp, vp, vp, u64, vp, i32]
        L.orc_trace_closest_brute.argtypes = [vp, vp, vp, u64, vp, vp, vp, i32]
        L.orc_trace_any_brute.argtypes = [vp, vp, vp, u64, vp, i32]
        L.orc_get_counters.argtypes = [vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class Oracle:
    """CPU restatement of the reference path, same call sequence as dartray_b200.Context."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.orc_create()

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error(self.h).decode())

    def set_triangles(self, P, idx, material=None, light=None, reverse=None):
        P = _arr(P, np.float32).reshape(-1, 3)
        idx = _arr(idx, np.uint32).reshape(-1, 3)
        m, l, r = _arr(material, np.int32), _arr(light, np.int32), _arr(reverse, np.uint8)
        self._ck(self.L.orc_set_triangles(self.h, _p(P), P.shape[0], _p(idx), idx.shape[0], _p(m), _p(l), _p(r)))

    def set_spheres(self, o2w, w2o, params, material=None, light=None, reverse=None):
        o2w = _arr(o2w, np.float32).reshape(-1, 16)
        w2o = _arr(w2o, np.float32).reshape(-1, 16)
        params = _arr(params, np.float64).reshape(-1, 4)
        m, l, r = _arr(material, np.int32), _arr(light, np.int32), _arr(reverse, np.uint8)
        self._ck(self.L.orc_set_spheres(self.h, o2w.shape[0], _p(o2w), _p(w2o), _p(params), _p(m), _p(l), _p(r)))

    def set_build_order(self, order):
        if order is None:
            self._ck(self.L.orc_set_build_order(self.h, None, 0))
        else:
            o = _arr(order, np.uint32)
            self._ck(self.L.orc_set_build_order(self.h, _p(o), o.shape[0]))

    def build_bvh(self, split=2, max_node_prims=4):
        self._ck(self.L.orc_build_bvh(self.h, split, max_node_prims))

    @property
    def build_seconds(self):
        return self.L.orc_build_seconds(self.h)

    def bvh_export(self):
        n = self.L.orc_bvh_num_nodes(self.h)
        npr = self.L.orc_num_prims(self.h)
        bounds = np.empty((n, 6), np.float32)
        offset = np.empty(n, np.int32)
        nprims = np.empty(n, np.int32)
        axis = np.empty(n, np.int32)
        ordered = np.empty(npr, np.uint32)
        self._ck(self.L.orc_bvh_export(self.h, _p(bounds), _p(offset), _p(nprims), _p(axis), _p(ordered)))
        return dict(bounds=bounds, offset=offset, n_primitives=nprims, axis=axis, ordered=ordered)

    def trace_closest(self, ro, rd, nthreads=1, want_t64=False):
        ro, rd = _arr(ro, np.float32), _arr(rd, np.float32)
        n = ro.shape[0]
        hits = np.empty(n, HIT_DTYPE)
        t64 = np.empty(n, np.float64) if want_t64 else None
        self._ck(self.L.orc_trace_closest(self.h, _p(ro), _p(rd), n, _p(hits), _p(t64), nthreads))
        return (hits, t64) if want_t64 else hits

    def trace_any(self, ro, rd, nthreads=1):
        ro, rd = _arr(ro, np.float32), _arr(rd, np.float32)
        n = ro.shape[0]
        occ = np.empty(n, np.uint8)
        self._ck(self.L.orc_trace_any(self.h, _p(ro), _p(rd), n, _p(occ), nthreads))
        return occ

    def trace_closest_brute(self, ro, rd, nthreads=8):
        ro, rd = _arr(ro, np.float32), _arr(rd, np.float32)
        n = ro.shape[0]
        hits = np.empty(n, HIT_DTYPE)
        nties = np.empty(n, np.int32)
        second = np.empty(n, np.float64)
        self._ck(self.L.orc_trace_closest_brute(self.h, _p(ro), _p(rd), n, _p(hits), _p(nties), _p(second), nthreads))
        return hits, nties, second

    def trace_any_brute(self, ro, rd, nthreads=8):
        ro, rd = _arr(ro, np.float32), _arr(rd, np.float32)
        n = ro.shape[0]
        occ = np.empty(n, np.uint8)
        self._ck(self.L.orc_trace_any_brute(self.h, _p(ro), _p(rd), n, _p(occ), nthreads))
        return occ

    def counters(self):
        out = np.zeros(3, np.uint64)
        self.L.orc_get_counters(self.h, _p(out))
        return dict(rays=int(out[0]), nodes_visited=int(out[1]), prims_tested=int(out[2]))
