"""ctypes binding of libdartray_gpu.so (include/drt.h) — the same symbols a dart:ffi binding looks up.

The library is the only execution path: if it is missing or no CUDA device is present the calls
raise; nothing here falls back to a CPU implementation."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
# DRT_LIB_PATH: an A/B build of the same library (tools/trace_ab.sh); the default is the in-tree build
LIB_PATH = os.environ.get("DRT_LIB_PATH") or os.path.join(PKG, "libdartray_gpu.so")

DEVICE_NONE = -1
SPLIT_MIDDLE, SPLIT_EQUAL_COUNTS, SPLIT_SAH = 0, 1, 2
KERNEL_FAST, KERNEL_EXACT_WALK, KERNEL_FAST_V1, KERNEL_FAST_Q = 0, 1, 2, 3

HIT_DTYPE = np.dtype([("t", np.float32), ("b1", np.float32), ("b2", np.float32), ("prim", np.int32)])

# every symbol include/drt.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "drt_version", "drt_create", "drt_create_multi", "drt_device_count", "drt_destroy", "drt_last_error", "drt_set_triangles", "drt_set_spheres", "drt_set_disks", "drt_set_quadrics", "drt_set_mesh_shading", "drt_set_infinite_light", "drt_set_lobe_wrappers", "drt_set_light_map", "drt_set_sample_table",
    "drt_set_instances", "drt_set_ray_times", "drt_set_build_order", "drt_build_bvh", "drt_bvh_info_get", "drt_bvh_export", "drt_trace_closest",
    "drt_trace_any", "drt_trace_closest_device", "drt_trace_any_device", "drt_set_counting", "drt_get_counters", "drt_set_kernel_variant",
    "drt_last_kernel_ms", "drt_kernel_launches",
    "drt_set_materials", "drt_set_material_lobes", "drt_set_measured", "drt_set_textures", "drt_set_material_programs", "drt_set_lights", "drt_set_spot_params", "drt_set_volumes", "drt_set_volume_integrator", "drt_set_shading_precision", "drt_set_camera", "drt_set_camera_motion", "drt_set_camera_kind", "drt_set_film", "drt_set_sampler", "drt_set_integrator",
    "drt_render", "drt_render_shard", "drt_set_batch_slots", "drt_film_clear", "drt_film_size", "drt_film_read",
    "drt_film_device", "drt_pixel_samples", "drt_render_stats_get", "drt_set_render_profiling", "drt_render_profile_get",
]


class DrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdartray_gpu error {code}: {msg}")
        self.code = code


class BvhInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_prims", C.c_uint32), ("n_leaves", C.c_uint32),
                ("max_leaf_prims", C.c_uint32), ("max_depth", C.c_uint32), ("device_bytes", C.c_uint64),
                ("build_seconds", C.c_double)]


class Counters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64),
                ("hits", C.c_uint64)]


class RenderProfile(C.Structure):
    _fields_ = [("ms", C.c_double * 7), ("launches", C.c_uint64 * 7), ("closest", Counters), ("any", Counters)]


PROFILE_TIME, PROFILE_WORK = 1, 2
PRECISION_F64, PRECISION_F32 = 0, 1
PROFILE_CLASSES = ["trace_closest", "trace_any", "integrator", "sampler", "resolve", "film", "other"]


class RenderStats(C.Structure):
    _fields_ = [("camera_samples", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("zeroed_samples", C.c_uint64)]


_lib = None


def load():
    """Load libdartray_gpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(dartray_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.drt_version.restype = i32
    L.drt_create.restype = vp
    L.drt_create.argtypes = [i32]
    L.drt_create_multi.restype = vp
    L.drt_create_multi.argtypes = [vp, i32]
    L.drt_device_count.argtypes = [vp]
    L.drt_destroy.argtypes = [vp]
    L.drt_last_error.restype = C.c_char_p
    L.drt_last_error.argtypes = [vp]
    L.drt_set_triangles.argtypes = [vp, vp, u32, vp, u32, vp, vp, vp]
    L.drt_set_spheres.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp]
    L.drt_set_disks.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp]
    L.drt_set_quadrics.argtypes = [vp, i32, u32, vp, vp, vp, vp, vp, vp]
    L.drt_set_mesh_shading.argtypes = [vp, vp, vp, vp, vp, u32, vp, vp, vp]
    L.drt_set_build_order.argtypes = [vp, vp, u32]
    L.drt_build_bvh.argtypes = [vp, i32, i32]
    L.drt_bvh_info_get.argtypes = [vp, C.POINTER(BvhInfo)]
    L.drt_bvh_export.argtypes = [vp, vp, vp, vp, vp, vp]
    L.drt_trace_closest.argtypes = [vp, vp, vp, u64, vp]
    L.drt_trace_any.argtypes = [vp, vp, vp, u64, vp]
    L.drt_trace_closest_device.argtypes = [vp, vp, vp, u64, vp, vp]
    L.drt_trace_any_device.argtypes = [vp, vp, vp, u64, vp, vp]
    L.drt_set_counting.argtypes = [vp, i32]
    L.drt_set_kernel_variant.argtypes = [vp, i32]
    L.drt_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.drt_last_kernel_ms.restype = C.c_double
    L.drt_last_kernel_ms.argtypes = [vp]
    L.drt_kernel_launches.restype = u64
    L.drt_kernel_launches.argtypes = [vp]
    dbl = C.c_double
    L.drt_set_materials.argtypes = [vp, u32, vp, vp, vp]
    L.drt_set_material_lobes.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp, vp]
    L.drt_set_spot_params.argtypes = [vp, u32, vp, vp]
    L.drt_set_sample_table.argtypes = [vp, vp, u32]
    L.drt_set_light_map.argtypes = [vp, u32, i32, i32, vp, vp, vp, vp, C.c_double]
    L.drt_set_lobe_wrappers.argtypes = [vp, u32, vp, vp]
    L.drt_set_textures.argtypes = [vp, u32, vp, vp, u64]
    L.drt_set_measured.argtypes = [vp, u32, vp, vp, vp, vp, u64]
    L.drt_set_instances.argtypes = [vp, u32, vp, vp, vp, vp, u32, vp, vp, vp, vp, vp, vp]
    L.drt_set_ray_times.argtypes = [vp, vp, u64]
    L.drt_set_material_programs.argtypes = [vp, u32, vp]
    L.drt_set_infinite_light.argtypes = [vp, u32, i32, i32, vp, vp, vp]
    L.drt_set_lights.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp]
    L.drt_set_camera.argtypes = [vp, vp, vp, dbl, dbl, dbl, dbl]
    L.drt_set_camera_kind.argtypes = [vp, i32]
    L.drt_set_camera_motion.argtypes = [vp, vp, C.c_double, C.c_double]
    L.drt_set_film.argtypes = [vp, i32, i32, vp, dbl, dbl, vp]
    L.drt_set_sampler.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, u64]
    L.drt_set_integrator.argtypes = [vp, i32, i32, i32, i32, dbl, dbl]
    L.drt_render.argtypes = [vp, i32, i32]
    L.drt_render_shard.argtypes = [vp, i32, i32]
    L.drt_set_batch_slots.argtypes = [vp, u64]
    L.drt_film_clear.argtypes = [vp]
    L.drt_film_size.argtypes = [vp, vp]
    L.drt_film_read.argtypes = [vp, vp, vp, vp]
    L.drt_film_device.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.drt_pixel_samples.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.drt_render_stats_get.argtypes = [vp, C.POINTER(RenderStats)]
    L.drt_set_volumes.argtypes = [vp, u32] + [vp] * 13
    L.drt_set_volume_integrator.argtypes = [vp, i32, dbl]
    L.drt_set_shading_precision.argtypes = [vp, i32]
    L.drt_set_render_profiling.argtypes = [vp, i32]
    L.drt_render_profile_get.argtypes = [vp, C.POINTER(RenderProfile)]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One drt_ctx: a scene resident on one CUDA device, or — `device` a list of ids — on several (drt_create_multi:
    one BVH build, renders split over the devices, films summed into the first device's)."""

    def __init__(self, device=0):
        self.L = load()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*device)
            self.h = self.L.drt_create_multi(ids, len(device))
            self.devices = list(device)
            device = device[0] if device else -1
        else:
            self.h = self.L.drt_create(device)
            self.devices = [device]
        if not self.h:
            raise DrtError(-4, self.L.drt_last_error(None).decode())
        self.device = device

    @property
    def device_count(self) -> int:
        return int(self.L.drt_device_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.drt_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise DrtError(rc, self.L.drt_last_error(self.h).decode())

    # -- scene ---------------------------------------------------------------------------------
    def set_triangles(self, P, idx, material=None, light=None, reverse=None):
        P = _arr(P, np.float32).reshape(-1, 3)
        idx = _arr(idx, np.uint32).reshape(-1, 3)
        m, l, r = _arr(material, np.int32), _arr(light, np.int32), _arr(reverse, np.uint8)
        self._ck(self.L.drt_set_triangles(self.h, _p(P), P.shape[0], _p(idx), idx.shape[0], _p(m), _p(l), _p(r)))

    def set_spheres(self, o2w, w2o, params, material=None, light=None, reverse=None):
        o2w = _arr(o2w, np.float32).reshape(-1, 16)
        w2o = _arr(w2o, np.float32).reshape(-1, 16)
        params = _arr(params, np.float64).reshape(-1, 4)
        m, l, r = _arr(material, np.int32), _arr(light, np.int32), _arr(reverse, np.uint8)
        self._ck(self.L.drt_set_spheres(self.h, o2w.shape[0], _p(o2w), _p(w2o), _p(params), _p(m), _p(l), _p(r)))

    def set_disks(self, o2w, w2o, params, material=None, light=None, reverse=None):
        """params: n x 4 (height, radius, innerradius, phimax degrees); call after set_spheres."""
        o2w = _arr(o2w, np.float32).reshape(-1, 16)
        w2o = _arr(w2o, np.float32).reshape(-1, 16)
        params = _arr(params, np.float64).reshape(-1, 4)
        m, l, r = _arr(material, np.int32), _arr(light, np.int32), _arr(reverse, np.uint8)
        self._ck(self.L.drt_set_disks(self.h, o2w.shape[0], _p(o2w), _p(w2o), _p(params), _p(m), _p(l), _p(r)))

    def set_quadrics(self, kind, o2w, w2o, params, material=None, light=None, reverse=None):
        """kind 2 cylinder / 3 cone / 4 paraboloid / 5 hyperboloid; params: n x 8 doubles (include/drt.h); appended to the
        quadric id range in call order, after set_spheres / set_disks."""
        o2w = _arr(o2w, np.float32).reshape(-1, 16)
        w2o = _arr(w2o, np.float32).reshape(-1, 16)
        params = _arr(params, np.float64).reshape(-1, 8)
        m, l, r = _arr(material, np.int32), _arr(light, np.int32), _arr(reverse, np.uint8)
        self._ck(self.L.drt_set_quadrics(self.h, int(kind), o2w.shape[0], _p(o2w), _p(w2o), _p(params), _p(m), _p(l), _p(r)))

    def set_mesh_shading(self, N, S, uv, mesh_of_tri, o2w, w2o, flags):
        """Per-vertex N / S (object space) / uv (any may be None), mesh index per triangle, per-mesh transforms and flags
        (bit 0 N, 1 S, 2 uv): what Triangle.getShadingGeometry reads (include/drt.h)."""
        N, S, uv = _arr(N, np.float32), _arr(S, np.float32), _arr(uv, np.float32)
        mot, flags = _arr(mesh_of_tri, np.uint32), _arr(flags, np.uint8)
        o2w, w2o = _arr(o2w, np.float32).reshape(-1, 16), _arr(w2o, np.float32).reshape(-1, 16)
        self._ck(self.L.drt_set_mesh_shading(self.h, _p(N), _p(S), _p(uv), _p(mot), o2w.shape[0], _p(o2w), _p(w2o), _p(flags)))

    def set_build_order(self, order):
        if order is None:
            self._ck(self.L.drt_set_build_order(self.h, None, 0))
        else:
            o = _arr(order, np.uint32)
            self._ck(self.L.drt_set_build_order(self.h, _p(o), o.shape[0]))

    def build_bvh(self, split: int = SPLIT_SAH, max_node_prims: int = 4):
        self._ck(self.L.drt_build_bvh(self.h, split, max_node_prims))

    def bvh_info(self) -> dict:
        info = BvhInfo()
        self._ck(self.L.drt_bvh_info_get(self.h, C.byref(info)))
        return {k: getattr(info, k) for k, _ in BvhInfo._fields_}

    def bvh_export(self) -> dict:
        info = self.bvh_info()
        n, npr = info["n_nodes"], info["n_prims"]
        bounds = np.empty((n, 6), np.float32)
        offset = np.empty(n, np.int32)
        nprims = np.empty(n, np.int32)
        axis = np.empty(n, np.int32)
        ordered = np.empty(npr, np.uint32)
        self._ck(self.L.drt_bvh_export(self.h, _p(bounds), _p(offset), _p(nprims), _p(axis), _p(ordered)))
        return dict(bounds=bounds, offset=offset, n_primitives=nprims, axis=axis, ordered=ordered)

    # -- queries (host buffers) ----------------------------------------------------------------
    def trace_closest(self, ro, rd, out=None):
        ro, rd = _arr(ro, np.float32), _arr(rd, np.float32)
        n = ro.shape[0]
        hits = out if out is not None else np.empty(n, HIT_DTYPE)
        self._ck(self.L.drt_trace_closest(self.h, _p(ro), _p(rd), n, _p(hits)))
        return hits

    def trace_any(self, ro, rd, out=None):
        ro, rd = _arr(ro, np.float32), _arr(rd, np.float32)
        n = ro.shape[0]
        occ = out if out is not None else np.empty(n, np.uint8)
        self._ck(self.L.drt_trace_any(self.h, _p(ro), _p(rd), n, _p(occ)))
        return occ

    # -- queries (device pointers, e.g. torch tensors' data_ptr()) ------------------------------
    def trace_closest_device(self, d_ro: int, d_rd: int, n: int, d_hits: int, stream: int = 0):
        self._ck(self.L.drt_trace_closest_device(self.h, d_ro, d_rd, n, d_hits, stream))

    def trace_any_device(self, d_ro: int, d_rd: int, n: int, d_occ: int, stream: int = 0):
        self._ck(self.L.drt_trace_any_device(self.h, d_ro, d_rd, n, d_occ, stream))

    def set_kernel_variant(self, variant: int):
        """0 = FAST (default), 1 = EXACT_WALK (f64 slab test at every node)."""
        self._ck(self.L.drt_set_kernel_variant(self.h, variant))

    def set_counting(self, enabled: bool):
        self._ck(self.L.drt_set_counting(self.h, 1 if enabled else 0))

    def counters(self) -> dict:
        c = Counters()
        self._ck(self.L.drt_get_counters(self.h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}

    @property
    def last_kernel_ms(self) -> float:
        return self.L.drt_last_kernel_ms(self.h)

    @property
    def kernel_launches(self) -> int:
        return int(self.L.drt_kernel_launches(self.h))

    # -- renderer: the same method names as the reference-side objects they replace ---------------
    def set_materials(self, kind, kd, sigma):
        kind, kd, sigma = _arr(kind, np.int32), _arr(kd, np.float32).reshape(-1, 3), _arr(sigma, np.float32)
        self._ck(self.L.drt_set_materials(self.h, kd.shape[0], _p(kind), _p(kd), _p(sigma)))

    def set_material_lobes(self, offsets, kind, rgb, fresnel, eta, k, scalars):
        """Materials as ordered BxDF lists (host.matte_lobes / mirror_lobes / glass_lobes / plastic_lobes / ...)."""
        offsets, kind, fresnel = _arr(offsets, np.uint32), _arr(kind, np.int32), _arr(fresnel, np.int32)
        rgb, eta, k = (_arr(v, np.float32).reshape(-1, 3) for v in (rgb, eta, k))
        scalars = _arr(scalars, np.float64).reshape(-1, 3)
        self._ck(self.L.drt_set_material_lobes(self.h, offsets.shape[0] - 1, _p(offsets), _p(kind), _p(rgb), _p(fresnel), _p(eta),
                                                _p(k), _p(scalars)))

    def set_lights(self, kind, L, pos, nsamples, shape_offsets, shape_prims):
        kind, L, pos = _arr(kind, np.int32), _arr(L, np.float32).reshape(-1, 3), _arr(pos, np.float32).reshape(-1, 3)
        ns, so, sp = _arr(nsamples, np.int32), _arr(shape_offsets, np.uint32), _arr(shape_prims, np.uint32)
        self._ck(self.L.drt_set_lights(self.h, kind.shape[0], _p(kind), _p(L), _p(pos), _p(ns), _p(so), _p(sp)))

    def set_spot_params(self, world_to_light, cosines):
        """worldToLight (n x 16) and (cosTotalWidth, cosFalloffStart) (n x 2) of the spot lights of the last set_lights."""
        w, cs = _arr(world_to_light, np.float32).reshape(-1, 16), _arr(cosines, np.float64).reshape(-1, 2)
        self._ck(self.L.drt_set_spot_params(self.h, w.shape[0], _p(w), _p(cs)))

    def set_light_map(self, index, texels, world_to_light, projection=None, screen=None, hither=1.0e-3):
        """Map (h x w x 3 float32, power-of-two, or None) and transforms of a projection (kind 5) / goniometric (kind 6) light."""
        t = _arr(texels, np.float32)
        w2l = _arr(world_to_light, np.float32).reshape(16)
        pr = _arr(projection, np.float32)
        sc = _arr(screen, np.float64)
        self._ck(self.L.drt_set_light_map(self.h, int(index), 0 if t is None else t.shape[1], 0 if t is None else t.shape[0], _p(t), _p(w2l),
                                          _p(pr), _p(sc), float(hither)))

    def set_lobe_wrappers(self, wrap, scale):
        """BRDFToBTDF (bit 0) / ScaledBxDF (bit 1, with its RGB scale) around the lobes of the last set_material_lobes."""
        w, sc = _arr(wrap, np.int32), _arr(scale, np.float32).reshape(-1, 3)
        self._ck(self.L.drt_set_lobe_wrappers(self.h, w.shape[0], _p(w), _p(sc)))

    def set_instances(self, object_offsets, object_prims, object_split, object_max_node_prims, instance_object, start_m, start_minv,
                      end_m, end_minv, times):
        """TransformedPrimitives (transformed_primitive.dart): objects (prim lists + nested accelerator parameters) and instances
        (object, world-to-primitive m / mInv at the start and end time, n x 2 times).  Before set_build_order / build_bvh."""
        oo, op = _arr(object_offsets, np.uint32), _arr(object_prims, np.uint32)
        os_, om = _arr(object_split, np.int32), _arr(object_max_node_prims, np.int32)
        io = _arr(instance_object, np.uint32)
        m = [_arr(x, np.float32).reshape(-1, 16) for x in (start_m, start_minv, end_m, end_minv)]
        tm = _arr(times, np.float64).reshape(-1, 2)
        self._ck(self.L.drt_set_instances(self.h, oo.shape[0] - 1, _p(oo), _p(op), _p(os_), _p(om), io.shape[0], _p(io), _p(m[0]), _p(m[1]),
                                          _p(m[2]), _p(m[3]), _p(tm)))

    def set_ray_times(self, times):
        """Ray i of the following trace_* calls travels at times[i] (None: every ray at time 0)."""
        t = _arr(times, np.float64)
        self._ck(self.L.drt_set_ray_times(self.h, _p(t), 0 if t is None else t.size))

    def set_measured(self, tables):
        """MeasuredMaterial data (measured_material.dart:76-205): list of (kind, array) — kind 0 = RegularHalfangleBRDF table
        (nThetaH x nThetaD x nPhiD x 3 float32), kind 1 = IrregIsotropicBRDFSamples (n x 6 float32: BRDFRemap point, RGB)."""
        kinds = np.array([k for k, _ in tables], np.int32)
        dims = np.zeros((len(tables), 3), np.int32)
        offs = np.zeros(len(tables), np.uint64)
        chunks, pos = [], 0
        for i, (k, a) in enumerate(tables):
            a = np.ascontiguousarray(a, np.float32)
            dims[i] = a.shape[:3] if k == 0 else (a.shape[0], 0, 0)
            offs[i] = pos
            pos += a.size
            chunks.append(a.ravel())
        data = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
        self._ck(self.L.drt_set_measured(self.h, len(tables), _p(kinds), _p(dims), _p(offs), _p(data), data.size))

    def set_textures(self, nodes, texels):
        """Texture nodes (host.TEX_DTYPE records = drt_texture) and the level-0 texels of their images."""
        from . import host
        n, t = np.ascontiguousarray(nodes, host.TEX_DTYPE), _arr(texels, np.float32)
        self._ck(self.L.drt_set_textures(self.h, n.shape[0], _p(n), _p(t), t.size))

    def set_material_programs(self, programs):
        from . import host
        pr = np.ascontiguousarray(programs, host.PROG_DTYPE)
        self._ck(self.L.drt_set_material_programs(self.h, pr.shape[0], _p(pr)))

    def set_infinite_light(self, index, texels, light_to_world, world_to_light):
        """Radiance map (h x w x 3 float32, power-of-two resolution: level 0 of the reference's MIPMap) and transforms of light
        `index`, which the last set_lights declared with kind 4."""
        t = _arr(texels, np.float32)
        l2w, w2l = _arr(light_to_world, np.float32).reshape(16), _arr(world_to_light, np.float32).reshape(16)
        self._ck(self.L.drt_set_infinite_light(self.h, int(index), t.shape[1], t.shape[0], _p(t), _p(l2w), _p(w2l)))

    def set_camera(self, raster_to_camera, camera_to_world, lens_radius=0.0, focal_distance=1e30, shutter_open=0.0,
                   shutter_close=1.0):
        r2c, c2w = _arr(raster_to_camera, np.float32).reshape(16), _arr(camera_to_world, np.float32).reshape(16)
        self._ck(self.L.drt_set_camera(self.h, _p(r2c), _p(c2w), lens_radius, focal_distance, shutter_open, shutter_close))

    def set_camera_motion(self, camera_to_world_end, start_time=0.0, end_time=1.0):
        """The end-time camera-to-world matrix of an animated camera (None: static); after set_camera."""
        m = _arr(camera_to_world_end, np.float32)
        self._ck(self.L.drt_set_camera_motion(self.h, _p(None if m is None else m.reshape(16)), float(start_time), float(end_time)))

    def set_camera_kind(self, kind):
        self._ck(self.L.drt_set_camera_kind(self.h, kind))

    def set_film(self, xres, yres, crop, xwidth, ywidth, table):
        crop, table = _arr(crop, np.float64), _arr(table, np.float32)
        self._ck(self.L.drt_set_film(self.h, xres, yres, _p(crop), xwidth, ywidth, _p(table)))

    def set_sampler(self, kind, xs, ys, spp, jitter, pixel_order, tile_size, seed, rng_mode=1):
        if rng_mode != 1:
            raise ValueError("the GPU path replays keyed (counter-based) sample streams only; the reference's single "
                             "serial RNG stream cannot be evaluated in parallel")
        self._ck(self.L.drt_set_sampler(self.h, kind, xs, ys, spp, jitter, pixel_order, tile_size, seed))

    def set_sample_table(self, table):
        """The bestcandidate sampler's 4096 x 5 pattern (doubles)."""
        t = _arr(table, np.float64).reshape(-1, 5)
        self._ck(self.L.drt_set_sample_table(self.h, _p(t), t.shape[0]))

    def set_integrator(self, kind, maxdepth, strategy, ao_nsamples, ao_mindist, ao_maxdist):
        self._ck(self.L.drt_set_integrator(self.h, kind, maxdepth, strategy, ao_nsamples, ao_mindist, ao_maxdist))

    def set_batch_slots(self, slots: int):
        self._ck(self.L.drt_set_batch_slots(self.h, slots))

    def render(self, task_num=0, task_count=1, nthreads=None):
        self._ck(self.L.drt_render(self.h, task_num, task_count))

    def render_shard(self, shard=0, n_shards=1):
        self._ck(self.L.drt_render_shard(self.h, shard, n_shards))

    def film_clear(self):
        self._ck(self.L.drt_film_clear(self.h))
        self._film_is_summed = False  # distributed.render_sharded

    def film_size(self):
        out = np.zeros(4, np.int32)
        self._ck(self.L.drt_film_size(self.h, _p(out)))
        return tuple(int(v) for v in out)  # left, top, width, height

    def film_read(self):
        _, _, w, h = self.film_size()
        rgb, xyz, wt = np.empty((h, w, 3), np.float32), np.empty((h, w, 3), np.float32), np.empty((h, w), np.float32)
        self._ck(self.L.drt_film_read(self.h, _p(rgb), _p(xyz), _p(wt)))
        return dict(rgb=rgb, xyz=xyz, weight=wt)

    def film_device(self):
        """(device pointer, number of float64 elements) of the film accumulators [h, w, (X, Y, Z, weight)]."""
        ptr, n = C.c_void_p(), C.c_uint64()
        self._ck(self.L.drt_film_device(self.h, C.byref(ptr), C.byref(n)))
        return int(ptr.value), int(n.value)

    def pixel_samples(self, x, y, cap=1 << 20):
        out = np.zeros(cap, np.float32)
        n, per = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.drt_pixel_samples(self.h, x, y, _p(out), cap, C.byref(n), C.byref(per)))
        return out[:per.value * n.value].reshape(n.value, per.value).copy()

    def set_volumes(self, v: dict):
        """v: host.pack_volumes(...) — the flat arrays of drt_set_volumes."""
        self._ck(self.L.drt_set_volumes(self.h, v["n"], _p(v["kind"]), _p(v["sigma_a"]), _p(v["sigma_s"]), _p(v["le"]), _p(v["g"]),
                                        _p(v["p0p1"]), _p(v["v2w"]), _p(v["w2v"]), _p(v["ab"]), _p(v["up"]), _p(v["dims"]),
                                        _p(v["density_offsets"]), _p(v["density"])))

    def set_volume_integrator(self, kind: int, step_size: float):
        self._ck(self.L.drt_set_volume_integrator(self.h, kind, float(step_size)))

    def set_shading_precision(self, precision: int):
        """PRECISION_F64 (0, default): the reference's arithmetic; PRECISION_F32 (1): float32 path-vertex kernels (3-sigma parity)."""
        self._ck(self.L.drt_set_shading_precision(self.h, int(precision)))

    def set_render_profiling(self, flags: int):
        """PROFILE_TIME: CUDA-event spans per kernel class; PROFILE_WORK: reference-walk counters of every traced queue."""
        self._ck(self.L.drt_set_render_profiling(self.h, flags))

    def render_profile(self) -> dict:
        p = RenderProfile()
        self._ck(self.L.drt_render_profile_get(self.h, C.byref(p)))
        cnt = lambda c: {k: int(getattr(c, k)) for k, _ in Counters._fields_}
        return {"ms": dict(zip(PROFILE_CLASSES, [float(v) for v in p.ms])),
                "launches": dict(zip(PROFILE_CLASSES, [int(v) for v in p.launches])),
                "closest": cnt(p.closest), "any": cnt(p.any)}

    def render_stats(self) -> dict:
        s = RenderStats()
        self._ck(self.L.drt_render_stats_get(self.h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in RenderStats._fields_}
