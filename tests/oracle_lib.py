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
        L.orc_set_disks.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp]
        L.orc_set_quadrics.argtypes = [vp, i32, u32, vp, vp, vp, vp, vp, vp]
        L.orc_set_mesh_shading.argtypes = [vp, vp, vp, vp, vp, u32, vp, vp, vp]
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
        dbl = C.c_double
        L.orc_set_materials.argtypes = [vp, u32, vp, vp, vp]
        L.orc_set_material_lobes.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp, vp]
        L.orc_set_spot_params.argtypes = [vp, u32, vp, vp]
        L.orc_set_sample_table.argtypes = [vp, vp, u32]
        L.orc_set_light_map.argtypes = [vp, u32, i32, i32, vp, vp, vp, vp, dbl]
        L.orc_set_lobe_wrappers.argtypes = [vp, u32, vp, vp]
        L.orc_set_infinite_light.argtypes = [vp, u32, i32, i32, vp, vp, vp]
        L.orc_set_textures.argtypes = [vp, u32, vp, vp, u64]
        L.orc_set_measured.argtypes = [vp, u32, vp, vp, vp, vp, u64]
        L.orc_set_instances.argtypes = [vp, u32, vp, vp, vp, vp, u32, vp, vp, vp, vp, vp, vp]
        L.orc_set_ray_times.argtypes = [vp, vp, u64]
        L.orc_instance_probe.argtypes = [vp, u32, dbl, vp, vp, vp, vp, vp, vp, vp]
        L.orc_set_material_programs.argtypes = [vp, u32, vp]
        L.orc_texture_eval.argtypes = [vp, i32, u32, vp, vp]
        L.orc_image_level.argtypes = [vp, i32, i32, vp, vp, vp]
        L.orc_set_lights.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp]
        L.orc_set_camera.argtypes = [vp, vp, vp, dbl, dbl, dbl, dbl]
        L.orc_set_camera_kind.argtypes = [vp, i32]
        L.orc_set_camera_motion.argtypes = [vp, vp, dbl, dbl]
        L.orc_set_film.argtypes = [vp, i32, i32, vp, dbl, dbl, vp]
        L.orc_set_sampler.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, u64, i32]
        L.orc_set_integrator.argtypes = [vp, i32, i32, i32, i32, dbl, dbl]
        L.orc_render.argtypes = [vp, i32, i32, i32]
        L.orc_film_clear.argtypes = [vp]
        L.orc_film_size.argtypes = [vp, vp]
        L.orc_film_read.argtypes = [vp, vp, vp, vp]
        L.orc_pixel_samples.argtypes = [vp, i32, i32, vp, i32, vp]
        L.orc_render_stats.argtypes = [vp, vp]
        L.orc_dart_random.argtypes = [C.c_int64, i32, vp, vp]
        L.orc_set_volumes.argtypes = [vp, u32] + [vp] * 13
        L.orc_set_volume_integrator.argtypes = [vp, i32, dbl]
        L.orc_bsdf_eval.argtypes = [vp, u32, u32, vp, vp, i32, vp, vp]
        L.orc_bsdf_sample.argtypes = [vp, u32, u32, vp, vp, i32, vp, vp, vp, vp]
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

    def set_disks(self, o2w, w2o, params, material=None, light=None, reverse=None):
        o2w = _arr(o2w, np.float32).reshape(-1, 16)
        w2o = _arr(w2o, np.float32).reshape(-1, 16)
        params = _arr(params, np.float64).reshape(-1, 4)
        m, l, r = _arr(material, np.int32), _arr(light, np.int32), _arr(reverse, np.uint8)
        self._ck(self.L.orc_set_disks(self.h, o2w.shape[0], _p(o2w), _p(w2o), _p(params), _p(m), _p(l), _p(r)))

    def set_quadrics(self, kind, o2w, w2o, params, material=None, light=None, reverse=None):
        o2w = _arr(o2w, np.float32).reshape(-1, 16)
        w2o = _arr(w2o, np.float32).reshape(-1, 16)
        params = _arr(params, np.float64).reshape(-1, 8)
        m, l, r = _arr(material, np.int32), _arr(light, np.int32), _arr(reverse, np.uint8)
        self._ck(self.L.orc_set_quadrics(self.h, int(kind), o2w.shape[0], _p(o2w), _p(w2o), _p(params), _p(m), _p(l), _p(r)))

    def set_mesh_shading(self, N, S, uv, mesh_of_tri, o2w, w2o, flags):
        N, S, uv = _arr(N, np.float32), _arr(S, np.float32), _arr(uv, np.float32)
        mot, flags = _arr(mesh_of_tri, np.uint32), _arr(flags, np.uint8)
        o2w, w2o = _arr(o2w, np.float32).reshape(-1, 16), _arr(w2o, np.float32).reshape(-1, 16)
        self._ck(self.L.orc_set_mesh_shading(self.h, _p(N), _p(S), _p(uv), _p(mot), o2w.shape[0], _p(o2w), _p(w2o), _p(flags)))

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

    # -- renderer (same method names as dartray_b200.capi.Context) --------------------------------
    def set_materials(self, kind, kd, sigma):
        kind, kd, sigma = _arr(kind, np.int32), _arr(kd, np.float32).reshape(-1, 3), _arr(sigma, np.float32)
        self._ck(self.L.orc_set_materials(self.h, kd.shape[0], _p(kind), _p(kd), _p(sigma)))

    def set_material_lobes(self, offsets, kind, rgb, fresnel, eta, k, scalars):
        """Materials as ordered BxDF lists (host.matte_lobes / mirror_lobes / glass_lobes / plastic_lobes / ...)."""
        offsets, kind, fresnel = _arr(offsets, np.uint32), _arr(kind, np.int32), _arr(fresnel, np.int32)
        rgb, eta, k = (_arr(v, np.float32).reshape(-1, 3) for v in (rgb, eta, k))
        scalars = _arr(scalars, np.float64).reshape(-1, 3)
        self._ck(self.L.orc_set_material_lobes(self.h, offsets.shape[0] - 1, _p(offsets), _p(kind), _p(rgb), _p(fresnel), _p(eta),
                                                _p(k), _p(scalars)))

    def set_lights(self, kind, L, pos, nsamples, shape_offsets, shape_prims):
        kind, L, pos = _arr(kind, np.int32), _arr(L, np.float32).reshape(-1, 3), _arr(pos, np.float32).reshape(-1, 3)
        ns, so, sp = _arr(nsamples, np.int32), _arr(shape_offsets, np.uint32), _arr(shape_prims, np.uint32)
        self._ck(self.L.orc_set_lights(self.h, kind.shape[0], _p(kind), _p(L), _p(pos), _p(ns), _p(so), _p(sp)))

    def set_spot_params(self, world_to_light, cosines):
        """worldToLight (n x 16) and (cosTotalWidth, cosFalloffStart) (n x 2) of the spot lights of the last set_lights."""
        w, cs = _arr(world_to_light, np.float32).reshape(-1, 16), _arr(cosines, np.float64).reshape(-1, 2)
        self._ck(self.L.orc_set_spot_params(self.h, w.shape[0], _p(w), _p(cs)))

    def set_light_map(self, index, texels, world_to_light, projection=None, screen=None, hither=1.0e-3):
        """Map (h x w x 3 float32, power-of-two, or None) and transforms of a projection (kind 5) / goniometric (kind 6) light."""
        t = _arr(texels, np.float32)
        w2l = _arr(world_to_light, np.float32).reshape(16)
        pr = _arr(projection, np.float32)
        sc = _arr(screen, np.float64)
        self._ck(self.L.orc_set_light_map(self.h, int(index), 0 if t is None else t.shape[1], 0 if t is None else t.shape[0], _p(t), _p(w2l),
                                          _p(pr), _p(sc), float(hither)))

    def set_lobe_wrappers(self, wrap, scale):
        """BRDFToBTDF (bit 0) / ScaledBxDF (bit 1, with its RGB scale) around the lobes of the last set_material_lobes."""
        w, sc = _arr(wrap, np.int32), _arr(scale, np.float32).reshape(-1, 3)
        self._ck(self.L.orc_set_lobe_wrappers(self.h, w.shape[0], _p(w), _p(sc)))

    def set_instances(self, object_offsets, object_prims, object_split, object_max_node_prims, instance_object, start_m, start_minv,
                      end_m, end_minv, times):
        """TransformedPrimitives (transformed_primitive.dart): objects (prim lists + nested accelerator parameters) and instances
        (object, world-to-primitive m / mInv at the start and end time, n x 2 times).  Before set_build_order / build_bvh."""
        oo, op = _arr(object_offsets, np.uint32), _arr(object_prims, np.uint32)
        os_, om = _arr(object_split, np.int32), _arr(object_max_node_prims, np.int32)
        io = _arr(instance_object, np.uint32)
        m = [_arr(x, np.float32).reshape(-1, 16) for x in (start_m, start_minv, end_m, end_minv)]
        tm = _arr(times, np.float64).reshape(-1, 2)
        self._ck(self.L.orc_set_instances(self.h, oo.shape[0] - 1, _p(oo), _p(op), _p(os_), _p(om), io.shape[0], _p(io), _p(m[0]), _p(m[1]),
                                          _p(m[2]), _p(m[3]), _p(tm)))

    def set_ray_times(self, times):
        """Ray i of the following trace_* calls travels at times[i] (None: every ray at time 0)."""
        t = _arr(times, np.float64)
        self._ck(self.L.orc_set_ray_times(self.h, _p(t), 0 if t is None else t.size))

    def instance_probe(self, inst, time=0.0):
        """Decompose / interpolate(time) / world bound of instance `inst`'s AnimatedTransform."""
        T, R, S = np.zeros((2, 3)), np.zeros((2, 4)), np.zeros((2, 16), np.float32)
        m, minv, bound = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(6, np.float32)
        an = C.c_int32(0)
        self._ck(self.L.orc_instance_probe(self.h, inst, float(time), _p(T), _p(R), _p(S), _p(m), _p(minv), _p(bound), C.byref(an)))
        return dict(T=T, R=R, S=S.reshape(2, 4, 4), m=m.reshape(4, 4), minv=minv.reshape(4, 4), bound=bound, animated=bool(an.value))

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
        self._ck(self.L.orc_set_measured(self.h, len(tables), _p(kinds), _p(dims), _p(offs), _p(data), data.size))

    def set_textures(self, nodes, texels):
        """Texture nodes (host.TEX_DTYPE records = drt_texture) and the level-0 texels of their images."""
        from dartray_b200 import host
        n, t = np.ascontiguousarray(nodes, host.TEX_DTYPE), _arr(texels, np.float32)
        self._ck(self.L.orc_set_textures(self.h, n.shape[0], _p(n), _p(t), t.size))

    def set_material_programs(self, programs):
        from dartray_b200 import host
        pr = np.ascontiguousarray(programs, host.PROG_DTYPE)
        self._ck(self.L.orc_set_material_programs(self.h, pr.shape[0], _p(pr)))

    def texture_eval(self, node, dgs):
        """Texture.evaluate of node at n DifferentialGeometry records (p 3, u, v, dudx, dvdx, dudy, dvdy, dpdx 3, dpdy 3)."""
        d = _arr(dgs, np.float64).reshape(-1, 15)
        out = np.zeros((d.shape[0], 3), np.float64)
        self._ck(self.L.orc_texture_eval(self.h, node, d.shape[0], _p(d), _p(out)))
        return out

    def image_levels(self, node, channels):
        """The MIPMap pyramid the oracle built for image texture `node`: list of (h, w) or (h, w, 3) float32 arrays."""
        w, h = C.c_int32(0), C.c_int32(0)
        n = self.L.orc_image_level(self.h, node, -1, None, None, None)
        out = []
        for lv in range(n):
            self.L.orc_image_level(self.h, node, lv, C.byref(w), C.byref(h), None)
            buf = np.zeros(w.value * h.value * channels, np.float32)
            self.L.orc_image_level(self.h, node, lv, C.byref(w), C.byref(h), _p(buf))
            out.append(buf.reshape((h.value, w.value) if channels == 1 else (h.value, w.value, 3)))
        return out

    def set_infinite_light(self, index, texels, light_to_world, world_to_light):
        """Radiance map (h x w x 3 float32, power-of-two resolution: level 0 of the reference's MIPMap) and transforms of light
        `index`, which the last set_lights declared with kind 4."""
        t = _arr(texels, np.float32)
        l2w, w2l = _arr(light_to_world, np.float32).reshape(16), _arr(world_to_light, np.float32).reshape(16)
        self._ck(self.L.orc_set_infinite_light(self.h, int(index), t.shape[1], t.shape[0], _p(t), _p(l2w), _p(w2l)))

    def set_camera(self, raster_to_camera, camera_to_world, lens_radius=0.0, focal_distance=1e30, shutter_open=0.0,
                   shutter_close=1.0):
        r2c, c2w = _arr(raster_to_camera, np.float32).reshape(16), _arr(camera_to_world, np.float32).reshape(16)
        self._ck(self.L.orc_set_camera(self.h, _p(r2c), _p(c2w), lens_radius, focal_distance, shutter_open, shutter_close))

    def set_camera_motion(self, camera_to_world_end, start_time=0.0, end_time=1.0):
        m = _arr(camera_to_world_end, np.float32)
        self._ck(self.L.orc_set_camera_motion(self.h, _p(None if m is None else m.reshape(16)), float(start_time), float(end_time)))

    def set_camera_kind(self, kind):
        self._ck(self.L.orc_set_camera_kind(self.h, kind))

    def set_film(self, xres, yres, crop, xwidth, ywidth, table):
        crop, table = _arr(crop, np.float64), _arr(table, np.float32)
        self._ck(self.L.orc_set_film(self.h, xres, yres, _p(crop), xwidth, ywidth, _p(table)))

    def set_sampler(self, kind, xs, ys, spp, jitter, pixel_order, tile_size, seed, rng_mode):
        self._ck(self.L.orc_set_sampler(self.h, kind, xs, ys, spp, jitter, pixel_order, tile_size, seed, rng_mode))

    def set_sample_table(self, table):
        """The bestcandidate sampler's 4096 x 5 pattern (doubles)."""
        t = _arr(table, np.float64).reshape(-1, 5)
        self._ck(self.L.orc_set_sample_table(self.h, _p(t), t.shape[0]))

    def set_integrator(self, kind, maxdepth, strategy, ao_nsamples, ao_mindist, ao_maxdist):
        self._ck(self.L.orc_set_integrator(self.h, kind, maxdepth, strategy, ao_nsamples, ao_mindist, ao_maxdist))

    def set_volumes(self, v):
        """v: the dict host.pack_volumes() returns (same arrays as Context.set_volumes)."""
        self._keep_vol = v
        self._ck(self.L.orc_set_volumes(self.h, v["n"], _p(v["kind"]), _p(v["sigma_a"]), _p(v["sigma_s"]), _p(v["le"]), _p(v["g"]),
                                        _p(v["p0p1"]), _p(v["v2w"]), _p(v["w2v"]), _p(v["ab"]), _p(v["up"]), _p(v["dims"]),
                                        _p(v["density_offsets"]), _p(v["density"])))

    def set_volume_integrator(self, kind, step_size):
        self._ck(self.L.orc_set_volume_integrator(self.h, kind, float(step_size)))

    def render(self, task_num=0, task_count=1, nthreads=8):
        self._ck(self.L.orc_render(self.h, task_num, task_count, nthreads))

    def film_clear(self):
        self._ck(self.L.orc_film_clear(self.h))

    def film_size(self):
        out = np.zeros(4, np.int32)
        self.L.orc_film_size(self.h, _p(out))
        return tuple(int(v) for v in out)  # left, top, width, height

    def film_read(self):
        _, _, w, h = self.film_size()
        rgb, xyz, wt = np.empty((h, w, 3), np.float32), np.empty((h, w, 3), np.float32), np.empty((h, w), np.float32)
        self._ck(self.L.orc_film_read(self.h, _p(rgb), _p(xyz), _p(wt)))
        return dict(rgb=rgb, xyz=xyz, weight=wt)

    def pixel_samples(self, x, y, cap=1 << 20):
        out = np.zeros(cap, np.float32)
        n = C.c_int(0)
        per = self.L.orc_pixel_samples(self.h, x, y, _p(out), cap, C.byref(n))
        return out[:per * n.value].reshape(n.value, per).copy()

    def bsdf_eval(self, material, wo, wi, flags=31):
        """BSDF.f / BSDF.pdf (bsdf.dart:128-198) of `material` in the canonical frame sn = +x, tn = +y, nn = ng = +z."""
        wo = np.ascontiguousarray(np.broadcast_to(np.asarray(wo, np.float64), np.asarray(wi).shape), np.float64).reshape(-1, 3)
        wi = np.ascontiguousarray(wi, np.float64).reshape(-1, 3)
        n = wi.shape[0]
        f, pdf = np.zeros((n, 3), np.float32), np.zeros(n, np.float64)
        self._ck(self.L.orc_bsdf_eval(self.h, material, n, _p(wo), _p(wi), flags, _p(f), _p(pdf)))
        return f, pdf

    def bsdf_sample(self, material, wo, u, flags=31):
        """BSDF.sample_f (bsdf.dart:53-126) for rows (u0, u1, component) of `u`; returns wi, f, pdf, sampled type."""
        u = np.ascontiguousarray(u, np.float64).reshape(-1, 3)
        wo = np.ascontiguousarray(np.broadcast_to(np.asarray(wo, np.float64), u.shape), np.float64)
        n = u.shape[0]
        wi, f, pdf, ty = np.zeros((n, 3), np.float64), np.zeros((n, 3), np.float32), np.zeros(n, np.float64), np.zeros(n, np.int32)
        self._ck(self.L.orc_bsdf_sample(self.h, material, n, _p(wo), _p(u), flags, _p(wi), _p(f), _p(pdf), _p(ty)))
        return wi, f, pdf, ty

    def render_stats(self):
        out = np.zeros(5, np.uint64)
        self.L.orc_render_stats(self.h, _p(out))
        return dict(zip(["camera_samples", "closest_rays", "shadow_rays", "nodes_visited", "prims_tested"], (int(v) for v in out)))

    def counters(self):
        out = np.zeros(3, np.uint64)
        self.L.orc_get_counters(self.h, _p(out))
        return dict(rays=int(out[0]), nodes_visited=int(out[1]), prims_tested=int(out[2]))
