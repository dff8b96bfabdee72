#!/usr/bin/env python
"""Writes tests/golden/*.npz — small fixed input/output vectors of the hot path.

Provenance: produced by the CPU oracle (oracle/, the C++ restatement of the DartRay path).  The
reference itself cannot run in this image (no Dart VM) and ships no golden vectors, so these are
REGRESSION vectors for the restatement and the CUDA path, not outputs of the Dart code: parity stays
"unpinned" in the sense of DESIGN.md §2.  Inputs are stored beside the outputs so a future run of the
real reference on the same inputs can be compared directly.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dartray_b200 import host, scenes  # noqa: E402
from tests.oracle_lib import Oracle  # noqa: E402
from tests.util import random_rays, random_soup, translate  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def trace_vectors():
    P, idx = random_soup(300, seed=11, extent=1.0, size=0.25)
    o2w, w2o = zip(*[translate(0.3, -0.2, 0.1), translate(-0.5, 0.4, -0.3)])
    sph = dict(o2w=np.stack(o2w), w2o=np.stack(w2o), params=np.array([[0.35, -0.35, 0.35, 360.0], [0.3, -0.1, 0.25, 270.0]]))
    ro, rd = random_rays(1024, seed=12, radius=2.5)
    ro[512:, 3] = 0.05  # a block of rays with a non-zero minDistance and a finite maxDistance
    rd[512:, 3] = 2.4
    o = Oracle()
    o.set_triangles(P, idx)
    o.set_spheres(sph["o2w"], sph["w2o"], sph["params"])
    o.build_bvh(2, 4)
    hits, t64 = o.trace_closest(ro, rd, want_t64=True)
    occ = o.trace_any(ro, rd)
    bvh = o.bvh_export()
    np.savez_compressed(os.path.join(OUT, "trace_soup300.npz"), P=P, idx=idx, sph_o2w=sph["o2w"], sph_w2o=sph["w2o"],
                        sph_params=sph["params"], ray_o=ro, ray_d=rd, hit_t=hits["t"], hit_b1=hits["b1"], hit_b2=hits["b2"],
                        hit_prim=hits["prim"], hit_t64=t64, occluded=occ, bvh_offset=bvh["offset"],
                        bvh_nprims=bvh["n_primitives"], bvh_axis=bvh["axis"], bvh_ordered=bvh["ordered"], bvh_bounds=bvh["bounds"])


RENDERS = {
    "path": (host.Sampler(kind=host.SAMPLER_LD, spp=4, seed=3), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)),
    "ao": (host.Sampler(kind=host.SAMPLER_LD, spp=1, seed=3), host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=16)),
    "direct": (host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=2, ys=2, seed=3), host.Integrator(kind=host.INTEGRATOR_DIRECT)),
}
FILM = (32, 24)


def render_vectors():
    sb, cam = scenes.cornell_synth()
    arrays = sb.arrays()
    out = {}
    for name, (sampler, integ) in RENDERS.items():
        o = Oracle()
        host.upload_scene(o, arrays)
        host.configure_render(o, cam, host.Film(*FILM), sampler, integ)
        o.render(0, 1, 1)
        f = o.film_read()
        out[f"{name}_rgb"] = f["rgb"]
        out[f"{name}_weight"] = f["weight"]
        out[f"{name}_samples_px_5_7"] = o.pixel_samples(5, 7)
        st = o.render_stats()
        out[f"{name}_rays"] = np.array([st["camera_samples"], st["closest_rays"], st["shadow_rays"]], np.int64)
    np.savez_compressed(os.path.join(OUT, "render_cornell_synth.npz"), **out)
    # the shipped scene itself (disk area light), web/scenes/cornell-path.pbrt
    sb, cam = scenes.cornell_path()
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    sampler, integ = RENDERS["path"]
    host.configure_render(o, cam, host.Film(*FILM), sampler, integ)
    o.render(0, 1, 1)
    f = o.film_read()
    st = o.render_stats()
    np.savez_compressed(os.path.join(OUT, "render_cornell_path.npz"), path_rgb=f["rgb"], path_weight=f["weight"],
                        path_rays=np.array([st["camera_samples"], st["closest_rays"], st["shadow_rays"]], np.int64))


def material_vectors():
    sb, cam = scenes.cornell_materials()
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    sampler, integ = RENDERS["path"]
    host.configure_render(o, cam, host.Film(*FILM), sampler, integ)
    o.render(0, 1, 1)
    f = o.film_read()
    st = o.render_stats()
    np.savez_compressed(os.path.join(OUT, "render_cornell_materials.npz"), path_rgb=f["rgb"], path_weight=f["weight"],
                        path_rays=np.array([st["camera_samples"], st["closest_rays"], st["shadow_rays"]], np.int64))


# Scenes of the widened path: every remaining quadric with a cylinder light; smooth-shaded meshes (per-vertex N / S /
# uv) under BxDF lists; an environment-mapped InfiniteAreaLight with glass and mirror spheres; the halton sampler.
FEATURES = {
    "quadrics_direct": ("_quadric_room", (), host.Sampler(kind=host.SAMPLER_LD, spp=2, seed=5), host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    "smooth_path": ("_smooth_room", ("lobes",), host.Sampler(kind=host.SAMPLER_LD, spp=2, seed=5), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4)),
    "sky_path": ("_sky_scene", ("lobes",), host.Sampler(kind=host.SAMPLER_LD, spp=2, seed=5), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4)),
    "sky_direct": ("_sky_scene", ("matte",), host.Sampler(kind=host.SAMPLER_LD, spp=2, seed=5), host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    "halton_direct": ("_cornell", (), host.Sampler(kind=host.SAMPLER_HALTON, spp=3, seed=5), host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    "adaptive_direct": ("_cornell", (), host.Sampler(kind=host.SAMPLER_ADAPTIVE, xs=2, ys=8, jitter=host.ADAPTIVE_CONTRAST, seed=5),
                        host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    "bestcandidate_direct": ("_cornell", (), host.Sampler(kind=host.SAMPLER_BEST_CANDIDATE, spp=4, seed=5), host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    "mapped_lights_direct": ("_mapped_lights_room", (), host.Sampler(kind=host.SAMPLER_LD, spp=2, seed=5), host.Integrator(kind=host.INTEGRATOR_DIRECT)),
    "wrapped_materials_path": ("_wrapped_room_arrays", (), host.Sampler(kind=host.SAMPLER_LD, spp=2, seed=5),
                               host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=4)),
}


def feature_scene(name):
    import tests.test_render_gpu as T
    fn, args, sampler, integ = FEATURES[name]
    arrays, cam = getattr(T, fn)(*args)
    if sampler.kind == host.SAMPLER_BEST_CANDIDATE:
        from tests.util import synthetic_sample_table
        sampler.sample_table = synthetic_sample_table()
    return arrays, cam, sampler, integ


def feature_vectors():
    out = {}
    for name in FEATURES:
        arrays, cam, sampler, integ = feature_scene(name)
        o = Oracle()
        host.upload_scene(o, arrays)
        host.configure_render(o, cam, host.Film(*FILM), sampler, integ)
        o.render(0, 1, 1)
        f = o.film_read()
        st = o.render_stats()
        out[f"{name}_rgb"] = f["rgb"]
        out[f"{name}_weight"] = f["weight"]
        out[f"{name}_rays"] = np.array([st["camera_samples"], st["closest_rays"], st["shadow_rays"]], np.int64)
    np.savez_compressed(os.path.join(OUT, "render_features.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    trace_vectors()
    render_vectors()
    material_vectors()
    feature_vectors()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
