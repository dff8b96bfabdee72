"""include/drt.h from a plain-C client: tools/abi_harness.c, compiled with gcc against the header and linked with
libdartray_gpu.so, drives the library in the call order of the Dart shim (dart/lib/gpu/gpu_sampler_renderer.dart) — the image has
no Dart SDK, so this is the shim's stand-in at run time — and must produce the film the ctypes path produces."""
import os
import struct
import subprocess

import numpy as np
import pytest

from dartray_b200 import capi, host, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "abi_harness")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tools", "abi_harness.c"), "-L", os.path.join(ROOT, "dartray_b200"), "-ldartray_gpu",
           "-Wl,-rpath," + os.path.join(ROOT, "dartray_b200")]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_header_compiles_as_c99_and_links(drt_lib, tmp_path):
    """CPU: the header is valid C (not only C++), every symbol the harness uses resolves at link time."""
    assert os.path.exists(_build(tmp_path))


def _write_blob(path, entries):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(entries)))
        for name, arr in entries.items():
            data = np.ascontiguousarray(arr).tobytes()
            f.write(name.encode().ljust(32, b"\0")[:32])
            f.write(struct.pack("<q", len(data)))
            f.write(data)


def _textured_cornell():
    """cornell-path.pbrt with bump-sphere.pbrt's material on the sphere and an image texture on the back wall's material."""
    sb, cam = scenes.cornell_path()
    rng = np.random.default_rng(5)
    kd = host.ScaleTexture(host.ImageTexture(rng.random((16, 16, 3)).astype(np.float32)), (0.8, 0.7, 0.4))
    bump = host.ScaleTexture(host.ImageTexture(rng.random((8, 8)).astype(np.float32), host.UVMapping(4.0, 4.0)), -0.1)
    m = sb.material_program("uber", kd=kd, ks=0.05, roughness=0.01, bumpmap=bump)
    sb.sph = [(s[0], s[1], s[2], m, s[4], s[5]) for s in sb.sph]
    return sb, cam


@pytest.mark.gpu
@pytest.mark.parametrize("textured", [False, True], ids=["matte", "textured"])
def test_c_client_renders_the_film_the_ctypes_client_renders(drt_lib, tmp_path, textured):
    exe = _build(tmp_path)
    sb, cam = _textured_cornell() if textured else scenes.cornell_path()  # the shipped scene: 22 triangles, a sphere, the DISK area light
    a = sb.arrays()
    assert bool(a.get("mat_general")) == textured
    film, smp = host.Film(96, 72), host.Sampler(kind=host.SAMPLER_LD, spp=4)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    xw, yw, table = film.table()
    d = lambda *v: np.asarray(v, np.float64)
    entries = {k: a[k] for k in ("P", "idx", "tri_mat", "tri_light", "tri_rev", "sph_o2w", "sph_w2o", "sph_params", "sph_mat", "sph_light",
                                 "sph_rev", "dsk_o2w", "dsk_w2o", "dsk_params", "dsk_mat", "dsk_light", "dsk_rev", "order", "mat_kind",
                                 "mat_kd", "mat_sigma", "light_kind", "light_L", "light_pos", "light_nsamples", "light_shape_offsets",
                                 "light_shape_prims")}
    if textured:  # the records go to the C client as raw bytes: its compiler's struct layout reads them
        entries.update({k: a[k] for k in ("mat_lobe_offsets", "lobe_kind", "lobe_rgb", "lobe_fresnel", "lobe_eta", "lobe_k", "lobe_scalars",
                                          "tex_nodes", "tex_texels", "mat_programs")})
    entries.update({
        "device": d(0), "bvh": d(2, 4),
        "raster_to_camera": np.asarray(cam.raster_to_camera(film.xres, film.yres), np.float32),
        "camera_to_world": np.asarray(cam.camera_to_world, np.float32),
        "camera": d(cam.lens_radius, cam.focal_distance, cam.shutter_open, cam.shutter_close, getattr(cam, "kind", 0)),
        "film": d(film.xres, film.yres, xw, yw), "crop": np.asarray(film.crop, np.float64),
        "filter_table": np.asarray(table, np.float32),
        "sampler": d(smp.kind, smp.xs, smp.ys, smp.spp, int(smp.jitter), smp.pixel_order, smp.tile_size, smp.seed),
        "integrator": d(integ.kind, integ.maxdepth, integ.strategy, integ.ao_nsamples, integ.ao_mindist, min(integ.ao_maxdist, 1e300)),
    })
    blob, out = str(tmp_path / "scene.blob"), str(tmp_path / "film.out")
    _write_blob(blob, entries)
    res = subprocess.run([exe, blob, out], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    print(res.stdout.strip())
    raw = open(out, "rb").read()
    left, top, w, h = struct.unpack("<4i", raw[:16])
    rgb = np.frombuffer(raw, np.float32, w * h * 3, 16).reshape(h, w, 3)
    wt = np.frombuffer(raw, np.float32, w * h, 16 + 4 * w * h * 3).reshape(h, w)
    g = capi.Context(0)
    host.upload_scene(g, a)
    host.configure_render(g, cam, film, smp, integ)
    g.render()
    ref = g.film_read()
    assert (w, h) == (96, 72)
    assert np.array_equal(wt, ref["weight"]) and np.array_equal(rgb, ref["rgb"])
