"""One process, one multi-device context (drt_create_multi) through ctypes: config 4 (cornell_synth 1080p x 256 spp path) and,
with --config5, soup_10m 4K (set-up time: one host build, per-device upload threads) on 1, 2, ... all GPUs of the box.
    python tools/multi_ctx_bench.py [--config5] [--spp5 64]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from dartray_b200 import capi, host, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--config5", action="store_true")
ap.add_argument("--spp5", type=int, default=64)
args = ap.parse_args()
ids = []
for d in range(16):
    try:
        capi.Context(d).close(); ids.append(d)
    except capi.DrtError:
        break
counts = [n for n in (1, 2, 4, 8) if n <= len(ids)]
sb, cam = scenes.cornell_synth()
arrays = sb.arrays()
film, smp = host.Film(1920, 1080), host.Sampler(kind=host.SAMPLER_LD, spp=256)
integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
base = None
for n in counts:
    c = capi.Context(ids[:n])
    host.upload_scene(c, arrays)
    host.configure_render(c, cam, film, smp, integ)
    c.render(); c.film_clear()  # warm-up: wavefront allocation
    t0 = time.perf_counter(); c.render(); dt = time.perf_counter() - t0
    st = c.render_stats(); rgb = c.film_read()["rgb"]
    if base is None: base = rgb
    print(json.dumps({"config": "4 (cornell_synth 1080p x 256 spp path)", "devices": n, "seconds": dt, "samples_per_s": st["camera_samples"] / dt,
                      "speedup": None if n == 1 else None, "max_abs_diff_vs_1gpu": float(np.abs(rgb - base).max())}), flush=True)
    c.close()
if args.config5:
    t0 = time.perf_counter(); sb5, cam5 = scenes.soup_render_scene(5120); arr5 = sb5.arrays(); t_gen = time.perf_counter() - t0
    for n in ([counts[0], counts[-1]] if len(counts) > 2 else counts):
        c = capi.Context(ids[:n])
        t0 = time.perf_counter(); host.upload_scene(c, arr5); t_up = time.perf_counter() - t0
        info = c.bvh_info()
        host.configure_render(c, cam5, host.Film(3840, 2160), host.Sampler(kind=host.SAMPLER_LD, spp=args.spp5), integ)
        c.render(); c.film_clear()  # warm-up with the timed configuration: the wavefront allocation is sized by it
        t0 = time.perf_counter(); c.render(); dt = time.perf_counter() - t0
        st = c.render_stats()
        print(json.dumps({"config": f"5 (soup_10m 4K x {args.spp5} spp path)", "devices": n, "scene_gen_s": t_gen, "set_triangles_build_upload_s": t_up,
                          "host_build_s": info["build_seconds"], "device_bytes_per_gpu": info["device_bytes"], "render_s": dt,
                          "samples_per_s": st["camera_samples"] / dt}), flush=True)
        c.close()
