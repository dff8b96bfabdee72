"""Times drt_render on the BASELINE.json render configs (run on the GPU box).

  python tools/render_bench.py ao      [xres yres spp ao_nsamples]     config 3 (soup_1m, AO)
  python tools/render_bench.py path    [xres yres spp]                 config 4 (cornell_synth, path maxdepth 5)
  python tools/render_bench.py soup    [xres yres spp n_spheres]       config 5 style (soup, path)
  python tools/render_bench.py materials | sky [xres yres spp]         BxDF-list materials / environment-lit scene (path)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dartray_b200 import capi, host, scenes  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "path"
    a = [int(v) for v in sys.argv[2:]]
    if what == "ao":
        xres, yres, spp, ns = a[:4] if len(a) >= 4 else (1920, 1080, 1, 64)
        sb, cam = scenes.soup_render_scene(512)
        sampler = host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=spp, ys=1, jitter=False)
        integ = host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=ns)
    elif what in ("materials", "sky"):  # BxDF-list materials (cornell_materials) / the environment-lit scene of the GPU tests
        xres, yres, spp = a[:3] if len(a) >= 3 else (1920, 1080, 64)
        if what == "materials":
            sb, cam = scenes.cornell_materials()
        else:
            from tests.test_render_gpu import _sky_scene

            class _SB:  # upload_scene only needs arrays()
                def __init__(self, arr):
                    self._a = arr

                def arrays(self):
                    return self._a
            arr, cam = _sky_scene("lobes")
            sb = _SB(arr)
        sampler = host.Sampler(kind=host.SAMPLER_LD, spp=spp)
        integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    elif what == "path":
        xres, yres, spp = a[:3] if len(a) >= 3 else (1920, 1080, 16)
        sb, cam = scenes.cornell_synth()
        sampler = host.Sampler(kind=host.SAMPLER_LD, spp=spp)
        integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    else:
        xres, yres, spp, nsph = a[:4] if len(a) >= 4 else (1920, 1080, 4, 512)
        sb, cam = scenes.soup_render_scene(nsph)
        sampler = host.Sampler(kind=host.SAMPLER_LD, spp=spp)
        integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    if os.environ.get("INTEG") == "whitted":
        integ = host.Integrator(kind=host.INTEGRATOR_WHITTED, maxdepth=3)
    if os.environ.get("INTEG") == "ao":  # the same scene through the ambient-occlusion integrator, 64 rays per hit
        integ = host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=64)
    if os.environ.get("INTEG") == "direct":  # the same scene through the directlighting integrator (strategy all)
        integ = host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=3)
    ctx = capi.Context(0)
    t0 = time.time()
    host.upload_scene(ctx, sb.arrays())
    t_build = time.time() - t0
    host.configure_render(ctx, cam, host.Film(xres, yres), sampler, integ)
    if os.environ.get("SLOTS"):
        ctx.set_batch_slots(int(os.environ["SLOTS"]))
    ctx.render()  # warm-up (allocations, first launches)
    ctx.film_clear()
    t0 = time.time()
    ctx.render()
    dt = time.time() - t0
    st = ctx.render_stats()
    rays = st["closest_rays"] + st["shadow_rays"]
    f = ctx.film_read()
    print(json.dumps({"what": what, "xres": xres, "yres": yres, "spp": spp, "seconds": dt, "build_seconds": t_build,
                      "samples_per_s": st["camera_samples"] / dt, "mrays_per_s": rays / dt / 1e6, "stats": st,
                      "mean_rgb": [float(v) for v in f["rgb"].mean(axis=(0, 1))], "launches": ctx.kernel_launches,
                      "bvh": ctx.bvh_info()}))


if __name__ == "__main__":
    main()
