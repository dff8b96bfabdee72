"""Where the time of one rank of an N-way sharded config-4 render goes: drt_render_shard(0, N) against 1/N of the whole render."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dartray_b200 import capi, host, scenes
sb, cam = scenes.cornell_synth()
c = capi.Context(0)
host.upload_scene(c, sb.arrays())
host.configure_render(c, cam, host.Film(1920, 1080), host.Sampler(kind=host.SAMPLER_LD, spp=256), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5))
for n in (1, 8, 8, 1, 8):
    c.film_clear()
    torch.cuda.synchronize()
    t = time.perf_counter()
    c.render_shard(0, n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    t2 = time.perf_counter()
    f = c.film_device() if hasattr(c, "film_device") else None
    torch.cuda.synchronize()
    print(f"shards {n}: render {dt * 1e3:.2f} ms (x{n} = {dt * n * 1e3:.1f} ms), film_device {1e3 * (time.perf_counter() - t2):.2f} ms, launches {c.kernel_launches}")
