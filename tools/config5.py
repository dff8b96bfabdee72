"""BASELINE.json configs[4] / SURVEY §8d config 5: ~10M-triangle procedural scene (soup_10m: 5,120 UV-spheres,
10,158,082 triangles + a quad light), 3840x2160, lowdiscrepancy 1024 spp, path maxdepth 5, on N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/config5.py [--spheres 5120] [--res 3840 2160] [--spp 1024] [--crop-parity]

Reports BVH bytes per GPU, build time, samples/s, per-GPU busy-time imbalance (max / mean); `--crop-parity` also
renders a 240x135 x 64 spp film of the same camera on rank 0 and compares it with the CPU oracle (the oracle cannot
afford the full size)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dartray_b200 import capi, distributed, host, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=5120)
    ap.add_argument("--res", type=int, nargs=2, default=[3840, 2160])
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--crop-parity", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    t0 = time.perf_counter()
    sb, cam = scenes.soup_render_scene(args.spheres)
    arrays = sb.arrays()
    t_gen = time.perf_counter() - t0
    ctx = capi.Context(local)
    t0 = time.perf_counter()
    host.upload_scene(ctx, arrays)
    t_upload = time.perf_counter() - t0
    info = ctx.bvh_info()
    film = host.Film(*args.res)
    integ = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)
    host.configure_render(ctx, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=min(args.spp, 4)), integ)
    distributed.render_sharded(ctx, rank, world)  # warm-up
    host.configure_render(ctx, cam, film, host.Sampler(kind=host.SAMPLER_LD, spp=args.spp), integ)
    ctx.film_clear()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t0 = time.perf_counter()
    ctx.render_shard(rank, world)
    busy = time.perf_counter() - t0  # this rank's own render time (before the film sum)
    if world > 1:
        distributed.sum_films(distributed.film_tensor(ctx))
        torch.cuda.synchronize()
    barrier()
    total = time.perf_counter() - t0
    st = ctx.render_stats()
    v = torch.tensor([busy, total, st["camera_samples"], st["closest_rays"], st["shadow_rays"]], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(v) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, v)
    else:
        gathered = [v]
    if rank == 0:
        g = torch.stack(gathered).cpu().numpy()
        sec = float(g[:, 1].max())
        out = {
            "config": f"soup ({args.spheres} spheres, {info['n_prims']} primitives), {args.res[0]}x{args.res[1]}, lowdiscrepancy "
                      f"{args.spp} spp, path maxdepth 5, {world} GPU(s)",
            "bvh": {"reference_nodes": info["n_nodes"], "device_bytes_per_gpu": info["device_bytes"],
                    "build_seconds": info["build_seconds"], "scene_gen_seconds": t_gen, "upload_total_seconds": t_upload},
            "seconds": sec, "camera_samples": int(g[:, 2].sum()), "samples_per_s": float(g[:, 2].sum() / sec),
            "mrays_per_s": float((g[:, 3].sum() + g[:, 4].sum()) / sec / 1e6),
            "busy_seconds_per_gpu": [float(x) for x in g[:, 0]],
            "imbalance_max_over_mean": float(g[:, 0].max() / g[:, 0].mean()),
            "mean_rgb": [float(x) for x in ctx.film_read()["rgb"].mean(axis=(0, 1))],
            "shading": "float32 (DRT_SHADE_F32=1)" if os.environ.get("DRT_SHADE_F32") == "1" else "binary64",
        }
        if args.crop_parity:
            from tests.oracle_lib import Oracle
            cfilm, csmp = host.Film(240, 135), host.Sampler(kind=host.SAMPLER_LD, spp=64)
            c2 = capi.Context(local)
            host.upload_scene(c2, arrays)
            host.configure_render(c2, cam, cfilm, csmp, integ)
            c2.render()
            o = Oracle()
            t0 = time.perf_counter()
            host.upload_scene(o, arrays)
            host.configure_render(o, cam, cfilm, csmp, integ)
            o.render(0, 1, os.cpu_count() or 1)
            a, b = c2.film_read()["rgb"], o.film_read()["rgb"]
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
            out["crop_parity"] = {"film": "240x135, 64 spp", "max_rel_err": float(err.max()), "q9999_rel_err": float(np.quantile(err, 0.9999)),
                                  "mean_gpu": float(a.mean()), "mean_oracle": float(b.mean()), "oracle_seconds": time.perf_counter() - t0}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
