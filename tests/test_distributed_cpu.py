"""Host-side logic of the multi-GPU render (dartray_b200/distributed.py) on CPU: shard partition
arithmetic and the film sum over a world_size-2 gloo group (the N > 1 path without a GPU)."""
import os
import socket

import numpy as np
import pytest

from dartray_b200 import distributed as D


@pytest.mark.parametrize("total", [0, 1, 1023, 1024, 1025, 4096, 50 * 38 + 89, 1921 * 1081])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 8])
def test_shards_partition_the_pixels(total, n):
    counts = [D.shard_pixel_count(total, s, n) for s in range(n)]
    assert sum(counts) == total
    if total <= 1 << 16:
        allidx = np.concatenate([D.shard_pixel_indices(total, s, n) for s in range(n)])
        assert allidx.shape[0] == total
        assert np.array_equal(np.sort(allidx), np.arange(total))
    # balance: no shard holds more than one block above the mean
    assert max(counts) - min(counts) <= D.BLOCK_PIXELS


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from dartray_b200 import host, scenes
    from tests.oracle_lib import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sb, cam = scenes.cornell_synth()
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, host.Film(40, 30), host.Sampler(kind=host.SAMPLER_LD, spp=2),
                          host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render(rank, world, 2)  # this rank's part of the sample extent into its own full-size film
    f = o.film_read()
    film = torch.from_numpy(np.concatenate([f["xyz"].astype(np.float64), f["weight"].astype(np.float64)[..., None]], axis=2))
    D.sum_films(film)  # what render_sharded does with the device film over NCCL
    np.save(os.path.join(out_dir, f"film{rank}.npy"), film.numpy())
    dist.destroy_process_group()


def test_film_sum_over_gloo_world_2(tmp_path):
    import torch.multiprocessing as mp

    from dartray_b200 import host, scenes
    from tests.oracle_lib import Oracle, build_oracle

    build_oracle()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "film0.npy"), np.load(tmp_path / "film1.npy")
    assert np.array_equal(a, b)  # every rank ends with the whole film
    sb, cam = scenes.cornell_synth()
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, host.Film(40, 30), host.Sampler(kind=host.SAMPLER_LD, spp=2),
                          host.Integrator(kind=host.INTEGRATOR_DIRECT))
    o.render(0, 1, 2)
    f = o.film_read()
    assert np.array_equal(a[..., 3], f["weight"].astype(np.float64))
    assert np.allclose(a[..., :3], f["xyz"], rtol=1e-6, atol=1e-7)
