"""BASELINE.json configs[0] / SURVEY §8d config 1: the Cornell scene of web/scenes/cornell-path.pbrt at 256x256,
path integrator, lowdiscrepancy 4 spp, box filter, tile pixel order, task 0 of 1.

Two variants: `cornell_path` — the scene as shipped, with its DISK area light (cornell-path.pbrt:15-19; `Shape
"disk"`, SURVEY §8f f2, is on the GPU path) — and `cornell_synth`, the same geometry with the disk replaced by a
2-triangle quad light (the config-4 benchmark scene).  The reference itself cannot be run (no Dart VM in the image):
the anchor is the CPU oracle in the reference's SERIAL stream mode; the GPU replays the KEYED streams."""
import numpy as np
import pytest

from dartray_b200 import capi, host, scenes
from tests.oracle_lib import Oracle

FILM = host.Film(256, 256)
INTEG = host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5)


def _oracle(mode, seed=0):
    sb, cam = scenes.cornell_synth()
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, FILM, host.Sampler(kind=host.SAMPLER_LD, spp=4, pixel_order=1, seed=seed, rng_mode=mode), INTEG)
    o.render(0, 1, 8 if mode == host.RNG_KEYED else 1)
    return o.film_read()["rgb"]


def _blocks(img, b=16):
    h, w, c = img.shape
    return img.reshape(h // b, b, w // b, b, c).mean(axis=(1, 3))


def test_config1_keyed_streams_match_the_serial_reference_stream_statistically():
    serial, keyed, keyed2 = _oracle(host.RNG_SERIAL), _oracle(host.RNG_KEYED), _oracle(host.RNG_KEYED, seed=1)
    assert abs(keyed.mean() - serial.mean()) <= 0.01 * serial.mean()  # image-wide mean within 1 %
    # 16x16-pixel block means: the serial-vs-keyed spread is the same size as the keyed-vs-keyed (seed) spread
    noise = np.abs(_blocks(keyed) - _blocks(keyed2)).mean()
    diff = np.abs(_blocks(keyed) - _blocks(serial)).mean()
    assert diff <= 1.5 * noise
    # per pixel: within 3 sigma of the Monte Carlo noise (sigma estimated from the two independent keyed renders)
    sigma = np.abs(keyed - keyed2).mean() / 1.128 + 1e-6  # E|a-b| = 1.128 sigma for two normal draws
    assert (np.abs(keyed - serial) <= 3 * sigma * np.sqrt(2)).mean() > 0.9


@pytest.mark.gpu
def test_config1_real_scene_with_disk_light_gpu_matches_oracle():
    """web/scenes/cornell-path.pbrt as shipped (disk area light), film / sampler overridden as configs[0] says."""
    sb, cam = scenes.cornell_path()
    arrays = sb.arrays()
    g, o = capi.Context(0), Oracle()
    for c in (g, o):
        host.upload_scene(c, arrays)
        host.configure_render(c, cam, FILM, host.Sampler(kind=host.SAMPLER_LD, spp=4, pixel_order=1), INTEG)
    g.render(0, 1)
    o.render(0, 1, 8)
    rgb, ref = g.film_read()["rgb"], o.film_read()["rgb"]
    err = np.abs(rgb - ref) / np.maximum(np.abs(ref), 1e-3)
    assert abs(rgb.mean() - ref.mean()) <= 1e-4 * ref.mean()
    assert err.max() <= 1e-3
    sg, so = g.render_stats(), o.render_stats()
    assert abs(sg["shadow_rays"] - so["shadow_rays"]) <= 1e-4 * so["shadow_rays"]


@pytest.mark.gpu
def test_config1_gpu_matches_the_keyed_oracle():
    sb, cam = scenes.cornell_synth()
    g = capi.Context(0)
    host.upload_scene(g, sb.arrays())
    host.configure_render(g, cam, FILM, host.Sampler(kind=host.SAMPLER_LD, spp=4, pixel_order=1), INTEG)
    g.render(0, 1)
    rgb, ref = g.film_read()["rgb"], _oracle(host.RNG_KEYED)
    err = np.abs(rgb - ref) / np.maximum(np.abs(ref), 1e-3)
    assert abs(rgb.mean() - ref.mean()) <= 1e-4 * ref.mean()
    assert err.max() <= 1e-3
