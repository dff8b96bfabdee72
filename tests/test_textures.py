"""Constant-valued texture trees fold on the host (dartray_b200/host.py), following core/texture/constant_texture.dart:23-37,
textures/scale_texture.dart:26-34 and textures/mix_texture.dart:26-31 with RGBColor's float32 stores (rgb_color.dart:136-151).
The expected values are recomputed here with struct-based float32 rounding, independently of numpy's casts."""
import struct

import numpy as np
import pytest

from dartray_b200 import host as H


def f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


def spec(v):
    return [f32(c) for c in v]


def test_constant_texture_keeps_doubles_and_rounds_spectra():
    assert H.ConstantTexture(0.1).evaluate() == 0.1  # a float texture is a Dart double
    got = H.ConstantTexture((0.1, 0.2, 0.3)).evaluate()
    assert got.dtype == np.float32 and got.tolist() == spec((0.1, 0.2, 0.3))


def test_scale_texture_float_times_float_is_a_double_product():
    assert H.ScaleTexture(0.1, 0.7).evaluate() == 0.1 * 0.7


@pytest.mark.parametrize("swap", [False, True])
def test_scale_texture_spectrum_by_number(swap):
    s, k = (0.1, 0.2, 0.3), 0.7
    t = H.ScaleTexture(k, s) if swap else H.ScaleTexture(s, k)  # `t1 is num` -> t2 * t1, else t1 * t2: the same spectrum * num
    assert t.evaluate().tolist() == [f32(c * k) for c in spec(s)]


def test_scale_texture_spectrum_by_spectrum():
    a, b = (0.1, 0.2, 0.3), (0.9, 0.5, 0.25)
    assert H.ScaleTexture(a, b).evaluate().tolist() == [f32(x * y) for x, y in zip(spec(a), spec(b))]


def test_mix_texture_rounds_each_operator():
    a, b, amt = (0.1, 0.2, 0.3), (0.9, 0.8, 0.7), 0.3
    want = [f32(f32(x * (1.0 - amt)) + f32(y * amt)) for x, y in zip(spec(a), spec(b))]
    assert H.MixTexture(a, b, amt).evaluate().tolist() == want
    assert H.MixTexture(0.2, 0.6, amt).evaluate() == 0.2 * (1.0 - amt) + 0.6 * amt  # float textures stay doubles


def test_mix_texture_defaults_are_the_plugin_defaults():
    assert H.MixTexture().evaluate() == 0.5  # tex1 0, tex2 1, amount 0.5 (mix_texture.dart:33-38)
    assert H.ScaleTexture().evaluate() == 1.0  # scale_texture.dart:36-39


def test_nested_trees_fold_into_material_lobes():
    amount = H.ScaleTexture(0.5, 0.5)
    kd = H.MixTexture((0.1, 0.2, 0.3), H.ScaleTexture((0.9, 0.8, 0.7), 0.5), amount)
    folded = kd.evaluate()
    lobes = H.matte_lobes(kd=kd, sigma=H.ConstantTexture(20.0))
    assert len(lobes) == 1 and lobes[0]["kind"] == H.LOBE_OREN_NAYAR and lobes[0]["param"] == 20.0
    assert lobes[0]["rgb"].tolist() == folded.tolist()
    same = H.matte_lobes(kd=tuple(float(c) for c in folded), sigma=20.0)
    assert same[0]["rgb"].tolist() == lobes[0]["rgb"].tolist()
    # every material whose parameters are textures in the reference takes them
    assert H.plastic_lobes(kd=kd, ks=H.ConstantTexture(0.25), roughness=H.ScaleTexture(0.2, 0.5))[1]["param"] == 10.0
    assert H.glass_lobes(kr=H.ConstantTexture(1.0), kt=kd)[1]["rgb"].tolist() == folded.tolist()
    mixed = H.mix_lobes(H.matte_lobes(kd=0.5), H.mirror_lobes(kr=0.9), amount=H.ConstantTexture(0.25))
    assert mixed[0]["scale"].tolist() == spec((0.25,) * 3) and mixed[1]["scale"].tolist() == spec((0.75,) * 3)


def test_scene_builder_material_takes_a_texture():
    b = H.SceneBuilder()
    kd = H.ScaleTexture((0.5, 0.5, 0.5), (0.2, 0.4, 0.8))
    b.material(kd)
    b.material(tuple(float(c) for c in kd.evaluate()))
    assert b.materials[0] == b.materials[1]


def test_textures_that_read_the_hit_are_refused():
    with pytest.raises(H.GpuUnsupported):
        H.matte_lobes(kd=H.OpaqueTexture("imagemap"))
    with pytest.raises(H.GpuUnsupported):
        H.ScaleTexture(H.OpaqueTexture("checkerboard"), 0.5).evaluate()
    with pytest.raises(ValueError):
        H.MixTexture(0.1, (0.1, 0.2, 0.3), 0.5).evaluate()


def test_constant_mesh_alpha_other_than_zero_is_accepted():
    # Triangle.intersect drops a hit only where alpha == 0.0 (triangle.dart:139-151): a constant 0.5 changes nothing
    P, idx = [(0, 0, 0), (1, 0, 0), (0, 1, 0)], [(0, 1, 2)]
    a, b = H.SceneBuilder(), H.SceneBuilder()
    a.material((0.5, 0.5, 0.5)); b.material((0.5, 0.5, 0.5))
    a.mesh(P, idx)
    b.mesh(P, idx, alpha=H.ScaleTexture(0.5, H.ConstantTexture(1.0)))
    aa, bb = a.arrays(), b.arrays()
    assert aa.keys() == bb.keys()
    for k in aa:
        assert np.array_equal(np.asarray(aa[k]), np.asarray(bb[k])), k
    with pytest.raises(H.GpuUnsupported):
        H.SceneBuilder().mesh(P, idx, alpha=H.MixTexture(0.0, 1.0, 0.0))  # folds to 0: transparent everywhere
    with pytest.raises(H.GpuUnsupported):
        H.SceneBuilder().mesh(P, idx, alpha=H.OpaqueTexture("imagemap"))
