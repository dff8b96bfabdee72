// @dart=2.9
// GpuSamplerRenderer: the Renderer (lib/core/renderer.dart:27-35) that replaces SamplerRenderer
// (lib/renderers/sampler_renderer.dart) when libdartray_gpu.so can take the scene.  It walks the objects
// DartRay.worldEnd has already constructed, flattens them into typed arrays and makes ONE coarse call
// sequence through dart:ffi (include/drt.h).  Scene parsing, plugin names and the CLI are untouched.
//
// REVIEWED, NOT RUN: no Dart SDK exists in the build image (SURVEY.md section 8c).  The Python twin of this
// file — dartray_b200/host.py (SceneBuilder.arrays, upload_scene, configure_render, *_lobes) — performs the
// same flattening against the same ABI and is what tests/ exercise.
library dartray_gpu;

import 'dart:async';
import 'dart:ffi';
import 'dart:math' as Math;
import 'dart:typed_data';
import 'package:ffi/ffi.dart';
import 'package:dartray/dartray_core.dart';
import 'drt_ffi.dart';

part 'gpu_textures.dart';

/// Thrown by the flattener when the scene uses something the GPU path does not cover; the caller then
/// keeps the stock SamplerRenderer (see `makeRenderer` at the bottom).
class GpuUnsupported implements Exception {
  final String what;
  GpuUnsupported(this.what);
  String toString() => 'not on the GPU path: $what';
}

const int _LAMBERTIAN = 0, _OREN_NAYAR = 1, _MICROFACET_BLINN = 2, _SPEC_REFLECTION = 3, _SPEC_TRANSMISSION = 4, _FRESNEL_BLEND = 5;
const int _FRESNEL_NOOP = 0, _FRESNEL_DIELECTRIC = 1, _FRESNEL_CONDUCTOR = 2;

class _Lobe {
  int kind, fresnel = _FRESNEL_NOOP;
  Spectrum rgb, eta, k;
  double param = 0.0, ei = 1.0, et = 1.0;
  int wrap = 0;     // drt_set_lobe_wrappers: bit 0 BRDFToBTDF, bit 1 ScaledBxDF
  Spectrum scale;   // ScaledBxDF's s
  _Lobe(this.kind, this.rgb, {this.fresnel: _FRESNEL_NOOP, this.eta, this.k, this.param: 0.0, this.ei: 1.0, this.et: 1.0,
        this.wrap: 0, this.scale});
}

class _Arena {
  final List<Pointer> _owned = [];
  Pointer<Float> floats(List<double> v) {
    final p = calloc<Float>(Math.max(v.length, 1));
    p.asTypedList(v.length).setAll(0, v);
    _owned.add(p);
    return p;
  }
  Pointer<Double> doubles(List<double> v) {
    final p = calloc<Double>(Math.max(v.length, 1));
    p.asTypedList(v.length).setAll(0, v);
    _owned.add(p);
    return p;
  }
  Pointer<Int32> ints(List<int> v) {
    final p = calloc<Int32>(Math.max(v.length, 1));
    p.asTypedList(v.length).setAll(0, v);
    _owned.add(p);
    return p;
  }
  Pointer<Uint32> uints(List<int> v) {
    final p = calloc<Uint32>(Math.max(v.length, 1));
    p.asTypedList(v.length).setAll(0, v);
    _owned.add(p);
    return p;
  }
  Pointer<Uint64> uint64s(List<int> v) {
    final p = calloc<Uint64>(Math.max(v.length, 1));
    p.asTypedList(v.length).setAll(0, v);
    _owned.add(p);
    return p;
  }
  Pointer<Uint8> bytes(List<int> v) {
    final p = calloc<Uint8>(Math.max(v.length, 1));
    p.asTypedList(v.length).setAll(0, v);
    _owned.add(p);
    return p;
  }
  // drt_texture records (include/drt.h: 280 bytes each, natural C alignment — the layout host.TEX_DTYPE asserts in Python)
  Pointer<Uint8> textures(List<_TexNode> nodes) {
    final p = calloc<Uint8>(Math.max(nodes.length, 1) * 280);
    final bd = p.asTypedList(nodes.length * 280).buffer.asByteData();
    for (int i = 0; i < nodes.length; ++i) {
      final n = nodes[i];
      final int o = i * 280;
      final ints = [n.kind, n.spectrum, n.tex1, n.tex2, n.amount, n.mapping, n.imageWidth, n.imageHeight, n.imageChannels, n.imageWrap,
                    n.imageTrilinear, n.aaMethod];
      for (int k = 0; k < 12; ++k) bd.setInt32(o + 4 * k, ints[k], Endian.host);
      bd.setUint64(o + 48, n.imageOffset, Endian.host);
      for (int k = 0; k < 3; ++k) bd.setFloat64(o + 56 + 8 * k, n.value[k], Endian.host);
      for (int k = 0; k < 9; ++k) bd.setFloat64(o + 80 + 8 * k, n.value2[k], Endian.host);
      final d = [n.su, n.sv, n.du, n.dv, n.maxAnisotropy];
      for (int k = 0; k < 5; ++k) bd.setFloat64(o + 152 + 8 * k, d[k], Endian.host);
      for (int k = 0; k < 16; ++k) bd.setFloat32(o + 192 + 4 * k, n.worldToTexture[k], Endian.host);
      for (int k = 0; k < 3; ++k) {
        bd.setFloat32(o + 256 + 4 * k, n.v1[k], Endian.host);
        bd.setFloat32(o + 268 + 4 * k, n.v2[k], Endian.host);
      }
    }
    _owned.add(p);
    return p;
  }
  // drt_material_program records (48 bytes: kind, tex[8], bump, m1, m2); materials without a program get kind -1
  Pointer<Int32> programs(int nMaterials, Map<int, _Program> progs) {
    final v = new List<int>.filled(12 * nMaterials, -1);
    progs.forEach((i, pr) {
      v[12 * i] = pr.kind;
      for (int k = 0; k < 8; ++k) v[12 * i + 1 + k] = pr.tex[k];
      v[12 * i + 9] = pr.bump;
      v[12 * i + 10] = pr.m1;
      v[12 * i + 11] = pr.m2;
    });
    return ints(v);
  }
  void free() {
    for (final p in _owned) {
      calloc.free(p);
    }
    _owned.clear();
  }
}

class GpuSamplerRenderer extends Renderer {
  final Sampler sampler;
  final Camera camera;
  final SurfaceIntegrator surfaceIntegrator;
  final int taskNum, taskCount;
  final String libraryPath;
  /// true: the path integrator's vertex kernels run in float32 (drt_set_shading_precision(DRT_PRECISION_F32)): the image agrees
  /// with the CPU render per pixel within 3 sigma of the Monte Carlo noise instead of to ~1e-6 per sample, 1.5 x faster on B200.
  final bool float32Shading;

  GpuSamplerRenderer(this.sampler, this.camera, this.surfaceIntegrator, this.taskNum, this.taskCount,
                     {this.libraryPath: 'libdartray_gpu.so', this.float32Shading: false});

  // The per-ray entry points of the interface are not used on the GPU path (the whole loop of
  // sampler_renderer.dart:118-218 runs inside drt_render).
  Spectrum Li(Scene scene, RayDifferential ray, Sample sample, RNG rng, [Intersection isect, Spectrum T]) =>
      throw new UnsupportedError('GpuSamplerRenderer renders whole tasks; Li is not called per ray');
  Spectrum transmittance(Scene scene, RayDifferential ray, Sample sample, RNG rng) => new Spectrum(1.0);

  // ---- materials: Material.getBSDF with constant textures, as ordered BxDF lists ------------------------
  // ConstantTexture, and ScaleTexture / MixTexture trees over constants (textures/scale_texture.dart:26-34,
  // textures/mix_texture.dart:26-31): none of them reads the hit, so evaluating them once with no DifferentialGeometry gives the
  // value every Material.getBSDF call would see — through the reference's own operators, so the float32 stores are the reference's.
  static dynamic _fold(Texture t, String what) {
    if (t is ConstantTexture) {
      return t.value;
    }
    if (t is ScaleTexture) {
      _fold(t.tex1, what);
      _fold(t.tex2, what);
      return t.evaluate(null);
    }
    if (t is MixTexture) {
      _fold(t.tex1, what);
      _fold(t.tex2, what);
      _fold(t.amount, what);
      return t.evaluate(null);
    }
    throw new GpuUnsupported('$what is not a constant texture');
  }
  static Spectrum _const(Texture t, String what) {
    final v = _fold(t, what);
    return v is Spectrum ? new Spectrum.from(v) : new Spectrum(v.toDouble());
  }
  static double _constF(Texture t, String what) {
    final v = _fold(t, what);
    if (v is num) {
      return v.toDouble();
    }
    throw new GpuUnsupported('$what is a spectrum texture where a float texture is expected');
  }
  static double _blinn(double roughness) {  // 1 / roughness, then blinn.dart:24-28
    double e = 1.0 / roughness;
    return (e > 10000.0 || e.isNaN) ? 10000.0 : e;
  }

  static List<_Lobe> _lobes(Material m) {
    final out = <_Lobe>[];
    void noBump(Texture b) {
      if (b != null) {
        throw new GpuUnsupported('bump maps');
      }
    }
    if (m is MatteMaterial) {  // matte_material.dart:41-65
      noBump(m.bumpMap);
      Spectrum r = _const(m.Kd, 'matte Kd').clamp();
      double sig = _constF(m.sigma, 'matte sigma').clamp(0.0, 90.0);
      if (!r.isBlack()) {
        out.add(sig == 0.0 ? new _Lobe(_LAMBERTIAN, r) : new _Lobe(_OREN_NAYAR, r, param: sig));
      }
    } else if (m is MirrorMaterial) {  // mirror_material.dart:26-43
      noBump(m.bumpMap);
      Spectrum r = _const(m.Kr, 'mirror Kr').clamp();
      if (!r.isBlack()) {
        out.add(new _Lobe(_SPEC_REFLECTION, r));
      }
    } else if (m is GlassMaterial) {  // glass_material.dart:26-52
      noBump(m.bumpMap);
      double ior = _constF(m.index, 'glass index');
      Spectrum r = _const(m.Kr, 'glass Kr').clamp(), t = _const(m.Kt, 'glass Kt').clamp();
      if (!r.isBlack()) {
        out.add(new _Lobe(_SPEC_REFLECTION, r, fresnel: _FRESNEL_DIELECTRIC, ei: 1.0, et: ior));
      }
      if (!t.isBlack()) {
        out.add(new _Lobe(_SPEC_TRANSMISSION, t, fresnel: _FRESNEL_DIELECTRIC, ei: 1.0, et: ior));
      }
    } else if (m is PlasticMaterial) {  // plastic_material.dart:26-53
      noBump(m.bumpMap);
      Spectrum kd = _const(m.Kd, 'plastic Kd').clamp(), ks = _const(m.Ks, 'plastic Ks').clamp();
      if (!kd.isBlack()) {
        out.add(new _Lobe(_LAMBERTIAN, kd));
      }
      if (!ks.isBlack()) {
        out.add(new _Lobe(_MICROFACET_BLINN, ks, fresnel: _FRESNEL_DIELECTRIC, ei: 1.5, et: 1.0,
                          param: _blinn(_constF(m.roughness, 'plastic roughness'))));
      }
    } else if (m is MetalMaterial) {  // metal_material.dart:26-46
      noBump(m.bumpMap);
      out.add(new _Lobe(_MICROFACET_BLINN, new Spectrum(1.0), fresnel: _FRESNEL_CONDUCTOR,
                        eta: _const(m.eta, 'metal eta'), k: _const(m.k, 'metal k'),
                        param: _blinn(_constF(m.roughness, 'metal roughness'))));
    } else if (m is UberMaterial) {  // uber_material.dart:27-75
      noBump(m.bumpMap);
      Spectrum op = _const(m.opacity, 'uber opacity').clamp();
      if (!op.isValue(1.0)) {
        out.add(new _Lobe(_SPEC_TRANSMISSION, -op + Spectrum.ONE, fresnel: _FRESNEL_DIELECTRIC, ei: 1.0, et: 1.0));
      }
      double e = _constF(m.eta, 'uber index');
      Spectrum kd = op * _const(m.Kd, 'uber Kd').clamp();
      if (!kd.isBlack()) {
        out.add(new _Lobe(_LAMBERTIAN, kd));
      }
      Spectrum ks = op * _const(m.Ks, 'uber Ks').clamp();
      if (!ks.isBlack()) {
        out.add(new _Lobe(_MICROFACET_BLINN, ks, fresnel: _FRESNEL_DIELECTRIC, ei: e, et: 1.0,
                          param: _blinn(_constF(m.roughness, 'uber roughness'))));
      }
      Spectrum kr = op * _const(m.Kr, 'uber Kr').clamp();
      if (!kr.isBlack()) {
        out.add(new _Lobe(_SPEC_REFLECTION, kr, fresnel: _FRESNEL_DIELECTRIC, ei: e, et: 1.0));
      }
      Spectrum kt = op * _const(m.Kt, 'uber Kt').clamp();
      if (!kt.isBlack()) {
        out.add(new _Lobe(_SPEC_TRANSMISSION, kt, fresnel: _FRESNEL_DIELECTRIC, ei: e, et: 1.0));
      }
    } else if (m is SubstrateMaterial) {  // substrate_material.dart:46-68: one FresnelBlend over an Anisotropic distribution
      noBump(m.bumpMap);
      Spectrum d = _const(m.Kd, 'substrate Kd').clamp(), sp = _const(m.Ks, 'substrate Ks').clamp();
      if (!d.isBlack() || !sp.isBlack()) {
        out.add(new _Lobe(_FRESNEL_BLEND, d, eta: sp, param: _blinn(_constF(m.nu, 'substrate uroughness')),
                          ei: _blinn(_constF(m.nv, 'substrate vroughness'))));  // Rs in the eta slot; anisotropic.dart:30-37 clamps like Blinn
      }
    } else if (m is ShinyMetalMaterial) {  // shiny_metal_material.dart:42-64
      noBump(m.bumpMap);
      Spectrum spec = _const(m.Ks, 'shinymetal Ks').clamp(), r = _const(m.Kr, 'shinymetal Kr').clamp();
      final Spectrum k = new Spectrum(0.0);
      if (!spec.isBlack()) {
        out.add(new _Lobe(_MICROFACET_BLINN, new Spectrum(1.0), fresnel: _FRESNEL_CONDUCTOR,
                          eta: ShinyMetalMaterial.FresnelApproxEta(spec), k: k,
                          param: _blinn(_constF(m.roughness, 'shinymetal roughness'))));
      }
      if (!r.isBlack()) {
        out.add(new _Lobe(_SPEC_REFLECTION, new Spectrum(1.0), fresnel: _FRESNEL_CONDUCTOR,
                          eta: ShinyMetalMaterial.FresnelApproxEta(r), k: k));
      }
    } else if (m is TranslucentMaterial) {  // translucent_material.dart:47-90: the transmissive halves are BRDFToBTDF wrappers
      noBump(m.bumpMap);
      Spectrum r = _const(m.reflect, 'translucent reflect').clamp(), t = _const(m.transmit, 'translucent transmit').clamp();
      if (!(r.isBlack() && t.isBlack())) {
        Spectrum kd = _const(m.Kd, 'translucent Kd').clamp();
        if (!kd.isBlack()) {
          if (!r.isBlack()) out.add(new _Lobe(_LAMBERTIAN, r * kd));
          if (!t.isBlack()) out.add(new _Lobe(_LAMBERTIAN, t * kd, wrap: 1));
        }
        Spectrum ks = _const(m.Ks, 'translucent Ks').clamp();
        if (!ks.isBlack()) {
          final double e = _blinn(_constF(m.roughness, 'translucent roughness'));
          if (!r.isBlack()) out.add(new _Lobe(_MICROFACET_BLINN, r * ks, fresnel: _FRESNEL_DIELECTRIC, ei: 1.5, et: 1.0, param: e));
          if (!t.isBlack()) out.add(new _Lobe(_MICROFACET_BLINN, t * ks, fresnel: _FRESNEL_DIELECTRIC, ei: 1.5, et: 1.0, param: e, wrap: 1));
        }
      }
    } else if (m is MixMaterial) {  // mix_material.dart:36-50: ScaledBxDF around every BxDF of both BSDFs, one level deep
      Spectrum s1 = _const(m.scale, 'mix amount').clamp();
      Spectrum s2 = (Spectrum.ONE - s1).clamp();
      for (final pair in [[m.m1, s1], [m.m2, s2]]) {
        for (final _Lobe l in _lobes(pair[0])) {
          if ((l.wrap & 2) != 0) {
            throw new GpuUnsupported('a mix of mix materials');
          }
          l.wrap |= 2;
          l.scale = pair[1];
          out.add(l);
        }
      }
      if (out.length > 8) {
        throw new GpuUnsupported('more than 8 BxDFs in one BSDF');
      }
    } else {
      throw new GpuUnsupported('material ${m.runtimeType}');
    }
    return out;
  }

  static List<double> _rgb(Spectrum s) {
    if (s == null) {
      return [0.0, 0.0, 0.0];
    }
    final c = s.toRGB();
    return [c.c[0], c.c[1], c.c[2]];
  }

  // ---- the render call ------------------------------------------------------------------------------
  Future<OutputImage> render(Scene scene) {
    final a = new _Arena();
    final drt = new Drt(libraryPath);
    try {
      drt.create(0);
      _uploadScene(drt, a, scene);
      _configure(drt, a);
      drt.render(taskNum, taskCount);  // blocking: replaces _SamplerRendererTask.run
      final size = calloc<Int32>(4);
      drt.filmSize(size);
      final int left = size[0], top = size[1], w = size[2], h = size[3];
      calloc.free(size);
      final rgb = calloc<Float>(w * h * 3);
      drt.filmRead(rgb, nullptr, nullptr);  // replaces ImageFilm.writeImage (image_film.dart:268-299)
      final film = camera.film;
      final out = new OutputImage(left, top, w, h, film.xResolution, film.yResolution,
                                  new Float32List.fromList(rgb.asTypedList(w * h * 3)));
      calloc.free(rgb);
      return new Future.value(out);
    } finally {
      drt.destroy();
      a.free();
    }
  }

  void _uploadScene(Drt drt, _Arena a, Scene scene) {
    if (scene.volumeRegion != null) {
      throw new GpuUnsupported('participating media');
    }
    if (scene.aggregate is! BVHAccel) {
      throw new GpuUnsupported('accelerator ${scene.aggregate.runtimeType} (only bvh)');
    }
    final BVHAccel bvh = scene.aggregate;
    // primitives arrive in the order BVHAccel's constructor refined them (bvh_accel.dart:46-52): that IS the build
    // order; ids are assigned per kind in upload order: triangles, then spheres, then disks.
    final tris = <GeometricPrimitive>[], sphs = <GeometricPrimitive>[], dsks = <GeometricPrimitive>[];
    final quads = <GeometricPrimitive>[];  // cylinder / cone / paraboloid / hyperboloid: drt_set_quadrics, ids after the disks
    // TransformedPrimitives (animated shapes, object instances: transformed_primitive.dart): the primitive each one wraps is an
    // "object" — a GeometricPrimitive, or a BVHAccel whose refined primitives join the lists above — shared by identity between
    // the instances of one ObjectBegin block (drt_set_instances)
    final instances = <TransformedPrimitive>[];
    final objects = <Primitive>[];
    final objectIndex = new Map<Primitive, int>.identity();
    final allGeometric = <Primitive>[];
    for (final Primitive p in bvh.primitives) {
      if (p is TransformedPrimitive) {
        instances.add(p);
        if (!objectIndex.containsKey(p.primitive)) {
          objectIndex[p.primitive] = objects.length;
          objects.add(p.primitive);
          if (p.primitive is BVHAccel) {
            allGeometric.addAll((p.primitive as BVHAccel).primitives);
          } else {
            allGeometric.add(p.primitive);
          }
        }
      } else {
        allGeometric.add(p);
      }
    }
    for (final Primitive p in allGeometric) {
      if (p is! GeometricPrimitive) {
        throw new GpuUnsupported('primitive ${p.runtimeType} (an instance inside an object, or an accelerator other than bvh inside one)');
      }
      final GeometricPrimitive g = p;
      if (g.shape is Triangle) {
        tris.add(g);
      } else if (g.shape is Sphere) {
        sphs.add(g);
      } else if (g.shape is Disk) {
        dsks.add(g);
      } else if (g.shape is Cylinder || g.shape is Cone || g.shape is Paraboloid || g.shape is Hyperboloid) {
        quads.add(g);
      } else {
        throw new GpuUnsupported('shape ${g.shape.runtimeType}');
      }
    }
    final ids = new Map<GeometricPrimitive, int>.identity();
    for (int i = 0; i < tris.length; ++i) ids[tris[i]] = i;
    for (int i = 0; i < sphs.length; ++i) ids[sphs[i]] = tris.length + i;
    for (int i = 0; i < dsks.length; ++i) ids[dsks[i]] = tris.length + sphs.length + i;
    for (int i = 0; i < quads.length; ++i) ids[quads[i]] = tris.length + sphs.length + dsks.length + i;

    // materials and area lights by identity
    final materials = <Material>[], lights = scene.lights;
    final matIndex = new Map<Material, int>.identity(), lightIndex = new Map<Light, int>.identity();
    for (int i = 0; i < lights.length; ++i) lightIndex[lights[i]] = i;
    int matOf(GeometricPrimitive g) => matIndex.putIfAbsent(g.material, () {
      materials.add(g.material);
      return materials.length - 1;
    });
    int lightOf(GeometricPrimitive g) => g.areaLight == null ? -1 : lightIndex[g.areaLight];

    // triangles: three world-space float32 vertices each (TriangleMesh.point, triangle_mesh.dart:39-42); no sharing
    final P = <double>[], idx = <int>[], tm = <int>[], tl = <int>[], tr = <int>[];
    final triKey = new Map<TriangleMesh, Map<int, int>>.identity();
    final meshes = <TriangleMesh>[], meshOfTri = <int>[];
    final meshIndex = new Map<TriangleMesh, int>.identity();
    final vN = <double>[], vS = <double>[], vUV = <double>[];
    for (int i = 0; i < tris.length; ++i) {
      final Triangle t = tris[i].shape;
      if (t.mesh.alphaTexture != null) {
        // Triangle.intersect / intersectP drop a hit only where alpha evaluates to exactly 0 (triangle.dart:139-151,196-237):
        // a constant-valued alpha other than 0 changes nothing
        if (_constF(t.mesh.alphaTexture, 'a mesh alpha texture') == 0.0) {
          throw new GpuUnsupported('a triangle mesh whose alpha texture is 0 everywhere');
        }
      }
      // per-vertex N / S / uv (triangle_mesh.dart:24-28) travel as they are stored: object space; drt_set_mesh_shading
      final TriangleMesh mesh = t.mesh;
      final int mi = meshIndex.putIfAbsent(mesh, () {
        meshes.add(mesh);
        return meshes.length - 1;
      });
      meshOfTri.add(mi);
      for (int k = 0; k < 3; ++k) {
        final int v = mesh.vertexIndex[3 * t.index + k];
        final Point p = mesh.point(v);
        P..add(p.x)..add(p.y)..add(p.z);
        idx.add(3 * i + k);
        if (mesh.n != null) { vN..add(mesh.n[v].x)..add(mesh.n[v].y)..add(mesh.n[v].z); } else { vN..add(0.0)..add(0.0)..add(0.0); }
        if (mesh.s != null) { vS..add(mesh.s[v].x)..add(mesh.s[v].y)..add(mesh.s[v].z); } else { vS..add(0.0)..add(0.0)..add(0.0); }
        if (mesh.uvs != null) { vUV..add(mesh.uvs[2 * v])..add(mesh.uvs[2 * v + 1]); } else { vUV..add(0.0)..add(0.0); }
      }
      tm.add(matOf(tris[i]));
      tl.add(lightOf(tris[i]));
      tr.add(t.reverseOrientation ? 1 : 0);
      triKey.putIfAbsent(t.mesh, () => <int, int>{})[t.index] = i;
    }
    drt.setTriangles(a.floats(P), P.length ~/ 3, a.uints(idx), tris.length, a.ints(tm), a.ints(tl), a.bytes(tr));
    if (meshes.any((m) => m.n != null || m.s != null || m.uvs != null)) {
      final mo2w = <double>[], mw2o = <double>[], flags = <int>[];
      for (final m in meshes) {
        mo2w.addAll(m.objectToWorld.m.data);
        mw2o.addAll(m.worldToObject.m.data);
        flags.add((m.n != null ? 1 : 0) | (m.s != null ? 2 : 0) | (m.uvs != null ? 4 : 0));
      }
      drt.setMeshShading(meshes.any((m) => m.n != null) ? a.floats(vN) : nullptr, meshes.any((m) => m.s != null) ? a.floats(vS) : nullptr,
                         meshes.any((m) => m.uvs != null) ? a.floats(vUV) : nullptr, a.uints(meshOfTri), meshes.length,
                         a.floats(mo2w), a.floats(mw2o), a.bytes(flags));
    }

    void quadrics(List<GeometricPrimitive> prims, bool disk) {
      if (prims.isEmpty) {
        return;
      }
      final o2w = <double>[], w2o = <double>[], prm = <double>[], qm = <int>[], ql = <int>[], qr = <int>[];
      for (final g in prims) {
        o2w.addAll(g.shape.objectToWorld.m.data);
        w2o.addAll(g.shape.worldToObject.m.data);
        if (disk) {
          final Disk d = g.shape;  // the PARAMETERS of disk.dart:157-166: phiMax is stored in radians there
          prm..add(d.height)..add(d.radius)..add(d.innerRadius)..add(Degrees(d.phiMax));
        } else {
          final Sphere s = g.shape;  // sphere.dart:24-32 stores clamped zmin / zmax and phiMax in radians
          prm..add(s.radius)..add(s.zmin)..add(s.zmax)..add(Degrees(s.phiMax));
        }
        qm.add(matOf(g));
        ql.add(lightOf(g));
        qr.add(g.shape.reverseOrientation ? 1 : 0);
      }
      if (disk) {
        drt.setDisks(prims.length, a.floats(o2w), a.floats(w2o), a.doubles(prm), a.ints(qm), a.ints(ql), a.bytes(qr));
      } else {
        drt.setSpheres(prims.length, a.floats(o2w), a.floats(w2o), a.doubles(prm), a.ints(qm), a.ints(ql), a.bytes(qr));
      }
    }
    quadrics(sphs, false);
    quadrics(dsks, true);
    // the remaining quadrics, one drt_set_quadrics call each (ids in list order).  The fields hold what the constructors
    // stored (cylinder.dart:24-31, cone.dart:23-27, paraboloid.dart:23-29, hyperboloid.dart:23-49): phiMax in radians, the
    // hyperboloid's points after its p2.z == 0 swap — handing those back reproduces the same object
    for (final g in quads) {
      final prm = new List<double>.filled(8, 0.0);
      int kind;
      final Shape sh = g.shape;
      if (sh is Cylinder) {
        kind = 2;
        prm[0] = sh.radius; prm[1] = sh.zmin; prm[2] = sh.zmax; prm[3] = Degrees(sh.phiMax);
      } else if (sh is Cone) {
        kind = 3;
        prm[0] = sh.height; prm[1] = sh.radius; prm[2] = Degrees(sh.phiMax);
      } else if (sh is Paraboloid) {
        kind = 4;
        prm[0] = sh.radius; prm[1] = sh.zmin; prm[2] = sh.zmax; prm[3] = Degrees(sh.phiMax);
      } else {
        final Hyperboloid hy = sh;
        kind = 5;
        prm[0] = hy.p1.x; prm[1] = hy.p1.y; prm[2] = hy.p1.z; prm[3] = hy.p2.x; prm[4] = hy.p2.y; prm[5] = hy.p2.z;
        prm[6] = Degrees(hy.phiMax);
      }
      drt.setQuadrics(kind, 1, a.floats(sh.objectToWorld.m.data), a.floats(sh.worldToObject.m.data), a.doubles(prm),
                      a.ints([matOf(g)]), a.ints([lightOf(g)]), a.bytes([sh.reverseOrientation ? 1 : 0]));
    }

    final int nGeometric = tris.length + sphs.length + dsks.length + quads.length;
    if (instances.isNotEmpty) {
      final offsets = <int>[0], prims = <int>[], split = <int>[], maxPrims = <int>[];
      for (final Primitive ob in objects) {
        if (ob is BVHAccel) {
          for (final Primitive q in ob.primitives) prims.add(ids[q]);  // the order its constructor refined them in
          split.add(ob.splitMethod);
          maxPrims.add(ob.maxPrimsInNode);
        } else {
          prims.add(ids[ob]);
          split.add(2);
          maxPrims.add(1);
        }
        offsets.add(prims.length);
      }
      final ob = <int>[], m0 = <double>[], i0 = <double>[], m1 = <double>[], i1 = <double>[], tt = <double>[];
      for (final TransformedPrimitive t in instances) {
        final AnimatedTransform w2p = t.worldToPrimitive;
        ob.add(objectIndex[t.primitive]);
        m0.addAll(w2p.startTransform.m.data); i0.addAll(w2p.startTransform.mInv.data);
        m1.addAll(w2p.endTransform.m.data); i1.addAll(w2p.endTransform.mInv.data);
        tt..add(w2p.startTime)..add(w2p.endTime);
      }
      drt.setInstances(objects.length, a.uints(offsets), a.uints(prims), a.ints(split), a.ints(maxPrims), instances.length, a.uints(ob),
                       a.floats(m0), a.floats(i0), a.floats(m1), a.floats(i1), a.doubles(tt));
    }
    final order = <int>[];
    int nextInstance = 0;
    for (final Primitive p in bvh.primitives) order.add(p is TransformedPrimitive ? nGeometric + (nextInstance++) : ids[p]);
    drt.setBuildOrder(a.uints(order), order.length);
    drt.buildBvh(bvh.splitMethod, bvh.maxPrimsInNode);

    // materials: constant parameters flatten into BxDF lists (_lobes); a material with a texture that reads the hit point or with
    // a bump map becomes a program (gpu_textures.dart) and owns an empty list.  A MixMaterial's two materials join the table.
    final flattener = new TextureFlattener();
    final programs = <int, _Program>{};
    int indexOfMaterial(Material sub) {
      int i = materials.indexOf(sub);
      if (i < 0) {
        materials.add(sub);
        i = materials.length - 1;
      }
      return i;
    }
    final lobeLists = <List<_Lobe>>[];
    for (int i = 0; i < materials.length; ++i) {  // the list may grow while a mix is flattened
      try {
        lobeLists.add(_lobes(materials[i]));
      } on GpuUnsupported {
        programs[i] = flattener.program(materials[i], indexOfMaterial);  // throws GpuUnsupported itself for what it cannot take
        lobeLists.add(<_Lobe>[]);
      }
    }
    final offsets = <int>[0], kind = <int>[], fres = <int>[], rgb = <double>[], eta = <double>[], kk = <double>[], scal = <double>[];
    for (final ll in lobeLists) {
      for (final l in ll) {
        kind.add(l.kind);
        fres.add(l.fresnel);
        rgb.addAll(_rgb(l.rgb));
        eta.addAll(_rgb(l.eta));
        kk.addAll(_rgb(l.k));
        scal..add(l.param)..add(l.ei)..add(l.et);
      }
      offsets.add(kind.length);
    }
    if (materials.isNotEmpty) {
      drt.setMaterialLobes(materials.length, a.uints(offsets), a.ints(kind), a.floats(rgb), a.ints(fres), a.floats(eta),
                           a.floats(kk), a.doubles(scal));
      final wraps = <int>[], scales = <double>[];
      for (final ll in lobeLists) {
        for (final l in ll) {
          wraps.add(l.wrap);
          scales.addAll(_rgb(l.scale == null ? new Spectrum(1.0) : l.scale));
        }
      }
      if (wraps.any((w) => w != 0)) {
        drt.setLobeWrappers(wraps.length, a.ints(wraps), a.floats(scales));
      }
      if (flattener.measuredTables.isNotEmpty) {
        final mt = flattener.measuredTables;
        drt.setMeasured(mt.length, a.ints([for (final t in mt) t.kind]), a.ints([for (final t in mt) ...t.dims]),
                        a.uint64s([for (final t in mt) t.offset]), a.floats(flattener.measuredData), flattener.measuredData.length);
      }
      if (programs.isNotEmpty) {
        drt.setTextures(a.textures(flattener.nodes), flattener.nodes.length, a.floats(flattener.texels), flattener.texels.length);
        drt.setMaterialPrograms(a.programs(materials.length, programs), materials.length);
      }
    }

    // lights
    final lk = <int>[], lL = <double>[], lpos = <double>[], lns = <int>[], so = <int>[0], sp = <int>[];
    final w2l = <double>[], cosines = <double>[];
    bool anySpot = false;
    final infinite = <int>[], mapped = <int>[];
    for (final Light l in lights) {
      lns.add(l.nSamples);
      w2l.addAll(l.worldToLight.m.data);
      if (l is DiffuseAreaLight) {
        lk.add(0);
        lL.addAll(_rgb(l.Lemit));
        lpos.addAll([0.0, 0.0, 0.0]);
        cosines.addAll([0.0, 0.0]);
        for (final Shape s in l.shapeSet.shapes) {  // shape_set.dart:26-41: its own refinement of the light's shape
          if (s is Triangle) {
            sp.add(triKey[s.mesh][s.index]);
          } else {
            final g = (sphs + dsks + quads).firstWhere((q) => identical(q.shape, s));
            sp.add(ids[g]);
          }
        }
      } else if (l is PointLight) {
        lk.add(1);
        lL.addAll(_rgb(l.intensity));
        lpos..add(l.lightPos.x)..add(l.lightPos.y)..add(l.lightPos.z);
        cosines.addAll([0.0, 0.0]);
      } else if (l is DistantLight) {
        lk.add(2);
        lL.addAll(_rgb(l.L));
        lpos..add(l.lightDir.x)..add(l.lightDir.y)..add(l.lightDir.z);
        cosines.addAll([0.0, 0.0]);
      } else if (l is SpotLight) {
        lk.add(3);
        anySpot = true;
        lL.addAll(_rgb(l.intensity));
        lpos..add(l.lightPos.x)..add(l.lightPos.y)..add(l.lightPos.z);
        cosines..add(l.cosTotalWidth)..add(l.cosFalloffStart);
      } else if (l is ProjectionLight) {
        lk.add(5);
        mapped.add(lk.length - 1);
        lL.addAll(_rgb(l.intensity));
        lpos..add(l.lightPos.x)..add(l.lightPos.y)..add(l.lightPos.z);
        cosines.addAll([0.0, 0.0]);
      } else if (l is GoniometricLight) {
        lk.add(6);
        mapped.add(lk.length - 1);
        lL.addAll(_rgb(l.intensity));
        lpos..add(l.lightPos.x)..add(l.lightPos.y)..add(l.lightPos.z);
        cosines.addAll([0.0, 0.0]);
      } else if (l is InfiniteAreaLight) {
        lk.add(4);
        infinite.add(lk.length - 1);
        lL.addAll(_rgb(l.L));
        lpos.addAll([0.0, 0.0, 0.0]);
        cosines.addAll([0.0, 0.0]);
      } else {
        throw new GpuUnsupported('light ${l.runtimeType}');
      }
      so.add(sp.length);
    }
    drt.setLights(lights.length, a.ints(lk), a.floats(lL), a.floats(lpos), a.ints(lns), a.uints(so), a.uints(sp));
    if (anySpot) {
      drt.setSpotParams(lights.length, a.floats(w2l), a.doubles(cosines));
    }
    for (final int i in mapped) {  // projection / goniometric lights: level 0 of their MIPMap (or none) + their transforms
      final Light l = lights[i];
      final MIPMap map = l is ProjectionLight ? l.projectionMap : (l as GoniometricLight).mipmap;
      final SpectrumImage level0 = map == null ? null : map.pyramid[0];
      if (l is ProjectionLight) {
        drt.setLightMap(i, level0 == null ? 0 : level0.width, level0 == null ? 0 : level0.height,
                        level0 == null ? nullptr : a.floats(level0.data), a.floats(l.worldToLight.m.data),
                        a.floats(l.lightProjection.m.data), a.doubles([l.screenX0, l.screenX1, l.screenY0, l.screenY1]), l.hither);
      } else {
        drt.setLightMap(i, level0 == null ? 0 : level0.width, level0 == null ? 0 : level0.height,
                        level0 == null ? nullptr : a.floats(level0.data), a.floats(l.worldToLight.m.data), nullptr, nullptr, 0.0);
      }
    }
    for (final int i in infinite) {
      // level 0 of the light's own MIPMap: already resampled to a power of two and multiplied by L by the reference's
      // constructor (infinite_area_light.dart:43-60, mipmap.dart:72-139); drt_set_infinite_light rebuilds the rest
      final InfiniteAreaLight l = lights[i];
      final SpectrumImage level0 = l.radianceMap.pyramid[0];
      drt.setInfiniteLight(i, level0.width, level0.height, a.floats(level0.data), a.floats(l.lightToWorld.m.data),
                           a.floats(l.worldToLight.m.data));
    }
  }

  void _configure(Drt drt, _Arena a) {
    // camera (perspective_camera.dart:46-57, orthographic_camera.dart, environment_camera.dart)
    final c2w = camera.cameraToWorld.startTransform.m.data;
    if (camera is ProjectiveCamera) {
      final ProjectiveCamera pc = camera;
      drt.setCamera(a.floats(pc.rasterToCamera.m.data), a.floats(c2w), pc.lensRadius, pc.focalDistance,
                    camera.shutterOpen, camera.shutterClose);
      drt.setCameraKind(camera is OrthographicCamera ? 1 : 0);
    } else if (camera is EnvironmentCamera) {
      drt.setCamera(a.floats(new Matrix4x4().data), a.floats(c2w), 0.0, 1.0e30, camera.shutterOpen, camera.shutterClose);
      drt.setCameraKind(2);
    } else {
      throw new GpuUnsupported('camera ${camera.runtimeType}');
    }
    // an animated camera: the end-time matrix and the transform times; the library decomposes and interpolates per camera ray
    final AnimatedTransform cw = camera.cameraToWorld;
    if (cw.actuallyAnimated) {
      drt.setCameraMotion(a.floats(cw.endTransform.m.data), cw.startTime, cw.endTime);
    }

    // film: the 16 x 16 filter table ImageFilm builds (image_film.dart:74-82), recomputed through the public Filter API
    if (camera.film is! ImageFilm) {
      throw new GpuUnsupported('film ${camera.film.runtimeType}');
    }
    final ImageFilm film = camera.film;
    final Filter f = film.filter;
    final table = <double>[];
    for (int y = 0; y < 16; ++y) {
      final double fy = (y + 0.5) * f.yWidth / 16;
      for (int x = 0; x < 16; ++x) {
        final double fx = (x + 0.5) * f.xWidth / 16;
        table.add(f.evaluate(fx, fy));
      }
    }
    drt.setFilm(film.xResolution, film.yResolution, a.doubles(film.cropWindow), f.xWidth, f.yWidth, a.floats(table));

    // sampler: seed = task number (sampler_renderer.dart:137)
    int order(PixelSampler p) => p is LinearPixelSampler ? 0 : 1;
    if (sampler is LowDiscrepancySampler) {
      final LowDiscrepancySampler s = sampler;
      drt.setSampler(0, 1, 1, s.nPixelSamples, 1, order(s.pixels), 32, taskNum);
    } else if (sampler is StratifiedSampler) {
      final StratifiedSampler s = sampler;
      drt.setSampler(1, s.xPixelSamples, s.yPixelSamples, s.xPixelSamples * s.yPixelSamples, s.jitterSamples ? 1 : 0,
                     order(s.pixels), 32, taskNum);
    } else if (sampler is RandomSampler) {
      final RandomSampler s = sampler;
      drt.setSampler(2, 1, 1, s.samplesPerPixel, 1, order(s.pixels), 32, taskNum);
    } else if (sampler is HaltonSampler) {
      // wantedSamples = samplesPerPixel * max(width, height)^2 (halton_sampler.dart:32-38): the library derives it from spp
      drt.setSampler(3, 1, 1, sampler.samplesPerPixel, 1, 1, 32, taskNum);
    } else if (sampler is BestCandidateSampler) {
      // the 4096 x 5 pattern is data of the reference (best_candidate_sampler.dart:163-4258): hand it over
      drt.setSampleTable(a.doubles(BestCandidateSampler.SAMPLE_TABLE), BestCandidateSampler.SAMPLE_TABLE_SIZE);
      drt.setSampler(5, 1, 1, sampler.samplesPerPixel, 1, 1, 32, taskNum);
    } else if (sampler is AdaptiveSampler) {
      final AdaptiveSampler s = sampler;  // minSamples / maxSamples are already normalised (adaptive_sampler.dart:52-84)
      drt.setSampler(4, s.minSamples, s.maxSamples, s.maxSamples, s.method, order(s.pixels), 32, taskNum);
    } else {
      throw new GpuUnsupported('sampler ${sampler.runtimeType}');
    }

    // surface integrator
    if (surfaceIntegrator is PathIntegrator) {
      final PathIntegrator p = surfaceIntegrator;
      drt.setIntegrator(0, p.maxDepth, 0, 1, 0.0, double.infinity);
      drt.setShadingPrecision(float32Shading ? 1 : 0);
    } else if (surfaceIntegrator is AmbientOcclusionIntegrator) {
      final AmbientOcclusionIntegrator ao = surfaceIntegrator;
      drt.setIntegrator(1, 0, 0, ao.nSamples, ao.minDist, ao.maxDist);
    } else if (surfaceIntegrator is DirectLightingIntegrator) {
      final DirectLightingIntegrator d = surfaceIntegrator;
      drt.setIntegrator(2, d.maxDepth, d.strategy, 1, 0.0, double.infinity);
    } else if (surfaceIntegrator is WhittedIntegrator) {
      final WhittedIntegrator w = surfaceIntegrator;
      drt.setIntegrator(3, w.maxDepth, 0, 1, 0.0, double.infinity);
    } else {
      throw new GpuUnsupported('surface integrator ${surfaceIntegrator.runtimeType}');
    }
  }
}

/// What DartRay._makeRenderer's else-branch (lib/dartray/dartray.dart:756-761) calls instead of constructing a
/// SamplerRenderer directly: the GPU renderer wrapped so that a scene it cannot take (GpuUnsupported) or a
/// DRT_E_UNSUPPORTED / missing library falls back to the stock Dart renderer with a warning.
class GpuOrDartRenderer extends Renderer {
  final GpuSamplerRenderer gpu;
  final SamplerRenderer dart;
  GpuOrDartRenderer(this.gpu, this.dart);

  Future<OutputImage> render(Scene scene) {
    try {
      return gpu.render(scene);
    } on GpuUnsupported catch (e) {
      LogWarning('GPU renderer: $e; rendering with the Dart SamplerRenderer');
    } on DrtError catch (e) {
      if (e.code != -6) {  // DRT_E_UNSUPPORTED falls back, anything else is a real error (log.dart:44-46)
        LogSevere(e.toString());
      }
      LogWarning('GPU renderer: ${e.message}; rendering with the Dart SamplerRenderer');
    } on ArgumentError catch (e) {  // DynamicLibrary.open failed
      LogWarning('GPU renderer: libdartray_gpu.so not loadable ($e); rendering with the Dart SamplerRenderer');
    }
    return dart.render(scene);
  }

  Spectrum Li(Scene scene, RayDifferential ray, Sample sample, RNG rng, [Intersection isect, Spectrum T]) =>
      dart.Li(scene, ray, sample, rng, isect, T);
  Spectrum transmittance(Scene scene, RayDifferential ray, Sample sample, RNG rng) =>
      dart.transmittance(scene, ray, sample, rng);
}
