// @dart=2.9
// dart:ffi binding of libdartray_gpu.so (include/drt.h).  One typedef pair per C entry point; the
// signatures are the header's, argument for argument.  REVIEWED, NOT RUN: the build image has no Dart
// SDK (SURVEY.md section 8c); dartray_b200/capi.py binds the same symbol table with ctypes and is the
// tested twin (tests/test_abi.py checks every symbol declared in include/drt.h is exported).
library dartray_gpu_ffi;

import 'dart:ffi';
import 'package:ffi/ffi.dart';

class DrtHit extends Struct { // drt_hit
  @Float() double t;
  @Float() double b1;
  @Float() double b2;
  @Int32() int prim;
}

class DrtRenderStats extends Struct { // drt_render_stats
  @Uint64() int cameraSamples;
  @Uint64() int closestRays;
  @Uint64() int shadowRays;
  @Uint64() int zeroedSamples;
}

typedef _Ctx = Pointer<Void>;
typedef _PF = Pointer<Float>;
typedef _PD = Pointer<Double>;
typedef _PI = Pointer<Int32>;
typedef _PU = Pointer<Uint32>;
typedef _PB = Pointer<Uint8>;

typedef _CreateC = _Ctx Function(Int32);
typedef _CreateD = _Ctx Function(int);
typedef _DestroyC = Void Function(_Ctx);
typedef _DestroyD = void Function(_Ctx);
typedef _LastErrorC = Pointer<Utf8> Function(_Ctx);
typedef _SetTrianglesC = Int32 Function(_Ctx, _PF, Uint32, _PU, Uint32, _PI, _PI, _PB);
typedef _SetTrianglesD = int Function(_Ctx, _PF, int, _PU, int, _PI, _PI, _PB);
typedef _SetQuadricsC = Int32 Function(_Ctx, Uint32, _PF, _PF, _PD, _PI, _PI, _PB);
typedef _SetQuadricsD = int Function(_Ctx, int, _PF, _PF, _PD, _PI, _PI, _PB);
typedef _SetKindQuadricsC = Int32 Function(_Ctx, Int32, Uint32, _PF, _PF, _PD, _PI, _PI, _PB);
typedef _SetKindQuadricsD = int Function(_Ctx, int, int, _PF, _PF, _PD, _PI, _PI, _PB);
typedef _SetTableC = Int32 Function(_Ctx, _PD, Uint32);
typedef _SetTableD = int Function(_Ctx, _PD, int);
typedef _SetLightMapC = Int32 Function(_Ctx, Uint32, Int32, Int32, _PF, _PF, _PF, _PD, Double);
typedef _SetLightMapD = int Function(_Ctx, int, int, int, _PF, _PF, _PF, _PD, double);
typedef _SetTexturesC = Int32 Function(_Ctx, Uint32, _PB, _PF, Uint64);
typedef _SetTexturesD = int Function(_Ctx, int, _PB, _PF, int);
typedef _SetMeasuredC = Int32 Function(_Ctx, Uint32, _PI, _PI, Pointer<Uint64>, _PF, Uint64);
typedef _SetMeasuredD = int Function(_Ctx, int, _PI, _PI, Pointer<Uint64>, _PF, int);
typedef _SetProgramsC = Int32 Function(_Ctx, Uint32, _PI);
typedef _SetProgramsD = int Function(_Ctx, int, _PI);
typedef _SetWrappersC = Int32 Function(_Ctx, Uint32, _PI, _PF);
typedef _SetWrappersD = int Function(_Ctx, int, _PI, _PF);
typedef _SetInfiniteC = Int32 Function(_Ctx, Uint32, Int32, Int32, _PF, _PF, _PF);
typedef _SetInfiniteD = int Function(_Ctx, int, int, int, _PF, _PF, _PF);
typedef _SetMeshShadingC = Int32 Function(_Ctx, _PF, _PF, _PF, _PU, Uint32, _PF, _PF, _PB);
typedef _SetMeshShadingD = int Function(_Ctx, _PF, _PF, _PF, _PU, int, _PF, _PF, _PB);
typedef _SetInstancesC = Int32 Function(_Ctx, Uint32, _PU, _PU, _PI, _PI, Uint32, _PU, _PF, _PF, _PF, _PF, _PD);
typedef _SetInstancesD = int Function(_Ctx, int, _PU, _PU, _PI, _PI, int, _PU, _PF, _PF, _PF, _PF, _PD);
typedef _SetOrderC = Int32 Function(_Ctx, _PU, Uint32);
typedef _SetOrderD = int Function(_Ctx, _PU, int);
typedef _BuildC = Int32 Function(_Ctx, Int32, Int32);
typedef _BuildD = int Function(_Ctx, int, int);
typedef _SetMaterialsC = Int32 Function(_Ctx, Uint32, _PI, _PF, _PF);
typedef _SetMaterialsD = int Function(_Ctx, int, _PI, _PF, _PF);
typedef _SetLobesC = Int32 Function(_Ctx, Uint32, _PU, _PI, _PF, _PI, _PF, _PF, _PD);
typedef _SetLobesD = int Function(_Ctx, int, _PU, _PI, _PF, _PI, _PF, _PF, _PD);
typedef _SetLightsC = Int32 Function(_Ctx, Uint32, _PI, _PF, _PF, _PI, _PU, _PU);
typedef _SetLightsD = int Function(_Ctx, int, _PI, _PF, _PF, _PI, _PU, _PU);
typedef _SetSpotC = Int32 Function(_Ctx, Uint32, _PF, _PD);
typedef _SetSpotD = int Function(_Ctx, int, _PF, _PD);
typedef _SetCameraC = Int32 Function(_Ctx, _PF, _PF, Double, Double, Double, Double);
typedef _SetCameraMotionC = Int32 Function(_Ctx, _PF, Double, Double);
typedef _SetCameraMotionD = int Function(_Ctx, _PF, double, double);
typedef _SetCameraD = int Function(_Ctx, _PF, _PF, double, double, double, double);
typedef _SetIntC = Int32 Function(_Ctx, Int32);
typedef _SetIntD = int Function(_Ctx, int);
typedef _SetFilmC = Int32 Function(_Ctx, Int32, Int32, _PD, Double, Double, _PF);
typedef _SetFilmD = int Function(_Ctx, int, int, _PD, double, double, _PF);
typedef _SetSamplerC = Int32 Function(_Ctx, Int32, Int32, Int32, Int32, Int32, Int32, Int32, Uint64);
typedef _SetSamplerD = int Function(_Ctx, int, int, int, int, int, int, int, int);
typedef _SetIntegratorC = Int32 Function(_Ctx, Int32, Int32, Int32, Int32, Double, Double);
typedef _SetIntegratorD = int Function(_Ctx, int, int, int, int, double, double);
typedef _SetPrecisionC = Int32 Function(_Ctx, Int32);
typedef _SetPrecisionD = int Function(_Ctx, int);
typedef _RenderC = Int32 Function(_Ctx, Int32, Int32);
typedef _RenderD = int Function(_Ctx, int, int);
typedef _FilmReadC = Int32 Function(_Ctx, _PF, _PF, _PF);
typedef _FilmReadD = int Function(_Ctx, _PF, _PF, _PF);
typedef _FilmSizeC = Int32 Function(_Ctx, _PI);
typedef _FilmSizeD = int Function(_Ctx, _PI);
typedef _StatsC = Int32 Function(_Ctx, Pointer<DrtRenderStats>);
typedef _StatsD = int Function(_Ctx, Pointer<DrtRenderStats>);
typedef _TraceC = Int32 Function(_Ctx, _PF, _PF, Uint64, Pointer<DrtHit>);
typedef _TraceD = int Function(_Ctx, _PF, _PF, int, Pointer<DrtHit>);
typedef _TraceAnyC = Int32 Function(_Ctx, _PF, _PF, Uint64, _PB);
typedef _TraceAnyD = int Function(_Ctx, _PF, _PF, int, _PB);

class DrtError implements Exception {
  final int code;
  final String message;
  DrtError(this.code, this.message);
  String toString() => 'libdartray_gpu error $code: $message';
}

/// Thin wrapper: one method per C entry point, `check` turns a negative return code into a [DrtError]
/// carrying drt_last_error (the shim maps it to LogSevere, lib/core/log.dart:44-46).
class Drt {
  final DynamicLibrary lib;
  _Ctx ctx;

  Drt([String path = 'libdartray_gpu.so']) : lib = DynamicLibrary.open(path);

  void create([int device = 0]) {
    ctx = lib.lookupFunction<_CreateC, _CreateD>('drt_create')(device);
    if (ctx == nullptr) {
      throw new DrtError(-4, 'drt_create failed: no CUDA device (there is no CPU fallback)');
    }
  }

  void destroy() {
    if (ctx != null && ctx != nullptr) {
      lib.lookupFunction<_DestroyC, _DestroyD>('drt_destroy')(ctx);
    }
    ctx = nullptr;
  }

  String lastError() => lib.lookupFunction<_LastErrorC, _LastErrorC>('drt_last_error')(ctx).toDartString();

  void check(int rc) {
    if (rc < 0) {
      throw new DrtError(rc, lastError());
    }
  }

  void setTriangles(_PF p, int nverts, _PU idx, int ntris, _PI mat, _PI light, _PB rev) =>
      check(lib.lookupFunction<_SetTrianglesC, _SetTrianglesD>('drt_set_triangles')(ctx, p, nverts, idx, ntris, mat, light, rev));
  void setSpheres(int n, _PF o2w, _PF w2o, _PD params, _PI mat, _PI light, _PB rev) =>
      check(lib.lookupFunction<_SetQuadricsC, _SetQuadricsD>('drt_set_spheres')(ctx, n, o2w, w2o, params, mat, light, rev));
  void setDisks(int n, _PF o2w, _PF w2o, _PD params, _PI mat, _PI light, _PB rev) =>
      check(lib.lookupFunction<_SetQuadricsC, _SetQuadricsD>('drt_set_disks')(ctx, n, o2w, w2o, params, mat, light, rev));
  /// kind 2 cylinder, 3 cone, 4 paraboloid, 5 hyperboloid; params: n x 8 doubles (include/drt.h)
  void setQuadrics(int kind, int n, _PF o2w, _PF w2o, _PD params, _PI mat, _PI light, _PB rev) =>
      check(lib.lookupFunction<_SetKindQuadricsC, _SetKindQuadricsD>('drt_set_quadrics')(ctx, kind, n, o2w, w2o, params, mat, light, rev));
  /// per-vertex N / S (object space) / uv, any may be nullptr; mesh index per triangle, per-mesh transforms and flags
  void setMeshShading(_PF n, _PF s, _PF uv, _PU meshOfTri, int nMeshes, _PF o2w, _PF w2o, _PB flags) =>
      check(lib.lookupFunction<_SetMeshShadingC, _SetMeshShadingD>('drt_set_mesh_shading')(ctx, n, s, uv, meshOfTri, nMeshes, o2w, w2o, flags));
  // TransformedPrimitives: objects (primitive lists + nested accelerator parameters) and instances (AnimatedTransform start / end)
  void setInstances(int nObjects, _PU offsets, _PU prims, _PI split, _PI maxPrims, int nInstances, _PU object, _PF m0, _PF i0, _PF m1,
                    _PF i1, _PD times) =>
      check(lib.lookupFunction<_SetInstancesC, _SetInstancesD>('drt_set_instances')(ctx, nObjects, offsets, prims, split, maxPrims,
                                                                                      nInstances, object, m0, i0, m1, i1, times));
  void setBuildOrder(_PU order, int n) => check(lib.lookupFunction<_SetOrderC, _SetOrderD>('drt_set_build_order')(ctx, order, n));
  void buildBvh(int split, int maxNodePrims) => check(lib.lookupFunction<_BuildC, _BuildD>('drt_build_bvh')(ctx, split, maxNodePrims));
  void setMaterials(int n, _PI kind, _PF kd, _PF sigma) =>
      check(lib.lookupFunction<_SetMaterialsC, _SetMaterialsD>('drt_set_materials')(ctx, n, kind, kd, sigma));
  void setMaterialLobes(int n, _PU offsets, _PI kind, _PF rgb, _PI fresnel, _PF eta, _PF k, _PD scalars) =>
      check(lib.lookupFunction<_SetLobesC, _SetLobesD>('drt_set_material_lobes')(ctx, n, offsets, kind, rgb, fresnel, eta, k, scalars));
  void setLights(int n, _PI kind, _PF L, _PF pos, _PI nsamples, _PU shapeOffsets, _PU shapePrims) =>
      check(lib.lookupFunction<_SetLightsC, _SetLightsD>('drt_set_lights')(ctx, n, kind, L, pos, nsamples, shapeOffsets, shapePrims));
  void setSpotParams(int n, _PF w2l, _PD cosines) => check(lib.lookupFunction<_SetSpotC, _SetSpotD>('drt_set_spot_params')(ctx, n, w2l, cosines));
  void setSampleTable(_PD table, int nEntries) =>
      check(lib.lookupFunction<_SetTableC, _SetTableD>('drt_set_sample_table')(ctx, table, nEntries));
  void setLightMap(int index, int width, int height, _PF rgb, _PF w2l, _PF projection, _PD screen, double hither) =>
      check(lib.lookupFunction<_SetLightMapC, _SetLightMapD>('drt_set_light_map')(ctx, index, width, height, rgb, w2l, projection, screen, hither));
  void setLobeWrappers(int nLobes, _PI wrap, _PF scaleRgb) =>
      check(lib.lookupFunction<_SetWrappersC, _SetWrappersD>('drt_set_lobe_wrappers')(ctx, nLobes, wrap, scaleRgb));
  // textures that read the hit point / materials bound to them (drt_texture, drt_material_program records packed by _Arena)
  void setTextures(_PB nodes, int n, _PF texels, int nTexelFloats) =>
      check(lib.lookupFunction<_SetTexturesC, _SetTexturesD>('drt_set_textures')(ctx, n, nodes, texels, nTexelFloats));
  void setMeasured(int n, _PI kind, _PI dims, Pointer<Uint64> offsets, _PF data, int nFloats) =>
      check(lib.lookupFunction<_SetMeasuredC, _SetMeasuredD>('drt_set_measured')(ctx, n, kind, dims, offsets, data, nFloats));
  void setMaterialPrograms(_PI programs, int n) =>
      check(lib.lookupFunction<_SetProgramsC, _SetProgramsD>('drt_set_material_programs')(ctx, n, programs));
  void setInfiniteLight(int index, int width, int height, _PF rgb, _PF l2w, _PF w2l) =>
      check(lib.lookupFunction<_SetInfiniteC, _SetInfiniteD>('drt_set_infinite_light')(ctx, index, width, height, rgb, l2w, w2l));
  void setCamera(_PF rasterToCamera, _PF cameraToWorld, double lensRadius, double focalDistance, double open, double close) =>
      check(lib.lookupFunction<_SetCameraC, _SetCameraD>('drt_set_camera')(ctx, rasterToCamera, cameraToWorld, lensRadius, focalDistance, open, close));
  void setCameraMotion(_PF cameraToWorldEnd, double startTime, double endTime) =>
      check(lib.lookupFunction<_SetCameraMotionC, _SetCameraMotionD>('drt_set_camera_motion')(ctx, cameraToWorldEnd, startTime, endTime));
  void setCameraKind(int kind) => check(lib.lookupFunction<_SetIntC, _SetIntD>('drt_set_camera_kind')(ctx, kind));
  void setFilm(int xres, int yres, _PD crop, double xw, double yw, _PF table) =>
      check(lib.lookupFunction<_SetFilmC, _SetFilmD>('drt_set_film')(ctx, xres, yres, crop, xw, yw, table));
  void setSampler(int kind, int xs, int ys, int spp, int jitter, int pixelOrder, int tileSize, int seed) =>
      check(lib.lookupFunction<_SetSamplerC, _SetSamplerD>('drt_set_sampler')(ctx, kind, xs, ys, spp, jitter, pixelOrder, tileSize, seed));
  void setIntegrator(int kind, int maxDepth, int strategy, int aoSamples, double aoMin, double aoMax) =>
      check(lib.lookupFunction<_SetIntegratorC, _SetIntegratorD>('drt_set_integrator')(ctx, kind, maxDepth, strategy, aoSamples, aoMin, aoMax));
  /// drt_set_shading_precision: 0 = DRT_PRECISION_F64 (the Dart VM's arithmetic, the default), 1 = DRT_PRECISION_F32 (path integrator
  /// only; per-pixel means within 3 sigma).
  void setShadingPrecision(int precision) =>
      check(lib.lookupFunction<_SetPrecisionC, _SetPrecisionD>('drt_set_shading_precision')(ctx, precision));
  void render(int taskNum, int taskCount) => check(lib.lookupFunction<_RenderC, _RenderD>('drt_render')(ctx, taskNum, taskCount));
  void filmSize(_PI out4) => check(lib.lookupFunction<_FilmSizeC, _FilmSizeD>('drt_film_size')(ctx, out4));
  void filmRead(_PF rgb, _PF xyz, _PF weight) => check(lib.lookupFunction<_FilmReadC, _FilmReadD>('drt_film_read')(ctx, rgb, xyz, weight));
  void renderStats(Pointer<DrtRenderStats> out) => check(lib.lookupFunction<_StatsC, _StatsD>('drt_render_stats_get')(ctx, out));
  void traceClosest(_PF o, _PF d, int n, Pointer<DrtHit> hits) => check(lib.lookupFunction<_TraceC, _TraceD>('drt_trace_closest')(ctx, o, d, n, hits));
  void traceAny(_PF o, _PF d, int n, _PB occluded) => check(lib.lookupFunction<_TraceAnyC, _TraceAnyD>('drt_trace_any')(ctx, o, d, n, occluded));
}
