// @dart=2.9
// Flattens the Texture objects and the texture-bound materials of a constructed scene into the records of drt_set_textures /
// drt_set_material_programs (include/drt.h).  REVIEWED, NOT RUN (no Dart SDK in the build image); the tested twin is
// dartray_b200/host.py: TextureTable / SceneBuilder.material_program, driven by tests/test_textures_gpu.py.
//
// Every value is read from the objects DartRay built (Texture.Create*, Material.Create): mapping parameters, the MIPMap's level 0
// exactly as MIPMap.texture left it (scale / gamma applied, float images converted, resampled to a power of two), the noise
// textures' octaves / roughness.  Nothing is re-derived from the scene file.
part of dartray_gpu;

class _TexNode {  // one drt_texture
  int kind = 0, spectrum = 0, tex1 = -1, tex2 = -1, amount = -1, mapping = 0;
  int imageWidth = 0, imageHeight = 0, imageChannels = 0, imageWrap = 0, imageTrilinear = 0, aaMethod = 0, imageOffset = 0;
  List<double> value = [0.0, 0.0, 0.0], value2 = new List<double>.filled(9, 0.0);
  double su = 1.0, sv = 1.0, du = 0.0, dv = 0.0, maxAnisotropy = 8.0;
  List<double> worldToTexture = [1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0];
  List<double> v1 = [0.0, 0.0, 0.0], v2 = [0.0, 0.0, 0.0];
}

class _Program {  // one drt_material_program
  int kind = -1, bump = -1, m1 = -1, m2 = -1;
  List<int> tex = new List<int>.filled(8, -1);
}

class _MeasuredTable {  // one table of drt_set_measured
  int kind = 0;  // 0 regular halfangle (.merl), 1 irregular isotropic (.brdf)
  List<int> dims = [0, 0, 0];
  int offset = 0;
}

class TextureFlattener {
  final List<_TexNode> nodes = [];
  final List<double> texels = [];
  // MeasuredMaterial data (measured_material.dart:76-205), shared between materials that loaded the same file
  final List<_MeasuredTable> measuredTables = [];
  final List<double> measuredData = [];
  final Map<Object, int> _measuredIds = {};

  int _measured(MeasuredMaterial m) {
    final Object key = m.regularHalfangleData != null ? m.regularHalfangleData : m.thetaPhiData;
    if (key == null) throw new GpuUnsupported('measured material without loaded data');
    if (_measuredIds.containsKey(key)) return _measuredIds[key];
    final t = new _MeasuredTable()..offset = measuredData.length;
    if (m.regularHalfangleData != null) {
      t.kind = 0;
      t.dims = [m.nThetaH, m.nThetaD, m.nPhiD];
      measuredData.addAll(m.regularHalfangleData);
    } else {
      // the samples the KdTree holds (kdtree.dart:35-57: nodeData is a permutation of the list it was built from); the library visits
      // the ones inside the search radius in this order, the reference in the order its tree yields them — the sum is the same set's
      final List samples = m.thetaPhiData.nodeData;
      t.kind = 1;
      t.dims = [samples.length, 0, 0];
      for (final IrregIsotropicBRDFSample s in samples) {
        measuredData..add(s.p.x)..add(s.p.y)..add(s.p.z);
        measuredData.addAll(GpuSamplerRenderer._rgb(s.v));
      }
    }
    measuredTables.add(t);
    return _measuredIds[key] = measuredTables.length - 1;
  }
  final Map<Texture, Map<bool, int>> _ids = {};

  void _map2D(_TexNode n, TextureMapping2D m) {
    if (m is UVMapping2D) {
      n.mapping = 0; n.su = m.su; n.sv = m.sv; n.du = m.du; n.dv = m.dv;
    } else if (m is SphericalMapping2D) {
      n.mapping = 1; n.worldToTexture = m.worldToTexture.m.data.toList();
    } else if (m is CylindricalMapping2D) {
      n.mapping = 2; n.worldToTexture = m.worldToTexture.m.data.toList();
    } else if (m is PlanarMapping2D) {
      n.mapping = 3; n.v1 = [m.vs.x, m.vs.y, m.vs.z]; n.v2 = [m.vt.x, m.vt.y, m.vt.z]; n.du = m.ds; n.dv = m.dt;
    } else {
      throw new GpuUnsupported('texture mapping ${m.runtimeType}');
    }
  }
  void _map3D(_TexNode n, TextureMapping3D m) {
    if (m is! IdentityMapping3D) throw new GpuUnsupported('texture mapping ${m.runtimeType}');
    n.mapping = 4;
    n.worldToTexture = (m as IdentityMapping3D).worldToTexture.m.data.toList();  // the Transform the mapping HOLDS (tex2world as handed over)
  }

  /// Node index of texture `t` used as a spectrum (true) or float (false) parameter.
  int add(Texture t, bool spectrum) {
    final known = _ids.putIfAbsent(t, () => {});
    if (known.containsKey(spectrum)) return known[spectrum];
    final n = new _TexNode()..spectrum = spectrum ? 1 : 0;
    if (t is ConstantTexture) {
      final v = t.value;
      n.value = v is Spectrum ? GpuSamplerRenderer._rgb(v) : [v.toDouble(), spectrum ? v.toDouble() : 0.0, spectrum ? v.toDouble() : 0.0];
    } else if (t is ScaleTexture) {
      n.kind = 1; n.tex1 = add(t.tex1, spectrum); n.tex2 = add(t.tex2, spectrum);
    } else if (t is MixTexture) {
      n.kind = 2; n.tex1 = add(t.tex1, spectrum); n.tex2 = add(t.tex2, spectrum); n.amount = add(t.amount, false);
    } else if (t is ImageTexture) {
      final MIPMap mip = t.mipmap;
      final SpectrumImage lv0 = mip.pyramid[0];  // level 0 as the constructor left it (mipmap.dart:72-147)
      n.kind = 3;
      n.imageWidth = lv0.width; n.imageHeight = lv0.height; n.imageChannels = lv0.samplesPerPixel;
      n.imageWrap = mip.wrapMode; n.imageTrilinear = mip.doTrilinear ? 1 : 0; n.maxAnisotropy = mip.maxAnisotropy;
      n.imageOffset = texels.length;
      texels.addAll(lv0.data);
      _map2D(n, t.mapping);
    } else if (t is CheckerboardTexture) {
      n.kind = 4; n.tex1 = add(t.tex1, spectrum); n.tex2 = add(t.tex2, spectrum); n.aaMethod = t.aaMethod;
      _map2D(n, t.mapping);
    } else if (t is UVTexture) {
      n.kind = 5; _map2D(n, t.mapping);
    } else if (t is BilerpTexture) {
      List<double> val(v) => v is Spectrum ? GpuSamplerRenderer._rgb(v) : [v.toDouble(), v.toDouble(), v.toDouble()];
      n.kind = 6; n.value = val(t.v00); n.value2 = []..addAll(val(t.v01))..addAll(val(t.v10))..addAll(val(t.v11));
      _map2D(n, t.mapping);
    } else if (t is FBmTexture) {
      n.kind = 7; n.aaMethod = t.octaves; n.value[0] = t.omega; _map3D(n, t.mapping);
    } else if (t is WrinkledTexture) {
      n.kind = 8; n.aaMethod = t.octaves; n.value[0] = t.omega; _map3D(n, t.mapping);
    } else if (t is WindyTexture) {
      n.kind = 9; _map3D(n, t.mapping);
    } else if (t is MarbleTexture) {
      n.kind = 10; n.aaMethod = t.octaves; n.value = [t.omega, t.scale, t.variation]; _map3D(n, t.mapping);
    } else if (t is DotsTexture) {
      n.kind = 11; n.tex1 = add(t.outsideDot, spectrum); n.tex2 = add(t.insideDot, spectrum); _map2D(n, t.mapping);
    } else if (t is Checkerboard3DTexture) {
      n.kind = 12; n.tex1 = add(t.tex1, spectrum); n.tex2 = add(t.tex2, spectrum); _map3D(n, t.mapping);
    } else {
      throw new GpuUnsupported('texture ${t.runtimeType}');
    }
    nodes.add(n);
    return known[spectrum] = nodes.length - 1;
  }

  /// The program of material `m`; `indexOf` maps a (sub)material to its index in the material table.
  _Program program(Material m, int indexOf(Material sub)) {
    final p = new _Program();
    void slots(int kind, List<Texture> tex, List<bool> spectrum, Texture bump) {
      p.kind = kind;
      for (int i = 0; i < tex.length; ++i) p.tex[i] = add(tex[i], spectrum[i]);
      if (bump != null) p.bump = add(bump, false);
    }
    const S = true, F = false;
    if (m is MatteMaterial) slots(0, [m.Kd, m.sigma], [S, F], m.bumpMap);
    else if (m is MirrorMaterial) slots(1, [m.Kr], [S], m.bumpMap);
    else if (m is GlassMaterial) slots(2, [m.Kr, m.Kt, m.index], [S, S, F], m.bumpMap);
    else if (m is PlasticMaterial) slots(3, [m.Kd, m.Ks, m.roughness], [S, S, F], m.bumpMap);
    else if (m is MetalMaterial) slots(4, [m.eta, m.k, m.roughness], [S, S, F], m.bumpMap);
    else if (m is ShinyMetalMaterial) slots(5, [m.Ks, m.Kr, m.roughness], [S, S, F], m.bumpMap);
    else if (m is SubstrateMaterial) slots(6, [m.Kd, m.Ks, m.nu, m.nv], [S, S, F, F], m.bumpMap);
    else if (m is TranslucentMaterial) slots(7, [m.Kd, m.Ks, m.reflect, m.transmit, m.roughness], [S, S, S, S, F], m.bumpMap);
    else if (m is UberMaterial) slots(8, [m.Kd, m.Ks, m.Kr, m.Kt, m.roughness, m.opacity, m.eta], [S, S, S, S, F, S, F], m.bumpMap);
    else if (m is SubsurfaceMaterial) slots(10, [m.Kr, m.eta], [S, F], m.bumpMap);
    else if (m is KdSubsurfaceMaterial) slots(10, [m.Kr, m.eta], [S, F], m.bumpMap);
    else if (m is MeasuredMaterial) {
      slots(11, [], [], m.bumpMap);
      p.m1 = _measured(m);
    } else if (m is MixMaterial) {
      slots(9, [m.scale], [S], null);
      p.m1 = indexOf(m.m1);
      p.m2 = indexOf(m.m2);
    } else {
      throw new GpuUnsupported('material ${m.runtimeType}');
    }
    return p;
  }
}
