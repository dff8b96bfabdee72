// Texture pass of the wavefront renderer: see texture_kernels.h and texture_device.cuh.
#include "texture_kernels.h"

#include <algorithm>
#include <cmath>

#include "texture_device.cuh"

namespace drt {

// ---- host: MIPMap.texture (mipmap.dart:63-181) from level 0 -----------------------------------------------------------------
namespace {
inline double log2Dart(double x) { return std::log(x) * (1.0 / std::log(2.0)); }  // common.dart:98-103
inline long long dmod(long long a, long long n) { return ((a % n) + n) % n; }
// texel() of the level being read while the pyramid is built (mipmap.dart:183-204)
inline const float* hostTexel(const float* lv, int W, int H, int ch, int wrap, long long s, long long t, const float* zero) {
  if (wrap == 0) { s = dmod(s, W); t = dmod(t, H); }
  else if (wrap == 2) { s = std::min<long long>(std::max<long long>(s, 0), W - 1); t = std::min<long long>(std::max<long long>(t, 0), H - 1); }
  else if (s < 0 || s >= W || t < 0 || t >= H) return zero;
  return lv + (size_t)(t * W + s) * ch;
}
}  // namespace

bool buildTextureTables(uint32_t n, const drt_texture* nodes, const float* texels, uint64_t nTexelFloats, std::vector<GTex>* out,
                        std::vector<float>* data, std::string* err) {
  out->clear();
  data->clear();
  data->resize(128);
  for (int i = 0; i < 128; ++i) {  // MIPMap.weightLut, a Float32List (:168-176)
    const double alpha = 2.0, r2 = (double)i / (128 - 1);
    (*data)[i] = (float)(std::exp(-alpha * r2) - std::exp(-alpha));
  }
  static const float kZero[3] = {0.f, 0.f, 0.f};
  for (uint32_t i = 0; i < n; ++i) {
    const drt_texture& d = nodes[i];
    GTex t{};
    t.kind = d.kind; t.spectrum = d.spectrum;
    t.tex1 = d.tex1; t.tex2 = d.tex2; t.amount = d.amount;
    t.mapping = d.mapping;
    t.aa = d.aa_method;
    for (int k = 0; k < 3; ++k) t.value[k] = d.value[k];
    for (int k = 0; k < 9; ++k) t.value2[k] = d.value2[k];
    t.su = d.su; t.sv = d.sv; t.du = d.du; t.dv = d.dv;
    t.maxAniso = d.max_anisotropy;
    for (int k = 0; k < 16; ++k) t.w2t[k] = d.world_to_texture[k];
    for (int k = 0; k < 3; ++k) { t.v1[k] = d.v1[k]; t.v2[k] = d.v2[k]; }
    if (d.kind < 0 || d.kind > 12 || d.mapping < 0 || d.mapping > 4) { *err = "texture kind / mapping out of range"; return false; }
    for (int child : {d.tex1, d.tex2, d.amount})
      if (child >= (int)i) { *err = "a texture node may only reference earlier nodes"; return false; }
    const bool two = d.kind == 1 || d.kind == 2 || d.kind == 4 || d.kind == 11 || d.kind == 12;
    if (two && (d.tex1 < 0 || d.tex2 < 0 || nodes[d.tex1].spectrum != d.spectrum || nodes[d.tex2].spectrum != d.spectrum)) {
      *err = "scale / mix / checkerboard / dots need two children of their own type";
      return false;
    }
    if (d.kind == 2 && (d.amount < 0 || nodes[d.amount].spectrum != 0)) { *err = "mix needs a float texture as amount"; return false; }
    if ((d.kind == 5 || d.kind == 10) && !d.spectrum) { *err = "'uv' and 'marble' have no float form (uv_texture.dart:39-41, marble_texture.dart:68-70)"; return false; }
    if ((d.kind == 7 || d.kind == 8 || d.kind == 10) && (d.aa_method < 0 || d.aa_method > 64)) { *err = "noise texture: octaves out of range"; return false; }
    if (d.kind == 3) {
      const int W = d.image_width, H = d.image_height, ch = d.image_channels;
      if (W < 1 || H < 1 || (W & (W - 1)) || (H & (H - 1)) || W > 32768 || H > 32768 || (ch != 1 && ch != 3)) {
        *err = "image: level 0 at power-of-two resolution (the reference resamples at load, mipmap.dart:72-139), 1 or 3 channels";
        return false;
      }
      if (ch != (d.spectrum ? 3 : 1)) { *err = "image channels must match the texture's type (image_texture.dart:38-42)"; return false; }
      if (d.image_wrap < 0 || d.image_wrap > 2) { *err = "image wrap mode out of range"; return false; }
      if (d.image_offset + (uint64_t)W * H * ch > nTexelFloats) { *err = "image beyond the texel array"; return false; }
      t.w = W; t.h = H; t.channels = ch; t.wrap = d.image_wrap; t.trilinear = d.image_trilinear ? 1 : 0;
      t.levels = 1 + (int)log2Dart((double)std::max(W, H));  // :143 as written: 8, 64, 128 ... come out one level short
      if (t.levels > 16) { *err = "image too large"; return false; }
      if (data->size() + 2 * (size_t)W * H * ch > 0xffffffffull) { *err = "texture data beyond 2^32 floats"; return false; }
      t.levelOffset[0] = (uint32_t)data->size();
      data->insert(data->end(), texels + d.image_offset, texels + d.image_offset + (size_t)W * H * ch);
      int pw = W, ph = H;
      for (int lv = 1; lv < t.levels; ++lv) {  // :152-166
        const int sw = std::max(1, pw / 2), sh = std::max(1, ph / 2);
        t.levelOffset[lv] = (uint32_t)data->size();
        data->resize(data->size() + (size_t)sw * sh * ch);
        const float* fine = data->data() + t.levelOffset[lv - 1];
        float* dst = data->data() + t.levelOffset[lv];
        for (int y = 0; y < sh; ++y)
          for (int x = 0; x < sw; ++x) {
            const float* a = hostTexel(fine, pw, ph, ch, t.wrap, 2 * x, 2 * y, kZero);
            const float* b = hostTexel(fine, pw, ph, ch, t.wrap, 2 * x + 1, 2 * y, kZero);
            const float* cc = hostTexel(fine, pw, ph, ch, t.wrap, 2 * x, 2 * y + 1, kZero);
            const float* dd = hostTexel(fine, pw, ph, ch, t.wrap, 2 * x + 1, 2 * y + 1, kZero);
            float* o = dst + ((size_t)y * sw + x) * ch;
            if (ch == 1) {  // Dart doubles: the plain average, stored into the level's Float32List
              o[0] = (float)(((double)a[0] + (double)b[0] + (double)cc[0] + (double)dd[0]) * 0.25);
            } else {
              // AS WRITTEN: SpectrumImage.operator[] returns ONE shared RGBColor (spectrum_image.dart:103-112,133-135); in
              // `texel(a) + texel(b)` the left operand is that same object, already overwritten by b when operator+ runs, so the
              // level holds (2 b + c + d) * 0.25, each operation a new float32 RGBColor (tests/test_oracle_textures.py pins it).
              for (int k = 0; k < 3; ++k) {
                float acc = (float)((double)b[k] + (double)b[k]);
                acc = (float)((double)acc + (double)cc[k]);
                acc = (float)((double)acc + (double)dd[k]);
                o[k] = (float)((double)acc * 0.25);
              }
              (void)a;
            }
          }
        pw = sw; ph = sh;
      }
    }
    out->push_back(t);
  }
  return true;
}

// ---- device ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) texturePassKernel(RenderParams rp, RenderScene rs, Wavefront wf, int cur, int hasDiff) {
  const uint32_t n = wf.counts[cur], cap = wf.cap;
  const TexCtx tc{rs.textures, rs.texData};
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const uint32_t slot = wf.extSlot[cur][q];
    const int prim = __float_as_int(wf.extHit[q].w);
    if (prim < 0) continue;
    const uint32_t mat = (uint32_t)primMaterial(rs, (uint32_t)prim);
    if (rs.programs[mat].kind < 0) { wf.hitCount[slot] = -1; continue; }
    const float4 o4 = wf.extO[cur][q], d4 = wf.extD[cur][q];
    const V3 o = V3{o4.x, o4.y, o4.z}, d = V3{d4.x, d4.y, d4.z};
    FullDG dg, dgs;
    const int inst = wf.extInst ? wf.extInst[q] : -1;
    if (inst >= 0) {
      // a hit through a TransformedPrimitive (transformed_primitive.dart:30-58): the shape's differential geometry in primitive
      // space, moved to world space with Inverse(w2p); objects carry no per-vertex N / S, so dgShading is the same record
      M4 m, inv;
      animInterpolate(rs.ts.instances[inst], wf.slotTime[slot], &m, &inv);
      fullGeometryCold(rs, (uint32_t)prim, XfPoint(m.d, o), XfVector(m.d, d), wf.extT[q], &dg, &dgs);
      if (!m4IsIdentity(m)) {
        dg.p = XfPoint(inv.d, dg.p);
        dg.nn = Normalize(XfNormal(m.d, dg.nn));
        dg.dpdu = XfVector(inv.d, dg.dpdu);
        dg.dpdv = XfVector(inv.d, dg.dpdv);
        dg.dndu = XfNormal(m.d, dg.dndu);
        dg.dndv = XfNormal(m.d, dg.dndv);
      }
      dgs = dg;
    } else {
      fullGeometryCold(rs, (uint32_t)prim, o, d, wf.extT[q], &dg, &dgs);
    }
    RayDiffs rd;
    rd.has = false;
    if (hasDiff) {
      const double2 xy = wf.camXY[slot], lens = wf.camLens[slot];
      cameraDifferentialsCold(rp, xy.x, xy.y, lens.x, lens.y, o, d, rp.diffScale, LerpD(wf.camTime[slot], rp.shutterOpen, rp.shutterClose), &rd);
    }
    // Intersection.getBSDF: dg.computeDifferentials(ray); getShadingGeometry copies the differentials into dgShading (triangle.dart:354-363)
    computeDifferentials(&dg, rd);
    dgs.dudx = dg.dudx; dgs.dvdx = dg.dvdx; dgs.dudy = dg.dudy; dgs.dvdy = dg.dvdy;
    dgs.dpdx = dg.dpdx; dgs.dpdy = dg.dpdy;
    HitBsdf hb;
    materialBsdfCold(rs, tc, rs.programs, mat, dg, dgs, &hb, 0);
    wf.hitCount[slot] = hb.n;
    for (int i = 0; i < hb.n; ++i) wf.hitLobes[(size_t)slot * 8 + i] = hb.lobes[i];
    wf.hitFrame[slot] = hb.nn.x; wf.hitFrame[cap + slot] = hb.nn.y; wf.hitFrame[2 * (size_t)cap + slot] = hb.nn.z;
    wf.hitFrame[3 * (size_t)cap + slot] = hb.sn.x; wf.hitFrame[4 * (size_t)cap + slot] = hb.sn.y; wf.hitFrame[5 * (size_t)cap + slot] = hb.sn.z;
  }
}

cudaError_t launchTexturePass(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int hasDiff, int numSMs,
                              cudaStream_t st) {
  // the texture tree is evaluated recursively (bounded depth) and the per-hit records live in local memory
  size_t cur_ = 0;
  cudaDeviceGetLimit(&cur_, cudaLimitStackSize);
  if (cur_ < 16384) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitStackSize, 16384);
    if (e != cudaSuccess) return e;
  }
  texturePassKernel<<<numSMs * 8, 128, 0, st>>>(rp, rs, wf, cur, hasDiff);
  return cudaGetLastError();
}

}  // namespace drt
