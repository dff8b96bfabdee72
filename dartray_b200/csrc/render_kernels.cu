// Wavefront render stages for sm_100a: sampler sequences -> camera rays -> (extend) -> shade ->
// (shadow / MIS rays) -> resolve -> film.  The traversal itself is trace_fast.cu; everything here is
// the per-sample arithmetic of the reference, restated per stage:
//
//   samplers     lib/samplers/{low_discrepancy,stratified,random}_sampler.dart, lib/core/montecarlo.dart:270-551
//   camera       lib/cameras/perspective_camera.dart:93-132
//   path         lib/surface_integrators/path_integrator.dart:29-131
//   AO           lib/surface_integrators/ambient_occlusion_integrator.dart:28-53
//   direct       lib/surface_integrators/direct_lighting_integrator.dart:30-96, lib/core/integrator.dart:39-185
//   renderer     lib/renderers/sampler_renderer.dart:67-98,173-193
//   film         lib/film/image_film.dart:99-185,268-299
//
// Random numbers: counter-based streams keyed by (pixel, array) for the sampler and (pixel, sample)
// for the integrators (shade_device.cuh); inside a stream the draw order is the reference's.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "render_kernels.h"
#include "shade_device.cuh"
#include "trace_device.cuh"
#include "volume_device.cuh"

#ifndef DRT_RK_NS
#define DRT_RK_NS extra  // this file as it stands; render_kernels_plain.cu includes it again with DRT_EXTRA = 0 / plain
#endif

namespace drt {
namespace DRT_RK_NS {

#define FULL 0xffffffffu

// Warp-aggregated queue append: one atomicAdd per warp, slots handed out by ballot prefix.
// Must be reached by all 32 lanes.
static __device__ __forceinline__ uint32_t warpPush(uint32_t* counter, bool want) {
  const unsigned m = __ballot_sync(FULL, want);
  if (m == 0) return 0;
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(FULL, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
}

// Two appends at once: lane 0 issues both atomics before either result is awaited (two warpPush calls in a row wait for the first
// counter's round trip to L2 before the second atomic leaves).
static __device__ __forceinline__ void warpPush2(uint32_t* counterA, bool wantA, uint32_t* counterB, bool wantB, uint32_t* atA, uint32_t* atB) {
  const unsigned mA = __ballot_sync(FULL, wantA), mB = __ballot_sync(FULL, wantB);
  const unsigned lane = threadIdx.x & 31u;
  uint32_t baseA = 0, baseB = 0;
  if (lane == 0) {
    if (mA) baseA = atomicAdd(counterA, (uint32_t)__popc(mA));
    if (mB) baseB = atomicAdd(counterB, (uint32_t)__popc(mB));
  }
  baseA = __shfl_sync(FULL, baseA, 0);
  baseB = __shfl_sync(FULL, baseB, 0);
  const unsigned lt = (1u << lane) - 1u;
  *atA = baseA + __popc(mA & lt);
  *atB = baseB + __popc(mB & lt);
}

static __device__ __forceinline__ Spec ld3(const float* a, uint32_t cap, uint32_t i) { return Spec{a[i], a[cap + i], a[2 * cap + i]}; }
static __device__ __forceinline__ void st3(float* a, uint32_t cap, uint32_t i, const Spec& s) { a[i] = s.r; a[cap + i] = s.g; a[2 * cap + i] = s.b; }
static __device__ __forceinline__ V3 ldv3(const float* a, uint32_t cap, uint32_t i) { return V3{a[i], a[cap + i], a[2 * cap + i]}; }
static __device__ __forceinline__ void stv3(float* a, uint32_t cap, uint32_t i, const V3& s) { a[i] = s.x; a[cap + i] = s.y; a[2 * cap + i] = s.z; }

static __device__ __forceinline__ void pixelOf(const PixelBatch& pb, uint32_t p, int* x, int* y) {
  uint64_t g = pb.firstPixel + p;
  if (pb.list) g = pb.list[g];
  else if (pb.nShards > 1) g = ((g / pb.blockPixels) * pb.nShards + pb.shard) * pb.blockPixels + (g % pb.blockPixels);
  *x = pb.x0 + (int)(g % (uint64_t)pb.w);
  *y = pb.y0 + (int)(g / (uint64_t)pb.w);
}

#ifndef DRT_PATH_ONLY  // the float32 path build (render_kernels_f32.cu) takes the path-vertex and resolve kernels only
// ---------------------------------------------------------------------------------------------------
// Low-discrepancy sampler: one group of G lanes (G = 1..32, a power of two chosen so that a block's
// arrays fit in shared memory) per (pixel, array).  LDShuffleScrambled1D/2D (montecarlo.dart:524-551):
// scrambled (0,2)-sequence values, a Fisher-Yates shuffle inside every block of nSamples and one
// across the nPixel blocks (Shuffle, montecarlo.dart:294-303).  Values are computed and written out by
// all lanes of the group; the swaps are order-dependent and run on the group's first lane in shared
// memory (the stream is counter-based, so every swap target is drawn in place).
__global__ void __launch_bounds__(128) samplerLDKernel(RenderParams rp, Wavefront wf, const SampleArray* __restrict__ arrays,
                                                       int nArrays, int maxVals, PixelBatch pb, int G) {
  extern __shared__ float smem[];
  const uint32_t groupsPerBlock = blockDim.x / G, grp = threadIdx.x / G, gl = threadIdx.x % G;
  float* buf = smem + (size_t)grp * (maxVals | 1);  // odd stride: groups start in different banks
  const uint64_t nTasks = (uint64_t)pb.nPixels * nArrays;
  const uint32_t nP = (uint32_t)rp.nPixelSamples;
  for (uint64_t t0 = (uint64_t)blockIdx.x * groupsPerBlock; t0 < nTasks; t0 += (uint64_t)gridDim.x * groupsPerBlock) {
    const uint64_t task = t0 + grp;
    const bool valid = task < nTasks;
    const uint32_t p = valid ? (uint32_t)(task / nArrays) : 0u;
    const SampleArray A = arrays[valid ? task % nArrays : 0];
    int x, y;
    pixelOf(pb, p, &x, &y);
    const uint64_t key = streamKey(rp.seed, x, y, rp.samplerKind == 4 ? pb.pass : 0u, A.streamId);  // adaptive: the visit keys the draw
    const uint32_t nS = (uint32_t)A.nSamples, total = valid ? nS * nP : 0u, dims = (uint32_t)A.dims;
    const uint32_t s0 = drawUint(key, 1), s1 = dims == 2 ? drawUint(key, 2) : 0u;
    const uint64_t base = dims;  // draws consumed by the scrambles
    for (uint32_t i = gl; i < total; i += G) {
      if (dims == 1) buf[i] = (float)VanDerCorput(i, s0);
      else { buf[2 * i] = (float)VanDerCorput(i, s0); buf[2 * i + 1] = (float)Sobol2(i, s1); }
    }
    __syncwarp();
    if (gl == 0 && valid) {
      if (nS > 1)
        for (uint32_t blk = 0; blk < nP; ++blk)
          for (uint32_t k = 0; k < nS; ++k) {
            const uint32_t e = blk * nS + k;
            const uint32_t o = k + drawUint(key, base + e + 1) % (nS - k);
            if (o != k)
              for (uint32_t j = 0; j < dims; ++j) {
                float a = buf[e * dims + j];
                buf[e * dims + j] = buf[(blk * nS + o) * dims + j];
                buf[(blk * nS + o) * dims + j] = a;
              }
          }
      const uint32_t bs = nS * dims;
      for (uint32_t i = 0; i < nP; ++i) {
        const uint32_t o = i + drawUint(key, base + nS * nP + i + 1) % (nP - i);
        if (o != i)
          for (uint32_t j = 0; j < bs; ++j) {
            float a = buf[i * bs + j];
            buf[i * bs + j] = buf[o * bs + j];
            buf[o * bs + j] = a;
          }
      }
    }
    __syncwarp();
    // Write-out by the whole warp, one group after the other: consecutive lanes store consecutive samples of one value
    // row (full 128-byte lines) instead of every group scattering 4-byte stores over its own rows.
    const uint32_t bs = nS * dims;
    const uint32_t slot0 = p * nP;
    const uint32_t lane = threadIdx.x & 31u, groupsPerWarp = 32u / (uint32_t)G;
    for (uint32_t gi = 0; gi < groupsPerWarp; ++gi) {
      const int src = (int)(gi * (uint32_t)G);
      const bool v_ = __shfl_sync(FULL, valid ? 1 : 0, src) != 0;
      const int dest_ = __shfl_sync(FULL, A.dest, src);
      const uint32_t slot0_ = __shfl_sync(FULL, slot0, src), bs_ = __shfl_sync(FULL, bs, src);
      const int x_ = __shfl_sync(FULL, x, src), y_ = __shfl_sync(FULL, y, src);
      if (!v_) continue;
      const float* b_ = smem + (size_t)((threadIdx.x >> 5) * groupsPerWarp + gi) * (maxVals | 1);
      if (dest_ >= 0) {
        for (uint32_t qv = 0; qv < bs_; ++qv)
          for (uint32_t i = lane; i < nP; i += 32u) wf.vals[(size_t)(dest_ + qv) * wf.cap + slot0_ + i] = b_[i * bs_ + qv];
      } else if (dest_ == -1) {  // montecarlo.dart:452-453: imageX = xPos + sample (f64 sum of an int and a float32)
        for (uint32_t i = lane; i < nP; i += 32u) wf.camXY[slot0_ + i] = make_double2(x_ + (double)b_[2 * i], y_ + (double)b_[2 * i + 1]);
      } else if (dest_ == -2) {
        for (uint32_t i = lane; i < nP; i += 32u) wf.camLens[slot0_ + i] = make_double2((double)b_[2 * i], (double)b_[2 * i + 1]);
      } else {
        for (uint32_t i = lane; i < nP; i += 32u) wf.camTime[slot0_ + i] = b_[i];
      }
    }
    __syncwarp();
  }
}

// The same sampler when every array holds ONE value per camera sample (path integrator, ambient occlusion): then
// LDShuffleScrambled1D/2D reduces to the shuffle across the nPixel samples, and because the i-th value of the
// (0,2)-sequence is a pure function of i, it is enough to shuffle INDICES: all lanes draw the swap targets (the stream
// is counter-based), one lane per group applies the swaps to a 16-bit permutation in shared memory (the only
// order-dependent part: four shared-memory accesses per step), then the whole warp evaluates VanDerCorput / Sobol2 at
// the permuted indices and stores full lines.  Shared memory per task: 4 bytes per pixel sample.
// Bit-identical short forms of montecarlo.dart:486-504 for this kernel: the (0,2)-sequence values are k * 2^-24 with k < 2^24, exact in
// float32 (the reference's min with OneMinusEpsilon = 1 - 2^-24 never binds), VanDerCorput's five swap steps are one bit reversal,
// and Sobol2's direction numbers XOR linearly, so two 256-entry tables (sobolLo / sobolHi, built per block) replace its loop.
static __device__ __forceinline__ float vdcValue(uint32_t n, uint32_t scramble) {
  return (float)((__brev(n) ^ scramble) >> 8) * 5.9604644775390625e-8f;
}
static __device__ __forceinline__ uint32_t sobolBits(uint32_t n) {  // Sobol2's XOR of direction numbers for the set bits of n
  uint32_t s = 0u;
  for (uint32_t v = 1u << 31; n != 0; n >>= 1, v ^= v >> 1)
    if (n & 0x1) s ^= v;
  return s;
}
// PT: the type of a permutation entry — uint8_t when nPixelSamples <= 256 (half the shared memory per task, so twice the tasks per
// block run the order-dependent swaps at once), uint16_t otherwise.
template <typename PT>
__global__ void __launch_bounds__(128) samplerLDPermKernel(RenderParams rp, Wavefront wf, const SampleArray* __restrict__ arrays,
                                                           int nArrays, PixelBatch pb, int G, int strideWords) {
  extern __shared__ float smem[];
  __shared__ uint32_t sobolLo[256], sobolHi[256];
  // x % d for the swap targets, d = nP - i: Lemire's exact remainder with M = floor((2^64 - 1) / d) + 1 (d = 1: M wraps to 0 and
  // the remainder is 0), tabulated per block when nP <= 256
  __shared__ unsigned long long modM[256];
  const uint32_t groupsPerBlock = blockDim.x / G, grp = threadIdx.x / G, gl = threadIdx.x % G;
  const uint32_t nP = (uint32_t)rp.nPixelSamples;
  const bool fastMod = nP <= 256u;
  for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
    sobolLo[i] = sobolBits(i);
    sobolHi[i] = sobolBits(i << 8);
    modM[i] = (fastMod && i < nP) ? (~0ull / (unsigned long long)(nP - i) + 1ull) : 0ull;
  }
  __syncthreads();
  PT* tgt = reinterpret_cast<PT*>(smem + (size_t)grp * strideWords);
  PT* perm = tgt + nP;
  const uint64_t nTasks = (uint64_t)pb.nPixels * nArrays;
  const uint32_t lane = threadIdx.x & 31u, groupsPerWarp = 32u / (uint32_t)G;
  for (uint64_t t0 = (uint64_t)blockIdx.x * groupsPerBlock; t0 < nTasks; t0 += (uint64_t)gridDim.x * groupsPerBlock) {
    const uint64_t task = t0 + grp;
    const bool valid = task < nTasks;
    const uint32_t p = valid ? (uint32_t)(task / nArrays) : 0u;
    const SampleArray A = arrays[valid ? task % nArrays : 0];
    int x, y;
    pixelOf(pb, p, &x, &y);
    const uint64_t key = streamKey(rp.seed, x, y, rp.samplerKind == 4 ? pb.pass : 0u, A.streamId);  // adaptive: the visit keys the draw
    const uint32_t dims = (uint32_t)A.dims;
    const uint32_t s0 = drawUint(key, 1), s1 = dims == 2 ? drawUint(key, 2) : 0u;
    // draws: dims scrambles, nP (unused: blocks of one sample are not shuffled, montecarlo.dart:528-530), then the nP swaps
    const uint64_t base = (uint64_t)dims + nP;
    if (valid)
      for (uint32_t i = gl; i < nP; i += G) {
        const uint32_t u = drawUint(key, base + i + 1);
        uint32_t rem;
        if (fastMod) rem = (uint32_t)__umul64hi(modM[i] * (unsigned long long)u, (unsigned long long)(nP - i));
        else rem = u % (nP - i);
        tgt[i] = (PT)(i + rem);
        perm[i] = (PT)i;
      }
    __syncwarp();
    if (gl == 0 && valid)
      for (uint32_t i = 0; i < nP; ++i) {  // Shuffle (montecarlo.dart:294-303) on the indices
        const uint32_t o = tgt[i];
        const PT a = perm[i];
        perm[i] = perm[o];
        perm[o] = a;
      }
    __syncwarp();
    const uint32_t slot0 = p * nP;
    for (uint32_t gi = 0; gi < groupsPerWarp; ++gi) {
      const int src = (int)(gi * (uint32_t)G);
      const bool v_ = __shfl_sync(FULL, valid ? 1 : 0, src) != 0;
      const int dest_ = __shfl_sync(FULL, A.dest, src);
      const uint32_t slot0_ = __shfl_sync(FULL, slot0, src), dims_ = __shfl_sync(FULL, dims, src);
      const uint32_t s0_ = __shfl_sync(FULL, s0, src), s1_ = __shfl_sync(FULL, s1, src);
      const int x_ = __shfl_sync(FULL, x, src), y_ = __shfl_sync(FULL, y, src);
      if (!v_) continue;
      const PT* pm = reinterpret_cast<const PT*>(smem + (size_t)((threadIdx.x >> 5) * groupsPerWarp + gi) * strideWords) + nP;
      for (uint32_t i = lane; i < nP; i += 32u) {
        const uint32_t e = pm[i];
        const float v0 = vdcValue(e, s0_);
        const float v1 = dims_ == 2 ? (float)((s1_ ^ sobolLo[e & 255u] ^ sobolHi[(e >> 8) & 255u]) >> 8) * 5.9604644775390625e-8f : 0.f;
        if (dest_ >= 0) {
          wf.vals[(size_t)dest_ * wf.cap + slot0_ + i] = v0;
          if (dims_ == 2) wf.vals[(size_t)(dest_ + 1) * wf.cap + slot0_ + i] = v1;
        } else if (dest_ == -1) {  // montecarlo.dart:452-453: imageX = xPos + sample (f64 sum of an int and a float32)
          wf.camXY[slot0_ + i] = make_double2(x_ + (double)v0, y_ + (double)v1);
        } else if (dest_ == -2) {
          wf.camLens[slot0_ + i] = make_double2((double)v0, (double)v1);
        } else {
          wf.camTime[slot0_ + i] = v0;
        }
      }
    }
    __syncwarp();
  }
}

// Stratified (stratified_sampler.dart:67-124, montecarlo.dart:270-325) and random
// (random_sampler.dart:47-88) samplers: one sequential stream per pixel visit -> one thread per pixel.
__global__ void __launch_bounds__(128) samplerSeqKernel(RenderParams rp, Wavefront wf, const SampleArray* __restrict__ arrays,
                                                        int nArrays, PixelBatch pb) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pb.nPixels) return;
  int x, y;
  pixelOf(pb, p, &x, &y);
  Stream rng{streamKey(rp.seed, x, y, pb.pass, DRT_STREAM_PIXEL), 0};
  const uint32_t n = (uint32_t)rp.nPixelSamples, slot0 = p * n, cap = wf.cap;
  if (rp.samplerKind == 1) {
    const bool jitter = rp.jitter != 0;
    const double dx = 1.0 / rp.xs, dy = 1.0 / rp.ys;
    for (int pass = 0; pass < 2; ++pass) {  // image samples, then lens samples (StratifiedSample2D)
      uint32_t si = 0;
      for (int yy = 0; yy < rp.ys; ++yy)
        for (int xx = 0; xx < rp.xs; ++xx, ++si) {
          double jx = jitter ? rng.randomFloat() : 0.5;
          double jy = jitter ? rng.randomFloat() : 0.5;
          float fx = (float)fmin((xx + jx) * dx, DRT_ONE_MINUS_EPS), fy = (float)fmin((yy + jy) * dy, DRT_ONE_MINUS_EPS);
          if (pass == 0) {  // stratified_sampler.dart:97-100: shifted to the pixel in float32
            fx = (float)((double)fx + x);
            fy = (float)((double)fy + y);
            wf.camXY[slot0 + si] = make_double2((double)fx, (double)fy);
          } else {
            wf.camLens[slot0 + si] = make_double2((double)fx, (double)fy);
          }
        }
    }
    const double invTot = 1.0 / n;
    for (uint32_t i = 0; i < n; ++i) {  // StratifiedSample1D(time)
      double delta = jitter ? rng.randomFloat() : 0.5;
      wf.camTime[slot0 + i] = (float)fmin((i + delta) * invTot, DRT_ONE_MINUS_EPS);
    }
    for (uint32_t i = 0; i < n; ++i) {  // Shuffle(lens, 2 dims)
      uint32_t o = i + rng.randomUint() % (n - i);
      double2 a = wf.camLens[slot0 + i];
      wf.camLens[slot0 + i] = wf.camLens[slot0 + o];
      wf.camLens[slot0 + o] = a;
    }
    for (uint32_t i = 0; i < n; ++i) {  // Shuffle(time)
      uint32_t o = i + rng.randomUint() % (n - i);
      const double a = wf.camTime[slot0 + i];
      wf.camTime[slot0 + i] = wf.camTime[slot0 + o];
      wf.camTime[slot0 + o] = a;
    }
    for (uint32_t i = 0; i < n; ++i) {  // LatinHypercube per integrator array, sample by sample
      for (int a = 3; a < nArrays; ++a) {
        const SampleArray A = arrays[a];
        const uint32_t nS = (uint32_t)A.nSamples, nDim = (uint32_t)A.dims;
        const double delta = 1.0 / nS;
        float* v = wf.vals + (size_t)A.dest * cap + slot0 + i;  // value q of the array lives at v[q * cap]
        for (uint32_t s = 0; s < nS; ++s)
          for (uint32_t j = 0; j < nDim; ++j) v[(size_t)(nDim * s + j) * cap] = (float)fmin((s + rng.randomFloat()) * delta, DRT_ONE_MINUS_EPS);
        for (uint32_t d = 0; d < nDim; ++d)
          for (uint32_t j = 0; j < nS; ++j) {
            uint32_t o = j + rng.randomUint() % (nS - j);
            float t = v[(size_t)(nDim * j + d) * cap];
            v[(size_t)(nDim * j + d) * cap] = v[(size_t)(nDim * o + d) * cap];
            v[(size_t)(nDim * o + d) * cap] = t;
          }
      }
    }
  } else {
    for (uint32_t si = 0; si < n; ++si) {
      double ix = rng.randomFloat() + x, iy = rng.randomFloat() + y;
      double lu = rng.randomFloat(), lv = rng.randomFloat();
      wf.camXY[slot0 + si] = make_double2(ix, iy);
      wf.camLens[slot0 + si] = make_double2(lu, lv);
      wf.camTime[slot0 + si] = rng.randomFloat();  // random_sampler.dart:71: a Dart double
      for (int a = 3; a < nArrays; ++a) {
        const SampleArray A = arrays[a];
        const uint32_t cnt = (uint32_t)(A.nSamples * A.dims);
        for (uint32_t q = 0; q < cnt; ++q) wf.vals[(size_t)(A.dest + q) * cap + slot0 + si] = (float)rng.randomFloat();
      }
    }
  }
}

// Halton sampler (halton_sampler.dart:59-104): a "pixel" of the batch is one index n of the sequence (pixelOf returns
// x = n mod 2^30, y = n / 2^30, which also key the sample's streams), one sample per index.  Image position from the radical
// inverses in bases 3 and 2 scaled by delta = max(width, height); samples beyond the window's INCLUSIVE right / bottom are
// rejected as written (camXY.x = NaN: raygen leaves them out of the ray queue, the film kernel skips them); lens and time from
// bases 5, 7, 11 at n + 1 (currentSample has been incremented by then); integrator arrays by LatinHypercube.
static __device__ inline double RadicalInverse(uint64_t n, int base) {  // montecarlo.dart:327-339: a truncated double product
  double val = 0.0;
  const double invBase = 1.0 / base;
  double invBi = invBase;
  while (n > 0) {
    const int d_i = (int)(n % (uint64_t)base);
    val += d_i * invBi;
    n = (uint64_t)((double)n * invBase);
    invBi *= invBase;
  }
  return val;
}
__global__ void __launch_bounds__(128) samplerHaltonKernel(RenderParams rp, Wavefront wf, const SampleArray* __restrict__ arrays,
                                                           int nArrays, PixelBatch pb) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pb.nPixels) return;
  int x, y;
  pixelOf(pb, p, &x, &y);
  const uint64_t n = ((uint64_t)(uint32_t)y << 30) | (uint32_t)x;
  const uint32_t cap = wf.cap;
  const double u = RadicalInverse(n, 3), v = RadicalInverse(n, 2);
  const double lerpDelta = (double)max(rp.winW, rp.winH), left = rp.winX, top = rp.winY;
  const double imageX = left * (1.0 - u) + (left + lerpDelta) * u, imageY = top * (1.0 - v) + (top + lerpDelta) * v;
  if (imageX > rp.winX + rp.winW - 1 || imageY > rp.winY + rp.winH - 1) {
    wf.camXY[p] = make_double2(CUDART_NAN, CUDART_NAN);
    return;
  }
  wf.camXY[p] = make_double2(imageX, imageY);
  wf.camLens[p] = make_double2(RadicalInverse(n + 1, 5), RadicalInverse(n + 1, 7));
  wf.camTime[p] = RadicalInverse(n + 1, 11);  // halton_sampler.dart:88: a Dart double
  Stream rng{streamKey(rp.seed, x, y, pb.pass, DRT_STREAM_PIXEL), 0};
  for (int a = 3; a < nArrays; ++a) {  // LatinHypercube (montecarlo.dart:305-325)
    const SampleArray A = arrays[a];
    const uint32_t nS = (uint32_t)A.nSamples, nDim = (uint32_t)A.dims;
    const double delta = 1.0 / nS;
    float* vv = wf.vals + (size_t)A.dest * cap + p;
    for (uint32_t s = 0; s < nS; ++s)
      for (uint32_t j = 0; j < nDim; ++j) vv[(size_t)(nDim * s + j) * cap] = (float)fmin((s + rng.randomFloat()) * delta, DRT_ONE_MINUS_EPS);
    for (uint32_t dd = 0; dd < nDim; ++dd)
      for (uint32_t j = 0; j < nS; ++j) {
        const uint32_t o = j + rng.randomUint() % (nS - j);
        const float t = vv[(size_t)(nDim * j + dd) * cap];
        vv[(size_t)(nDim * j + dd) * cap] = vv[(size_t)(nDim * o + dd) * cap];
        vv[(size_t)(nDim * o + dd) * cap] = t;
      }
  }
}

// Best-candidate sampler (best_candidate_sampler.dart:74-132): like halton a "pixel" of the batch is one index
// n = tile * 4096 + tableOffset.  The lowdiscrepancy kernels have already written the integrator arrays of the index (one pixel
// sample each, LDShuffleScrambled1D / 2D, :125-131); this kernel sets the camera sample from the pattern and rejects — as
// written, both coordinates against left and right (:117-118).
__global__ void __launch_bounds__(128) samplerBestCandidateKernel(RenderParams rp, Wavefront wf, PixelBatch pb) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pb.nPixels) return;
  int x, y;
  pixelOf(pb, p, &x, &y);
  const uint64_t n = ((uint64_t)(uint32_t)y << 30) | (uint32_t)x;
  const uint64_t tile = n / 4096;
  const double* T = rp.bcTable + (size_t)(n % 4096) * 5;
  const double* so = rp.bcTileShifts + 3 * tile;
  const int xTile = rp.bcXTileStart + (int)(tile % (uint64_t)rp.bcTilesX), yTile = rp.bcYTileStart + (int)(tile / (uint64_t)rp.bcTilesX);
  const double imageX = (xTile + T[0]) * rp.bcTableWidth, imageY = (yTile + T[1]) * rp.bcTableWidth;
  const int left = rp.winX, right = rp.winX + rp.winW - 1;
  if (imageX < left || imageX > right || imageY < left || imageY > right) {
    wf.camXY[p] = make_double2(CUDART_NAN, CUDART_NAN);
    return;
  }
  const double t = so[0] + T[2], lu = so[1] + T[3], lv = so[2] + T[4];
  wf.camXY[p] = make_double2(imageX, imageY);
  wf.camLens[p] = make_double2(lu > 1 ? (lu - 1) : lu, lv > 1 ? (lv - 1) : lv);
  wf.camTime[p] = t > 1 ? (t - 1) : t;  // best_candidate_sampler.dart:111: a Dart double
}

#endif  // DRT_PATH_ONLY
// The fourth lane of a renderer ray's origin: scenes with TransformedPrimitives carry the wavefront slot there (the traversal looks the
// ray's time up in Wavefront::slotTime — every ray a camera sample spawns inherits its time, ray.dart:59); the interval itself always
// travels in the f64 range arrays.
static __device__ __forceinline__ float rayLaneW(const Wavefront& wf, uint32_t slot, float informational) {
  return wf.slotTime ? __uint_as_float(slot) : informational;
}

#ifndef DRT_PATH_ONLY  // the float32 path build (render_kernels_f32.cu) takes the path-vertex and resolve kernels only
// ---------------------------------------------------------------------------------------------------
// Camera rays (perspective_camera.dart:93-132; ray differentials only feed texture filtering and are
// not generated) + per-slot state reset.  Extension queue 0 = all slots in slot order.
__global__ void __launch_bounds__(256) raygenKernel(RenderParams rp, Wavefront wf, PixelBatch pb, uint32_t nSlots, RenderCounters* rc) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t qi = s;  // position in extension queue 0
  if (rp.samplerKind == 3 || rp.samplerKind == 5) {  // halton / bestcandidate: rejected indices never become rays (warp-uniform branch)
    const bool accepted = s < nSlots && !isnan(wf.camXY[s].x);
    qi = warpPush(&wf.counts[Q_EXT0], accepted);
    const unsigned m = __ballot_sync(FULL, accepted);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&rc->cameraSamples, (unsigned long long)__popc(m));
    if (!accepted) return;
  } else {
    if (s == 0) wf.counts[Q_EXT0] = nSlots;
    if (s >= nSlots) return;
  }
  const uint32_t n = (uint32_t)rp.nPixelSamples, cap = wf.cap;
  int x, y;
  pixelOf(pb, s / n, &x, &y);
  wf.pixX[s] = x;
  wf.pixY[s] = y;
  wf.sampleIdx[s] = pb.pass * n + (s % n);
  const double2 im = wf.camXY[s];
  V3 o = V3{0.f, 0.f, 0.f}, d;
  if (rp.cameraKind == 2) {  // environment_camera.dart:42-52
    const double theta = DRT_PI * im.y / rp.yres, phi = 2 * DRT_PI * im.x / rp.xres;
    d = mkv(sin(theta) * cos(phi), cos(theta), sin(theta) * sin(phi));
  } else {
    V3 Pras = mkv(im.x, im.y, 0.0);
    V3 Pcamera = XfPoint(rp.rasterToCamera, Pras);
    if (rp.cameraKind == 1) { o = Pcamera; d = V3{0.f, 0.f, 1.f}; }  // orthographic_camera.dart:52-58
    else d = Normalize(Pcamera);
  }
  if (rp.cameraKind != 2 && rp.lensRadius > 0.0) {
    const double2 ln = wf.camLens[s];
    double lu, lv;
    ConcentricSampleDisk(ln.x, ln.y, &lu, &lv);
    lu *= rp.lensRadius;
    lv *= rp.lensRadius;
    double ft = rp.focalDistance / d.z;
    V3 Pfocus = RayAt(o, d, ft);
    o = mkv(lu, lv, 0.0);
    d = Normalize(Pfocus - o);
  }
  const float* c2w = rp.cameraToWorld;
  M4 camM, camInv;
  if (rp.cameraMotion) {  // cameraToWorld.transformRay[Differential]: interpolate(ray.time) first (animated_transform.dart:138-169)
    animInterpolate(*rp.cameraMotion, LerpD(wf.camTime[s], rp.shutterOpen, rp.shutterClose), &camM, &camInv);
    c2w = camM.d;
  }
  V3 wo = XfPoint(c2w, o), wd = XfVector(c2w, d);
  if (wf.slotTime) wf.slotTime[s] = LerpD(wf.camTime[s], rp.shutterOpen, rp.shutterClose);  // the samplers' Lerp(time sample, shutterOpen, shutterClose)
  wf.extO[0][qi] = make_float4(wo.x, wo.y, wo.z, rayLaneW(wf, s, 0.f));
  wf.extD[0][qi] = make_float4(wd.x, wd.y, wd.z, CUDART_INF_F);
  wf.extRange[0][qi] = make_double2(0.0, CUDART_INF);
  wf.extSlot[0][qi] = s;
  st3(wf.L, cap, s, Spec{0.f, 0.f, 0.f});
  st3(wf.T, cap, s, Spec{1.f, 1.f, 1.f});
  wf.shIdx[s] = -1;
  wf.misIdx[s] = -1;
}

__global__ void resetCountsKernel(Wavefront wf, unsigned mask) {
  if (threadIdx.x < Q_COUNT && ((mask >> threadIdx.x) & 1u)) wf.counts[threadIdx.x] = 0;
}

#endif  // DRT_PATH_ONLY
// Sample-record access helpers (Sample.oneD / twoD, sample.dart:23-79)
static __device__ __forceinline__ float val(const Wavefront& wf, int v, uint32_t slot) { return wf.vals[(size_t)v * wf.cap + slot]; }

// integrator stream of a slot (sampler_renderer.dart:137 shares one RNG; here keyed per camera sample)
static __device__ __forceinline__ uint64_t integratorKey(const RenderParams& rp, const Wavefront& wf, uint32_t slot) {
  return streamKey(rp.seed, wf.pixX[slot], wf.pixY[slot], wf.sampleIdx[slot], DRT_STREAM_INTEGRATOR);
}

static __device__ __forceinline__ void pushDirectWork(const Wavefront& wf, uint32_t slot, bool valid, const DirectWork& w,
                                                      const V3& p, double rayEps, int lightNum) {
  const bool wantSh = valid && w.hasShadow, wantMis = valid && w.hasMis;
  uint32_t si, mi;
  warpPush2(&wf.counts[Q_SHADOW], wantSh, &wf.counts[Q_MIS], wantMis, &si, &mi);
  const uint32_t cap = wf.cap;
  if (wantSh) {
    wf.shO[si] = make_float4(w.shO.x, w.shO.y, w.shO.z, rayLaneW(wf, slot, (float)w.shMin));
    wf.shD[si] = make_float4(w.shD.x, w.shD.y, w.shD.z, (float)w.shMax);
#if !DRT_REAL32  // the float32 build's intervals ARE float32 values: they travel in the .w lanes of the ray records alone
    wf.shRange[si] = make_double2(w.shMin, w.shMax);
#endif
    st3(wf.pendSh, cap, slot, w.shContribution);
  }
  if (wantMis) {
    wf.misO[mi] = make_float4(p.x, p.y, p.z, rayLaneW(wf, slot, (float)rayEps));
    wf.misD[mi] = make_float4(w.misD.x, w.misD.y, w.misD.z, CUDART_INF_F);
#if !DRT_REAL32  // the float32 build's intervals ARE float32 values: they travel in the .w lanes of the ray records alone
    wf.misRange[mi] = make_double2(rayEps, CUDART_INF);
#endif
    st3(wf.pendMisF, cap, slot, w.misF);
    wf.pendMisScale[slot] = w.misScale;
    wf.misLight[slot] = lightNum;
  }
  if (valid) {
    wf.shIdx[slot] = wantSh ? (int32_t)si : -1;
    wf.misIdx[slot] = wantMis ? (int32_t)mi : -1;
  }
}

// tHit of a traced queue entry: the binary64 build reads the traversal kernels' f64 output (ray.maxDistance after the hit, ray.dart:36);
// the float32 build would round it to float32 first, which is what the hit record's first lane already holds — its trace calls ask
// for no f64 array at all (render_api.cu)
#if DRT_REAL32
#define DRT_EXT_THIT(wf, q) ((wf).extHit[q].x)
#define DRT_MIS_THIT(wf, mi) ((wf).misHit[mi].x)
#else
#define DRT_EXT_THIT(wf, q) ((wf).extT[q])
#define DRT_MIS_THIT(wf, mi) ((wf).misT[mi])
#endif
// ---------------------------------------------------------------------------------------------------
// Path integrator, one vertex (path_integrator.dart:44-119 loop body for `bounces` = bounce).
#ifndef DRT_SHADE_MIN_BLOCKS
#define DRT_SHADE_MIN_BLOCKS 4  // 128 registers: 16 warps/SM; 321 vs 277 Msamples/s at 2 (tools/shade_sweep.sh)
#endif
// GENERAL = false: every material is matte (one diffuse BxDF, never a specular bounce); true: BxDF lists.
// Counting sort of extension queue `cur` by material (BxDF-list scenes): histogram, one-block exclusive scan, scatter.
#define DRT_SORT_MAX_MATERIALS 1023
// keyMode 0: the material (BxDF-list scenes: a warp evaluates one BxDF list).  keyMode 1 (matte-only scenes): the SHAPE CLASS of the
// hit — triangle, then the quadric kinds — so that a warp rebuilds one kind of differential geometry; in both modes the rays that
// missed sort last, so the warps of the shading kernel are full of live vertices (a quarter of config 4's lanes idled on misses:
// profiles/r02u_shade_lines.txt shows at most 25 of 32 threads per instruction).
#define DRT_SHAPE_CLASSES 7  // triangle + GSphere::shape 0..5
#ifndef DRT_SHAPE_SORT_BUILD
#define DRT_SHAPE_SORT_BUILD 0  // 1: the matte-only path kernel can walk the queue in shape-class order (env DRT_SHAPE_SORT=1); see launchShadePath
#endif
static __device__ __forceinline__ uint32_t sortBins(const RenderScene& rs, int keyMode) {
  return keyMode ? (uint32_t)DRT_SHAPE_CLASSES + 1u : (uint32_t)rs.nMaterials + 1u;
}
static __device__ __forceinline__ uint32_t shadeKey(const RenderScene& rs, const Wavefront& wf, uint32_t q, int keyMode) {
  const int prim = __float_as_int(wf.extHit[q].w);
  if (keyMode) {
    if (prim < 0) return (uint32_t)DRT_SHAPE_CLASSES;
    if ((uint32_t)prim < rs.ntris) return 0u;
    const int sh = rs.ts.spheres[(uint32_t)prim - rs.ntris].shape;
    return 1u + (uint32_t)(sh < 0 ? 0 : (sh > 5 ? 5 : sh));
  }
  return prim < 0 ? (uint32_t)rs.nMaterials : (uint32_t)primMaterial(rs, (uint32_t)prim);
}
// Both passes aggregate per warp first (__match_any_sync on the key): a scene has a handful of materials, so per-entry atomics
// would all land on the same few counters.  AGG = false keeps the per-entry atomics (DRT_SORT_PLAIN_ATOMICS, for A/B runs).
template <bool AGG>
__global__ void __launch_bounds__(256) matHistKernel(RenderScene rs, Wavefront wf, int cur, int keyMode) {
  __shared__ uint32_t h[DRT_SORT_MAX_MATERIALS + 1];
  const uint32_t n = wf.counts[cur], bins = sortBins(rs, keyMode), lane = threadIdx.x & 31u;
  for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) h[b] = 0;
  __syncthreads();
  for (uint32_t q0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); q0 < n; q0 += gridDim.x * blockDim.x) {  // warp-uniform trip count
    const uint32_t q = q0 + lane;
    const bool v = q < n;
    const uint32_t key = v ? shadeKey(rs, wf, q, keyMode) : 0xffffffffu;
    if (AGG) {
      const uint32_t peers = __match_any_sync(FULL, key);
      if (v && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&h[key], (uint32_t)__popc(peers));
    } else if (v) {
      atomicAdd(&h[key], 1u);
    }
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x)
    if (h[b]) atomicAdd(&wf.matHist[b], h[b]);
}
__global__ void matScanKernel(RenderScene rs, Wavefront wf, int keyMode) {  // one thread: at most 1024 bins
  uint32_t acc = 0;
  const int bins = (int)sortBins(rs, keyMode);
  for (int b = 0; b < bins; ++b) {
    const uint32_t c = wf.matHist[b];
    wf.matHist[b] = acc;
    acc += c;
  }
}
template <bool AGG>
__global__ void __launch_bounds__(256) matScatterKernel(RenderScene rs, Wavefront wf, int cur, int keyMode) {
  const uint32_t n = wf.counts[cur], lane = threadIdx.x & 31u;
  for (uint32_t q0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); q0 < n; q0 += gridDim.x * blockDim.x) {
    const uint32_t q = q0 + lane;
    const bool v = q < n;
    const uint32_t key = v ? shadeKey(rs, wf, q, keyMode) : 0xffffffffu;
    if (AGG) {  // one atomic per (warp, material): the leader reserves the run, the lanes take consecutive places in lane order
      const uint32_t peers = __match_any_sync(FULL, key);
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (v && (int)lane == leader) base = atomicAdd(&wf.matHist[key], (uint32_t)__popc(peers));
      base = __shfl_sync(FULL, base, leader);
      if (v) wf.shadeOrder[base + __popc(peers & ((1u << lane) - 1u))] = q;
    } else if (v) {
      wf.shadeOrder[atomicAdd(&wf.matHist[key], 1u)] = q;
    }
  }
}

// EXTRA: see hitGeometry (shade_device.cuh).  SORTED: work item i of the launch is queue entry shadeOrder[i].
// PART: 0 the whole vertex in one kernel; 1 / 2 the same work as two launches — emitted light + the direct-lighting set-up (shadow and
// MIS rays), then the next direction + Russian roulette + the extension ray — each recomputing the cheap hit geometry / BSDF frame.
// Smaller kernels (the single one is 11.5 K instructions: ncu shows `no_instruction` stalls of 2.3 warps per issue) with fewer live
// registers; the per-slot draws keep their positions in the integrator stream, so the results are the single kernel's.
#ifndef DRT_SHADE_NOLOOP
#define DRT_SHADE_NOLOOP 0  // 1 (timing experiment, ray statistics not counted): one queue entry per thread instead of the grid-stride
                             // loop, tried because the float32 kernel moves as many local-memory sectors as global ones (ncu; a 712-byte
                             // frame, the RenderScene parameter copied to the stack at entry for the out-of-line callees).  Measured on B200 (profiles/r02z13_noloop_ab.log): 0.5822 s against
                             // 0.4658 s — 131 K short CTAs cost far more than the spills.  Rejected, off
#endif
#ifndef DRT_SHADE_PREFETCH
#define DRT_SHADE_PREFETCH 0  // 1: prefetch the thread's next queue entry.  Measured on B200 with the float32 kernel (config 4,
                              // profiles/r02z11_prefetch_ab.log): 0.4716 s against 0.4672 s — rejected, off
#endif
#ifndef DRT_SHADE_SPLIT_MIN_BLOCKS
#define DRT_SHADE_SPLIT_MIN_BLOCKS 5
#endif
template <bool GENERAL, bool EXTRA, int PART>
__global__ void __launch_bounds__(128, PART == 0 ? DRT_SHADE_MIN_BLOCKS : DRT_SHADE_SPLIT_MIN_BLOCKS) shadePathKernel(RenderParams rp, RenderScene rs, Wavefront wf, int bounce, int cur,
                                                       RenderCounters* rc, int sortedOrder) {
  const uint32_t n = wf.counts[cur], cap = wf.cap;
  const int nxt = cur ^ 1;
  unsigned long long nShadow = 0, nClosest = 0;
  const bool sorted = (GENERAL || DRT_SHAPE_SORT_BUILD) && sortedOrder != 0;
#if DRT_SHADE_NOLOOP  // EXPERIMENT (timing only: the ray statistics are not counted): one queue entry per thread, no grid-stride loop
  for (uint32_t q0 = blockIdx.x * blockDim.x; q0 < n; q0 = 0xffffffffu) {
#else
  for (uint32_t q0 = blockIdx.x * blockDim.x; q0 < n; q0 += gridDim.x * blockDim.x) {
#endif
    uint32_t q = q0 + threadIdx.x;
    bool valid = q < n;
#if DRT_SHADE_PREFETCH  // the queue entry this thread shades in its NEXT trip, on its way to L1 while this one is shaded (A/B knob)
    {
      const uint32_t qn = q + gridDim.x * blockDim.x;
      if (qn < n && !sorted) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(wf.extSlot[cur] + qn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(wf.extHit + qn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(wf.extO[cur] + qn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(wf.extD[cur] + qn));
      }
    }
#endif
    uint32_t slot = 0;
    int prim = -1;
    if (valid) {
      if (sorted) q = wf.shadeOrder[q];
      slot = wf.extSlot[cur][q];
      prim = __float_as_int(wf.extHit[q].w);
      if (PART != 2) {
        wf.shIdx[slot] = -1;
        wf.misIdx[slot] = -1;
      }
      valid = prim >= 0;  // miss: the path ends; area/point lights add no Le along escaping rays
    }
    if (DRT_SHAPE_SORT_BUILD && sorted && !__any_sync(FULL, valid)) continue;  // the misses sort last: their warps have nothing to shade or to push
    DirectWork dw;
    dw.hasShadow = dw.hasMis = false;
    bool cont = false;
    V3 p = V3{0.f, 0.f, 0.f}, wi = V3{0.f, 0.f, 0.f};
    double rayEps = 0.0;
    int lightNum = 0;
    if (valid) {
      const float4 o4 = wf.extO[cur][q], d4 = wf.extD[cur][q];
      const V3 o = V3{o4.x, o4.y, o4.z}, d = V3{d4.x, d4.y, d4.z};
      ShapeHit h;
      hitGeometryQ<EXTRA>(rs, wf, q, slot, (uint32_t)prim, o, d, DRT_EXT_THIT(wf, q), &h);
      Spec T = ld3(wf.T, cap, slot);
      if (EXTRA && rs.nVolumes > 0 && bounce > 0) {  // pathThroughput *= renderer.transmittance(ray) once the ray found this vertex (:116)
        uint32_t ctr = wf.trCtr[slot];
        Spec tr;
        volTransmittanceDrawCold(rs, streamKey(rp.seed, wf.pixX[slot], wf.pixY[slot], wf.sampleIdx[slot], DRT_STREAM_TRANSMITTANCE), &ctr, o, d,
                                 wf.extRange[cur][q].x, wf.extT[q], &tr);
        wf.trCtr[slot] = ctr;
        T = T * tr;
      }
      const V3 wo = -d;
      // emitted light at the first vertex or after a specular bounce (:46-48); matte BSDFs never set specularBounce
      if (PART != 2 && (bounce == 0 || (GENERAL && wf.specBounce[slot]))) {
        const int li = primLight(rs, (uint32_t)prim);
        if (li >= 0) {
          Spec L = ld3(wf.L, cap, slot);
          L = L + T * areaL(rs.lights[li], h.nn, wo);
          st3(wf.L, cap, slot, L);
        }
      }
      typename BsdfOf<GENERAL>::type bsdf = makeBsdfT<GENERAL, EXTRA>(rs, (uint32_t)prim, h, o, d);
      if (GENERAL && EXTRA) applyHitBsdf(rs, wf, slot, &bsdf);
      p = h.p;
      rayEps = h.rayEps;
      const V3 nrm = bsdf.nn;
      Stream rng{integratorKey(rp, wf, slot), 0};
      if (bounce > 3) {  // draws already consumed by bounces 3..bounce-1: 10 (or 3 without lights) each, +1 RR from bounce 4 on
        const uint64_t per = rs.nLights > 0 ? 10 : 3;
        rng.ctr = (uint64_t)(bounce - 3) * per + (uint64_t)(bounce - 4);
      }
      // direct lighting: UniformSampleOneLight (integrator.dart:79-117)
      if (PART == 2) {
        if (rs.nLights > 0 && bounce >= 3) rng.ctr += 7;  // the draws the direct-lighting launch consumed: light number, LightSample, BSDFSample
      } else if (rs.nLights > 0) {
        float lu0, lu1, bu0, bu1;
        double lcomp, bcomp;
        if (bounce < 3) {
          lightNum = (int)floor((double)val(wf, rp.pLightNum[bounce], slot) * (double)rs.nLights);
          lu0 = val(wf, rp.pLightPos[bounce], slot); lu1 = val(wf, rp.pLightPos[bounce] + 1, slot);
          lcomp = val(wf, rp.pLightComp[bounce], slot);
          bu0 = val(wf, rp.pBsdfPos[bounce], slot); bu1 = val(wf, rp.pBsdfPos[bounce] + 1, slot);
          bcomp = GENERAL ? (double)val(wf, rp.pBsdfComp[bounce], slot) : 0.0;
        } else {
          lightNum = (int)floor(rng.randomFloat() * rs.nLights);
          lu0 = (float)rng.randomFloat(); lu1 = (float)rng.randomFloat(); lcomp = rng.randomFloat();  // LightSample.random
          bu0 = (float)rng.randomFloat(); bu1 = (float)rng.randomFloat(); bcomp = rng.randomFloat();  // BSDFSample.random
        }
        lightNum = min(lightNum, rs.nLights - 1);
        estimateDirectSetup(rs, lightNum, p, nrm, wo, rayEps, bsdf, lu0, lu1, lcomp, bu0, bu1, bcomp, BSDF_ALL & ~BSDF_SPECULAR, &dw);
        st3(wf.pendT, cap, slot, T);
      }
      // next direction (:63-92)
      if (PART != 1) {
      float pu0, pu1;
      double pcomp;
      if (bounce < 3) {
        pu0 = val(wf, rp.pPathPos[bounce], slot); pu1 = val(wf, rp.pPathPos[bounce] + 1, slot);
        pcomp = GENERAL ? (double)val(wf, rp.pPathComp[bounce], slot) : 0.0;
      } else {
        pu0 = (float)rng.randomFloat(); pu1 = (float)rng.randomFloat(); pcomp = rng.randomFloat();
      }
      double pdf = 0.0;
      int flags = 0;
      Spec f = bsdfSampleF(bsdf, wo, &wi, pu0, pu1, pcomp, &pdf, BSDF_ALL, &flags);
      if (!(IsBlack(f) || pdf == 0.0)) {
        if (GENERAL) wf.specBounce[slot] = (flags & BSDF_SPECULAR) != 0 ? 1 : 0;
        T = T * (f * AbsDot(wi, nrm) / pdf);
        cont = true;
        if (bounce > 3) {  // Russian roulette (:93-99)
          const double lumT = Luminance(T);
          double continueProbability = lumT != lumT ? lumT : fmin(0.5, lumT);  // Dart's Math.min propagates NaN (path_integrator.dart:94)
          if (rng.randomFloat() > continueProbability) cont = false;
          else T = T / continueProbability;
        }
        if (bounce == rp.maxDepth) cont = false;
        if (cont) st3(wf.T, cap, slot, T);
      }
      }  // PART != 1
    }
    if (PART != 2) pushDirectWork(wf, slot, valid, dw, p, rayEps, lightNum);
    const uint32_t ei = PART == 1 ? 0u : warpPush(&wf.counts[nxt], cont);
    if (cont) {
      wf.extO[nxt][ei] = make_float4(p.x, p.y, p.z, rayLaneW(wf, slot, (float)rayEps));
      wf.extD[nxt][ei] = make_float4(wi.x, wi.y, wi.z, CUDART_INF_F);
#if !DRT_REAL32  // the float32 build's intervals ARE float32 values: they travel in the .w lanes of the ray records alone
      wf.extRange[nxt][ei] = make_double2(rayEps, CUDART_INF);
#endif
      wf.extSlot[nxt][ei] = slot;
    }
#if !DRT_SHADE_NOLOOP
    nShadow += (valid && dw.hasShadow) ? 1 : 0;
    nClosest += ((valid && dw.hasMis) ? 1 : 0) + (cont ? 1 : 0);
#endif
  }
#if DRT_SHADE_NOLOOP
  return;
#endif
  // ray statistics (stats.dart:541-555): one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    nShadow += __shfl_down_sync(FULL, nShadow, o);
    nClosest += __shfl_down_sync(FULL, nClosest, o);
  }
  if ((threadIdx.x & 31) == 0 && (nShadow | nClosest)) {
    atomicAdd(&rc->shadowRays, nShadow);
    atomicAdd(&rc->closestRays, nClosest);
  }
}

// Finishes Integrator.EstimateDirect (integrator.dart:119-185) once the shadow and MIS rays of the
// vertices shaded from extension queue `cur` are traced, and adds the estimate where the integrator
// adds it.  mode: see RESOLVE_* (render_kernels.h).
template <bool EXTRA>
__global__ void __launch_bounds__(128) resolveDirectKernel(RenderParams rp, RenderScene rs, Wavefront wf, int cur, int mode,
                                                           int nSamplesOfLight) {
  const uint32_t n = wf.counts[cur], cap = wf.cap;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const uint32_t slot = wf.extSlot[cur][q];
    const int si = wf.shIdx[slot], mi = wf.misIdx[slot];
    const bool direct = (mode & 1) != 0;
    if (direct) {
      if (__float_as_int(wf.extHit[q].w) < 0) continue;  // camera ray missed
    } else if (si < 0 && mi < 0) {
      continue;
    }
    Spec Ld = mks1(0.0);
    const bool media = EXTRA && rs.nVolumes > 0;  // Li *= transmittance along the unoccluded shadow ray (integrator.dart:137)
    uint64_t trKey = 0;
    uint32_t trCtr = 0;
    if (media) {
      trKey = streamKey(rp.seed, wf.pixX[slot], wf.pixY[slot], wf.sampleIdx[slot], DRT_STREAM_TRANSMITTANCE);
      trCtr = wf.trCtr[slot];
    }
    if (si >= 0 && !wf.shOcc[si]) {
      Spec c = ld3(wf.pendSh, cap, slot);
      if (media) {
        const float4 so = wf.shO[si], sd = wf.shD[si];
        const double2 sr = wf.shRange[si];
        Spec tr;
        if (mode & 64) volTransmittanceSampleCold(rs, (double)val(wf, rp.pTauSample, slot), V3{so.x, so.y, so.z}, V3{sd.x, sd.y, sd.z}, sr.x, sr.y, &tr);
        else volTransmittanceDrawCold(rs, trKey, &trCtr, V3{so.x, so.y, so.z}, V3{sd.x, sd.y, sd.z}, sr.x, sr.y, &tr);
        c = c * tr;  // every factor of the contribution is a product with Li: the order of the float32 roundings is Li's first
      }
      Ld = Ld + c;
    }
    if (mi >= 0) {
      const int prim = __float_as_int(wf.misHit[mi].w);
      const int light = wf.misLight[slot];
      if (prim >= 0 && primLight(rs, (uint32_t)prim) == light) {
        const float4 o4 = wf.misO[mi], d4 = wf.misD[mi];
        const V3 o = V3{o4.x, o4.y, o4.z}, wi = V3{d4.x, d4.y, d4.z};
        ShapeHit h;
        hitGeometry<EXTRA>(rs, (uint32_t)prim, o, wi, DRT_MIS_THIT(wf, mi), &h);
        Spec Li = areaL(rs.lights[light], h.nn, -wi);  // Intersection.Le, intersection.dart:62-65
        if (!IsBlack(Li)) {
          if (media) {  // renderer.transmittance along the ray up to the light's surface (integrator.dart:178)
            Spec tr;
            volTransmittanceDrawCold(rs, trKey, &trCtr, o, wi, wf.misRange[mi].x, wf.misT[mi], &tr);
            Li = Li * tr;
          }
          Ld = Ld + ld3(wf.pendMisF, cap, slot) * Li * wf.pendMisScale[slot];
        }
      } else if (EXTRA && prim < 0 && rs.lights[light].kind == 4) {  // the BSDF-sampled ray escaped: Li = light.Le(ray), integrator.dart:172-174
        const float4 d4 = wf.misD[mi];
        Spec Li = infiniteLe(rs, rs.lights[light], V3{d4.x, d4.y, d4.z});
        if (!IsBlack(Li)) {
          if (media) {
            const float4 o4 = wf.misO[mi];
            Spec tr;
            volTransmittanceDrawCold(rs, trKey, &trCtr, V3{o4.x, o4.y, o4.z}, V3{d4.x, d4.y, d4.z}, wf.misRange[mi].x, CUDART_INF, &tr);
            Li = Li * tr;
          }
          Ld = Ld + ld3(wf.pendMisF, cap, slot) * Li * wf.pendMisScale[slot];
        }
      }
    }
    wf.shIdx[slot] = -1;
    wf.misIdx[slot] = -1;
    if (media) wf.trCtr[slot] = trCtr;
    if (!direct) {  // path: L += pathThroughput * (EstimateDirect * nLights)
      Spec L = ld3(wf.L, cap, slot);
      L = L + ld3(wf.pendT, cap, slot) * (Ld * (double)rs.nLights);
      st3(wf.L, cap, slot, L);
    } else if (mode & (16 | 64)) {  // directlighting, strategy one (integrator.dart:79-117); whitted: the sample as it is
      Spec L = ld3(wf.L, cap, slot);
      const Spec est = (mode & 64) ? Ld : Ld * (double)rs.nLights;
      L = L + ((mode & 32) ? ld3(wf.pendT, cap, slot) * est : est);
      st3(wf.L, cap, slot, L);
    } else {  // UniformSampleAllLights (integrator.dart:39-77): Ld over the light's samples, L over lights (kept in T)
      Spec acc = (mode & 2) ? mks1(0.0) : ld3(wf.Ld, cap, slot);
      acc = acc + Ld;
      if (mode & 4) {
        Spec all = ld3(wf.T, cap, slot);
        all = all + acc / (double)nSamplesOfLight;
        if (mode & 8) {
          Spec L = ld3(wf.L, cap, slot);
          st3(wf.L, cap, slot, L + ((mode & 32) ? ld3(wf.pendT, cap, slot) * all : all));
        } else {
          st3(wf.T, cap, slot, all);
        }
      } else {
        st3(wf.Ld, cap, slot, acc);
      }
    }
  }
}

// Rays of queue `cur` that escaped the scene pick up the infinite lights' Le (launched only when the scene has one):
//   ESCAPE_CAMERA    Li = sum of lights.Le(ray)                          sampler_renderer.dart:86-92
//   ESCAPE_PATH      L += pathThroughput * Le per light, after a specular bounce only   path_integrator.dart:106-114
//   ESCAPE_WEIGHTED  the same sum through renderer.Li at the end of a specular chain, times the chain weight (integrator.dart:187-290)
__global__ void __launch_bounds__(128) escapeKernel(RenderScene rs, Wavefront wf, int cur, int mode) {
  const uint32_t n = wf.counts[cur], cap = wf.cap;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    if (__float_as_int(wf.extHit[q].w) >= 0) continue;
    const uint32_t slot = wf.extSlot[cur][q];
    if (mode == ESCAPE_PATH && !wf.specBounce[slot]) continue;
    const float4 d4 = wf.extD[cur][q];
    const V3 d = V3{d4.x, d4.y, d4.z};
    Spec L = ld3(wf.L, cap, slot);
    if (mode == ESCAPE_PATH) {
      const Spec T = ld3(wf.T, cap, slot);
      for (int i = 0; i < rs.nLights; ++i)
        if (rs.lights[i].kind == 4) L = L + T * infiniteLe(rs, rs.lights[i], d);
    } else {
      Spec Li = mks1(0.0);
      for (int i = 0; i < rs.nLights; ++i)
        if (rs.lights[i].kind == 4) Li = Li + infiniteLe(rs, rs.lights[i], d);
      // T * Li + Lvi of SamplerRenderer.Li with T = 1, Lvi = 0, as the oracle writes it
      Li = mks1(1.0) * Li + mks1(0.0);
      L = (mode == ESCAPE_WEIGHTED) ? L + ld3(wf.pendT, cap, slot) * Li : Li;
    }
    st3(wf.L, cap, slot, L);
  }
}

#ifndef DRT_PATH_ONLY  // the float32 path build (render_kernels_f32.cu) takes the path-vertex and resolve kernels only
// ---------------------------------------------------------------------------------------------------
// Ambient occlusion (ambient_occlusion_integrator.dart:28-53)
__global__ void __launch_bounds__(128) aoSetupKernel(RenderParams rp, RenderScene rs, Wavefront wf) {
  const uint32_t n = wf.counts[Q_EXT0], cap = wf.cap;
  for (uint32_t q0 = blockIdx.x * blockDim.x; q0 < n; q0 += gridDim.x * blockDim.x) {
    const uint32_t q = q0 + threadIdx.x;
    bool hit = false;
    uint32_t slot = 0;
    if (q < n) {
      slot = wf.extSlot[0][q];
      const int prim = __float_as_int(wf.extHit[q].w);
      hit = prim >= 0;
      if (hit) {
        const float4 o4 = wf.extO[0][q], d4 = wf.extD[0][q];
        const V3 o = V3{o4.x, o4.y, o4.z}, d = V3{d4.x, d4.y, d4.z};
        ShapeHit h;
        hitGeometryQ(rs, wf, q, slot, (uint32_t)prim, o, d, wf.extT[q], &h);
        stv3(wf.hitP, cap, slot, h.p);
        stv3(wf.hitN, cap, slot, FaceForward(h.nn, -d));
        Stream rng{integratorKey(rp, wf, slot), 0};
        wf.aoScramble[slot] = rng.randomUint();
        wf.aoScramble[cap + slot] = rng.randomUint();
        wf.nClear[slot] = 0;
      }
    }
    const uint32_t hi = warpPush(&wf.counts[Q_HITS], hit);
    if (hit) wf.hitList[hi] = slot;
  }
}

// Where the AO rays of a chunk sit in the shadow queue.  The reference's order — hit after hit, sample after sample — puts the 2^k
// samples of ONE hit side by side, and a (0,2)-sequence spreads consecutive samples over the whole hemisphere: a warp of the any-hit
// kernel then holds 32 rays of one origin pointing everywhere.  The occlusion count of a hit does not depend on the order its rays are
// traced in, so the queue is laid out for the traversal instead: the first 2^k points of the scrambled (0,2)-sequence form a
// (0, k, 2)-net (montecarlo.dart:486-504: base-2 digit scrambling keeps the property), i.e. every cell of the 2^a x 2^b grid over
// [0,1)^2, a + b = k, holds exactly one sample of a hit — its cell number is a bijection of the sample index.  Blocks of 2^16
// consecutive hits (neighbouring pixels) are transposed: consecutive queue entries = the SAME cell of consecutive hits, rays of
// nearby origins and directions.  DRT_AO_PLAIN_ORDER=1 keeps the reference's order (A/B runs).
static __device__ __forceinline__ __attribute__((unused)) uint32_t aoCell(uint32_t i, uint32_t s0, uint32_t s1, int k) {
  const int a = (k + 1) >> 1, b = k >> 1;
  const uint32_t v0 = __brev(i) ^ s0, v1 = sobolBits(i) ^ s1;  // the integers VanDerCorput / Sobol2 turn into [0, 1) values
  const uint32_t cx = a ? v0 >> (32 - a) : 0u, cy = b ? v1 >> (32 - b) : 0u;
  return (cx << b) | cy;
}
#ifndef DRT_AO_BLOCK_LOG2
#define DRT_AO_BLOCK_LOG2 16  // hits per transposed block.  Measured on B200 (profiles/r02z15 - r02z17_ao_block_ab.log; config 3 / AO on the
                              // Cornell scene): 2^5 19.12 / 34.4 ms, 2^8 18.98 / 33.5, 2^13 18.89 / 33.0, 2^16 18.85 / 32.3, 2^20 18.96 / 33.0
#endif
#define DRT_AO_BLOCK (1u << DRT_AO_BLOCK_LOG2)
static __device__ __forceinline__ uint64_t aoRayPos(uint32_t hI, uint32_t cell, uint32_t hitsHere, uint32_t nS, int plain) {
  if (plain) return (uint64_t)hI * nS + cell;
  const uint32_t B = hI >> DRT_AO_BLOCK_LOG2, first = B << DRT_AO_BLOCK_LOG2;
  const uint32_t m = min(DRT_AO_BLOCK, hitsHere - first);  // the last block of a chunk is shorter
  return (uint64_t)first * nS + (uint64_t)cell * m + (hI - first);
}

// AO rays of hits [firstHit, firstHit + maxHits) of the hit list, nSamples each, into the shadow queue.  One thread per QUEUE ENTRY, so
// that the ray records are written coalesced: in the transposed part of the queue entry r is cell (r / 32) % nS of hit 32 * block +
// r % 32, and the sample index that falls into that cell follows from the two generator matrices — the low a bits of i from the
// cell's x (the van der Corput digits are the reversed index), the high b bits through the inverse of the linear map
// x -> top b bits of Sobol2's direction-number sum of (x << a), tabulated per block (aoCell is the forward map; the films of
// tests/test_render_gpu.py::test_ambient_occlusion_ray_queue_layout_keeps_every_sample are bit-identical to the oracle's only if
// the two agree for every sample).
#define DRT_AO_MAX_B 10  // cells grids up to 2^11 x 2^10 (nSamples <= 2^21); beyond that the plain layout
__global__ void __launch_bounds__(256) aoGenKernel(RenderParams rp, Wavefront wf, uint32_t firstHit, uint32_t maxHits, int nS, int plain) {
  __shared__ uint16_t invHi[1 << DRT_AO_MAX_B];
  const uint32_t nHits = wf.counts[Q_HITS], cap = wf.cap;
  const uint32_t hitsHere = nHits > firstHit ? min(nHits - firstHit, maxHits) : 0u;
  const uint64_t nRays = (uint64_t)hitsHere * nS;
  const int k = 31 - __clz(nS), a = (k + 1) >> 1, b = k >> 1;
  if (!plain) {
    for (uint32_t x = threadIdx.x; x < (1u << b); x += blockDim.x) invHi[b ? sobolBits(x << a) >> (32 - b) : 0u] = (uint16_t)x;
    __syncthreads();
  }
  const uint64_t blockRays = (uint64_t)DRT_AO_BLOCK * (uint64_t)nS;
  if (blockIdx.x == 0 && threadIdx.x == 0) wf.counts[Q_SHADOW] = (uint32_t)nRays;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nRays; r += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t hI, i, slot, s0, s1;
    if (!plain) {
      const uint32_t B = (uint32_t)(r / blockRays), first = B << DRT_AO_BLOCK_LOG2;
      const uint32_t m = min(DRT_AO_BLOCK, hitsHere - first);
      const uint64_t rem = r - (uint64_t)first * nS;
      const uint32_t cell = (uint32_t)(rem / m);
      hI = first + (uint32_t)(rem - (uint64_t)cell * m);
      slot = wf.hitList[firstHit + hI];
      s0 = wf.aoScramble[slot]; s1 = wf.aoScramble[cap + slot];
      const uint32_t cx = cell >> b, cy = cell & ((1u << b) - 1u);
      const uint32_t iLo = a ? __brev(cx ^ (s0 >> (32 - a))) >> (32 - a) : 0u;
      const uint32_t tgt = b ? (cy ^ ((sobolBits(iLo) ^ s1) >> (32 - b))) : 0u;
      i = iLo | ((uint32_t)invHi[tgt] << a);
#ifdef DRT_AO_SELFCHECK  // the inverse against the forward map
      if (aoCell(i, s0, s1, k) != cell) __trap();
#endif
    } else {
      hI = (uint32_t)(r / nS);
      i = (uint32_t)(r % nS);
      slot = wf.hitList[firstHit + hI];
      s0 = wf.aoScramble[slot]; s1 = wf.aoScramble[cap + slot];
    }
    const double u0 = VanDerCorput(i, s0), u1 = Sobol2(i, s1);
    V3 w = UniformSampleSphere(u0, u1);
    const V3 nrm = ldv3(wf.hitN, cap, slot), p = ldv3(wf.hitP, cap, slot);
    if (Dot(w, nrm) < 0.0) w = -w;
    // new Ray(p, w, minDist, maxDist) (ambient_occlusion_integrator.dart:45): no time argument, the ray travels at time 0
    wf.shO[r] = make_float4(p.x, p.y, p.z, wf.slotTime ? __uint_as_float(0xffffffffu) : (float)rp.aoMinDist);
    wf.shD[r] = make_float4(w.x, w.y, w.z, (float)rp.aoMaxDist);
    wf.shRange[r] = make_double2(rp.aoMinDist, rp.aoMaxDist);
  }
}

__global__ void __launch_bounds__(256) aoCountKernel(RenderParams rp, Wavefront wf, uint32_t firstHit, uint32_t maxHits, int nS,
                                                     RenderCounters* rc, int plain) {
  const uint32_t nHits = wf.counts[Q_HITS], cap = wf.cap;
  const uint32_t hitsHere = nHits > firstHit ? min(nHits - firstHit, maxHits) : 0u;
  for (uint32_t hI = blockIdx.x * blockDim.x + threadIdx.x; hI < hitsHere; hI += gridDim.x * blockDim.x) {
    const uint32_t slot = wf.hitList[firstHit + hI];
    int nClear = 0;
    for (int c = 0; c < nS; ++c) nClear += wf.shOcc[aoRayPos(hI, (uint32_t)c, hitsHere, (uint32_t)nS, plain)] ? 0 : 1;
    st3(wf.L, cap, slot, mks1((double)nClear / nS));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&rc->shadowRays, (unsigned long long)hitsHere * nS);
}

// ---------------------------------------------------------------------------------------------------
// Direct lighting (direct_lighting_integrator.dart:30-68): emitted light at the camera hit, then one
// launch per (light, sample) of UniformSampleAllLights, or one launch of UniformSampleOneLight.
// `cur`: the extension queue holding the vertices; `weighted`: the vertices are those of a specular chain
// (integrator.dart:187-290) and contribute with the chain's weight (kept in pendT), added to what L already holds.
__global__ void __launch_bounds__(128) directSetupKernel(RenderParams rp, RenderScene rs, Wavefront wf, int cur, int weighted) {
  const uint32_t n = wf.counts[cur], cap = wf.cap;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const uint32_t slot = wf.extSlot[cur][q];
    const int prim = __float_as_int(wf.extHit[q].w);
    st3(wf.T, cap, slot, Spec{0.f, 0.f, 0.f});  // L of UniformSampleAllLights
    if (prim < 0) continue;
    const int li = primLight(rs, (uint32_t)prim);
    if (li < 0) continue;
    const float4 o4 = wf.extO[cur][q], d4 = wf.extD[cur][q];
    const V3 o = V3{o4.x, o4.y, o4.z}, d = V3{d4.x, d4.y, d4.z};
    ShapeHit h;
    hitGeometry(rs, (uint32_t)prim, o, d, wf.extT[q], &h);
    const Spec Le = mks1(0.0) + areaL(rs.lights[li], h.nn, -d);
    if (!weighted) st3(wf.L, cap, slot, Le);
    else st3(wf.L, cap, slot, ld3(wf.L, cap, slot) + ld3(wf.pendT, cap, slot) * Le);
  }
}

// Whitted integrator (whitted_integrator.dart:26-78) on the vertices of queue `cur`: emitted light and the stream
// position of the vertex's light samples (one LightSample.random(rng) per light, drawn before the specular branch calls)
__global__ void __launch_bounds__(128) whittedSetupKernel(RenderParams rp, RenderScene rs, Wavefront wf, int cur, int weighted) {
  const uint32_t n = wf.counts[cur], cap = wf.cap;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const uint32_t slot = wf.extSlot[cur][q];
    const int prim = __float_as_int(wf.extHit[q].w);
    if (prim < 0) continue;
    const uint32_t base = wf.specCtr[slot];
    wf.aoScramble[slot] = base;  // free in this integrator: the counter the vertex's first light sample starts from
    wf.specCtr[slot] = base + 3u * (uint32_t)rs.nLights;
    const int li = primLight(rs, (uint32_t)prim);
    if (li < 0) continue;
    const float4 o4 = wf.extO[cur][q], d4 = wf.extD[cur][q];
    const V3 o = V3{o4.x, o4.y, o4.z}, d = V3{d4.x, d4.y, d4.z};
    ShapeHit h;
    hitGeometry(rs, (uint32_t)prim, o, d, wf.extT[q], &h);
    const Spec Le = mks1(0.0) + areaL(rs.lights[li], h.nn, -d);
    if (!weighted) st3(wf.L, cap, slot, Le);
    else st3(wf.L, cap, slot, ld3(wf.L, cap, slot) + ld3(wf.pendT, cap, slot) * Le);
  }
}

template <bool GENERAL>
__global__ void __launch_bounds__(128) whittedSampleKernel(RenderParams rp, RenderScene rs, Wavefront wf, int light, int cur,
                                                           RenderCounters* rc, int sortedOrder) {
  const uint32_t n = wf.counts[cur];
  unsigned long long nShadow = 0;
  const bool sorted = GENERAL && sortedOrder != 0;
  for (uint32_t q0 = blockIdx.x * blockDim.x; q0 < n; q0 += gridDim.x * blockDim.x) {
    uint32_t q = q0 + threadIdx.x;
    bool valid = q < n;
    uint32_t slot = 0;
    int prim = -1;
    if (valid) {
      if (sorted) q = wf.shadeOrder[q];
      slot = wf.extSlot[cur][q];
      prim = __float_as_int(wf.extHit[q].w);
      valid = prim >= 0;
    }
    DirectWork dw;
    dw.hasShadow = dw.hasMis = false;
    V3 p = V3{0.f, 0.f, 0.f};
    double rayEps = 0.0;
    if (valid) {
      const float4 o4 = wf.extO[cur][q], d4 = wf.extD[cur][q];
      const V3 o = V3{o4.x, o4.y, o4.z}, d = V3{d4.x, d4.y, d4.z};
      ShapeHit h;
      hitGeometryQ(rs, wf, q, slot, (uint32_t)prim, o, d, wf.extT[q], &h);
      typename BsdfOf<GENERAL>::type bsdf = makeBsdfT<GENERAL>(rs, (uint32_t)prim, h, o, d);
      if (GENERAL) applyHitBsdf(rs, wf, slot, &bsdf);
      p = h.p;
      rayEps = h.rayEps;
      Stream rng{integratorKey(rp, wf, slot), (uint64_t)wf.aoScramble[slot] + 3ull * (uint64_t)light};
      const float lu0 = (float)rng.randomFloat(), lu1 = (float)rng.randomFloat();  // LightSample.random (light_sample.dart:45-51)
      const double lcomp = rng.randomFloat();
      whittedLightSetup(rs, light, p, bsdf.nn, -d, rayEps, bsdf, lu0, lu1, lcomp, &dw);
    }
    pushDirectWork(wf, slot, valid, dw, p, rayEps, light);
    nShadow += (valid && dw.hasShadow) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) nShadow += __shfl_down_sync(FULL, nShadow, o);
  if ((threadIdx.x & 31) == 0 && nShadow) atomicAdd(&rc->shadowRays, nShadow);
}

// One SpecularReflect / SpecularTransmit call (integrator.dart:187-290) for every vertex of queue `cur`: draws the
// BSDFSample.random of the call from the slot's integrator stream, samples the BSDF with `flags`, and, where the
// reference recurses, appends the child ray to the other queue and multiplies the chain weight (pendT) by
// f * |wi.n| / pdf.  `level` = 1 for the call at the camera vertex; `isNew`: this call has not been made before for
// this prefix (chains re-walk their prefixes with the counters recorded the first time).
__global__ void __launch_bounds__(128) specularStepKernel(RenderParams rp, RenderScene rs, Wavefront wf, int cur, int flags, int level,
                                                          int isNew, RenderCounters* rc) {
  const uint32_t n = wf.counts[cur], cap = wf.cap;
  const int nxt = cur ^ 1;
  unsigned long long nClosest = 0;
  for (uint32_t q0 = blockIdx.x * blockDim.x; q0 < n; q0 += gridDim.x * blockDim.x) {
    const uint32_t q = q0 + threadIdx.x;
    bool cont = false;
    uint32_t slot = 0;
    V3 p = V3{0.f, 0.f, 0.f}, wi = V3{0.f, 0.f, 0.f};
    double rayEps = 0.0;
    if (q < n) {
      slot = wf.extSlot[cur][q];
      const int prim = __float_as_int(wf.extHit[q].w);
      if (prim >= 0) {
        uint32_t base;
        if (isNew) {
          base = wf.specCtr[slot];
          wf.specCtrAt[(size_t)level * cap + slot] = base;
          wf.specCtr[slot] = base + 3;
        } else {
          base = wf.specCtrAt[(size_t)level * cap + slot];
        }
        Stream rng{integratorKey(rp, wf, slot), base};
        const float u0 = (float)rng.randomFloat(), u1 = (float)rng.randomFloat();
        const double comp = rng.randomFloat();
        const float4 o4 = wf.extO[cur][q], d4 = wf.extD[cur][q];
        const V3 o = V3{o4.x, o4.y, o4.z}, d = V3{d4.x, d4.y, d4.z};
        ShapeHit h;
        hitGeometryQ(rs, wf, q, slot, (uint32_t)prim, o, d, wf.extT[q], &h);
        BsdfG bsdf = makeBsdfG(rs, (uint32_t)prim, h, o, d);
        applyHitBsdf(rs, wf, slot, &bsdf);
        double pdf = 0.0;
        int type = 0;
        const Spec f = bsdfSampleF(bsdf, -d, &wi, u0, u1, comp, &pdf, flags, &type);
        const double an = AbsDot(wi, bsdf.nn);
        if (pdf > 0.0 && !IsBlack(f) && an != 0.0) {
          cont = true;
          p = h.p;
          rayEps = h.rayEps;
          const Spec w = f * (an / pdf);
          st3(wf.pendT, cap, slot, level == 1 ? w : ld3(wf.pendT, cap, slot) * w);
        }
      }
    }
    const uint32_t ei = warpPush(&wf.counts[nxt], cont);
    if (cont) {
      wf.extO[nxt][ei] = make_float4(p.x, p.y, p.z, rayLaneW(wf, slot, (float)rayEps));
      wf.extD[nxt][ei] = make_float4(wi.x, wi.y, wi.z, CUDART_INF_F);
      wf.extRange[nxt][ei] = make_double2(rayEps, CUDART_INF);
      wf.extSlot[nxt][ei] = slot;
    }
    nClosest += (cont && isNew) ? 1 : 0;  // re-walked prefix rays are an artefact of the chain evaluation, not reference rays
  }
  for (int o = 16; o > 0; o >>= 1) nClosest += __shfl_down_sync(FULL, nClosest, o);
  if ((threadIdx.x & 31) == 0 && nClosest) atomicAdd(&rc->closestRays, nClosest);
}

template <bool GENERAL>
__global__ void __launch_bounds__(128) directSampleKernel(RenderParams rp, RenderScene rs, Wavefront wf, int light, int j, int cur,
                                                          RenderCounters* rc, int sortedOrder) {
  const uint32_t n = wf.counts[cur];
  unsigned long long nShadow = 0, nClosest = 0;
  const bool sorted = GENERAL && sortedOrder != 0;
  for (uint32_t q0 = blockIdx.x * blockDim.x; q0 < n; q0 += gridDim.x * blockDim.x) {
    uint32_t q = q0 + threadIdx.x;
    bool valid = q < n;
    uint32_t slot = 0;
    int prim = -1;
    if (valid) {
      if (sorted) q = wf.shadeOrder[q];
      slot = wf.extSlot[cur][q];
      prim = __float_as_int(wf.extHit[q].w);
      valid = prim >= 0;
    }
    DirectWork dw;
    dw.hasShadow = dw.hasMis = false;
    V3 p = V3{0.f, 0.f, 0.f};
    double rayEps = 0.0;
    int lightNum = light;
    if (valid) {
      const float4 o4 = wf.extO[cur][q], d4 = wf.extD[cur][q];
      const V3 o = V3{o4.x, o4.y, o4.z}, d = V3{d4.x, d4.y, d4.z};
      ShapeHit h;
      hitGeometryQ(rs, wf, q, slot, (uint32_t)prim, o, d, wf.extT[q], &h);
      typename BsdfOf<GENERAL>::type bsdf = makeBsdfT<GENERAL>(rs, (uint32_t)prim, h, o, d);
      if (GENERAL) applyHitBsdf(rs, wf, slot, &bsdf);
      p = h.p;
      rayEps = h.rayEps;
      DirectOffsets off;
      if (light < 0) {  // strategy one: light number from the sample record (integrator.dart:92-97)
        off = rp.direct[0];
        lightNum = min((int)floor((double)val(wf, rp.dlLightNum, slot) * (double)rs.nLights), rs.nLights - 1);
      } else {
        off = rp.direct[light];
      }
      const float lu0 = val(wf, off.lightPos + 2 * j, slot), lu1 = val(wf, off.lightPos + 2 * j + 1, slot);
      const double lcomp = val(wf, off.lightComp + j, slot);
      const float bu0 = val(wf, off.bsdfPos + 2 * j, slot), bu1 = val(wf, off.bsdfPos + 2 * j + 1, slot);
      const double bcomp = GENERAL ? (double)val(wf, off.bsdfComp + j, slot) : 0.0;
      estimateDirectSetup(rs, lightNum, p, bsdf.nn, -d, rayEps, bsdf, lu0, lu1, lcomp, bu0, bu1, bcomp, BSDF_ALL & ~BSDF_SPECULAR, &dw);
    }
    pushDirectWork(wf, slot, valid, dw, p, rayEps, lightNum);
    nShadow += (valid && dw.hasShadow) ? 1 : 0;
    nClosest += (valid && dw.hasMis) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    nShadow += __shfl_down_sync(FULL, nShadow, o);
    nClosest += __shfl_down_sync(FULL, nClosest, o);
  }
  if ((threadIdx.x & 31) == 0 && (nShadow | nClosest)) {
    atomicAdd(&rc->shadowRays, nShadow);
    atomicAdd(&rc->closestRays, nClosest);
  }
}

// ---------------------------------------------------------------------------------------------------
// SamplerRenderer's radiance checks (sampler_renderer.dart:181-193) + ImageFilm.addSample
// (image_film.dart:99-150).  Accumulators are float64 and updated atomically, so the sums do not
// depend on the order samples arrive in (the reference adds float32 in pixel order).
// Adaptive sampler: what SamplerRenderer's loop hands to reportResults (sampler_renderer.dart:173-207) — the radiance values
// after the NaN / negative / infinite clean-up — and the camera rays' primitives.
__global__ void saveCameraPrimsKernel(Wavefront wf, uint32_t nSlots) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nSlots) wf.camPrim[wf.extSlot[0][q]] = __float_as_int(wf.extHit[q].w);
}
__global__ void __launch_bounds__(128) adaptiveDecideKernel(RenderParams rp, Wavefront wf, PixelBatch pb, uint32_t* list, uint32_t* listCount) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pb.nPixels) return;
  const uint32_t n = (uint32_t)rp.nPixelSamples, s0 = p * n;
  bool needs = false;
  if (rp.adaptiveMethod == 0) {  // adaptive_sampler.dart:163-171; an escaped ray compares as -1 (see the oracle)
    for (uint32_t i = 0; i + 1 < n; ++i) needs = needs || wf.camPrim[s0 + i] != wf.camPrim[s0 + i + 1];
  } else {  // :172-186
    double Lavg = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
      const Spec L = ld3(wf.L, wf.cap, s0 + i);
      const double lum = Luminance(L);
      const bool zeroed = isnan(L.r) || isnan(L.g) || isnan(L.b) || lum < -1e-5 || isinf(lum);
      Lavg += zeroed ? 0.0 : lum;
    }
    Lavg /= n;
    for (uint32_t i = 0; i < n; ++i) {
      const Spec L = ld3(wf.L, wf.cap, s0 + i);
      double lum = Luminance(L);
      if (isnan(L.r) || isnan(L.g) || isnan(L.b) || lum < -1e-5 || isinf(lum)) lum = 0.0;
      needs = needs || fabs(lum - Lavg) / Lavg > 0.5;
    }
  }
  wf.adaptFlag[p] = needs ? 1 : 0;
  if (needs) {
    uint64_t g = pb.firstPixel + p;
    if (pb.nShards > 1) g = ((g / pb.blockPixels) * pb.nShards + pb.shard) * pb.blockPixels + (g % pb.blockPixels);
    list[atomicAdd(listCount, 1u)] = (uint32_t)g;
  }
}

// ---------------------------------------------------------------------------------------------------
// Participating media: VolumeIntegrator.Li along the camera rays (sampler_renderer.dart:93-97).
#if DRT_EXTRA
// Light.sampleLAtPoint as the single-scattering integrator calls it (single_scatter_integrator.dart:105-109): Li, wi, pdf and the
// VisibilityTester's ray.  Same light code as estimateDirectSetup (shade_device.cuh), without a BSDF.
static __device__ __noinline__ void volLightSampleCold(const RenderScene& rs, int lightIndex, V3 p, float lu0, float lu1, double lcomp, Spec* LiOut,
                                                       V3* wiOut, double* pdfOut, V3* shD, double* shMax) {
  const GLight l = rs.lights[lightIndex];
  V3 wi, segTo = p;
  double lightPdf = 1.0, eps2 = 0.0;
  Spec Li = mks1(0.0);
  const bool infinite = l.kind == 4;
  const bool distant = l.kind == 2 || infinite;
  if (infinite) {
    infiniteSampleCold(rs, l, lu0, lu1, &wi, &lightPdf, &Li);
  } else if (distant) {
    wi = V3{l.pos[0], l.pos[1], l.pos[2]};
    Li = lightRadiance(l);
  } else if (l.kind != 0) {
    const V3 pos = V3{l.pos[0], l.pos[1], l.pos[2]};
    wi = Normalize(pos - p);
    segTo = pos;
    if (l.kind == 1) {
      Li = lightRadiance(l) / DistanceSquared(pos, p);
    } else if (l.kind >= 5) {
      mappedPointLightCold(rs, l, -wi, DistanceSquared(pos, p), &Li);
    } else {  // spot_light.dart:36-70
      const V3 wn = -wi;
      const V3 wl = Normalize(mkv((double)l.w2l[0] * wn.x + (double)l.w2l[1] * wn.y + (double)l.w2l[2] * wn.z,
                                  (double)l.w2l[3] * wn.x + (double)l.w2l[4] * wn.y + (double)l.w2l[5] * wn.z,
                                  (double)l.w2l[6] * wn.x + (double)l.w2l[7] * wn.y + (double)l.w2l[8] * wn.z));
      const double costheta = wl.z;
      double falloff;
      if (costheta < l.cosTotalWidth) falloff = 0.0;
      else if (costheta > l.cosFalloffStart) falloff = 1.0;
      else {
        const double dl = (costheta - l.cosTotalWidth) / (l.cosFalloffStart - l.cosTotalWidth);
        falloff = dl * dl * dl * dl;
      }
      Li = lightRadiance(l) * falloff / DistanceSquared(pos, p);
    }
  } else {  // diffuse_area_light.dart:59-70
    V3 ns;
    const V3 ps = shapeSetSample(rs, l, p, lu0, lu1, lcomp, &ns);
    wi = Normalize(ps - p);
    lightPdf = shapeSetPdf(rs, l, p, wi);
    segTo = ps;
    eps2 = 1.0e-3;
    Li = areaL(l, ns, -wi);
  }
  *LiOut = Li;
  *wiOut = wi;
  *pdfOut = lightPdf;
  if (distant) {
    *shD = wi;
    *shMax = CUDART_INF;
  } else {
    const double dist = Distance(p, segTo);  // visibility_tester.dart:26-29 with eps1 = 0
    *shD = (segTo - p) / dist;
    *shMax = dist * (1.0 - eps2);
  }
}

// Shuffle (montecarlo.dart:294-303) of `count` entries of `dims` floats each, entry i at a[(dims * i + j) * stride]
static __device__ inline void volShuffle(float* a, size_t stride, int count, int dims, Stream& rng) {
  for (int i = 0; i < count; ++i) {
    const int other = i + (int)(rng.randomUint() % (uint32_t)(count - i));
    for (int j = 0; j < dims; ++j) {
      const float t = a[(size_t)(dims * i + j) * stride];
      a[(size_t)(dims * i + j) * stride] = a[(size_t)(dims * other + j) * stride];
      a[(size_t)(dims * other + j) * stride] = t;
    }
  }
}

// EmissionIntegrator.Li (emission_integrator.dart:31-83) / SingleScatteringIntegrator.Li (single_scatter_integrator.dart:52-133): one
// thread marches one camera ray.  The single-scattering shadow rays are traced in place (anyHitWalk: the reference's walk for one
// ray); their sample arrays (LDShuffleScrambled1D / 2D over the march's steps) live in a per-slot scratch column.
__global__ void __launch_bounds__(128) volumeLiKernel(RenderParams rp, RenderScene rs, Wavefront wf, RenderCounters* rc) {
  const uint32_t n = wf.counts[Q_EXT0], cap = wf.cap;
  unsigned long long nShadow = 0;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const uint32_t slot = wf.extSlot[0][q];
    const float4 o4 = wf.extO[0][q], d4 = wf.extD[0][q];
    const bool hit = __float_as_int(wf.extHit[q].w) >= 0;
    VRay ray{V3{o4.x, o4.y, o4.z}, V3{d4.x, d4.y, d4.z}, wf.extRange[0][q].x, hit ? wf.extT[q] : CUDART_INF};
    double t0 = 0.0, t1 = 0.0;
    Spec Tr = mks1(1.0), Lv = mks1(0.0);
    if (volIntersectP(rs, ray, &t0, &t1) && (t1 - t0) != 0.0) {
      Stream rng{streamKey(rp.seed, wf.pixX[slot], wf.pixY[slot], wf.sampleIdx[slot], DRT_STREAM_VOLUME_LI), 0};
      const double stepSize = rs.volStep;
      const bool single = rs.volIntegrator == 1;
      int nSamples = (int)ceil((t1 - t0) / stepSize);
      const double step = (t1 - t0) / nSamples;
      V3 p = VRayAt(ray, t0), pPrev;
      const V3 w = -ray.d;
      t0 += (double)val(wf, rp.pScatterSample, slot) * step;
      float* aNum = nullptr; float* aComp = nullptr; float* aPos = nullptr;
      bool overflow = false;
      if (single) {
        if (nSamples > (int)wf.volMaxSteps) { overflow = true; nSamples = 0; }  // the host sized the scratch from the regions' bound
        aNum = wf.volScratch + slot;
        aComp = aNum + (size_t)wf.volMaxSteps * cap;
        aPos = aComp + (size_t)wf.volMaxSteps * cap;
        {  // LDShuffleScrambled1D(1, nSamples, lightNum, rng): montecarlo.dart:524-536
          const uint32_t sc = rng.randomUint();
          for (int i = 0; i < nSamples; ++i) aNum[(size_t)i * cap] = (float)VanDerCorput((uint32_t)i, sc);
          for (int i = 0; i < nSamples; ++i) rng.randomUint();  // Shuffle of one element per block: a draw, no move
          volShuffle(aNum, cap, nSamples, 1, rng);
        }
        {
          const uint32_t sc = rng.randomUint();
          for (int i = 0; i < nSamples; ++i) aComp[(size_t)i * cap] = (float)VanDerCorput((uint32_t)i, sc);
          for (int i = 0; i < nSamples; ++i) rng.randomUint();
          volShuffle(aComp, cap, nSamples, 1, rng);
        }
        {  // LDShuffleScrambled2D: montecarlo.dart:539-551
          const uint32_t s0 = rng.randomUint(), s1 = rng.randomUint();
          for (int i = 0; i < nSamples; ++i) {
            aPos[(size_t)(2 * i) * cap] = (float)VanDerCorput((uint32_t)i, s0);
            aPos[(size_t)(2 * i + 1) * cap] = (float)Sobol2((uint32_t)i, s1);
          }
          for (int i = 0; i < nSamples; ++i) rng.randomUint();
          volShuffle(aPos, cap, nSamples, 2, rng);
        }
      }
      for (int i = 0; i < nSamples; ++i, t0 += step) {
        pPrev = p;
        p = VRayAt(ray, t0);
        VRay tauRay{pPrev, p - pPrev, 0.0, 1.0};
        const Spec stepTau = volTau(rs, tauRay, 0.5 * stepSize, rng.randomFloat());
        Tr = Tr * expNeg(stepTau);
        if (Luminance(Tr) < 1.0e-3) {  // possibly terminate the march
          const double continueProb = 0.5;
          if (rng.randomFloat() > continueProb) {
            Tr = mks1(0.0);
            break;
          }
          Tr = Tr / continueProb;
        }
        Lv = Lv + Tr * volCoeff(rs, p, VOL_LVE);
        if (single) {
          const Spec ss = volCoeff(rs, p, VOL_SIG_S);
          if (!IsBlack(ss) && rs.nLights > 0) {
            const int ln = min((int)floor((double)aNum[(size_t)i * cap] * rs.nLights), rs.nLights - 1);
            Spec L;
            V3 wo, shD;
            double pdf = 0.0, shMax = 0.0;
            volLightSampleCold(rs, ln, p, aPos[(size_t)(2 * i) * cap], aPos[(size_t)(2 * i + 1) * cap], (double)aComp[(size_t)i * cap], &L, &wo,
                               &pdf, &shD, &shMax);
            if (!IsBlack(L) && pdf > 0.0) {
              ++nShadow;
              if (!anyHitWalk(rs.ts, make_float4(p.x, p.y, p.z, 0.f), make_float4(shD.x, shD.y, shD.z, 0.f), 0.0, shMax,
                              wf.slotTime ? wf.slotTime[slot] : 0.0)) {
                // vis.transmittance(scene, renderer, null, rng): step 4 x stepSize, the offset drawn from this march's own stream
                VRay sray{p, shD, 0.0, shMax};
                const Spec Ld = L * expNeg(volTau(rs, sray, 4.0 * stepSize, rng.randomFloat()));
                Lv = Lv + Tr * ss * volPhase(rs, p, w, -wo) * Ld * (double)rs.nLights / pdf;
              }
            }
          }
        }
      }
      if (overflow) { Tr = mks1(CUDART_NAN); }  // cannot happen with a host bound that holds; visible (zeroed sample counter) if it does
      Lv = Lv * step;
    }
    st3(wf.volT, cap, slot, Tr);
    st3(wf.volL, cap, slot, Lv);
  }
  for (int o = 16; o > 0; o >>= 1) nShadow += __shfl_down_sync(FULL, nShadow, o);
  if ((threadIdx.x & 31) == 0 && nShadow) atomicAdd(&rc->shadowRays, nShadow);
}

__global__ void __launch_bounds__(256) volumeCombineKernel(Wavefront wf, uint32_t nSlots) {  // T * Li + Lvi
  const uint32_t cap = wf.cap;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nSlots; slot += gridDim.x * blockDim.x)
    st3(wf.L, cap, slot, ld3(wf.volT, cap, slot) * ld3(wf.L, cap, slot) + ld3(wf.volL, cap, slot));
}
#endif  // DRT_EXTRA

// WARPSUM: when all 32 samples of a warp fall on ONE pixel (the usual case: a pixel's samples are consecutive slots and a box filter
// of half a pixel covers one pixel), the warp adds its terms with a butterfly and lane 0 issues the four atomics instead of 128 on the
// same four addresses.  The terms are float32 values times the filter weight, summed in f64 — exact for unit weights — so the film
// does not depend on the grouping.
template <bool WARPSUM>
__global__ void __launch_bounds__(256) filmKernel(RenderParams rp, Wavefront wf, uint32_t nSlots, int skipFlagged, RenderCounters* rc) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = s < nSlots;
  if (live && skipFlagged && wf.adaptFlag[s / (uint32_t)rp.nPixelSamples]) live = false;  // adaptive: reportResults returned false (:148-151)
  if (live && (rp.samplerKind == 3 || rp.samplerKind == 5) && isnan(wf.camXY[s].x)) live = false;  // halton / bestcandidate: a rejected index
  int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
  float X = 0.f, Y = 0.f, Z = 0.f;
  double dimageX = 0.0, dimageY = 0.0;
  if (live) {
    Spec L = ld3(wf.L, wf.cap, s);
    const double lum = Luminance(L);
    if (isnan(L.r) || isnan(L.g) || isnan(L.b) || lum < -1e-5 || isinf(lum)) {
      L = Spec{0.f, 0.f, 0.f};
      atomicAdd(&rc->zeroedSamples, 1ull);
    }
    const double2 im = wf.camXY[s];
    dimageX = im.x - 0.5; dimageY = im.y - 0.5;
    x0 = (int)ceil(dimageX - rp.xWidth); x1 = (int)floor(dimageX + rp.xWidth);
    y0 = (int)ceil(dimageY - rp.yWidth); y1 = (int)floor(dimageY + rp.yWidth);
    x0 = max(x0, rp.left); x1 = min(x1, rp.left + rp.width - 1);
    y0 = max(y0, rp.top); y1 = min(y1, rp.top + rp.height - 1);
    if ((x1 - x0) < 0 || (y1 - y0) < 0) live = false;
    const double r = L.r, g = L.g, b = L.b;  // RGBColor.toXYZ -> XYZColor (float32), spectrum.dart:293-297
    X = (float)(0.412453 * r + 0.357580 * g + 0.180423 * b); Y = (float)(0.212671 * r + 0.715160 * g + 0.072169 * b);
    Z = (float)(0.019334 * r + 0.119193 * g + 0.950227 * b);
  }
  if (WARPSUM) {
    // one pixel for the whole warp?  (every lane live, a one-pixel footprint, the same pixel as lane 0)
    const int px0 = __shfl_sync(FULL, x0, 0), py0 = __shfl_sync(FULL, y0, 0);
    const bool same = live && x0 == x1 && y0 == y1 && x0 == px0 && y0 == py0;
    if (__all_sync(FULL, same)) {
      const double fy = fabs((y0 - dimageY) * rp.invYWidth * 16), fx = fabs((x0 - dimageX) * rp.invXWidth * 16);
      const double wt = rp.filterTable[min((int)floor(fy), 15) * 16 + min((int)floor(fx), 15)];
      double a0 = wt * X, a1 = wt * Y, a2 = wt * Z, a3 = wt;
      for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(FULL, a0, o);
        a1 += __shfl_xor_sync(FULL, a1, o);
        a2 += __shfl_xor_sync(FULL, a2, o);
        a3 += __shfl_xor_sync(FULL, a3, o);
      }
      if ((threadIdx.x & 31) == 0) {
        double* px = rp.film + 4 * ((size_t)(y0 - rp.top) * rp.width + (x0 - rp.left));
        atomicAdd(px + 0, a0);
        atomicAdd(px + 1, a1);
        atomicAdd(px + 2, a2);
        atomicAdd(px + 3, a3);
      }
      return;
    }
  }
  if (!live) return;
  for (int y = y0; y <= y1; ++y) {
    const double fy = fabs((y - dimageY) * rp.invYWidth * 16);
    const int iy = min((int)floor(fy), 15);
    for (int x = x0; x <= x1; ++x) {
      const double fx = fabs((x - dimageX) * rp.invXWidth * 16);
      const int ix = min((int)floor(fx), 15);
      const double wt = rp.filterTable[iy * 16 + ix];
      double* px = rp.film + 4 * ((size_t)(y - rp.top) * rp.width + (x - rp.left));
      atomicAdd(px + 0, wt * X);
      atomicAdd(px + 1, wt * Y);
      atomicAdd(px + 2, wt * Z);
      atomicAdd(px + 3, wt);
    }
  }
}

// ImageFilm.writeImage (image_film.dart:268-299): XYZ -> RGB, divide by the weight sum, clamp at 0
__global__ void filmConvertKernel(RenderParams rp, float* rgb, float* xyz, float* weight) {
  const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= (size_t)rp.width * rp.height) return;
  const float X = (float)rp.film[4 * pi], Y = (float)rp.film[4 * pi + 1], Z = (float)rp.film[4 * pi + 2], W = (float)rp.film[4 * pi + 3];
  if (xyz) { xyz[3 * pi] = X; xyz[3 * pi + 1] = Y; xyz[3 * pi + 2] = Z; }
  if (weight) weight[pi] = W;
  if (rgb) {
    const double x = X, y = Y, z = Z, w = W;
    const double c0 = 3.240479 * x - 1.537150 * y - 0.498535 * z;
    const double c1 = -0.969256 * x + 1.875991 * y + 0.041556 * z;
    const double c2 = 0.055648 * x - 0.204043 * y + 1.057311 * z;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    if (w != 0.0) {
      const double invWt = 1.0 / w;
      o0 = (float)fmax(0.0, c0 * invWt); o1 = (float)fmax(0.0, c1 * invWt); o2 = (float)fmax(0.0, c2 * invWt);
    }
    rgb[3 * pi] = o0; rgb[3 * pi + 1] = o1; rgb[3 * pi + 2] = o2;
  }
}

#endif  // DRT_PATH_ONLY
// ---------------------------------------------------------------------------------------------------
// Grid of a grid-stride kernel: one resident wave — as many CTAs per SM as the occupancy API says fit (cached per kernel).  A fixed
// "8 per SM" was a partial second wave for kernels of which 5 - 7 fit and half the machine's warps for kernels of which 16 fit.
template <class K>
static inline int residentGrid(K kernel, uint64_t n, int block, int numSMs) {
  static int perSm = 0;  // one instance per kernel type AND per call site's function pointer type; keyed below by the pointer itself
  static const void* cachedFor = nullptr;
  if (cachedFor != (const void*)kernel) {
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, block, 0) != cudaSuccess || b < 1) b = 8;
    perSm = b;
    cachedFor = (const void*)kernel;
  }
  uint64_t want = (n + block - 1) / block;
  uint64_t cap = (uint64_t)numSMs * perSm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

static inline int gridFor(uint64_t n, int block, int numSMs, int perSm) {
  uint64_t want = (n + block - 1) / block;
  uint64_t cap = (uint64_t)numSMs * perSm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

#ifndef DRT_PATH_ONLY  // the float32 path build (render_kernels_f32.cu) takes the path-vertex and resolve kernels only
cudaError_t launchSampler(const RenderParams& rp, const Wavefront& wf, const SampleArray* dArrays, int nArrays, int maxVals,
                          int maxOthers, const PixelBatch& pb, int numSMs, cudaStream_t st) {
  if (pb.nPixels == 0) return cudaSuccess;
  const bool ld = rp.samplerKind == 0 || rp.samplerKind == 4 || rp.samplerKind == 5;  // adaptive draws LDPixelSample too; bestcandidate
                                                                                      // its integrator arrays (one pixel sample)
  if (ld && rp.ldAllSingle && rp.nPixelSamples <= 2048) {
    const int block = 128;
    const bool bytes = rp.nPixelSamples <= 256;  // permutation entries of 8 bits: half the shared memory per task
    // two arrays of nPixelSamples entries of 16 (8) bits; odd word stride
    const int strideWords = (bytes ? (rp.nPixelSamples + 1) / 2 : rp.nPixelSamples) | 1;
    const size_t taskBytes = (size_t)strideWords * sizeof(float);
    // shared memory per block for the tasks' permutation arrays: the kernel's own tables (Sobol, remainders) take 4 KB of the 48 KB
    // a block gets without opting in; env DRT_SAMPLER_SMEM_KB asks for more (fewer lanes per task, fewer blocks per SM) for A/B runs
    static const size_t budgetKb = std::getenv("DRT_SAMPLER_SMEM_KB") ? (size_t)std::atoi(std::getenv("DRT_SAMPLER_SMEM_KB")) : 48;
    const size_t budget = std::min<size_t>(std::max<size_t>(budgetKb, 16), 200) * 1024 - 4 * 1024;
    int tasksPerBlock = (int)std::min<size_t>(block, std::max<size_t>(4, budget / taskBytes));
    int G = 1;
    while (block / G > tasksPerBlock) G <<= 1;
    const size_t smem = (size_t)(block / G) * taskBytes;
    uint64_t tasks = (uint64_t)pb.nPixels * nArrays;
    int grid = gridFor(tasks * G, block, numSMs, 16);
    if (smem + 4 * 1024 > 48 * 1024) {
      cudaError_t e = bytes ? cudaFuncSetAttribute(samplerLDPermKernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(samplerLDPermKernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    if (bytes) samplerLDPermKernel<uint8_t><<<grid, block, smem, st>>>(rp, wf, dArrays, nArrays, pb, G, strideWords);
    else samplerLDPermKernel<uint16_t><<<grid, block, smem, st>>>(rp, wf, dArrays, nArrays, pb, G, strideWords);
  } else if (ld) {
    const int block = 128;
    // lanes per (pixel, array) task: as few as shared memory allows (48 KB of arrays per block), because the
    // order-dependent shuffle runs on one lane per task
    const size_t taskBytes = (size_t)(maxVals | 1) * sizeof(float);
    int tasksPerBlock = (int)std::min<size_t>(block, std::max<size_t>(4, (48 * 1024) / taskBytes));
    int G = 1;
    while (block / G > tasksPerBlock) G <<= 1;
    const size_t smem = (size_t)(block / G) * taskBytes;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(samplerLDKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    uint64_t tasks = (uint64_t)pb.nPixels * nArrays;
    int grid = gridFor(tasks * G, block, numSMs, 16);
    samplerLDKernel<<<grid, block, smem, st>>>(rp, wf, dArrays, nArrays, maxVals, pb, G);
  } else if (rp.samplerKind == 3) {
    samplerHaltonKernel<<<(pb.nPixels + 127) / 128, 128, 0, st>>>(rp, wf, dArrays, nArrays, pb);
  } else {
    samplerSeqKernel<<<(pb.nPixels + 127) / 128, 128, 0, st>>>(rp, wf, dArrays, nArrays, pb);
  }
  // bestcandidate: the camera sample comes from the pattern, after the lowdiscrepancy kernel wrote the integrator arrays
  if (rp.samplerKind == 5) samplerBestCandidateKernel<<<(pb.nPixels + 127) / 128, 128, 0, st>>>(rp, wf, pb);
  return cudaGetLastError();
}

cudaError_t launchRaygen(const RenderParams& rp, const Wavefront& wf, const PixelBatch& pb, RenderCounters* rc, cudaStream_t st) {
  const uint32_t nSlots = pb.nPixels * (uint32_t)rp.nPixelSamples;
  if (nSlots == 0) return cudaSuccess;
  raygenKernel<<<(nSlots + 255) / 256, 256, 0, st>>>(rp, wf, pb, nSlots, rc);
  return cudaGetLastError();
}

cudaError_t launchResetCounts(const Wavefront& wf, unsigned mask, cudaStream_t st) {
  resetCountsKernel<<<1, 32, 0, st>>>(wf, mask);
  return cudaGetLastError();
}

#endif  // DRT_PATH_ONLY
// Counting sort of extension queue `cur` by material into wf.shadeOrder; *sorted = 0 when the scene has one material or more than
// the sort's bins (the shading kernels then walk the queue as it is).  DRT_NO_MATERIAL_SORT turns it off.
static cudaError_t launchQueueSort(const RenderScene& rs, const Wavefront& wf, int cur, int keyMode, int numSMs, cudaStream_t st) {
  const size_t bins = keyMode ? (size_t)DRT_SHAPE_CLASSES + 1 : (size_t)rs.nMaterials + 1;
  cudaError_t e = cudaMemsetAsync(wf.matHist, 0, bins * sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  const int g2 = gridFor(wf.cap, 256, numSMs, 4);
  static const bool plainAtomics = std::getenv("DRT_SORT_PLAIN_ATOMICS") != nullptr;
  if (plainAtomics) matHistKernel<false><<<g2, 256, 0, st>>>(rs, wf, cur, keyMode);
  else matHistKernel<true><<<g2, 256, 0, st>>>(rs, wf, cur, keyMode);
  matScanKernel<<<1, 1, 0, st>>>(rs, wf, keyMode);
  if (plainAtomics) matScatterKernel<false><<<g2, 256, 0, st>>>(rs, wf, cur, keyMode);
  else matScatterKernel<true><<<g2, 256, 0, st>>>(rs, wf, cur, keyMode);
  return cudaGetLastError();
}
cudaError_t launchMaterialSort(const RenderScene& rs, const Wavefront& wf, int cur, int numSMs, int* sorted, cudaStream_t st) {
  static const bool sortOff = std::getenv("DRT_NO_MATERIAL_SORT") != nullptr;
  *sorted = (!sortOff && rs.general && rs.nMaterials > 1 && rs.nMaterials <= DRT_SORT_MAX_MATERIALS) ? 1 : 0;
  if (!*sorted) return cudaSuccess;
  return launchQueueSort(rs, wf, cur, 0, numSMs, st);
}

cudaError_t launchShadePath(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int bounce, int cur,
                            RenderCounters* rc, int numSMs, cudaStream_t st) {
  // one resident wave of the grid-stride kernel: a grid that is not a multiple of what fits (e.g. 8 CTAs per SM asked for, 7 resident)
  // ends in a partial second wave
  static int perSm[2] = {0, 0};
  if (!perSm[0]) {
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, shadePathKernel<false, DRT_EXTRA != 0, 0>, 128, 0) != cudaSuccess || b < 1) b = 8;
    perSm[0] = b;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, shadePathKernel<true, DRT_EXTRA != 0, 0>, 128, 0) != cudaSuccess || b < 1) b = 8;
    perSm[1] = b;
  }
  const int grid = DRT_SHADE_NOLOOP ? (int)((wf.cap + 127u) / 128u) : gridFor(wf.cap, 128, numSMs, perSm[rs.general ? 1 : 0]);
  if (rs.general) {
    // material-coherent warps: sort the queue by material first (skipped for a single material or more than the sort's bins)
    // (directSampleKernel walks the same order since the sort's atomics are warp-aggregated: 64 -> 46.5 ms on cornell_materials at
    // 16 spp; with one atomic per entry the sort cost more than it saved there, 64 -> 74 ms)
    int sorted = 0;
    cudaError_t e = launchMaterialSort(rs, wf, cur, numSMs, &sorted, st);
    if (e != cudaSuccess) return e;
    shadePathKernel<true, DRT_EXTRA != 0, 0><<<grid, 128, 0, st>>>(rp, rs, wf, bounce, cur, rc, sorted);
  } else {
    // matte-only scenes: DRT_SHAPE_SORT=1 walks the queue in shape-class order, misses last (full, type-coherent warps).  Measured on
    // B200 (profiles/r02v_shape_sort_ab.log, r02x_path_*.csv): the kernel itself gains 11 % (22 % on the bounces that have misses:
    // 25.9 instead of 18.9 threads per instruction), but the three sort passes (3.0 ms per 12 launches) and the less coalesced
    // shadow / MIS queues in resolveDirectKernel (+1.4 ms) take it back: config 4 0.986 s against 0.978 s.  Off by default.
    static const bool shapeSortOn = DRT_SHAPE_SORT_BUILD && std::getenv("DRT_SHAPE_SORT") != nullptr;
    const int sorted = shapeSortOn ? 1 : 0;
    if (sorted) {
      cudaError_t e = launchQueueSort(rs, wf, cur, 1, numSMs, st);
      if (e != cudaSuccess) return e;
    }
    // DRT_SHADE_SPLIT=1 (scenes without media): the vertex as two smaller launches, see shadePathKernel.  Measured on B200 (config 4,
    // profiles/r02y_shade_split_ab.log): 96 registers each instead of 128, and SLOWER, 1.085 s against 0.986 s — the two launches read
    // the queue and rebuild the hit geometry twice, and the extra warps do not buy that back.  Off by default.
    static const bool split = std::getenv("DRT_SHADE_SPLIT") != nullptr;
    if (split && rs.nVolumes == 0) {
      shadePathKernel<false, DRT_EXTRA != 0, 1><<<grid, 128, 0, st>>>(rp, rs, wf, bounce, cur, rc, sorted);
      shadePathKernel<false, DRT_EXTRA != 0, 2><<<grid, 128, 0, st>>>(rp, rs, wf, bounce, cur, rc, sorted);
    } else {
      shadePathKernel<false, DRT_EXTRA != 0, 0><<<grid, 128, 0, st>>>(rp, rs, wf, bounce, cur, rc, sorted);
    }
  }
  return cudaGetLastError();
}

cudaError_t launchResolveDirect(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int mode,
                                int nSamplesOfLight, int numSMs, cudaStream_t st) {
  static int perSm = 0;  // one resident wave, as many CTAs as fit (the kernel waits on memory: more warps in flight)
  if (!perSm) {
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, resolveDirectKernel<DRT_EXTRA != 0>, 128, 0) != cudaSuccess || b < 1) b = 8;
    perSm = b;
  }
  resolveDirectKernel<DRT_EXTRA != 0><<<gridFor(wf.cap, 128, numSMs, perSm), 128, 0, st>>>(rp, rs, wf, cur, mode, nSamplesOfLight);
  return cudaGetLastError();
}

#ifndef DRT_PATH_ONLY  // the float32 path build (render_kernels_f32.cu) takes the path-vertex and resolve kernels only
cudaError_t launchEscape(const RenderScene& rs, const Wavefront& wf, int cur, int mode, int numSMs, cudaStream_t st) {
  escapeKernel<<<residentGrid(escapeKernel, wf.cap, 128, numSMs), 128, 0, st>>>(rs, wf, cur, mode);
  return cudaGetLastError();
}

cudaError_t launchAoSetup(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int numSMs, cudaStream_t st) {
  aoSetupKernel<<<residentGrid(aoSetupKernel, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf);
  return cudaGetLastError();
}

static inline int roundUpPow2(int v) {  // common.dart:117-125
  v--;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v + 1;
}

cudaError_t launchAoGen(const RenderParams& rp, const Wavefront& wf, uint32_t firstHit, uint32_t maxHits, int numSMs,
                        cudaStream_t st) {
  const int nS = roundUpPow2(rp.aoSamples);
  static const int plainEnv = std::getenv("DRT_AO_PLAIN_ORDER") != nullptr ? 1 : 0;
  const int plain = (plainEnv || (31 - __builtin_clz((unsigned)nS)) / 2 > DRT_AO_MAX_B) ? 1 : 0;
  aoGenKernel<<<gridFor((uint64_t)maxHits * nS, 256, numSMs, 8), 256, 0, st>>>(rp, wf, firstHit, maxHits, nS, plain);
  return cudaGetLastError();
}

cudaError_t launchAoCount(const RenderParams& rp, const Wavefront& wf, uint32_t firstHit, uint32_t maxHits, RenderCounters* rc,
                          int numSMs, cudaStream_t st) {
  const int nS = roundUpPow2(rp.aoSamples);
  static const int plainEnv = std::getenv("DRT_AO_PLAIN_ORDER") != nullptr ? 1 : 0;
  const int plain = (plainEnv || (31 - __builtin_clz((unsigned)nS)) / 2 > DRT_AO_MAX_B) ? 1 : 0;
  aoCountKernel<<<gridFor(maxHits, 256, numSMs, 8), 256, 0, st>>>(rp, wf, firstHit, maxHits, nS, rc, plain);
  return cudaGetLastError();
}

cudaError_t launchDirectSetup(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int weighted, int numSMs,
                              cudaStream_t st) {
  directSetupKernel<<<residentGrid(directSetupKernel, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf, cur, weighted);
  return cudaGetLastError();
}

cudaError_t launchDirectSample(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int light, int j, int cur,
                               RenderCounters* rc, int sorted, int numSMs, cudaStream_t st) {
  if (rs.general) directSampleKernel<true><<<residentGrid(directSampleKernel<true>, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf, light, j, cur, rc, sorted);
  else directSampleKernel<false><<<residentGrid(directSampleKernel<false>, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf, light, j, cur, rc, 0);
  return cudaGetLastError();
}

cudaError_t launchWhittedSetup(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int weighted, int numSMs,
                               cudaStream_t st) {
  whittedSetupKernel<<<residentGrid(whittedSetupKernel, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf, cur, weighted);
  return cudaGetLastError();
}

cudaError_t launchWhittedSample(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int light, int cur,
                                RenderCounters* rc, int sorted, int numSMs, cudaStream_t st) {
  if (rs.general) whittedSampleKernel<true><<<residentGrid(whittedSampleKernel<true>, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf, light, cur, rc, sorted);
  else whittedSampleKernel<false><<<residentGrid(whittedSampleKernel<false>, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf, light, cur, rc, 0);
  return cudaGetLastError();
}

cudaError_t launchSpecularStep(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int flags, int level,
                               int isNew, RenderCounters* rc, int numSMs, cudaStream_t st) {
  specularStepKernel<<<residentGrid(specularStepKernel, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf, cur, flags, level, isNew, rc);
  return cudaGetLastError();
}

cudaError_t launchSaveCameraPrims(const Wavefront& wf, uint32_t nSlots, cudaStream_t st) {
  if (nSlots == 0) return cudaSuccess;
  saveCameraPrimsKernel<<<(nSlots + 255) / 256, 256, 0, st>>>(wf, nSlots);
  return cudaGetLastError();
}

cudaError_t launchAdaptiveDecide(const RenderParams& rp, const Wavefront& wf, const PixelBatch& pb, uint32_t* list, uint32_t* listCount,
                                 cudaStream_t st) {
  if (pb.nPixels == 0) return cudaSuccess;
  adaptiveDecideKernel<<<(pb.nPixels + 127) / 128, 128, 0, st>>>(rp, wf, pb, list, listCount);
  return cudaGetLastError();
}

cudaError_t launchFilm(const RenderParams& rp, const Wavefront& wf, uint32_t nSlots, int skipFlagged, RenderCounters* rc, cudaStream_t st) {
  if (nSlots == 0) return cudaSuccess;
  static const bool perSample = std::getenv("DRT_FILM_PLAIN_ATOMICS") != nullptr;  // A/B knob: four atomics per (sample, pixel)
  if (perSample) filmKernel<false><<<(nSlots + 255) / 256, 256, 0, st>>>(rp, wf, nSlots, skipFlagged, rc);
  else filmKernel<true><<<(nSlots + 255) / 256, 256, 0, st>>>(rp, wf, nSlots, skipFlagged, rc);
  return cudaGetLastError();
}

cudaError_t launchVolumeLi(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, RenderCounters* rc, int numSMs, cudaStream_t st) {
#if DRT_EXTRA
  volumeLiKernel<<<residentGrid(volumeLiKernel, wf.cap, 128, numSMs), 128, 0, st>>>(rp, rs, wf, rc);
  return cudaGetLastError();
#else
  (void)rp; (void)rs; (void)wf; (void)rc; (void)numSMs; (void)st;
  return cudaErrorNotSupported;
#endif
}

cudaError_t launchVolumeCombine(const Wavefront& wf, uint32_t nSlots, cudaStream_t st) {
#if DRT_EXTRA
  if (nSlots == 0) return cudaSuccess;
  volumeCombineKernel<<<(nSlots + 255) / 256, 256, 0, st>>>(wf, nSlots);
  return cudaGetLastError();
#else
  (void)wf; (void)nSlots; (void)st;
  return cudaErrorNotSupported;
#endif
}

cudaError_t launchFilmConvert(const RenderParams& rp, float* rgb, float* xyz, float* weight, cudaStream_t st) {
  const size_t n = (size_t)rp.width * rp.height;
  filmConvertKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rp, rgb, xyz, weight);
  return cudaGetLastError();
}

#endif  // DRT_PATH_ONLY
}  // namespace DRT_RK_NS
}  // namespace drt
