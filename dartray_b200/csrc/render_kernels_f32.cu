// The path-vertex and resolve kernels in float32 arithmetic (namespace drt::plainf): what DRT_PRECISION_F32 runs for the path
// integrator on scenes of the `plain` build (BASELINE.json config 4).  The sources are the SAME files as the binary64 build,
// lowered textually by dartray_b200/gen_f32.py into _gen/ before nvcc runs (double -> float, literals suffixed); this unit is
// compiled with contraction and the fast division / square root on, there being no bit-replay to protect: north_star asks the
// path tracer for per-pixel agreement within 3 sigma of its Monte Carlo variance (tests/test_precision_gpu.py).
#define DRT_EXTRA 0
#define DRT_RK_NS plainf
#define DRT_PATH_ONLY 1
#define DRT_REAL32 1
// CTAs of 128 threads per SM the path-vertex kernel is compiled for.  Measured on B200 (config 4, profiles/r02z_f32_ab.log):
// 4 (128 registers) 0.6647 s, 5 (96) 0.6899 s, 6 (80) 0.7036 s, 8 (64 registers) 0.6599 s; the binary64 kernel: 0.9709 s
#ifndef DRT_SHADE_MIN_BLOCKS_F32
#define DRT_SHADE_MIN_BLOCKS_F32 8
#endif
#define DRT_SHADE_MIN_BLOCKS DRT_SHADE_MIN_BLOCKS_F32
#include "_gen/render_kernels_f32.inc"
