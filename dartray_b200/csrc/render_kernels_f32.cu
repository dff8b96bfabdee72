// The path-vertex and resolve kernels in float32 arithmetic (namespace drt::plainf): what DRT_PRECISION_F32 runs for the path
// integrator on scenes of the `plain` build (BASELINE.json config 4).  The sources are the SAME files as the binary64 build,
// lowered textually by dartray_b200/gen_f32.py into _gen/ before nvcc runs (double -> float, literals suffixed); this unit is
// compiled with contraction and the fast division / square root on, there being no bit-replay to protect: north_star asks the
// path tracer for per-pixel agreement within 3 sigma of its Monte Carlo variance (tests/test_precision_gpu.py).
#define DRT_EXTRA 0
#define DRT_RK_NS plainf
#define DRT_PATH_ONLY 1
#define DRT_REAL32 1
// CTAs of 128 threads per SM the path-vertex kernel is compiled for.  Measured on B200 (config 4) with the grid sized to one resident
// wave (profiles/r02z20_f32mb_ab.log): 4 (128 registers) 0.4631 s, 5 (96) 0.4572, 6 (80) 0.4521, 7 (72) 0.4570, 8 (64) 0.4639,
// 10 (48) 0.5300.  (The first sweep, r02z_f32_ab.log, ran every variant on a grid of 8 CTAs per SM: 5, 6 and 7 paid for a partial
// second wave and 8 looked best.)
#ifndef DRT_SHADE_MIN_BLOCKS_F32
#define DRT_SHADE_MIN_BLOCKS_F32 6
#endif
#define DRT_SHADE_MIN_BLOCKS DRT_SHADE_MIN_BLOCKS_F32
#include "_gen/render_kernels_f32.inc"
