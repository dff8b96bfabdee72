// The float32 path-vertex and resolve kernels (see render_kernels_f32.cu) of the `extra` build: per-vertex mesh attributes (N / S / uv)
// and the cylinder / cone / paraboloid / hyperboloid shapes (namespace drt::extraf).
#define DRT_EXTRA 1
#define DRT_RK_NS extraf
#define DRT_PATH_ONLY 1
#define DRT_REAL32 1
#ifndef DRT_SHADE_MIN_BLOCKS_F32X
#define DRT_SHADE_MIN_BLOCKS_F32X 6
#endif
#define DRT_SHADE_MIN_BLOCKS DRT_SHADE_MIN_BLOCKS_F32X
#include "_gen/render_kernels_f32.inc"
