// traceSmallKernel (trace_fast.cu) in float32 arithmetic, for the renderer's queues under DRT_PRECISION_F32: the same source lowered
// textually by dartray_b200/gen_f32.py (double -> float: the "exact" slab test, the Moeller-Trumbore triangle test and the quadric
// tests of trace_device.cuh all run in float32).  Compiled with contraction and the fast division / square root, like the float32
// shading units.  The kernel and launcher are renamed so that they cannot be merged with the binary64 instantiations at link time.
#define DRT_SMALL_ONLY 1
#define DRT_REAL32 1
#define traceSmallKernel traceSmallKernelF32
#define launchSmall launchSmallF32
#include "_gen/trace_fast_f32.inc"
