// trace_fast.cu in float32 arithmetic, for the path integrator's ray queues under DRT_PRECISION_F32: the same source lowered textually
// by dartray_b200/gen_f32.py (double -> float: the "exact" slab test, the Moeller-Trumbore triangle test and the quadric tests of
// trace_device.cuh all run in float32).  launchTraceFastF32 picks like launchTraceFast: the leaf-list kernel on small scenes, the
// quantised-node kernel with a float32 leaf phase (trace_q_f32.cu) from 65,536 primitives up, the float32-box tree kernel between.
// Compiled with contraction and the fast division / square root, like the float32 shading units.  Kernels and launchers are renamed so
// that they cannot be merged with the binary64 instantiations at link time.
#define DRT_REAL32 1
#define traceFastKernel traceFastKernelF32
#define traceSmallKernel traceSmallKernelF32
#define launchSmall launchSmallF32
#define launchOne launchOneF32
#define launchTraceFast launchTraceFastF32
#define launchTraceQ launchTraceQF32
#include "_gen/trace_fast_f32.inc"
