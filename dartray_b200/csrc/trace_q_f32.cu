// traceQKernel (trace_fast2.cu) with a float32 leaf phase, for the path integrator's queues under DRT_PRECISION_F32 on scenes that run
// the quantised-node kernel: the same source lowered textually by dartray_b200/gen_f32.py.  The node steps are float32 already; here the
// leaf box test, the triangle test and the quadric tests run in float32 too (trace_device.cuh lowered), the ray's maxDistance is a
// float32 register.  "Slow" rays still decode and test every box in binary64 (the block kept verbatim).  Kernel and launchers renamed
// so that they cannot be merged with the binary64 instantiations at link time.
#define DRT_REAL32 1
#define traceQKernel traceQKernelF32
#define launchOneQ launchOneQF32
#define launchTraceQ launchTraceQF32
#define slabExactQ slabExactQF32
#define slowSlotQ slowSlotQF32
#include "_gen/trace_fast2_f32.inc"
