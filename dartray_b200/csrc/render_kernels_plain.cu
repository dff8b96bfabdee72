// The render stage kernels compiled WITHOUT per-vertex mesh attributes and the cylinder / cone / paraboloid / hyperboloid
// shapes (DRT_EXTRA = 0, namespace drt::plain): what scenes without those features run — BASELINE.json configs 3 and 4
// among them.  Carrying the rare features' out-of-line code in one build cost config 4 4 % (A/B on B200).
#define DRT_EXTRA 0
#define DRT_RK_NS plain
#include "render_kernels.cu"
