"""dartray_b200/gen_f32.py: the textual lowering that produces the float32 shading build (csrc/render_kernels_f32.cu)."""
import os
import re

from dartray_b200 import gen_f32


def test_literals_types_and_intrinsics_are_lowered():
    src = "\n".join([
        "static DRT_HD inline double f(double x) { return 1.0 / x + 2.5e-3 * .5 + 1e300 + 3.f + 7 + 0x9E3779B97F4A7C15ull; }  // 1.0 stays in prose",
        "double2 r = make_double2(CUDART_INF, CUDART_NAN); float y = __double2float_rn(z);",
        "V3 a = b * s.radius + rs.lights[li].area * l.cosTotalWidth; h->area2 = radius;",
    ])
    out = gen_f32.lower(src).split("\n")
    assert out[0].startswith("static DRT_HD inline float f(float x) { return 1.0f / x + 2.5e-3f * .5f + 1e300f + 3.f + 7 + 0x9E3779B97F4A7C15ull; }")
    assert out[0].endswith("// 1.0 stays in prose")
    assert out[1] == "double2 r = make_double2(CUDART_INF_F, CUDART_NAN_F); float y = (float)(z);"
    assert out[2] == "V3 a = b * ((float)s.radius) + ((float)rs.lights[li].area) * ((float)l.cosTotalWidth); h->area2 = radius;"


def test_keep_blocks_and_includes():
    src = "\n".join(['#include "shade_device.cuh"', '#include "gpu_types.h"', "// f32-keep-begin", "double keep = 1.0;", "// f32-keep-end",
                     "double lowered = 1.0;"])
    out = gen_f32.lower(src).split("\n")
    assert out[0] == '#include "shade_device_f32.cuh"' and out[1] == '#include "gpu_types.h"'
    assert out[3] == "double keep = 1.0;" and out[5] == "float lowered = 1.0f;"


def test_generated_sources_hold_no_binary64_arithmetic_types():
    """After lowering, `double` survives only inside identifiers of the shared storage types (double2, make_double2, __double_as_...)."""
    for path in gen_f32.generate():
        text = open(path).read()
        text = re.sub(r"// f32-keep-begin.*?// f32-keep-end", "", text, flags=re.S)  # blocks copied verbatim on purpose
        code = "\n".join(line.split("//")[0] for line in text.split("\n"))
        assert not re.search(r"\bdouble\b", code), path
        assert os.path.basename(path).endswith(("_f32.cuh", "_f32.inc"))
    assert "f32-keep-begin" in open(os.path.join(gen_f32.GEN, "trace_fast2_f32.inc")).read()


def test_assignment_to_a_narrowed_member_is_refused():
    import pytest
    with pytest.raises(ValueError):
        gen_f32.lower("g.radius = 2.0;")
    assert "==" in gen_f32.lower("if (s.radius == 0.0) return;")
