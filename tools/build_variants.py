"""Builds A/B variants of libdartray_gpu.so (traversal tuning macros) into dartray_b200/variants/ for one gpurun call.
usage: python tools/build_variants.py name1:DEF=1,DEF2=3 name2:...   (the defines apply to trace_fast.cu / trace_fast2.cu)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from dartray_b200 import build

out_dir = os.path.join(build.PKG, "variants")
os.makedirs(out_dir, exist_ok=True)
only = ("trace_fast.cu", "trace_fast2.cu")
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    defines = tuple(d for d in defs.split(",") if d)
    lib = os.path.join(out_dir, f"lib_{name}.so")
    build.build(defines=defines, lib=lib, only=only if "--all" not in sys.argv else ())
    print(name, defines, lib)
