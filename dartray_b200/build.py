"""In-tree build of libdartray_gpu.so for sm_100a (nvcc cross-compiles without a GPU).

Every translation unit is compiled to an object file on its own (in parallel, only when it or a header
changed) and the objects are linked into the shared library; no relocatable device code is needed, each
.cu launches only its own kernels.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

try:
    from . import gen_f32
except ImportError:  # run as a script
    import gen_f32

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_obj")
LIB = os.path.join(PKG, "libdartray_gpu.so")

COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # Dart doubles never contract a*b+c; the kernels call fma()/fmaf() explicitly where it is safe.
    "-fmad=false",
    "-Xcompiler", "-fPIC,-pthread,-ffp-contract=off",
]
# render_kernels_f32.cu (the float32 shading build, gen_f32.py): no bit replay to protect, so contraction and the fast division /
# square root / sincos are on.  DRT_F32_FLAGS overrides the extra flags for A/B builds.
F32_UNITS = ("render_kernels_f32.cu", "render_kernels_f32x.cu", "trace_fast_f32.cu", "trace_q_f32.cu")
F32_FLAGS = os.environ.get("DRT_F32_FLAGS", "-use_fast_math").split()


def _flags_for(src: str) -> list:
    flags = list(COMPILE_FLAGS) + ["-I", CSRC]
    if os.path.basename(src) in F32_UNITS:
        flags = [f for f in flags if f != "-fmad=false"] + F32_FLAGS
    return flags


LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-pthread"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    return hs + [os.path.join(PKG, "..", "include", "drt.h"), __file__]


def _obj_of(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src) + ".o")


def _stale_obj(src: str, hdr_time: float, extra_key: str) -> bool:
    o = _obj_of(src)
    if not os.path.exists(o):
        return True
    key = o + ".flags"
    if not os.path.exists(key) or open(key).read() != extra_key:
        return True
    t = os.path.getmtime(o)
    # render_kernels_plain.cu includes render_kernels.cu
    deps = [src] + ([os.path.join(CSRC, "render_kernels.cu"), os.path.join(CSRC, "trace_fast.cu"), os.path.join(CSRC, "trace_fast2.cu")]
                    if src.endswith(("render_kernels_plain.cu",) + F32_UNITS) else [])
    return any(os.path.getmtime(d) > t for d in deps) or hdr_time > t


def build(force: bool = False, verbose: bool = False, defines: tuple = (), lib: str = LIB, only: tuple = ()) -> str:
    """Builds (if stale) and returns the path of the shared library.  `defines` (e.g. ("DRT_MIN_BLOCKS=4",)) and a
    different `lib` path give A/B variants for one gpurun call; objects of variant builds live in their own directory.
    `only`: file names the defines apply to — the other objects are taken from the default build."""
    global OBJ
    if defines and only:
        build()  # the shared objects
        base_objs = {s: _obj_of(s) for s in sources() if os.path.basename(s) not in only}
    else:
        base_objs = {}
    extra_key = " ".join(sorted(defines)) + " | " + " ".join(F32_FLAGS)
    gen_f32.generate()
    obj_dir = OBJ if not defines else OBJ + "_" + "".join(ch if ch.isalnum() else "_" for ch in " ".join(sorted(defines)))
    saved, OBJ = OBJ, obj_dir
    try:
        os.makedirs(OBJ, exist_ok=True)
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        hdr_time = max(os.path.getmtime(h) for h in _headers())
        todo = [s for s in sources() if s not in base_objs and (force or _stale_obj(s, hdr_time, extra_key))]
        dflags = [f"-D{d}" for d in defines]

        def compile_one(src):
            cmd = [nvcc] + _flags_for(src) + dflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj_of(src), src]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
            with open(_obj_of(src) + ".flags", "w") as f:
                f.write(extra_key)
            return res.stderr

        if todo:
            with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
                for log in ex.map(compile_one, todo):
                    if verbose:
                        print(log)
        objs = [base_objs.get(s) or _obj_of(s) for s in sources()]
        if todo or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
            cmd = [nvcc] + LINK_FLAGS + ["-o", lib] + objs
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return lib
    finally:
        OBJ = saved


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    out = LIB
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            out = a[6:]
    only = tuple(a[7:] for a in sys.argv[1:] if a.startswith("--only="))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, lib=out, only=only))
