"""Build the REFERENCE's own CUDA op for sm_100a from the sources where they lie.

    python oracle/build_ref.py          # -> oracle/_ref/msda_reference_sm100*.so (git-ignored)

TEST / BASELINE INFRASTRUCTURE ONLY: the product never loads this module.  It is the
reference kernels (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh) recompiled for
Blackwell -- the GPU bar to beat and a second parity witness on the B200 box
(tests/test_gpu_vs_reference_cuda.py, bench.py's `ref_cuda_*` fields).

The reference's own build (models/ops/setup.py) is not run: it refuses without a visible GPU
(:47) and torch 2.11 no longer converts ``value.type()`` inside ``AT_DISPATCH_FLOATING_TYPES``
(src/cuda/ms_deform_attn_cuda.cu:64,134).  No reference source is copied or edited: the three
translation units are compiled in place and the missing overload
``detail::scalar_type(const at::DeprecatedTypeProperties&)`` is supplied by a force-included
header (oracle/ref_compat.h).  The pybind11 module is named ``msda_reference_sm100`` so that it
can never shadow this repo's ``MultiScaleDeformableAttention``.
"""
from __future__ import annotations

import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF_SRC = Path("/root/reference/models/ops/src")
NAME = "msda_reference_sm100"


def target() -> Path:
    return OUT / f"{NAME}{sysconfig.get_config_var('EXT_SUFFIX')}"


def build(force: bool = False) -> Path | None:
    if not REF_SRC.exists():
        return target() if target().exists() else None
    import torch
    tdir = Path(torch.__file__).parent
    so = target()
    if so.exists() and not force:  # cuda.o from an interrupted run is rebuilt
        return so
    OUT.mkdir(exist_ok=True)
    inc = [f"-I{REF_SRC}", "-I/usr/local/cuda/include", f"-I{tdir / 'include'}", f"-I{tdir / 'include/torch/csrc/api/include'}",
           f"-I{sysconfig.get_paths()['include']}"]
    defs = ["-DWITH_CUDA", f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    compat = str(HERE / "ref_compat.h")
    objs = []
    steps = [
        (["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
          "-include", compat, "-Xcompiler", "-fPIC", "-c", str(REF_SRC / "cuda/ms_deform_attn_cuda.cu")], "cuda.o"),
        (["g++", "-O2", "-std=c++17", "-fPIC", "-include", compat, "-c", str(REF_SRC / "cpu/ms_deform_attn_cpu.cpp")], "cpu.o"),
        (["g++", "-O2", "-std=c++17", "-fPIC", "-include", compat, "-c", str(REF_SRC / "vision.cpp")], "vision.o"),
    ]
    for cmd, obj in steps:
        o = OUT / obj
        res = subprocess.run(cmd + defs + inc + ["-w", "-o", str(o)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"reference build step failed: {' '.join(cmd)}\n{res.stderr[-4000:]}")
        objs.append(str(o))
    libdir = tdir / "lib"
    subprocess.run(["g++", "-shared", "-o", str(so), *objs, f"-L{libdir}", "-L/usr/local/cuda/lib64",
                    "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-ltorch_python", "-lc10", "-lc10_cuda", "-lcudart",
                    f"-Wl,-rpath,{libdir}"], check=True)
    for o in objs:
        Path(o).unlink()
    return so


def load():
    """Import the built module (None when it was never built, e.g. reference sources absent)."""
    so = target()
    if not so.exists():
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
