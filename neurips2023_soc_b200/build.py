"""Compile csrc/ into the in-tree C-ABI library ``libmsda_b200.so`` (sm_100a only).

Plain ``nvcc``: no torch headers, no JIT cache.  Every ``.cu`` is one translation unit, compiled in
parallel into ``csrc/_obj/`` and linked with ``nvcc --shared``.  The library stays inside the package
directory (git-ignored) so that it travels with a snapshot of the repo.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_obj"
LIB = PKG / "libmsda_b200.so"
SOURCES = sorted(CSRC.glob("*.cu"))
HEADERS = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG.parent / "include" / "msda_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; libmsda_b200.so cannot be built")
    return cand


def is_stale(target: Path = LIB, sources=None) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(p.stat().st_mtime > t for p in (sources if sources is not None else SOURCES + HEADERS))


def _compile(src: Path, out: Path, extra, verbose: bool) -> str:
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-c", "-o", str(out), str(src)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n" + res.stdout + res.stderr)
    return res.stderr


def build_library(force: bool = False, verbose: bool = False, defines=(), out: Path = LIB) -> Path:
    """``defines``: extra ``-DNAME=VALUE`` macros (kernel-variant builds, tools/build_variant.sh); a variant goes to
    its own ``out`` and object directory."""
    variant = out != LIB
    if not force and not variant and not is_stale():
        return LIB
    objdir = OBJ / (out.stem if variant else "default")
    objdir.mkdir(parents=True, exist_ok=True)
    extra = [f"-D{d}" for d in defines]
    objs = [objdir / (s.stem + ".o") for s in SOURCES]
    todo = [(s, o) for s, o in zip(SOURCES, objs) if force or variant or is_stale(o, [s] + HEADERS)]
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        logs = list(ex.map(lambda so: _compile(so[0], so[1], extra, verbose), todo))
    if verbose:
        print("\n".join(logs))
    res = subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", str(out),
                          *map(str, objs)], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if variant:
        shutil.rmtree(objdir, ignore_errors=True)      # a variant's objects are never reused
    return out


if __name__ == "__main__":
    import sys
    # python -m neurips2023_soc_b200.build [--force] [-v] [--variant NAME -DMACRO=VALUE ...]
    #   --variant: kernel-variant build for A/B runs -> variants/NAME.so (git-ignored, shipped by gpurun;
    #   select with MSDA_LIB=$PWD/variants/NAME.so)
    argv = sys.argv[1:]
    if "--variant" in argv:
        name = argv[argv.index("--variant") + 1]
        out = PKG.parent / "variants" / f"{name}.so"
        out.parent.mkdir(exist_ok=True)
        print(build_library(force=True, verbose="-v" in argv, defines=[a[2:] for a in argv if a.startswith("-D")], out=out))
    else:
        print(build_library(force="--force" in argv, verbose="-v" in argv))
