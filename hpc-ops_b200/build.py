"""Build the C-ABI library `hpc/_C.so` in-tree with nvcc for sm_100a.

    python hpc-ops_b200/build.py [--force] [--verbose]

Each csrc/*.cu is compiled to an object (only when stale) and linked into one shared library that
exports the `extern "C"` launchers declared in include/hpc_b200.h. No torch headers are involved:
the host side (hpc/*.py) binds the library with ctypes.
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OBJ = ROOT / "build" / "obj"
OUT = ROOT / "hpc" / "_C.so"
REPO = ROOT.parent

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(REPO / "include"),
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found; cannot build hpc/_C.so")
    return nvcc


def _stale(src: Path, obj: Path, deps) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(d.stat().st_mtime > t for d in [src, *deps])


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((REPO / "include").glob("*.h"))
    headers.append(Path(__file__))
    sources = sorted(CSRC.glob("*.cu"))
    built_json = json.dumps({"arch": "sm_100a", "nvcc": nvcc, "built": time.strftime("%Y-%m-%d")})
    info = OBJ / "built_info.h"
    text = "#define HPC_B200_BUILT_JSON " + json.dumps(built_json) + "\n"
    if not info.exists() or info.read_text() != text:
        info.write_text(text)
    defs = ["-include", str(info)]

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        if not force and not _stale(src, obj, headers):
            return obj, None
        cmd = [nvcc, *NVCC_FLAGS, *(defs if src.stem == "version" else []), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj, (r.stdout + r.stderr)

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    relink = force or not OUT.exists() or any(o.stat().st_mtime > OUT.stat().st_mtime for o in objs)
    if relink:
        cmd = [nvcc, "-shared", "-o", str(OUT), *map(str, objs), "-gencode",
               "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
