"""Build the sm_100a shared library `liblabelanything_b200.so` in-tree with nvcc.

The library is plain CUDA C++ behind a C ABI (include/labelanything_b200.h): no torch headers, no pybind,
statically linked cudart, cuTensorMapEncodeTiled resolved from the driver at run time.  nvcc cross-compiles
for sm_100a without a GPU, so this runs in the CPU-only build container; the .so travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "liblabelanything_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; labelanything_b200 needs the CUDA 12.9 toolkit to build")
    return cand


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _fingerprint(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "labelanything_b200.h"]:
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: Path, verbose: bool) -> Path:
    obj = BUILD_DIR / (src.stem + ".o")
    stamp = BUILD_DIR / (src.stem + ".sha")
    fp = _fingerprint(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == fp:
        return obj
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = BUILD_DIR / (src.stem + ".log")
    log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        for line in (res.stdout + res.stderr).splitlines():
            if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                print(f"[labelanything_b200.build] {src.name}: {line.strip()}", file=sys.stderr)
    stamp.write_text(fp)
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compile every csrc/*.cu for sm_100a and link liblabelanything_b200.so (incremental)."""
    nvcc = _nvcc()
    BUILD_DIR.mkdir(exist_ok=True)
    if force:
        for f in BUILD_DIR.glob("*.sha"):
            f.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


def build_variant(tag: str, defines: list[str], sources: tuple[str, ...] = ("la_attention.cu",)) -> Path:
    """Experiment builds: recompile `sources` with extra -D flags, link them with the regular objects of the other
    sources into _variants/liblabelanything_b200_<tag>.so.  Loaded instead of the product library when the environment
    variable LA_B200_LIB points at it (tools/ only; the product path never sets it)."""
    nvcc = _nvcc()
    build()
    vdir = PKG_DIR / "_variants"
    vdir.mkdir(exist_ok=True)
    objs = []
    for src in _sources():
        if src.name in sources:
            obj = vdir / f"{src.stem}_{tag}.o"
            cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(src), "-o", str(obj)]
            res = subprocess.run(cmd, capture_output=True, text=True)
            (vdir / f"{src.stem}_{tag}.log").write_text(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for variant {tag}:\n{res.stdout}\n{res.stderr}")
            objs.append(obj)
        else:
            objs.append(BUILD_DIR / (src.stem + ".o"))
    out = vdir / f"liblabelanything_b200_{tag}.so"
    cmd = [nvcc, "-shared", "-o", str(out), *map(str, objs), "-cudart", "static",
           "-gencode", "arch=compute_100a,code=sm_100a"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":   # python -m labelanything_b200.build --variant tag D1=1 D2=0
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force="--force" in sys.argv))
