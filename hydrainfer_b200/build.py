"""Builds libhi_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA behind `extern "C"` (include/hi_b200.h), so the build is a
handful of `nvcc -c` calls plus one link, and the .so travels with the source tree to the GPU box.
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
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libhi_b200.so"
OBJ_DIR = CSRC / "build"
SOURCES = ["api.cu", "scatter.cu", "rope.cu", "attn_simt.cu", "attn_tc.cu", "attn_tc2.cu", "attn_decode_tc.cu", "migrate.cu"]
HEADERS = ["common.cuh", "ptx_sm100.cuh", "attn_common.cuh", "tma_maps.h", "../../include/hi_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


# Extra -D flags for debug builds, e.g. HI_BUILD_DEFINES="-DHI_MBAR_DEBUG" (part of the build stamp).
NVCC_FLAGS += [f for f in os.environ.get("HI_BUILD_DEFINES", "").split() if f]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def _stamp() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    """True when lib/libhi_b200.so was built from the sources as they are now."""
    stamp_file = LIB_DIR / "build.stamp"
    return LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == _stamp()


def build_variant(name: str, defines: list[str]) -> Path:
    """Dev builds (tracing / debug instrumentation): lib/libhi_b200_<name>.so compiled with extra -D flags.  Loaded
    instead of the product library when HI_B200_LIB points at it (hydrainfer_b200/_lib.py); never the default."""
    nvcc = _nvcc()
    obj_dir = CSRC / f"build_{name}"
    obj_dir.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    out = LIB_DIR / f"libhi_b200_{name}.so"
    objs = []
    for src in SOURCES:
        obj = obj_dir / (Path(src).stem + ".o")
        res = subprocess.run([nvcc, *NVCC_FLAGS, *defines, "-c", str(CSRC / src), "-o", str(obj)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        objs.append(str(obj))
    res = subprocess.run([nvcc, "-shared", "-o", str(out), *objs, "-cudart", "static"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile (if sources changed) and return the path of the shared library."""
    stamp_file = LIB_DIR / "build.stamp"
    stamp = _stamp()
    if not force and LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB_PATH
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)

    def compile_one(name: str) -> tuple[str, str]:
        obj = OBJ_DIR / (Path(name).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / name), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{res.stdout}\n{res.stderr}")
        return str(obj), res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    (OBJ_DIR / "ptxas.log").write_text("\n".join(log for _, log in results))
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *[o for o, _ in results], "-cudart", "static"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp_file.write_text(stamp)
    return LIB_PATH


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m hydrainfer_b200.build --variant trace -DHI_PAIR_TRACE
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        path = build(force="--force" in sys.argv, verbose=True)
        print(path)
