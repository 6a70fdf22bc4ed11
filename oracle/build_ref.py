"""TEST INFRASTRUCTURE — builds `oracle/_ref/`, the reference "as built", from the sources where they lie under
/root/reference.  Nothing under hydrainfer_b200/ imports it; tests/, smoke() and bench.py's reference legs are its
only users.  Runs in the build container only (the GPU box has no /root/reference; it receives the finished tree).

    python oracle/build_ref.py            # the four small pybind modules + the python files of the path
    python oracle/build_ref.py --fa2      # also the reference's in-tree FlashAttention-2 `mha_varlen_fwd` (minutes)

What it produces (all git-ignored, NOT gpurun-ignored, so it travels to the GPU box like our own .so):

    oracle/_ref/hydrainfer/{__init__,layer/*,memory/*,utils/*,_C/**/__init__}.py     the reference's own python files of the
                                                                                    path (import closure of layer + memory)
    oracle/_ref/hydrainfer/_C/kernel/kv_cache_kernels.<abi>.so      csrc/kernel/kv_cache_kernels/{kv_cache_kernels.cu,*_pybind.cpp}
    oracle/_ref/hydrainfer/_C/kernel/cache_kernels.<abi>.so         csrc/kernel/cache_kernels/{cache_kernels.cu,*_pybind.cpp}
    oracle/_ref/hydrainfer/_C/kernel/position_embedding.<abi>.so    csrc/kernel/position_embedding/{rope.cu,*_pybind.cpp}
    oracle/_ref/hydrainfer/_C/data_transfer/block_migration.<abi>.so  csrc/data_transfer/{block_migration.cpp,*_pybind.cpp}
    oracle/_ref/hydrainfer/_C/kernel/flash_attn.<abi>.so (--fa2)    csrc/kernel/flash_attn/{flash_api.cpp,*_pybind.cpp} + the
                                      instantiation files written by the reference's own generate_instantiation_cu.py
                                      (run unmodified, output under oracle/_ref/build/), cutlass from third_party/cutlass

i.e. exactly where the reference's CMake drops them (csrc/CMakeLists.txt:4-11), compiled UNMODIFIED with nvcc 12.9 for
sm_100 against this image's torch headers (the reference's own build downloads libtorch 2.4.0 and cannot configure
offline, CMakeLists.txt:78-104).  No reference source is copied into the tracked tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = ROOT / "oracle" / "_ref"
PKG = OUT / "hydrainfer"
BUILD = OUT / "build"
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")

# python files on the path: the import closure of hydrainfer.layer.{causal_attention,multihead_attention,rotary_embedding}
# and hydrainfer.memory (oracle/make_golden.py lists the entry points)
PY_GLOBS = ["__init__.py", "_C/__init__.py", "_C/kernel/__init__.py", "_C/data_transfer/__init__.py",
            "layer/*.py", "memory/*.py", "utils/*.py"]

MODULES = {
    # name: (destination package dir, sources relative to csrc/, extra include dirs relative to csrc/)
    "kv_cache_kernels": ("_C/kernel", ["kernel/kv_cache_kernels/kv_cache_kernels.cu", "kernel/kv_cache_kernels/kv_cache_kernels_pybind.cpp"], ["kernel/kv_cache_kernels"]),
    "cache_kernels": ("_C/kernel", ["kernel/cache_kernels/cache_kernels.cu", "kernel/cache_kernels/cache_kernels_pybind.cpp"], ["kernel/cache_kernels"]),
    "position_embedding": ("_C/kernel", ["kernel/position_embedding/rope.cu", "kernel/position_embedding/position_embedding_pybind.cpp"], ["kernel/position_embedding", "kernel"]),
    "block_migration": ("_C/data_transfer", ["data_transfer/block_migration.cpp", "data_transfer/block_migration_pybind.cpp"], ["data_transfer"]),
}


def _torch_flags() -> tuple[list[str], list[str]]:
    import torch
    from torch.utils import cpp_extension as ce

    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    lib_dir = Path(torch.__file__).parent / "lib"
    link = [f"-L{lib_dir}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            "-Xlinker", f"-rpath={lib_dir}"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc + [f"-D_GLIBCXX_USE_CXX11_ABI={abi}"], link


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


COMMON = ["-gencode", "arch=compute_100,code=sm_100", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-w",
          "--expt-relaxed-constexpr", "--expt-extended-lambda", "-DTORCH_API_INCLUDE_EXTENSION_H"]


def _stamp(sources: list[Path], extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for s in sources:
        h.update(s.read_bytes())
    return h.hexdigest()


def _compile(name: str, sources: list[Path], includes: list[Path], dest: Path, extra_flags: list[str] = ()) -> Path:
    out = dest / f"{name}{EXT_SUFFIX}"
    stamp_file = BUILD / f"{name}.stamp"
    stamp = _stamp(sources, " ".join(COMMON + list(extra_flags)))
    if out.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return out
    inc, link = _torch_flags()
    obj_dir = BUILD / name
    obj_dir.mkdir(parents=True, exist_ok=True)
    dest.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    flags = [*COMMON, *extra_flags, f"-DPY_MODULE_NAME={name}", f"-DTORCH_EXTENSION_NAME={name}", *inc, *[f"-I{p}" for p in includes]]

    host_flags = ["-O2", "-std=c++17", "-fPIC", "-w", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-DPY_MODULE_NAME={name}", f"-DTORCH_EXTENSION_NAME={name}",
                  *inc, *[f"-I{p}" for p in includes]]
    cxx = os.environ.get("CXX") or shutil.which("g++")

    def one(src: Path) -> str:
        obj = obj_dir / (src.stem + ".o")
        if src.suffix == ".cpp" and cxx is not None:
            # host-only translation units (pybind glue, block_migration.cpp: CUDA runtime API calls, no device code): g++ is
            # several times faster than nvcc's front end on the torch headers
            cmd = [cxx, *host_flags, "-c", str(src), "-o", str(obj)]
        else:
            cmd = [nvcc, *flags, "-x", "cu", "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"reference source {src} did not compile:\n{res.stderr[-4000:]}")
        return str(obj)

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        objs = list(pool.map(one, sources))
    res = subprocess.run([nvcc, "-shared", "-o", str(out), *objs, *link], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link of {name} failed:\n{res.stderr[-4000:]}")
    stamp_file.write_text(stamp)
    return out


def copy_python_tree() -> None:
    src_pkg = REF / "hydrainfer"
    for pattern in PY_GLOBS:
        for src in sorted(src_pkg.glob(pattern)):
            dst = PKG / src.relative_to(src_pkg)
            dst.parent.mkdir(parents=True, exist_ok=True)
            if not dst.exists() or dst.read_bytes() != src.read_bytes():
                shutil.copyfile(src, dst)


def build_fa2() -> Path:
    """The reference's vendored FlashAttention-2 (a7): its generator script is executed as it is, with oracle/_ref/build
    as the working directory, and the 16 instantiation files it writes are compiled with the reference's headers."""
    gen_dir = BUILD / "fa2"
    gen_dir.mkdir(parents=True, exist_ok=True)
    fa = REF / "csrc" / "kernel" / "flash_attn"
    subprocess.run([sys.executable, str(fa / "generate_instantiation_cu.py")], cwd=gen_dir, check=True)
    generated = sorted((gen_dir / "generated").glob("*.cu"))
    sources = [fa / "flash_api.cpp", fa / "flash_attn_pybind.cpp", *generated]
    includes = [fa, fa / "src", REF / "third_party" / "cutlass" / "include"]
    return _compile("flash_attn", sources, includes, PKG / "_C" / "kernel",
                    extra_flags=["-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
                                 "-U__CUDA_NO_HALF2_OPERATORS__", "--use_fast_math"])


def build(fa2: bool = False, verbose: bool = True) -> dict[str, Path]:
    if not REF.exists():
        raise RuntimeError("/root/reference is not present: oracle/_ref can only be built in the build container")
    BUILD.mkdir(parents=True, exist_ok=True)
    copy_python_tree()
    built = {}
    csrc = REF / "csrc"
    with ThreadPoolExecutor(max_workers=len(MODULES)) as pool:  # the modules are independent: compile them side by side
        futures = {name: pool.submit(_compile, name, [csrc / s for s in sources], [csrc / i for i in includes], PKG / dest)
                   for name, (dest, sources, includes) in MODULES.items()}
        for name, fut in futures.items():
            built[name] = fut.result()
            if verbose:
                print(f"oracle/_ref: {built[name].relative_to(ROOT)}", file=sys.stderr)
    if fa2:
        built["flash_attn"] = build_fa2()
        if verbose:
            print(f"oracle/_ref: {built['flash_attn'].relative_to(ROOT)}", file=sys.stderr)
    return built


if __name__ == "__main__":
    build(fa2="--fa2" in sys.argv)
