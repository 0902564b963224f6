"""Build libpcm_b200.so (all sm_100a kernels + the C ABI of include/pcm_b200.h) in-tree.

    python -m pointcloudmatters_b200.build [--force]

One nvcc invocation per .cu (run in parallel), explicit
`-gencode arch=compute_100a,code=sm_100a -lineinfo`; the shared object lands next to this file
so it travels to the GPU box with the repository snapshot.  No torch involvement: the library
has a plain C ABI and is loaded with ctypes (pointcloudmatters_b200/_lib.py).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libpcm_b200.so"
OBJ = HERE / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _needs(src: Path, obj: Path) -> bool:
    if not obj.exists():
        return True
    deps = [src] + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "pcm_b200.h"]
    return any(d.stat().st_mtime > obj.stat().st_mtime for d in deps)


def _compile(src: Path, obj: Path, verbose: bool) -> str:
    cmd = [_nvcc(), *NVCC_FLAGS, "-Xptxas", "-v", "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sources()
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    todo = [(s, o) for s, o in zip(srcs, objs) if force or _needs(s, o)]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            logs = list(ex.map(lambda so: _compile(so[0], so[1], verbose), todo))
        (OBJ / "ptxas.log").write_text("\n".join(logs))
        if verbose:
            print("\n".join(logs))
    if todo or not OUT.exists():
        cmd = [_nvcc(), "-shared", "-o", str(OUT), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
