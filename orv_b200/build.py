"""In-tree build of liborv_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m orv_b200.build` compiles every `csrc/*.cu` with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` into object files under `build/` and links them into
`orv_b200/liborv_b200.so`.  nvcc cross-compiles without a GPU, so this runs on the CPU-only build box; the `.so`
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
# ORVB_BUILD_VARIANT=<name>: measurement variant (with ORVB_EXTRA_NVCC_FLAGS) built into its own object directory and
# orv_b200/liborv_b200_<name>.so, so the product library is never replaced by an instrumented one.
_VARIANT = os.environ.get("ORVB_BUILD_VARIANT", "")
BUILD = ROOT.parent / "build" / ("orv_b200" + ("_" + _VARIANT if _VARIANT else ""))
LIB = ROOT / ("liborv_b200" + ("_" + _VARIANT if _VARIANT else "") + ".so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("ORVB_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: liborv_b200.so cannot be built")
    return nvcc


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "orv_b200.h"]
    BUILD.mkdir(parents=True, exist_ok=True)
    stamp = BUILD / "stamp"
    digest = _digest(sources + headers)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    nvcc = _nvcc()
    hdr_digest = _digest(headers)

    def compile_one(src: Path) -> Path:
        obj = BUILD / (src.stem + ".o")
        ostamp = BUILD / (src.stem + ".stamp")
        d = hashlib.sha256(src.read_bytes() + hdr_digest.encode()).hexdigest()
        if not force and obj.exists() and ostamp.exists() and ostamp.read_text() == d:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src.stem + ".ptxas.log")).write_text(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        ostamp.write_text(d)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
