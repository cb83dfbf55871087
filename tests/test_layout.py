"""Structural guarantees the judge checks: the product never touches the oracle, the C ABI is what the header says."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for base, _, files in os.walk(d):
        if "__pycache__" in base:
            continue
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(base, f)


def test_product_never_imports_the_oracle_or_the_reference():
    for path in _py_files(os.path.join(ROOT, "orv_b200")):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), path
        assert "flat_oracle" not in src, path
        assert "/root/reference" not in src, path


def test_header_symbols_are_exported_by_the_library():
    from orv_b200 import _lib as L
    header = open(os.path.join(ROOT, "include", "orv_b200.h")).read()
    declared = set(re.findall(r"\b(orvb_[a-z0-9_]+)\s*\(", header))
    typedef_names = set(re.findall(r"typedef struct (orvb_[a-z0-9_]+)", header))
    declared -= typedef_names
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    lib = L.load()  # builds with nvcc if the .so is missing; loading needs no GPU
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"liborv_b200.so does not export {sym}"
    assert lib.orvb_version() == 104
    assert isinstance(lib.orvb_last_error(), bytes)
    # ... and nothing else: every orvb_* symbol the library exports is declared in the header
    import subprocess
    nm = subprocess.run(["nm", "-D", "--defined-only", str(L.LIB_PATH)], capture_output=True, text=True, check=True)
    exported = {ln.split()[-1] for ln in nm.stdout.splitlines() if ln.split() and ln.split()[-1].startswith("orvb_")}
    assert exported == declared, exported ^ declared


def test_ctypes_structs_match_header_sizes():
    """The ctypes mirrors must have the C layout (sizes computed from the header's field lists by gcc)."""
    import subprocess
    import tempfile
    from orv_b200 import _lib as L
    prog = r'''
#include <stdio.h>
#include "orv_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(orvb_rowmap), sizeof(orvb_gemm_args), sizeof(orvb_ln_args),
         sizeof(orvb_config), sizeof(orvb_block_weights), sizeof(orvb_weights), sizeof(orvb_shape),
         sizeof(orvb_forward_args), sizeof(orvb_sampler_step_args), sizeof(orvb_attention_args),
         sizeof(orvb_voxelize_args), sizeof(orvb_gs_args), sizeof(orvb_conv_args), sizeof(orvb_spatial_norm_args));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirrors = [L.RowMap, L.GemmArgs, L.LnArgs, L.Config, L.BlockWeights, L.Weights, L.Shape, L.ForwardArgs,
               L.SamplerStepArgs, L.AttentionArgs, L.VoxelizeArgs, L.GsArgs, L.ConvArgs, L.SpatialNormArgs]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_no_reference_sources_in_repo():
    for d in ("orv_b200", "oracle", "tests"):
        for path in _py_files(os.path.join(ROOT, d)):
            assert os.path.basename(os.path.dirname(path)) != "orv", path
