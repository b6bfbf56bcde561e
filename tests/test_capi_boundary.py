"""CPU tier: the C-ABI library loads and exports every symbol include/mps_capi.h declares (no compute without a GPU),
and refuses to run without a device instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from openmps_b200 import capi

HEADER = os.path.join(ROOT, "include", "mps_capi.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mps_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("mps_create", "mps_destroy", "mps_add_particles", "mps_download", "mps_forward_time", "mps_forward_time_auto",
                 "mps_search_neighbor", "mps_compute_density", "mps_set_ppe", "mps_solve_ppe", "mps_pressure_gradient",
                 "mps_dynamic_stabilize", "mps_get_neighbors", "mps_get_csr"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "libopenmps_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/mps_capi.h but not exported: {missing}"


def test_no_cpu_fallback_without_a_device():
    lib = capi.load_library()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU-only build box")
    from openmps_b200 import scenes
    with pytest.raises(capi.MpsError) as ei:
        capi.GpuComputer(scenes.dambreak2d().env)
    assert ei.value.code == capi.MPS_CUDA_ERROR
    assert lib.mps_stage_name(0) == b"sort"


def test_product_does_not_reference_the_oracle():
    """The shipped path (package + headers) must not import, link or mention anything under oracle/."""
    bad = []
    for base in ("openmps_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "_obj" in dp or "__pycache__" in dp:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|mps_oracle|libref_", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
