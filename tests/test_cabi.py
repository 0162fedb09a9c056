"""The C-ABI library loads without a GPU and exports every symbol that
include/justpic_c.h declares (no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

from justpic.jl_b200 import _build, _cabi

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "justpic_c.h"


def header_symbols():
    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jp_[a-z0-9_]+)\s*\(", txt)))


def test_header_matches_binding_table():
    assert header_symbols() == sorted(_cabi.SYMBOLS)


def test_library_builds_and_exports_every_symbol():
    lib_path = _build.build()          # nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(str(lib_path))
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in justpic_c.h but not exported"
    assert _cabi.load().jp_version() >= 100


def test_sm100a_cubin_embedded():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(_build.build())], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device():
    import torch
    from justpic.jl_b200 import CUDABackend, LinRange, expand_range, init_particles
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    xv = LinRange(0, 1, 5); dx = xv[1] - xv[0]; xc = LinRange(dx / 2, 1 - dx / 2, 4)
    g = ((xv, expand_range(xc)), (expand_range(xc), xv))
    with pytest.raises(Exception):
        init_particles(CUDABackend, 4, 8, 2, *g)
    with pytest.raises(RuntimeError):
        init_particles(CUDABackend, 4, 8, 2, *g, device="cpu")


def test_product_never_imports_oracle():
    pkg = ROOT / "justpic"
    for f in pkg.rglob("*"):
        if f.suffix in (".py", ".cu", ".h", ".cuh", ".cpp"):
            txt = f.read_text()
            assert "oracle" not in txt.replace("oracle twin", ""), f"{f} mentions the oracle"


def header_enum_constants():
    """every `NAME = value` inside the typedef enums of the header"""
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    out = {}
    for body in re.findall(r"typedef\s+enum\s*\{(.*?)\}", txt, flags=re.S):
        for name, val in re.findall(r"\b(JP_[A-Z0-9_]+)\s*=\s*(-?\d+)", body):
            out[name] = int(val)
    return out


def test_enum_constants_match_the_binding():
    """the Python mirror passes plain integers to jp_set_option & co.: they must be the header's"""
    consts = header_enum_constants()
    assert len(consts) > 30
    checked = 0
    for name, val in consts.items():
        if hasattr(_cabi, name):
            assert getattr(_cabi, name) == val, f"{name}: header {val}, _cabi {getattr(_cabi, name)}"
            checked += 1
    for must in ("JP_OPT_MOVE_POLICY", "JP_MOVE_POLICY_DENSE", "JP_OPT_GRAPH_STEP_OFFSET", "JP_OPT_MOVE_INTERP", "JP_OPT_ADVECT_CLASSIFY",
                 "JP_OPT_PROFILE", "JP_P2G_TWOPASS_FASTW", "JP_MOVE_DIRECT"):
        assert must in consts and hasattr(_cabi, must), must
    assert checked >= 20


def test_integration_doc_names_every_entry_point():
    doc = (ROOT / "INTEGRATION.md").read_text()
    missing = [s for s in header_symbols() if s not in doc]
    assert not missing, f"INTEGRATION.md does not mention {missing}"
