"""CPU-side checks of the drop-in boundary: libspindyn_cuda.so loads, exports
every symbol include/spindyn.h declares, and fails loudly (no CPU fallback) when
no CUDA device is present.  No compute calls are made here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, sd


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "spindyn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(sd.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spindyn.h but not exported"


def test_ctypes_binding_covers_the_header():
    bound = set(sd.SIGNATURES) | set(sd.OTHER_SYMBOLS)
    assert set(declared_symbols()) == bound


def test_no_cpu_fallback_without_a_device():
    if sd.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(sd.SpinDynError) as e:
        sd.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under spindynamics.jl_b200/ may name it."""
    pkg = os.path.join(ROOT, "spindynamics.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
                assert "libsd_emul" not in text, f


def test_host_side_validation_mirrors_reference_errors():
    """Basis.jl:9-20 / SpinModel.jl:84-86 raise ArgumentError before any device work."""
    with pytest.raises(ValueError):
        sd.build_model(0)
    with pytest.raises(ValueError):
        sd.build_model(64)
    with pytest.raises(ValueError):
        sd.build_model(4, nup=5)
    with pytest.raises(ValueError):
        sd.XXZChain(4, boundary="twisted")
    assert sd.nn_hopping(4, 0.5) == [(1, 2, 0.5), (2, 3, 0.5), (3, 4, 0.5)]
    assert len(sd.long_range_hopping(5, lambda i, j: 1.0)) == 10
    assert sd.bit_at(0b0101, 0) == 1 and sd.bit_at(0b0101, 1) == 0
    assert sd.sz_value(1) == 0.5 and sd.sz_value(0) == -0.5
    assert sd.flip_bits(0b0101, 0, 1) == 0b0110
    a, b = sd._rescaling_from_bounds(-3.0, 5.0)                      # test_KPM.jl:32-41
    assert abs((-3.0 - b) / a + 0.99) < 1e-12 and abs((5.0 - b) / a - 0.99) < 1e-12


def _header_arg_counts():
    src = open(os.path.join(ROOT, "include", "spindyn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(sd_[A-Za-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_julia_binding_matches_the_header():
    """spindynamics.jl_b200/julia/SpinDynamicsCUDA.jl cannot run here (no julia): check statically that
    every symbol it ccalls is declared in include/spindyn.h with the same number of arguments."""
    jl = open(os.path.join(ROOT, "spindynamics.jl_b200", "julia", "SpinDynamicsCUDA.jl")).read()
    hdr = _header_arg_counts()
    calls = re.findall(r"ccall\(\(:(sd_[A-Za-z0-9_]+),\s*\w+\),\s*\w+,\s*\(([^)]*)\)", jl)
    assert len(calls) >= 25
    for name, argt in calls:
        assert name in hdr, f"{name} is ccalled from Julia but not declared in include/spindyn.h"
        n = len([a for a in argt.split(",") if a.strip()])
        assert n == hdr[name], f"{name}: Julia passes {n} arguments, the header declares {hdr[name]}"


def test_ctypes_signatures_have_the_headers_argument_counts():
    """Every entry of spindyn._lib.SIGNATURES lists as many argument types as include/spindyn.h declares parameters."""
    hdr = _header_arg_counts()
    for name, argtypes in sd.SIGNATURES.items():
        assert name in hdr, name
        assert len(argtypes) == hdr[name], f"{name}: ctypes binds {len(argtypes)} arguments, the header declares {hdr[name]}"
