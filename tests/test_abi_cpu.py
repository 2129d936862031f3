"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, rejects what the north star says must be rejected, and fails loudly without a GPU."""
import os
import re
import pytest
from mcmcglm_b200 import _lib, Engine, CggError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cggibbs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgg_[a-z0-9_]+)\s*\(", src)) - {"cgg_exchange_fn"})


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(_lib.EXPORTS) == syms
    assert L.cgg_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.Config) == 136
    assert C.sizeof(_lib.Stats) == 136


def test_unsupported_inputs_are_rejected_before_touching_cuda():
    with pytest.raises(CggError) as e:
        Engine(10, 2, family="Gamma")
    assert e.value.code == _lib.E_UNSUPPORTED
    with pytest.raises(CggError) as e:
        Engine(10, 2, family="binomial", link="cloglog")
    assert e.value.code == _lib.E_UNSUPPORTED
    with pytest.raises(CggError) as e:
        Engine(10, 2, family="poisson", link="probit")          # a link the engine knows, on a family it does not go with
    assert e.value.code == _lib.E_UNSUPPORTED
    with pytest.raises(CggError) as e:
        Engine(10, 2, prior="beta")
    assert e.value.code == _lib.E_UNSUPPORTED
    with pytest.raises(CggError) as e:
        Engine(10, 2, prior="gamma", prior_mu=-1.0)              # gamma shape must be positive
    assert e.value.code == _lib.E_ARG


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(CggError) as e:
        Engine(10, 2)
    assert e.value.code == _lib.E_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mcmcglm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_r_shim_compiles_against_stub_headers():
    """r/src/rshim.c (the .Call layer a maintainer of the R package would ship) against include/cggibbs.h and minimal stand-ins
    for R's headers (tests/r_stub): every C-ABI call in it has the right arity and types.  Syntax only: there is no R here."""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "r_stub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "rshim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    shim = open(os.path.join(ROOT, "r", "src", "rshim.c")).read()
    rcode = "".join(open(os.path.join(ROOT, "r", "R", f)).read() for f in ("front.R", "engine.R"))
    registered = set(re.findall(r'\{"(C_cgg_[a-z0-9_]+)"', shim))
    called = set(re.findall(r"\.Call\((C_cgg_[a-z0-9_]+)", rcode))
    assert called <= registered and len(registered) >= 11, called - registered
