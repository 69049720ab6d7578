"""The C ABI: header <-> library <-> Python bindings agree, the library loads on a CPU box, and there is no CPU fallback."""
import os
import re
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "agcgpu.h")).read()
    return sorted(set(re.findall(r"\b(agcgpu_[a-z_0-9]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    import agc_b200
    lib = agc_b200.lib()
    syms = header_symbols()
    assert len(syms) >= 29
    out = subprocess.run(["nm", "-D", "--defined-only", agc_b200.lib_path()], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (agcgpu_[a-z_0-9]+)", out)))
    assert exported == syms, (set(syms) ^ set(exported))
    for s in syms:
        getattr(lib, s)


def test_bindings_cover_the_kernel_abi():
    import agc_b200
    for s in agc_b200.EXPORTED_SYMBOLS:
        assert s in header_symbols()


def test_no_cpu_fallback():
    import torch
    import agc_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(agc_b200.AgcGpuError) as e:
        agc_b200.Device()
    assert "no CUDA device" in str(e.value) or "fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """only tests/, tools/, bench.py and __graft_entry__.py may reference oracle/"""
    bad = []
    for d, _, fs in os.walk(os.path.join(ROOT, "agc_b200")):
        if "build" in d:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"agc_oracle|orc_lz|libzstd_ref|liblzdiff_ref|oracle/_ref|import orc", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_sass_is_sm100a():
    import agc_b200
    out = subprocess.run(["cuobjdump", "-lelf", agc_b200.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_[89]\d", out)
