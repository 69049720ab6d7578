"""agc_b200 -- B200 (sm_100a) implementation of AGC's compression hot path behind a C ABI (include/agcgpu.h).

This package is only a thin ctypes view of libagcgpu.so for tests, bench.py and the Python-side tooling; the host
pipeline that mirrors CAGCCompressor lives in agc_b200/csrc/host (C++), as the reference's host code is C++.
There is no CPU fallback: importing works everywhere (so the symbol table can be checked on a CPU box), creating a
context without an sm_100 GPU raises.
"""
from ._lib import (Device, AgcGpuError, lib, lib_path, Cut, SegReq, Assign, Stats, Params, FKmer, EXPORTED_SYMBOLS)

__all__ = ["Device", "AgcGpuError", "lib", "lib_path", "Cut", "SegReq", "Assign", "Stats", "Params", "FKmer", "EXPORTED_SYMBOLS"]
