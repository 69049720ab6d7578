"""Diagonal-streaming LZ-diff encoder (agc_b200/csrc/lz_diag_core.cuh: warp per segment, windows along the current diagonal, states
resolved with real index probes) built for the host (tests/lzd_host, same source as the device kernel; the 32 lanes of every phase
run one after the other) against the C oracle's CLZDiff_V2::Encode: every delta must be byte-identical, whatever the window size."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc
from test_lzc_host import make_case, mutate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u8p = C.POINTER(C.c_uint8); u32p = C.POINTER(C.c_uint32)
NAMES = "rounds windows path_tokens stops defers opens multi probes".split()


def build(defs):
    d = os.path.join(ROOT, "tests", "lzd_host")
    subprocess.check_call(["make", "-C", d, "-s", "-B"] + ([f"LZD_DEFS={defs}"] if defs else []))
    # a private copy: the library is rebuilt with other window sizes by the next parameter
    import shutil, tempfile
    tmp = tempfile.NamedTemporaryFile(suffix=".so", delete=False); tmp.close()
    shutil.copy(os.path.join(d, "liblzd_host.so"), tmp.name)
    L = C.CDLL(tmp.name)
    L.lzd_host_encode.restype = C.c_long
    L.lzd_host_encode.argtypes = [u8p, C.c_uint, u8p, C.c_uint, u32p, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, u8p, C.c_uint]
    return L


def enc(L, text, ref, mml, z, is_rc=0, lead=0):
    ht, short = z.ht()
    text = np.ascontiguousarray(text, np.uint8); ref = np.ascontiguousarray(ref, np.uint8)
    cap = len(text) * 3 // 2 + 64
    out = np.zeros(cap, np.uint8)
    r = L.lzd_host_encode(text.ctypes.data_as(u8p), len(text), ref.ctypes.data_as(u8p), len(ref), ht.ctypes.data_as(u32p), len(ht),
                          int(short), mml, is_rc, lead, out.ctypes.data_as(u8p), cap)
    return r, (out[:r].tobytes() if r >= 0 else None)


def counters(L):
    a = (C.c_ulonglong * 8)(); L.lzd_host_counters(a)
    return dict(zip(NAMES, list(a)))


@pytest.mark.parametrize("defs", ["", "-DLZD_ITERS=1u"])      # whole-segment windows / 2048-base windows (every boundary case)
def test_diag_encoder_fuzz(defs):
    L = build(defs)
    for s in range(320):
        rng, mml, ref, t = make_case(s)
        z = orc.LZ(ref, mml)
        r, got = enc(L, t, ref, mml, z, int(rng.random() < 0.4), int(rng.integers(0, 70)))
        assert r >= 0 and got == z.encode(t), f"seed {s} (kind {s % 8}): delta differs from the oracle"


@pytest.mark.parametrize("p", [0.001, 0.01, 0.03])
def test_diag_encoder_realistic_segments(p):
    """60 kb segments with SNPs (and a few indels): identical deltas, and nearly all tokens come from the windows"""
    L = build("")
    rng = np.random.default_rng(int(p * 1e4))
    counters(L)
    for i in range(30):
        ref = rng.integers(0, 4, 60031).astype(np.uint8); t = mutate(rng, ref, p, i % 3)
        z = orc.LZ(ref, 20)
        r, got = enc(L, t, ref, 20, z, i % 2, i % 37)
        assert got == z.encode(t), f"segment {i}: delta differs from the oracle"
    c = counters(L)
    if p <= 0.01:
        assert c["rounds"] <= 6 * 30 and c["path_tokens"] >= 25 * 30 * (p / 0.001) ** 0.5, c


def test_diag_encoder_edges():
    L = build("")
    rng = np.random.default_rng(5)
    ref = rng.integers(0, 4, 5000).astype(np.uint8)
    z = orc.LZ(ref, 20)
    cases = [ref.copy(), ref[:17], ref[:20], ref[:21], ref[:40], np.zeros(0, np.uint8), ref[100:], ref[:-100], np.concatenate([ref, ref[:300]]),
             rng.integers(0, 4, 3000).astype(np.uint8), np.concatenate([ref[:2000], ref[2003:]]), np.concatenate([ref[:2000], ref[1990:]])]
    t = ref.copy(); t[0] = (t[0] + 1) % 4; cases.append(t)
    t = ref.copy(); t[-1] = (t[-1] + 1) % 4; cases.append(t)
    t = ref.copy(); t[1000:1040] = (t[1000:1040] + 1) % 4; cases.append(t)               # a run of mismatches
    t = ref.copy(); t[np.arange(1000, 1400, 7)] = (t[np.arange(1000, 1400, 7)] + 2) % 4; cases.append(t)   # mismatches 7 apart
    for i, t in enumerate(cases):
        for rc in (0, 1):
            r, got = enc(L, t, ref, 20, z, rc, 3 * i)
            assert got == z.encode(t), f"edge case {i} rc={rc}"
