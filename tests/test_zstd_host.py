"""Host build of the residual coder (tests/zstd_host, same source as the device build, one lane) vs the reference's own
libzstd (oracle/_ref/libzstd_ref.so): frames must be byte-identical.  CPU-only; the GPU suite repeats this on the device."""
import ctypes as C
import os
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import agc_parts
from test_gpu_zstd import _gen, _gen_adv

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libzstd_ref.so")), reason="libzstd_ref.so not built")


@pytest.fixture(scope="module")
def ze():
    d = os.path.join(ROOT, "tests", "zstd_host")
    so = os.path.join(d, "libze_host.so")
    src = [os.path.join(d, "ze_host.cpp"), os.path.join(ROOT, "agc_b200", "csrc", "zstd_enc.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src[0]])
    L = C.CDLL(so)
    for f in (L.ze_host_compress, L.ze_host_compress_narrow):
        f.restype = C.c_long
        f.argtypes = [C.c_char_p, C.c_ulong, C.c_int, C.c_char_p, C.c_ulong]
    L.ze_host_bound.restype = C.c_ulong
    L.ze_host_bound.argtypes = [C.c_ulong]

    def compress(raw, level, narrow=False):
        cap = L.ze_host_bound(len(raw)) + 64
        out = C.create_string_buffer(cap)
        buf = raw + b"\0" * 64                       # the coder may read (never use) a few bytes past the input
        r = (L.ze_host_compress_narrow if narrow else L.ze_host_compress)(buf, len(raw), level, out, cap)
        assert r >= 0, f"ze_host_compress failed: {r}"
        return out.raw[:r]
    return compress


def test_host_frames_match_reference(ze):
    rng = np.random.default_rng(11)
    for n in (0, 1, 7, 9, 64, 300, 4000, 16384, 16385, 40000, 131072, 131073):
        for kind in range(7):
            for level in (13, 17, 18, 19):
                if n > 4000 and (kind + level) % 3:
                    continue
                raw = _gen(rng, kind, n)
                assert ze(raw, level) == agc_parts.zstd_compress(raw, level), f"{n} bytes kind {kind} level {level}"


def test_host_multiblock(ze):
    rng = np.random.default_rng(12)
    for kind, n, level in ((2, 300000, 17), (1, 262145, 19), (1, 400000, 17)):
        raw = _gen(rng, kind, n)
        assert ze(raw, level) == agc_parts.zstd_compress(raw, level)


def test_host_near_duplicate_text(ze):
    """raw-group packs: concatenated near-copies of one sequence (long matches, insert-heavy parse)"""
    rng = np.random.default_rng(13)
    ref = rng.integers(0, 4, 27000).astype(np.uint8)
    parts = []
    for _ in range(12):
        t = ref.copy(); m = rng.random(len(t)) < 0.01; t[m] = (t[m] + 1) % 4
        parts.append(bytes(t))
    raw = b"".join(parts)
    for level in (17, 19):
        assert ze(raw, level) == agc_parts.zstd_compress(raw, level)


def test_host_window_engine_cases(ze):
    """deep bucket chains, skipped positions, compare-limit and end-of-block walks, copies across block boundaries"""
    rng = np.random.default_rng(8)
    for kind in range(6):
        for n, level in ((3000, 17), (16000, 13), (16384, 19), (40000, 17), (100000, 17), (131080, 19), (200000, 18), (300000, 17)):
            raw = _gen_adv(rng, kind, n)
            assert ze(raw, level) == agc_parts.zstd_compress(raw, level), f"kind {kind}, {n} bytes, level {level}"


def test_host_narrow_coder(ze):
    """the 32-slot-window instantiation (device: one warp per frame, inputs <= 32 KB) produces the same frames"""
    rng = np.random.default_rng(14)
    for n in (0, 1, 9, 300, 4000, 16384, 16385, 32768):
        for kind in range(7):
            for level in (13, 17, 19):
                raw = _gen(rng, kind, n)
                assert ze(raw, level, narrow=True) == agc_parts.zstd_compress(raw, level), f"{n} bytes kind {kind} level {level}"
    for kind in range(6):
        for n, level in ((3000, 17), (16000, 13), (30000, 19)):
            raw = _gen_adv(rng, kind, n)
            assert ze(raw, level, narrow=True) == agc_parts.zstd_compress(raw, level), f"adv kind {kind}, {n} bytes, level {level}"


def test_host_btlazy2_class(ze):
    """level 13 above 256 KB = ZSTD_btlazy2 (lazy parser at depth 2 over the delayed-update binary tree)"""
    rng = np.random.default_rng(21)
    for kind, n in ((0, 262145), (1, 300000), (2, 262145), (2, 700000), (6, 300000)):
        raw = _gen(rng, kind, n)
        assert ze(raw, 13) == agc_parts.zstd_compress(raw, 13), f"kind {kind}, {n} bytes"
    for kind, n in ((2, 270000), (5, 1200000), (0, 270000)):
        raw = _gen_adv(rng, kind, n)
        assert ze(raw, 13) == agc_parts.zstd_compress(raw, 13), f"adv kind {kind}, {n} bytes"
