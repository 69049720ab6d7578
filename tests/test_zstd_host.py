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


# ---- the decoding side (agc_b200/csrc/zstd_dec.cuh, host build) -------------------------------------------------------
@pytest.fixture(scope="module")
def zd():
    d = os.path.join(ROOT, "tests", "zstd_host")
    so = os.path.join(d, "libzd_host.so")
    src = [os.path.join(d, "zd_host.cpp"), os.path.join(ROOT, "agc_b200", "csrc", "zstd_dec.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src[0]])
    L = C.CDLL(so)
    L.zd_host_decompress.restype = C.c_long
    L.zd_host_decompress.argtypes = [C.c_char_p, C.c_ulong, C.c_char_p, C.c_ulong]

    def decompress(frame, cap):
        out = C.create_string_buffer(cap + 64)
        r = L.zd_host_decompress(frame, len(frame), out, cap)
        return r, out.raw[:max(r, 0)]
    return decompress


def test_decoder_against_libzstd_frames(zd):
    """frames written by the reference's libzstd at fast, lazy and optimal levels (raw / RLE / Huffman / treeless literals,
    predefined / RLE / FSE / repeat sequence tables, multi-block) decode to the input"""
    rng = np.random.default_rng(31)
    for level in (1, 3, 6, 9, 13, 17, 18, 19):
        for gen, kinds in ((_gen, range(7)), (_gen_adv, range(6))):
            for kind in kinds:
                for n in (0, 1, 5, 200, 4000, 40000, 300000):
                    if n > 40000 and (kind + level) % 2:
                        continue
                    raw = gen(rng, kind, n)
                    r, out = zd(agc_parts.zstd_compress(raw, level), len(raw))
                    assert r == len(raw) and out == raw, f"level {level} {gen.__name__} kind {kind} n {n}: {r}"


def test_decoder_roundtrips_the_coder(ze, zd):
    """frames of this repository's coder (both instantiations) decode to the input"""
    rng = np.random.default_rng(32)
    for kind in range(7):
        for n, level, narrow in ((300, 17, True), (9000, 13, True), (30000, 19, True), (50000, 17, False), (300000, 19, False)):
            raw = _gen(rng, kind, n)
            r, out = zd(ze(raw, level, narrow=narrow), len(raw))
            assert r == len(raw) and out == raw


def test_decoder_rejects_damaged_frames(zd):
    """truncated or corrupted input and short output buffers give an error code, never a wrong answer of the right size"""
    rng = np.random.default_rng(33)
    raw = _gen(rng, 2, 20000)
    fr = agc_parts.zstd_compress(raw, 17)
    assert zd(fr, len(raw))[0] == len(raw)
    assert zd(fr, len(raw) - 1)[0] < 0                           # output buffer too small
    for cut in (0, 3, 5, 9, len(fr) // 2, len(fr) - 1):
        assert zd(fr[:cut], len(raw))[0] < 0
    assert zd(b"\x00" * 20, 100)[0] < 0                          # not a zstd frame
    hits = 0
    for _ in range(200):                                         # random single-byte damage: an error, or (rarely) a different text
        i = int(rng.integers(4, len(fr)))
        bad = bytearray(fr); bad[i] ^= 1 << int(rng.integers(0, 8))
        r, out = zd(bytes(bad), len(raw))
        hits += (r < 0) or (out != raw)
    assert hits >= 190
