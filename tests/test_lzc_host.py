"""Chunk-parallel LZ-diff encoder (agc_b200/csrc/lz_chunk_core.cuh: thread-per-chunk parse + per-segment stitch) built for the
host (tests/lzc_host, same source as the device kernels) against the C oracle's CLZDiff_V2::Encode: every delta the stitcher
produces must be byte-identical; segments it hands to the sequential kernel must stay rare on realistic data."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u8p = C.POINTER(C.c_uint8); u32p = C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def lzc():
    d = os.path.join(ROOT, "tests", "lzc_host")
    subprocess.check_call(["make", "-C", d, "-s"])
    L = C.CDLL(os.path.join(d, "liblzc_host.so"))
    L.lzc_host_encode.restype = C.c_long
    L.lzc_host_encode.argtypes = [u8p, C.c_uint, u8p, C.c_uint, u32p, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, u8p, C.c_uint]
    L.lzc_host_costs.restype = C.c_long
    L.lzc_host_costs.argtypes = [u8p, C.c_uint, u8p, C.c_uint, u32p, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, C.c_int, u32p]
    return L


def costs(L, text, ref, mml, z, prefix, is_rc=0, lead=0):
    ht, short = z.ht()
    text = np.ascontiguousarray(text, np.uint8); ref = np.ascontiguousarray(ref, np.uint8)
    out = np.zeros(max(len(text), 1), np.uint32)
    r = L.lzc_host_costs(text.ctypes.data_as(u8p), len(text), ref.ctypes.data_as(u8p), len(ref), ht.ctypes.data_as(u32p), len(ht),
                         int(short), mml, is_rc, lead, int(prefix), out.ctypes.data_as(u32p))
    return r, out[:len(text)]


def enc(L, text, ref, mml, z, is_rc=0, lead=0):
    ht, short = z.ht()
    text = np.ascontiguousarray(text, np.uint8); ref = np.ascontiguousarray(ref, np.uint8)
    cap = len(text) * 3 // 2 + 64
    out = np.zeros(cap, np.uint8)
    r = L.lzc_host_encode(text.ctypes.data_as(u8p), len(text), ref.ctypes.data_as(u8p), len(ref), ht.ctypes.data_as(u32p), len(ht),
                          int(short), mml, is_rc, lead, out.ctypes.data_as(u8p), cap)
    return r, (out[:r].tobytes() if r >= 0 else None)


def mutate(rng, ref, p_snp, n_indel=0, max_indel=50):
    t = ref.copy()
    m = rng.random(len(t)) < p_snp
    t[m] = (t[m] + rng.integers(1, 4, int(m.sum()))) % 4
    for _ in range(n_indel):
        pos = int(rng.integers(0, max(1, len(t)))); ln = int(rng.integers(1, max_indel + 1))
        t = np.delete(t, slice(pos, pos + ln)) if rng.random() < 0.5 else np.insert(t, pos, rng.integers(0, 4, ln).astype(np.uint8))
    return t.astype(np.uint8)


def make_case(s):
    rng = np.random.default_rng(s)
    mml = int(rng.choice([15, 18, 20, 24, 32]))
    m = int(rng.choice([100, 3000, 10000, 30000, 60031]))
    ref = rng.integers(0, 4, m).astype(np.uint8)
    kind = s % 8
    if kind == 0: t = ref.copy()                                         # equal sequences -> empty delta
    elif kind == 1: t = mutate(rng, ref, 0.001)
    elif kind == 2: t = mutate(rng, ref, 0.01)
    elif kind == 3: t = mutate(rng, ref, 0.01, 5)
    elif kind == 4: t = mutate(rng, ref, 0.05, 3)
    elif kind == 5:                                                      # novel insertion (chunks without a single match)
        t = mutate(rng, ref, 0.002, 2); a = int(rng.integers(0, len(t)))
        t = np.insert(t, a, rng.integers(0, 4, int(rng.integers(100, 6000))).astype(np.uint8))
    elif kind == 6:                                                      # repeat inside the reference (several candidates)
        if m > 2000: ref[m // 2: m // 2 + 500] = ref[100:600]
        t = mutate(rng, ref, 0.005, 2)
    else: t = mutate(rng, ref[int(rng.integers(0, m // 2)):], 0.003, 1)   # trimmed start (matches off the main diagonal)
    return rng, mml, ref, t


def test_chunk_encoder_fuzz(lzc):
    n_seq = 0
    for s in range(320):
        rng, mml, ref, t = make_case(s)
        z = orc.LZ(ref, mml)
        r, got = enc(lzc, t, ref, mml, z, int(rng.random() < 0.4), int(rng.integers(0, 70)))
        assert r != -2 and r != -3, f"seed {s}: output bound exceeded"
        if r <= -10: n_seq += 1
        else: assert got == z.encode(t), f"seed {s}: delta differs from the oracle"
    assert n_seq <= 16, f"{n_seq} of 320 segments went to the sequential kernel"


@pytest.mark.parametrize("p", [0.001, 0.01, 0.03])
def test_chunk_encoder_realistic_segments(lzc, p):
    """60 kb segments with SNPs (and a few indels): identical deltas, and the stitcher resolves (almost) all of them itself"""
    rng = np.random.default_rng(int(p * 1e4))
    n_seq = 0
    for i in range(40):
        ref = rng.integers(0, 4, 60031).astype(np.uint8); t = mutate(rng, ref, p, i % 3)
        z = orc.LZ(ref, 20)
        r, got = enc(lzc, t, ref, 20, z, i % 2, i % 37)
        if r <= -10: n_seq += 1
        else: assert got == z.encode(t)
    assert n_seq <= 1


def test_chunk_encoder_small_chunks(lzc):
    """small batches use smaller chunks (LzcReq::chunk): same deltas and cost vectors"""
    lzc.lzc_host_set_chunk.argtypes = [C.c_uint]
    try:
        for chunk in (512, 1024):
            lzc.lzc_host_set_chunk(chunk)
            n_seq = 0
            for s in range(160):
                rng, mml, ref, t = make_case(s)
                z = orc.LZ(ref, mml)
                r, got = enc(lzc, t, ref, mml, z, int(rng.random() < 0.4), int(rng.integers(0, 70)))
                if r <= -10: n_seq += 1
                else: assert got == z.encode(t), f"chunk {chunk} seed {s}: delta differs from the oracle"
                r, cv = costs(lzc, t, ref, mml, z, s & 1, 0, 3)
                if r > -10: assert np.array_equal(cv, z.cost_vector(t, s & 1)), f"chunk {chunk} seed {s}: cost vector differs"
            assert n_seq <= 12
    finally:
        lzc.lzc_host_set_chunk(2048)


def test_chunk_encoder_edges(lzc):
    rng = np.random.default_rng(3)
    ref = rng.integers(0, 4, 5000).astype(np.uint8)
    z = orc.LZ(ref, 20)
    cases = [ref[:0], ref[:5], ref[:17], ref[:18], ref[:2048], ref[:2049], ref[1:2049], ref[:4096], ref[3:4099],
             np.concatenate([ref, ref]), rng.integers(0, 4, 7000).astype(np.uint8), np.zeros(3000, np.uint8)]
    for i, t in enumerate(cases):
        for rc in (0, 1):
            r, got = enc(lzc, t, ref, 20, z, rc, 13)
            if r > -10: assert got == z.encode(np.ascontiguousarray(t)), f"edge case {i} rc={rc}"


def test_chunk_cost_vectors_fuzz(lzc):
    """GetCodingCostVector through the chunk parse in cost mode (the missing-middle path): identical vectors, both modes"""
    n_seq = 0
    for s in range(240):
        rng, mml, ref, t = make_case(s)
        if s % 8 == 5 and len(ref) > 4000:          # the missing-middle shape: the second half of the text does not match this reference
            t = np.concatenate([t[:len(t) // 2], rng.integers(0, 4, int(rng.integers(500, 9000))).astype(np.uint8)])
        z = orc.LZ(ref, mml)
        for prefix in (0, 1):
            r, got = costs(lzc, t, ref, mml, z, prefix, int(rng.random() < 0.4), int(rng.integers(0, 70)))
            if r <= -10: n_seq += 1
            else: assert np.array_equal(got, z.cost_vector(t, prefix)), f"seed {s} prefix {prefix}: cost vector differs from the oracle"
    assert n_seq <= 24
