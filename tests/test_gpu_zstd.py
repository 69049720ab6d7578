"""Device residual coder vs the reference's own libzstd (oracle/_ref/libzstd_ref.so = vendored zstd "1.5.5"):
frames must be byte-identical for every level AGC uses and every size class of clevels.h."""
import os
import sys
import numpy as np
import pytest
import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import agc_parts

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libzstd_ref.so")),
                                                  reason="libzstd_ref.so not built")]


def _gen(rng, kind, n):
    if kind == 0: return bytes(rng.integers(0, 256, n, dtype=np.uint8))
    if kind == 1: return bytes(rng.integers(0, 4, n, dtype=np.uint8))
    if kind == 2:
        ref = rng.integers(0, 4, 3000).astype(np.uint8); z = orc.LZ(ref, 20); parts = []
        while sum(map(len, parts)) < n:
            t = ref.copy(); m = rng.random(len(t)) < 0.01; t[m] = (t[m] + 1) % 4
            parts.append(z.encode(t) + b"\xff")
        return b"".join(parts)[:n]
    if kind == 3: return bytes(np.repeat(rng.integers(0, 256, max(1, n // 50), dtype=np.uint8), 50)[:n])
    if kind == 4: return (b"chr1 some description\x00" * (n // 22 + 1))[:n]
    if kind == 5: return bytes([7]) * n
    a = rng.integers(0, 256, n // 3 + 1, dtype=np.uint8)
    return bytes(np.concatenate([a, a, a])[:n])


def _gen_adv(rng, kind, n):
    """inputs aimed at the match finder's window engine: deep bucket chains, positions the reference skips, walks that run
    out of compares or reach the end of a block, near-copies across block boundaries"""
    if kind == 0:                                            # long runs of few symbols (period 1, very deep chains)
        out = bytearray()
        while len(out) < n:
            out += bytes([int(rng.integers(33, 36))]) * int(rng.integers(1, 400))
        return bytes(out[:n])
    if kind == 1:                                            # short periods 2..9 with sparse errors
        out = bytearray()
        while len(out) < n:
            per = bytes(rng.integers(48, 58, int(rng.integers(2, 10)), dtype=np.uint8))
            seg = bytearray(per * int(rng.integers(3, 120)))
            for _ in range(len(seg) // 200):
                seg[int(rng.integers(0, len(seg)))] = 65
            out += seg
        return bytes(out[:n])
    if kind == 2:                                            # near-copies of one 20 kb sequence over 4 symbols (raw-group packs)
        ref = rng.integers(0, 4, 20000).astype(np.uint8); parts = []
        while sum(map(len, parts)) < n:
            t = ref.copy(); m = rng.random(len(t)) < 0.02; t[m] = (t[m] + rng.integers(1, 4, int(m.sum()))) % 4
            parts.append(bytes(t) + b"\xff")
        return b"".join(parts)[:n]
    if kind == 3:                                            # exact repeats of a block much longer than ZSTD_OPT_NUM
        blk = bytes(rng.integers(0, 256, 9000, dtype=np.uint8))
        return (blk * (n // 9000 + 1))[:n]
    if kind == 4:                                            # two-symbol text: every bucket is deep, walks hit the compare limit
        return bytes(rng.integers(0, 2, n, dtype=np.uint8) + 97)
    # delta-pack-like text with long '!' runs
    out = bytearray()
    while len(out) < n:
        out += b"!" * int(rng.integers(0, 60)) + str(int(rng.integers(0, 3000))).encode() + b"," + str(int(rng.integers(0, 900))).encode() + b"." + bytes([int(rng.integers(65, 69))])
    return bytes(out[:n])


def test_zstd_window_engine_cases(dev_factory):
    rng = np.random.default_rng(8)
    dev = dev_factory(k=21, min_match_len=20)
    inputs, levels = [], []
    for kind in range(6):
        for n, level in ((3000, 17), (16000, 13), (16384, 19), (40000, 17), (100000, 17), (131080, 19), (200000, 18), (300000, 17)):
            inputs.append(_gen_adv(rng, kind, n)); levels.append(level)
    got = dev.zstd_compress(inputs, levels)
    for i, (raw, lv) in enumerate(zip(inputs, levels)):
        assert got[i] == agc_parts.zstd_compress(raw, lv), f"frame {i}: kind {i // 8}, {len(raw)} bytes, level {lv}"


def test_zstd_frames_match_reference(dev_factory):
    rng = np.random.default_rng(5)
    dev = dev_factory(k=21, min_match_len=20)
    inputs, levels = [], []
    for n in (0, 1, 6, 7, 8, 9, 63, 64, 300, 4000, 16384, 16385, 40000, 65536, 65537, 131072, 131073, 150000):
        for kind in range(7):
            for level in (13, 17, 18, 19):
                if n > 40000 and (kind + level) % 3:       # keep the slow big cases to a sample
                    continue
                inputs.append(_gen(rng, kind, n)); levels.append(level)
    got = dev.zstd_compress(inputs, levels)
    for i, (raw, lv) in enumerate(zip(inputs, levels)):
        assert got[i] == agc_parts.zstd_compress(raw, lv), f"frame {i}: {len(raw)} bytes, level {lv}"
        assert agc_parts.zstd_decompress(got[i]) == raw


def test_zstd_large_multiblock(dev_factory):
    rng = np.random.default_rng(6)
    dev = dev_factory(k=21, min_match_len=20)
    inputs = [_gen(rng, 2, 300000), _gen(rng, 1, 262145), _gen(rng, 2, 1400000)]
    levels = [17, 19, 17]
    got = dev.zstd_compress(inputs, levels)
    for raw, lv, g in zip(inputs, levels, got):
        assert g == agc_parts.zstd_compress(raw, lv)


def test_zstd_btlazy2_class(dev_factory):
    """level 13 above 256 KB selects ZSTD_btlazy2 (tuple-packed references of very long segments)"""
    rng = np.random.default_rng(7)
    dev = dev_factory(k=21, min_match_len=20)
    inputs = [_gen(rng, 2, 262145), _gen(rng, 1, 300000), _gen_adv(rng, 2, 400000), _gen_adv(rng, 0, 300000)]
    got = dev.zstd_compress(inputs, [13] * len(inputs))
    for raw, g in zip(inputs, got):
        assert g == agc_parts.zstd_compress(raw, 13)


def test_zstd_unsupported_is_loud(dev_factory):
    import agc_b200
    dev = dev_factory(k=21, min_match_len=20)
    with pytest.raises(agc_b200.AgcGpuError):
        dev.zstd_compress([b"abc"], [3])             # only the levels AGC uses (13 / 17 / 18 / 19) exist: refused, not approximated


def test_zstd_async_batches(dev_factory):
    """agcgpu_zstd_submit / agcgpu_zstd_collect: several batches in flight (wide and narrow frames, an empty input, a batch of one),
    other device work in between, frames back in submission order and identical to libzstd's"""
    rng = np.random.default_rng(9)
    dev = dev_factory(k=21, min_match_len=20)
    batches = [([_gen(rng, 2, 50000), _gen_adv(rng, 5, 70000), b"", _gen(rng, 1, 900)], [17, 17, 19, 13]),
               ([_gen(rng, 3, 16000)], [18]),
               ([_gen_adv(rng, k, 20000) for k in range(6)], [17, 13, 19, 18, 17, 17])]
    assert dev.zstd_collect() == []
    for inputs, levels in batches:
        dev.zstd_submit(inputs, levels)
        dev.zstd_compress([b"between the batches" * 10], [19])           # the synchronous call while batches are in flight
    got = dev.zstd_collect()
    exp = [agc_parts.zstd_compress(x, lv) for inputs, levels in batches for x, lv in zip(inputs, levels)]
    assert got == exp
    dev.zstd_submit([b"abc" * 100], [17])                                 # a second round on the same context
    assert dev.zstd_collect() == [agc_parts.zstd_compress(b"abc" * 100, 17)]


def test_host_alloc_roundtrip():
    import ctypes as C
    import agc_b200
    L = agc_b200.lib()
    cap = C.c_uint64(0)
    p = L.agcgpu_host_alloc(1 << 20, C.byref(cap))
    assert p and cap.value >= (1 << 20)
    C.memset(p, 0x5A, 1 << 20)
    L.agcgpu_host_free(p, cap.value)
    cap2 = C.c_uint64(0)
    q = L.agcgpu_host_alloc(1 << 19, C.byref(cap2))                      # comes back from the pool
    assert q and cap2.value >= (1 << 19)
    L.agcgpu_host_free(q, cap2.value)
