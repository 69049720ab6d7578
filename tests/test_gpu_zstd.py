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


def test_zstd_unsupported_is_loud(dev_factory):
    import agc_b200
    dev = dev_factory(k=21, min_match_len=20)
    with pytest.raises(agc_b200.AgcGpuError):
        dev.zstd_compress([bytes(300000)], [13])     # level 13 above 256 KB = btlazy2: refused, not approximated
    with pytest.raises(agc_b200.AgcGpuError):
        dev.zstd_compress([b"abc"], [3])
