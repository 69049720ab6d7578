"""The decoding side on the device and what is built on it: agcgpu_lz_decode_batch, agcgpu_zstd_decompress_batch, `create --verify`
and `append`.  (Kept in a file of its own, after the parity / pipeline / residual-coder suites: these entry points were written
after the round's GPU budget was spent -- their logic is checked on the CPU suite against the mocked device ABI and the host
build of the decoder, their first run on hardware is the next round's first GPU call.)"""
import os
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import agc_parts
from test_gpu_parity import _mk_pairs, _run_pairs
from test_gpu_zstd import _gen
from test_host_pipeline import collection, run_append_case

pytestmark = pytest.mark.gpu
REF_AGC = os.path.join(ROOT, "oracle", "_ref", "agc")
OUR_AGC = os.path.join(ROOT, "agc_b200", "bin", "agc-b200")


def test_lz_decode_roundtrip(dev_factory):
    """agcgpu_lz_decode_batch (CLZDiff_V2::Decode): clean, reverse-complemented and non-ACGT pairs"""
    rng = np.random.default_rng(23)
    for mml, dirty, use_rc in ((20, False, False), (15, False, True), (24, True, False)):
        dev = dev_factory(k=31, min_match_len=mml, segment_size=60000)
        _run_pairs(dev, rng, _mk_pairs(rng, 24, dirty=dirty), mml, use_rc, check_decode=True)


def test_zstd_decode_on_device(dev_factory):
    """agcgpu_zstd_decompress_batch: frames of the reference's libzstd (several levels) and of the device coder decode to
    their inputs in one batch; a truncated frame fails the call"""
    import agc_b200
    rng = np.random.default_rng(41)
    dev = dev_factory(k=31)
    raws, frames = [], []
    for level in (1, 5, 13, 17, 19):
        for kind in range(7):
            for n in (0, 1, 200, 4000, 40000, 200000):
                if n > 4000 and (kind + level) % 3:
                    continue
                raw = _gen(rng, kind, n)
                raws.append(raw); frames.append(agc_parts.zstd_compress(raw, level))
    assert dev.zstd_decompress(frames) == raws
    mine = [_gen(rng, k, n) for k in range(7) for n in (300, 30000, 150000)]
    coded = dev.zstd_compress(mine, [17] * len(mine))
    assert dev.zstd_decompress(coded) == mine
    with pytest.raises(agc_b200.AgcGpuError):                   # a truncated frame fails the call
        dev.zstd_decompress([frames[0], frames[-1][:-5]])


def test_create_with_self_check(tmp_path):
    """--verify: every frame and every LZ delta is decoded again on the device and compared; same archive"""
    tmp = str(tmp_path)
    files, flags = collection("complex", tmp)
    a = os.path.join(tmp, "a.agc"); b = os.path.join(tmp, "b.agc")
    subprocess.check_call([OUR_AGC, "create", "-o", a] + flags + files)
    subprocess.check_call([OUR_AGC, "create", "--verify", "-o", b] + flags + files)
    assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.skipif(not os.path.exists(REF_AGC), reason="reference binary not built (make -f oracle/Makefile.ref)")
@pytest.mark.parametrize("case,n_first,steps", [("viral", 12, 1), ("complex", 3, 2), ("complex_n", 7, 1), ("fallback", 4, 2), ("concatenated", 2, 1), ("adaptive", 4, 1), ("fallback_adaptive", 4, 1)])
def test_append_matches_reference(tmp_path, case, n_first, steps):
    """`agc-b200 append` on the device (frames decoded by k_zstd_decode, references re-indexed, packs continued) vs the reference's"""
    a, b, files = run_append_case(str(tmp_path), OUR_AGC, case, n_first, steps)
    assert a == b, f"appended archives differ: {len(a)} vs {len(b)} bytes"
