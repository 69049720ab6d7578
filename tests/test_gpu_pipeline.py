"""Whole-pipeline parity: agc-b200 create (GPU path) vs the reference binary (oracle/_ref/agc) on the same FASTA files.
While the device residual coder is incomplete the comparison is made on the pre-zstd content of every part
(--dump-parts vs the reference archive decoded with the reference's own libzstd); once agcgpu_zstd_compress_batch
exists the archives are compared byte for byte."""
import os
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_data
import agc_parts
from test_host_pipeline import collection, ALL_CASES

pytestmark = pytest.mark.gpu

REF_AGC = os.path.join(ROOT, "oracle", "_ref", "agc")
OUR_AGC = os.path.join(ROOT, "agc_b200", "bin", "agc-b200")


def _run_both(tmp, files, flags):
    ref_out = os.path.join(tmp, "ref.agc"); our_out = os.path.join(tmp, "our.agc"); dump = os.path.join(tmp, "our.dump")
    subprocess.check_call([REF_AGC, "create", "-t", "4", "-o", ref_out] + flags + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call([OUR_AGC, "create", "-o", our_out, "--dump-parts", dump] + flags + files)
    return agc_parts.compare_dump_to_archive(dump, ref_out)


@pytest.mark.skipif(not os.path.exists(REF_AGC), reason="reference binary not built (make -f oracle/Makefile.ref)")
@pytest.mark.parametrize("case", ALL_CASES)
def test_parts_match_reference(tmp_path, case):
    tmp = str(tmp_path)
    files, flags = collection(case, tmp)          # the same collections the CPU suite runs through the mocked device ABI
    bad = _run_both(tmp, files, flags)
    assert not bad, "\n".join(bad[:10])
    # and the real thing: the archive written through the device residual coder is byte-identical to the reference's
    our = os.path.join(tmp, "our_full.agc")
    subprocess.check_call([OUR_AGC, "create", "-o", our] + flags + files)
    a = open(our, "rb").read(); b = open(os.path.join(tmp, "ref.agc"), "rb").read()
    assert len(a) == len(b) and a == b, f"archives differ: {len(a)} vs {len(b)} bytes, first diff at {next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), -1)}"
    # the reference decompressor accepts it and returns the input
    last = open(files[-1], "rb").read()
    if "-c" in flags:                # every contig is a sample named after the contig: fetch the last one
        last = last[last.rindex(b">"):]
        sample = last[1:last.index(b"\n")].split()[0].decode()
    else:
        sample = os.path.splitext(os.path.basename(files[-1]))[0]
    out = subprocess.run([REF_AGC, "getset", our, sample], capture_output=True).stdout
    assert out == last


@pytest.mark.skipif(not os.path.exists(REF_AGC), reason="reference binary not built (make -f oracle/Makefile.ref)")
@pytest.mark.parametrize("case", ["complex", "adaptive"])
def test_async_coder_every_flush(tmp_path, case):
    """AGCGPU_ZSTD_ASYNC_MIN=1: every flush point submits its parts to the residual coder at once (many small batches in flight
    while the next samples are processed) -- the archive is still the reference's"""
    tmp = str(tmp_path)
    files, flags = collection(case, tmp)
    ref_out = os.path.join(tmp, "ref.agc"); our = os.path.join(tmp, "our.agc")
    subprocess.check_call([REF_AGC, "create", "-t", "4", "-o", ref_out] + flags + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call([OUR_AGC, "create", "-o", our] + flags + files, env=dict(os.environ, AGCGPU_ZSTD_ASYNC_MIN="1"))
    assert open(our, "rb").read() == open(ref_out, "rb").read()
