"""Host side of the path on the CPU: agc_b200/csrc/host (CAGCCompressor / CCollection_V3 / CArchive mirrors, unchanged)
linked against tests/mock/libagcgpu_mock.so -- the device entry points of include/agcgpu.h restated over the C oracle
and the host build of the residual coder (TEST INFRASTRUCTURE, see tests/mock/agcgpu_mock.cpp) -- must write archives
that are byte-identical to the reference binary's (oracle/_ref/agc).  This is what checks add_segment's rare branches,
registration order, pack bookkeeping, the -a (adaptive) flow and the container without a GPU; tests/test_gpu_pipeline.py
repeats the same collections on the device."""
import os
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_data

REF_AGC = os.path.join(ROOT, "oracle", "_ref", "agc")
MOCK_DIR = os.path.join(ROOT, "tests", "mock")
MOCK_AGC = os.path.join(MOCK_DIR, "agc-mock")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_AGC), reason="reference binary not built (make -f oracle/Makefile.ref)")


def collection(case, tmp):
    """(files, flags) of a named collection; shared with tests/test_gpu_pipeline.py"""
    d = os.path.join(tmp, "d")
    if case == "viral":
        return gen_data.viral(d, n_samples=40, ref_len=30000, p=0.01, seed=1)[0], ["-k", "25"]
    if case == "complex":
        return gen_data.complex_collection(d, seed=5), ["-k", "21", "-s", "2000", "-b", "5"]
    if case == "complex_n":
        return gen_data.complex_collection(d, seed=6, with_n=True), ["-k", "31", "-s", "3000", "-l", "18", "-b", "4"]
    if case == "tiny":
        rng = np.random.default_rng(3)
        os.makedirs(d)
        files = []
        for i, nm in enumerate(["ref", "a", "b", "c"]):
            fn = os.path.join(d, nm + ".fa")
            gen_data.write_fasta(fn, [(f"{nm}{j}", rng.integers(0, 4, int(rng.integers(5, 28)), dtype=np.uint8)) for j in range(1 + i)])
            files.append(fn)
        return files, ["-k", "29", "-l", "22"]
    if case == "smallpacks":
        return gen_data.viral(d, n_samples=25, ref_len=9000, p=0.02, seed=9)[0], ["-k", "17", "-s", "1000", "-b", "3", "-l", "15"]
    if case == "adaptive":          # BASELINE configs[2] in small: novel contigs force the new-splitter path
        return gen_data.adaptive_collection(d, seed=2), ["-a", "-k", "21", "-s", "2000", "-b", "3"]
    if case == "adaptive_big_segments":
        return gen_data.adaptive_collection(d, seed=7, n_samples=6, ref_len=120000, n_ctg=3, novel_len=30000), ["-a", "-k", "29", "-s", "10000"]
    if case == "adaptive_complex":  # -a on a collection whose novel contigs are mostly shorter than segment_size
        return gen_data.complex_collection(d, seed=5), ["-a", "-k", "21", "-s", "2000", "-b", "5"]
    if case == "concatenated":      # -c: 3 + 23 contigs, units of 5, a partial last unit
        return gen_data.concatenated_collection(d, seed=1, n_ctg=23, per_file=9), ["-c", "-k", "21", "-s", "2000", "-b", "5"]
    if case == "concatenated_full_units":   # 3 + 17 = 20 contigs = 5 complete units: the trailing token registers nothing (1142-1156)
        return gen_data.concatenated_collection(d, seed=2, n_ctg=17, per_file=6), ["-c", "-k", "21", "-s", "2000", "-b", "4"]
    if case == "concatenated_adaptive":
        return gen_data.concatenated_collection(d, seed=3, n_ctg=22, per_file=8), ["-c", "-a", "-k", "21", "-s", "2000", "-b", "6"]
    raise KeyError(case)


ALL_CASES = ["viral", "complex", "complex_n", "tiny", "smallpacks", "adaptive", "adaptive_big_segments", "adaptive_complex",
             "concatenated", "concatenated_full_units", "concatenated_adaptive"]


@pytest.fixture(scope="module")
def mock_agc():
    subprocess.check_call(["make", "-C", MOCK_DIR, "-j4"], stdout=subprocess.DEVNULL)
    return MOCK_AGC


@pytest.mark.parametrize("case", ALL_CASES)
def test_host_pipeline_archives_match_reference(tmp_path, mock_agc, case):
    tmp = str(tmp_path)
    files, flags = collection(case, tmp)
    ref_out = os.path.join(tmp, "ref.agc"); our_out = os.path.join(tmp, "our.agc")
    subprocess.check_call([REF_AGC, "create", "-t", "4", "-o", ref_out] + flags + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call([mock_agc, "create", "-o", our_out] + flags + files)
    a = open(our_out, "rb").read(); b = open(ref_out, "rb").read()
    assert a == b, f"archives differ: {len(a)} vs {len(b)} bytes, first diff at {next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), -1)}"
    if "-a" in flags:               # the case must really exercise the adaptive path: the non-adaptive archive is a different one
        plain = os.path.join(tmp, "plain.agc")
        subprocess.check_call([REF_AGC, "create", "-t", "4", "-o", plain] + [f for f in flags if f != "-a"] + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert open(plain, "rb").read() != b


def test_refused_modes_fail_loudly(tmp_path, mock_agc):
    files, flags = collection("tiny", str(tmp_path))
    for extra in (["-f", "0.1"],):
        r = subprocess.run([mock_agc, "create", "-o", os.path.join(str(tmp_path), "x.agc")] + extra + flags + files, capture_output=True)
        assert r.returncode != 0 and b"not implemented" in r.stderr
