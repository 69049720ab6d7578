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
    if case == "fallback":          # -f, segment_size <= 10000: candidates are decided by shared-minimizer counts (1908-1913)
        return gen_data.fallback_collection(d, seed=11, seg=2000), ["-f", "0.05", "-k", "21", "-s", "2000", "-b", "4"]
    if case == "fallback_estimates":    # -f, long segments: candidates are decided by CSegment::estimate (1919)
        return gen_data.fallback_collection(d, seed=12, seg=12000, n_samples=6, ref_len=200000), ["-f", "0.02", "-k", "25", "-s", "12000"]
    if case == "fallback_adaptive":
        return gen_data.fallback_collection(d, seed=13, seg=2000), ["-f", "0.05", "-a", "-k", "21", "-s", "2000", "-b", "4"]
    raise KeyError(case)


ALL_CASES = ["viral", "complex", "complex_n", "tiny", "smallpacks", "adaptive", "adaptive_big_segments", "adaptive_complex",
             "concatenated", "concatenated_full_units", "concatenated_adaptive", "fallback", "fallback_estimates", "fallback_adaptive"]


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
    for mode in ("-a", "-f"):       # the case must really exercise the mode: without the flag the reference writes a different archive
        if mode in flags:
            plain = os.path.join(tmp, "plain.agc")
            i = flags.index(mode)
            rest = flags[:i] + flags[i + (2 if mode == "-f" else 1):]
            subprocess.check_call([REF_AGC, "create", "-t", "4", "-o", plain] + rest + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            assert open(plain, "rb").read() != b


def append_flags(flags):
    """`append` takes the archive's own k / l / s / b"""
    out, i = [], 0
    while i < len(flags):
        if flags[i] in ("-k", "-s", "-l", "-b"):
            i += 2
            continue
        out.append(flags[i]); i += 1
    return out


APPEND_CASES = [("viral", 12, 1), ("complex", 5, 1), ("complex", 3, 2), ("complex_n", 7, 1), ("smallpacks", 10, 2), ("smallpacks", 7, 1),
                ("concatenated", 2, 1), ("fallback", 4, 2), ("tiny", 2, 1),
                # append -a: the reference sample is decoded from the archive to rebuild the reference k-mer list (828-847)
                ("adaptive", 4, 1), ("adaptive", 2, 2), ("adaptive_big_segments", 3, 1), ("adaptive_complex", 5, 1), ("fallback_adaptive", 4, 1),
                ("concatenated_adaptive", 2, 1)]


def run_append_case(tmp, agc, case, n_first, steps):
    """reference: create(first files) then append the rest in `steps` runs; `agc` extends the same base; -> (ours, reference) bytes"""
    files, flags = collection(case, tmp)
    base = os.path.join(tmp, "base.agc")
    subprocess.check_call([REF_AGC, "create", "-t", "3", "-o", base] + flags + files[:n_first], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    rest = files[n_first:]
    chunks = [rest] if steps == 1 else [rest[:len(rest) // 2], rest[len(rest) // 2:]]
    rb = ob = base
    for i, ch in enumerate(chunks):
        r2 = os.path.join(tmp, f"ref{i}.agc"); o2 = os.path.join(tmp, f"our{i}.agc")
        subprocess.check_call([REF_AGC, "append", "-t", "3", "-o", r2] + append_flags(flags) + [rb] + ch, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.check_call([agc, "append", "-o", o2] + append_flags(flags) + [ob] + ch)
        rb, ob = r2, o2
    return open(ob, "rb").read(), open(rb, "rb").read(), files


@pytest.mark.parametrize("case,n_first,steps", APPEND_CASES)
def test_append_matches_reference(tmp_path, mock_agc, case, n_first, steps):
    """CAGCCompressor::Append: the archive state is reloaded (zstd frames decoded, references handed back to the device, last
    packs and the last contig batch unpacked) and the extended archive is byte-identical to the reference's -- including the
    reference's behaviour that a reloaded group estimates to 0 until something is added to it (segment.cpp:84-86)"""
    a, b, files = run_append_case(str(tmp_path), mock_agc, case, n_first, steps)
    assert a == b, f"appended archives differ: {len(a)} vs {len(b)} bytes"


def test_append_survives_damaged_archives(tmp_path, mock_agc):
    """bit flips, truncation and footer damage of the input archive: `append` fails with a message (or succeeds when the damage
    sits in a part that is copied verbatim), it never crashes"""
    tmp = str(tmp_path)
    files, flags = collection("complex", tmp)
    base = os.path.join(tmp, "base.agc")
    subprocess.check_call([REF_AGC, "create", "-t", "3", "-o", base] + flags + files[:5], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    good = open(base, "rb").read()
    rng = np.random.default_rng(1)
    failed = 0
    for it in range(45):
        b = bytearray(good)
        if it % 3 == 0:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        elif it % 3 == 1:
            b = b[:int(rng.integers(0, len(b)))]
        else:
            b[int(rng.integers(len(b) - 600, len(b)))] = int(rng.integers(0, 256))
        bad = os.path.join(tmp, "bad.agc")
        open(bad, "wb").write(bytes(b))
        r = subprocess.run([mock_agc, "append", "-o", os.path.join(tmp, "o.agc"), bad] + files[5:7], capture_output=True, timeout=120)
        assert 0 <= r.returncode < 128, f"append crashed (exit {r.returncode}) on damaged archive, iteration {it}"
        failed += r.returncode != 0
    assert failed >= 15


def test_golden_archives(tmp_path, mock_agc):
    """the committed hashes of the reference's archives (tools/make_golden_modes.py): create in every mode, and the appended
    archives -- this check does not execute the reference binary for the create cases"""
    import hashlib
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "archives_modes.json")))
    for case, g in gold["create"].items():
        tmp = os.path.join(str(tmp_path), case); os.makedirs(tmp)
        files, flags = collection(case, tmp)
        assert hashlib.sha256(b"".join(open(f, "rb").read() for f in files)).hexdigest() == g["fasta_sha256"], f"{case}: generator drifted"
        out = os.path.join(tmp, "o.agc")
        subprocess.check_call([mock_agc, "create", "-o", out] + flags + files)
        b = open(out, "rb").read()
        assert len(b) == g["agc_size"] and hashlib.sha256(b).hexdigest() == g["agc_sha256"], f"{case}: archive differs from the reference's golden hash"
    for (case, n_first, steps) in APPEND_CASES:
        g = gold["append"][f"{case}:{n_first}:{steps}"]
        tmp = os.path.join(str(tmp_path), f"app_{case}_{n_first}_{steps}"); os.makedirs(tmp)
        a, _, _ = run_append_case(tmp, mock_agc, case, n_first, steps)
        assert len(a) == g["agc_size"] and hashlib.sha256(a).hexdigest() == g["agc_sha256"], f"append {case}:{n_first}:{steps} differs from the golden hash"


def test_self_check_mode(tmp_path, mock_agc):
    """--verify: every coded frame is decoded again (device decoder; here its host build) and compared before it is written;
    the archive is the same one"""
    tmp = str(tmp_path)
    files, flags = collection("complex", tmp)
    a = os.path.join(tmp, "a.agc"); b = os.path.join(tmp, "b.agc")
    subprocess.check_call([mock_agc, "create", "-o", a] + flags + files)
    subprocess.check_call([mock_agc, "create", "--verify", "-o", b] + flags + files)
    assert open(a, "rb").read() == open(b, "rb").read()


def test_gzipped_inputs(tmp_path, mock_agc):
    """.fa.gz inputs (CGenomeIO through zlib; sample names lose .gz and .fa, application.cpp:606-630): same archive"""
    import gzip
    import shutil
    files, flags = collection("smallpacks", str(tmp_path))
    gz = []
    for f in files:
        with open(f, "rb") as i, gzip.open(f + ".gz", "wb") as o:
            shutil.copyfileobj(i, o)
        gz.append(f + ".gz")
    a = subprocess.run([mock_agc, "create"] + flags + gz, capture_output=True).stdout
    b = subprocess.run([REF_AGC, "create", "-t", "2"] + flags + gz, capture_output=True).stdout
    assert a == b and len(a) > 1000


def test_archive_to_stdout(tmp_path, mock_agc):
    """no -o: the archive goes to stdout (COutFile::Open with an empty name, src/common/io.h:281-300)"""
    files, flags = collection("smallpacks", str(tmp_path))
    a = subprocess.run([mock_agc, "create"] + flags + files, capture_output=True).stdout
    b = subprocess.run([REF_AGC, "create", "-t", "2"] + flags + files, capture_output=True).stdout
    assert a == b and len(a) > 1000


def test_cli_clamps_options_like_the_reference(tmp_path, mock_agc):
    """b_value<T>::assign (src/app/application.h:23-47): -f 0.2 means -f 0.05, -k 40 means -k 32, ..."""
    tmp = str(tmp_path)
    files, _ = collection("fallback", tmp)
    outs = []
    for flags in (["-f", "0.2", "-k", "40", "-s", "50", "-l", "3"], ["-f", "0.05", "-k", "32", "-s", "100", "-l", "15"]):
        for exe, name in ((REF_AGC, "ref"), (mock_agc, "our")):
            out = os.path.join(tmp, f"{name}{len(outs)}.agc")
            subprocess.check_call([exe, "create", "-o", out] + flags + files[:3], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            outs.append(open(out, "rb").read())
    assert outs[0] == outs[1] == outs[2] == outs[3]


def test_fasta_layout_variants(mock_agc):
    """the in-place record cutter of the ingest path (CAGCCompressor::add_sample_files_arena) against CGenomeIO::ReadContigRaw's rules:
    CRLF, lower case, N / IUPAC / junk symbols, tabs and spaces in headers, no final newline, blank lines, empty records, gzipped
    members -- archives identical to the reference binary's (tools/fuzz_fasta_layout.py)"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_fasta_layout
    n, bad = fuzz_fasta_layout.run(30, seed=5, our=mock_agc, verbose=False)
    assert n == 30 and bad == 0
