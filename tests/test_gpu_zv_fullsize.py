"""BASELINE.json configurations at FULL size on the device, through the CLI of this repo (agc-b200 create) against the reference
binary (oracle/_ref/agc create) on the same files: the archives must be byte-identical (cmp).
  configs[1]  C2: 1000 x 30 kb viral genomes, k=25
  configs[2]  C3: 63 x (5 Mb + 120 kb novel contig), -a, k=29
  configs[3]  C4 in shape: one 250 Mb contig with repeats + one sample (0.1 % SNP, an indel every 10 kb): 250 Mb contigs, a 250 M
              k-mer sort in splitter determination, ~4200 segments per contig, u32 positions in one contig
AGC_FULLSIZE_C4_LEN shrinks the C4 contig (bases) for quick runs."""
import hashlib
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF_AGC = os.path.join(ROOT, "oracle", "_ref", "agc")
OUR_AGC = os.path.join(ROOT, "agc_b200", "bin", "agc-b200")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF_AGC), reason="reference binary not built (make -f oracle/Makefile.ref)")]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def _both(d, files, flags):
    lst = os.path.join(d, "list.txt")
    open(lst, "w").write("\n".join(files[1:]) + "\n")
    t0 = time.time()
    subprocess.check_call([OUR_AGC, "create"] + flags + ["-o", os.path.join(d, "our.agc"), "-i", lst, files[0]], stdout=subprocess.DEVNULL)
    t1 = time.time()
    subprocess.check_call([REF_AGC, "create"] + flags + ["-t", str(os.cpu_count() or 1), "-o", os.path.join(d, "ref.agc"), "-i", lst, files[0]],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t2 = time.time()
    print(f"ours {t1 - t0:.2f} s (incl. process start + CUDA context), reference {t2 - t1:.2f} s")
    return _sha(os.path.join(d, "our.agc")), _sha(os.path.join(d, "ref.agc"))


def _tmp(name):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    import tempfile
    return tempfile.mkdtemp(prefix=f"agc_{name}_", dir=base)


def test_c2_full_size():
    import gen_data, shutil
    d = _tmp("c2")
    try:
        files, _ = gen_data.viral(os.path.join(d, "data"), n_samples=1000, ref_len=30000, p=0.01, seed=1)
        a, b = _both(d, files, ["-k", "25"])
        assert a == b, "C2: archive differs from the reference's"
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_c3_full_size():
    import gen_data, shutil
    d = _tmp("c3")
    try:
        files = gen_data.bacterial_adaptive(os.path.join(d, "data"), seed=2, n_samples=63, ref_len=5_000_000)
        a, b = _both(d, files, ["-a", "-k", "29"])
        assert a == b, "C3: archive differs from the reference's"
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_c4_shape_250mb_contig():
    import gen_data, shutil
    d = _tmp("c4")
    try:
        n = int(os.environ.get("AGC_FULLSIZE_C4_LEN", 250_000_000))
        files = gen_data.human_chromosome(os.path.join(d, "data"), seed=3, n_samples=1, ctg_len=n)
        a, b = _both(d, files, [])
        assert a == b, "C4 shape: archive differs from the reference's"
    finally:
        shutil.rmtree(d, ignore_errors=True)
