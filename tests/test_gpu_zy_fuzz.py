"""Differential run of random small collections and random `create` flags (k, l, s, b, -a, -c, -f) through agc-b200 on the
device against the reference binary: whole archives must be byte-identical.  The same generator (tools/fuzz_host_pipeline.py)
drives the CPU suite's mocked-device run; the seeds below cover every flag combination and were all green there."""
import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fuzz_host_pipeline as fz

pytestmark = pytest.mark.gpu
OUR_AGC = os.path.join(ROOT, "agc_b200", "bin", "agc-b200")
SEEDS = list(range(9000, 9024))


@pytest.mark.skipif(not os.path.exists(fz.REF), reason="reference binary not built (make -f oracle/Makefile.ref)")
def test_random_collections_match_reference(tmp_path):
    bad = []
    combos = set()
    for seed in SEEDS:
        d = os.path.join(str(tmp_path), str(seed))
        files, flags = fz.make_case(os.path.join(d, "d"), seed)
        combos.add(tuple(f for f in flags if f in ("-a", "-c", "-f")))
        ref = os.path.join(d, "ref.agc"); our = os.path.join(d, "our.agc")
        subprocess.check_call([fz.REF, "create", "-t", "3", "-o", ref] + flags + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        r = subprocess.run([OUR_AGC, "create", "-o", our] + flags + files, capture_output=True)
        if r.returncode != 0 or open(ref, "rb").read() != open(our, "rb").read():
            bad.append((seed, " ".join(flags), r.returncode, r.stderr.decode()[-200:]))
    assert not bad, bad
    assert len(combos) >= 6, combos          # the seed range really mixes the modes
