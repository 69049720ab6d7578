"""CPU tests of the oracle (oracle/agc_oracle.c): against the committed golden vectors generated from the unmodified
reference (tools/make_golden.py), and -- when oracle/_ref is built -- against the reference itself on fresh random inputs."""
import json
import os
import hashlib
import sys
import numpy as np
import pytest
import orc
from conftest import mutate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_lz_golden_vectors():
    cases = json.load(open(os.path.join(GOLD, "lz_golden.json")))
    assert len(cases) >= 40
    for c in cases:
        ref = np.frombuffer(bytes.fromhex(c["ref"]), np.uint8)
        text = np.frombuffer(bytes.fromhex(c["text"]), np.uint8)
        z = orc.LZ(ref, c["mml"])
        assert z.encode(text).hex() == c["encode"]
        assert z.estimate(text) == c["estimate"]
        assert z.estimate(text, c["bound"]) == c["estimate_bounded"]
        assert z.cost_vector(text, 1).tolist() == c["cost_prefix"]
        assert z.cost_vector(text, 0).tolist() == c["cost_suffix"]


def test_reference_quirks():
    """a planted exact run of min_match_len bases is not a match, min_match_len+1 is; Estimate != Encode size"""
    rng = np.random.default_rng(1)
    ref = rng.integers(0, 4, 4000).astype(np.uint8)
    for run, expect_match in ((20, False), (21, True), (24, True)):
        text = rng.integers(0, 4, 300).astype(np.uint8)
        text[100:100 + run] = ref[2000:2000 + run]            # ref position 2000 is hashed (multiple of 4)
        text[99] = (ref[1999] + 1) % 4; text[100 + run] = (ref[2000 + run] + 1) % 4
        enc = orc.LZ(ref, 20).encode(text)
        assert (b"," in enc) == expect_match, (run, enc)
    text = ref.copy(); text[::53] = (text[::53] + 1) % 4
    z = orc.LZ(ref, 20)
    assert z.estimate(text) != len(z.encode(text))


def test_equal_sequence_is_empty():
    ref = np.random.default_rng(2).integers(0, 4, 1000).astype(np.uint8)
    z = orc.LZ(ref, 20)
    assert z.encode(ref) == b"" and z.estimate(ref) == 0


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
def test_lz_fuzz_against_reference():
    rng = np.random.default_rng(77)
    for it in range(120):
        m = int(rng.choice([0, 5, 30, 100, 1000, 5000]))
        mml = int(rng.choice([15, 18, 20, 22, 32]))
        ref = rng.integers(0, 4, m).astype(np.uint8)
        if it % 5 == 0:
            ref = mutate(rng, ref, 0, 0, 3)
        text = mutate(rng, ref, float(rng.choice([0, 0.001, 0.01, 0.1])), int(rng.integers(0, 4)), int(rng.integers(0, 3)) if it % 3 == 0 else 0)
        z = orc.LZ(ref, mml)
        assert z.encode(text) == orc.ref_encode(ref, text, mml)
        assert z.estimate(text) == orc.ref_estimate(ref, text, mml)
        b = int(rng.integers(0, 50))
        assert z.estimate(text, b) == orc.ref_estimate(ref, text, mml, b)
        for pf in (0, 1):
            assert np.array_equal(z.cost_vector(text, pf), orc.ref_cost_vector(ref, text, mml, pf))


def test_preprocess_table():
    raw = bytes(range(128)) * 2 + b"ACGTacgtNnRYKM\n\r>@"
    out = orc.preprocess(raw)
    exp = [orc_c for orc_c in out]
    assert len(out) == sum(1 for c in raw if c >= 64)
    assert list(orc.preprocess(b"ACGTN acgtn\n@`")) == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 32, 32]


def test_scan_and_splitters_properties():
    rng = np.random.default_rng(3)
    k = 21
    ref = rng.integers(0, 4, 50000).astype(np.uint8)
    spl, singles = orc.determine_splitters([ref], k, 2000)
    assert len(spl) >= 20 and np.all(np.diff(spl.astype(np.float64)) >= 0)
    assert np.all(np.isin(spl, singles))
    cuts = orc.scan_contig(ref, k, spl)
    # segments tile the contig with k-base overlaps
    assert cuts[0].start == 0
    for a, b in zip(cuts, cuts[1:]):
        assert b.start == a.start + a.len - k
    assert cuts[-1].start + cuts[-1].len == len(ref)
    # the reverse complement is cut at the mirrored places
    rc = orc.revcomp(ref)
    cuts_rc = orc.scan_contig(rc, k, spl)
    assert sorted(c.len for c in cuts) == sorted(c.len for c in cuts_rc)


def test_bytes2tuples_roundtrip_marker():
    for n in range(0, 13):
        b = (np.arange(n) % 4).astype(np.uint8)
        t = orc.bytes2tuples(b)
        assert t[-1] == (4 << 4) + n % 4 and len(t) == n // 4 + 2


def test_generator_is_deterministic(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    arch = json.load(open(os.path.join(GOLD, "archives.json")))
    spec = [a for a in arch if a["name"] == "viral40"][0]
    kw = {k: v for k, v in spec["spec"].items() if k != "kind"}
    files, _ = gen_data.viral(str(tmp_path), **kw)
    sha = hashlib.sha256(b"".join(open(f, "rb").read() for f in files)).hexdigest()
    assert sha == spec["fasta_sha256"]
