"""Golden archives of the UNMODIFIED reference binary (oracle/_ref/agc) for every collection of tests/test_host_pipeline.py
(create with -a / -c / -f in the combinations listed there) and for the create-then-append runs: sha256 of the inputs and of the
archive, committed as tests/golden/archives_modes.json.  Run in the build container:
    make -f oracle/Makefile.ref && python tools/make_golden_modes.py
tests/test_host_pipeline.py::test_golden_archives compares the host pipeline (mocked device ABI) with these hashes, so the
oracle functions behind the modes stay pinned to the reference's output where the reference binary is not available."""
import hashlib, json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from test_host_pipeline import collection, ALL_CASES, APPEND_CASES, REF_AGC, run_append_case


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    out = {"create": {}, "append": {}}
    for case in ALL_CASES:
        with tempfile.TemporaryDirectory() as tmp:
            files, flags = collection(case, tmp)
            hs = set()
            for t in ("1", "4"):
                o = os.path.join(tmp, "o.agc")
                subprocess.check_call([REF_AGC, "create", "-t", t, "-o", o] + flags + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                hs.add(sha(open(o, "rb").read())); size = os.path.getsize(o)
            assert len(hs) == 1, f"{case}: reference output depends on the thread count"
            out["create"][case] = dict(flags=flags, fasta_sha256=sha(b"".join(open(f, "rb").read() for f in files)), agc_sha256=hs.pop(), agc_size=size)
    for case, n_first, steps in APPEND_CASES:
        with tempfile.TemporaryDirectory() as tmp:
            _, ref_bytes, files = run_append_case(tmp, REF_AGC, case, n_first, steps)      # "ours" = the reference itself here
            out["append"][f"{case}:{n_first}:{steps}"] = dict(agc_sha256=sha(ref_bytes), agc_size=len(ref_bytes))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "archives_modes.json"), "w"), indent=1)
    print("golden mode archives written:", len(out["create"]), "create,", len(out["append"]), "append")


if __name__ == "__main__":
    main()
