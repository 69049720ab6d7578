"""Generate the committed golden vectors under tests/golden/ from the UNMODIFIED reference built in oracle/_ref/
(liblzdiff_ref.so = src/common/lz_diff.cpp, agc = the CLI).  Run in the build container (needs /root/reference):
    make -f oracle/Makefile.ref && python tools/make_golden.py
"""
import hashlib, json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import orc, gen_data

def lz_cases():
    rng = np.random.default_rng(2024)
    cases = []
    for it in range(40):
        m = int(rng.choice([0, 9, 60, 400, 1500, 3000]))
        mml = int(rng.choice([15, 18, 20, 24, 32]))
        ref = rng.integers(0, 4, m).astype(np.uint8)
        if it % 5 == 1 and m > 50: ref[20:40] = 4
        t = ref.copy()
        if len(t):
            msk = rng.random(len(t)) < float(rng.choice([0, 0.002, 0.02, 0.2]))
            t[msk] = (t[msk] + rng.integers(1, 4, int(msk.sum()))) % 4
        t = list(t)
        for _ in range(int(rng.integers(0, 3))):
            pos = int(rng.integers(0, max(1, len(t))))
            if rng.random() < 0.5: del t[pos:pos + int(rng.integers(1, 25))]
            else: t[pos:pos] = list(rng.integers(0, 4, int(rng.integers(1, 25))))
        t = np.array(t, np.uint8)
        if it % 4 == 2 and len(t) > 30: t[10:10 + int(rng.integers(3, 15))] = 4
        if it % 9 == 3 and len(t) > 5: t[3] = 11
        if it % 7 == 0: t = ref.copy()
        bound = int(rng.integers(0, 40))
        cases.append(dict(mml=mml, ref=ref.tobytes().hex(), text=t.tobytes().hex(), bound=bound,
                          encode=orc.ref_encode(ref, t, mml).hex(), estimate=orc.ref_estimate(ref, t, mml),
                          estimate_bounded=orc.ref_estimate(ref, t, mml, bound),
                          cost_prefix=orc.ref_cost_vector(ref, t, mml, 1).tolist(), cost_suffix=orc.ref_cost_vector(ref, t, mml, 0).tolist()))
    return cases

def archive_cases():
    out = []
    agc = os.path.join(ROOT, "oracle", "_ref", "agc")
    specs = [("viral40", dict(kind="viral", n_samples=40, ref_len=30000, p=0.01, seed=1), ["-k", "25"]),
             ("complex5", dict(kind="complex", seed=5), ["-k", "21", "-s", "2000", "-b", "5"]),
             ("complexN6", dict(kind="complex", seed=6, with_n=True), ["-k", "31", "-s", "3000", "-l", "18", "-b", "4"]),
             ("smallpacks", dict(kind="viral", n_samples=25, ref_len=9000, p=0.02, seed=9), ["-k", "17", "-s", "1000", "-b", "3", "-l", "15"])]
    for name, spec, flags in specs:
        with tempfile.TemporaryDirectory() as tmp:
            kw = {k: v for k, v in spec.items() if k != "kind"}
            files = gen_data.viral(tmp, **kw)[0] if spec["kind"] == "viral" else gen_data.complex_collection(tmp, **kw)
            fasta_sha = hashlib.sha256(b"".join(open(f, "rb").read() for f in files)).hexdigest()
            shas = set()
            for t in ("1", "4"):
                o = os.path.join(tmp, "o.agc")
                subprocess.check_call([agc, "create", "-t", t, "-o", o] + flags + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                shas.add(hashlib.sha256(open(o, "rb").read()).hexdigest()); size = os.path.getsize(o)
            assert len(shas) == 1, "reference output depends on thread count?!"
            out.append(dict(name=name, spec=spec, flags=flags, fasta_sha256=fasta_sha, agc_sha256=shas.pop(), agc_size=size))
    return out

if __name__ == "__main__":
    json.dump(lz_cases(), open(os.path.join(ROOT, "tests", "golden", "lz_golden.json"), "w"))
    json.dump(archive_cases(), open(os.path.join(ROOT, "tests", "golden", "archives.json"), "w"), indent=1)
    print("golden vectors written")
