"""GPU diagnostic: time the device residual coder on the real parts of the C2 archive, one input per call"""
import sys, os, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import agc_b200, gen_data, agc_parts as ap
d = "/dev/shm/c2p"
files, _ = gen_data.viral(d, n_samples=1000, ref_len=30000, p=0.01, seed=1)
open(d + "/list.txt", "w").write("\n".join(files[1:]) + "\n")
subprocess.check_call([os.path.join(ROOT, "oracle/_ref/agc"), "create", "-k", "25", "-t", "8", "-o", d + "/ref.agc", "-i", d + "/list.txt", d + "/ref.fa"],
                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
names, parts = ap.read_archive(d + "/ref.agc")
dev = agc_b200.Device(k=21, min_match_len=20)
sel = [("x0d", 0, 17), ("xHd", 0, 17), ("xHd", 1, 17), ("xHr", 0, 13)]
only = os.environ.get("ZS_ONLY")
if only:
    sel = [x for x in sel if f"{x[0]}{x[1]}" == only]
inputs = {}
for pt in parts:
    for nm, ix, lv in sel:
        if pt["name"] == nm and pt["index"] == ix:
            inputs[(nm, ix)] = (ap.zstd_decompress(pt["payload"][:-1]), lv, pt["payload"][:-1])
dev.zstd_compress([b"hello hello hello hello"], [17])
for (nm, ix), (raw, lv, frame) in inputs.items():
    for r in range(1 if only else 2):
        t0 = time.time(); out = dev.zstd_compress([raw], [lv]); dt = time.time() - t0
    t0 = time.time(); ap.zstd_compress(raw, lv); ct = time.time() - t0
    print(f"{nm}[{ix}] L{lv} {len(raw)} B -> {len(out[0])} B  gpu {dt*1e3:.1f} ms ({dt*1e6/len(raw):.2f} us/B)  cpu libzstd {ct*1e3:.1f} ms  identical={out[0]==frame}", flush=True)
if only:
    sys.exit(0)
allraw = [v[0] for k, v in inputs.items() if k[0] == "xHd"] * 10
t0 = time.time(); dev.zstd_compress(allraw, [17] * len(allraw)); dt = time.time() - t0
print(f"20 delta packs together: {dt*1e3:.1f} ms")
