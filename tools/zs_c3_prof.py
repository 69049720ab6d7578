"""GPU diagnostic: time the device residual coder on real delta packs / references of the C3 archive (built by the reference)"""
import sys, os, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import agc_b200, gen_data, agc_parts as ap
d = "/dev/shm/c3p"
ns = int(os.environ.get("N_SAMPLES", 63))
files = gen_data.bacterial_adaptive(d, seed=2, n_samples=ns, ref_len=int(os.environ.get("REF_LEN", 5000000)))
open(d + "/list.txt", "w").write("\n".join(files[1:]) + "\n")
subprocess.check_call([os.path.join(ROOT, "oracle/_ref/agc"), "create", "-a", "-k", "29", "-t", "16", "-o", d + "/ref.agc", "-i", d + "/list.txt", d + "/ref.fa"],
                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
names, parts = ap.read_archive(d + "/ref.agc")[:2]
ds = sorted([p for p in parts if p["name"].startswith("x") and p["name"].endswith("d") and p["metadata"] > 0], key=lambda p: -p["metadata"])
rs = [p for p in parts if p["name"].startswith("x") and p["name"].endswith("r") and p["metadata"] > 0]
sel = [ds[0], ds[len(ds) // 2], ds[-20], rs[0]]
dev = agc_b200.Device(k=21, min_match_len=20)
dev.zstd_compress([b"hello hello hello hello"], [17])
for p in sel:
    raw = ap.zstd_decompress(p["payload"][:-1]); lv = 17 if p["name"].endswith("d") else (13 if p["payload"][-1] == 1 else 19)
    t0 = time.time(); out = dev.zstd_compress([raw], [lv]); dt = time.time() - t0
    t0 = time.time(); ap.zstd_compress(raw, lv); ct = time.time() - t0
    print(f"{p['name']} L{lv} {len(raw)} B -> {len(out[0])} B  gpu {dt*1e3:.1f} ms ({dt*1e6/len(raw):.2f} us/B)  cpu libzstd {ct*1e3:.1f} ms ({ct*1e6/len(raw):.2f} us/B) identical={out[0]==p['payload'][:-1]}", flush=True)
    if os.environ.get("ZS_DUMP"):
        open(f"gpurun_out/zs_{p['name']}.bin", "wb").write(raw)
