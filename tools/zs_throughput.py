"""GPU diagnostic: residual-coder throughput on many small frames (the HPP-scale population: ~15 KB tuple-packed references at
level 13, few-KB delta packs at level 17), frames per second and input MB/s per batch"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, agc_b200, orc
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(0)
refs = [bytes(rng.integers(0, 256, 15008, dtype=np.uint8)) for _ in range(n_frames)]          # 60 kb reference, 4 bases / byte
ref = rng.integers(0, 4, 60000).astype(np.uint8); z = orc.LZ(ref, 20)
deltas = []
for i in range(64):
    parts = []
    for s in range(20):
        t = ref.copy(); m = rng.random(len(t)) < 0.001; t[m] = (t[m] + rng.integers(1, 4, int(m.sum()))) % 4
        parts.append(z.encode(t) + b"\xff")
    deltas.append(b"".join(parts))
deltas = [deltas[i % 64] for i in range(n_frames)]
dev = agc_b200.Device(k=21, min_match_len=20)
dev.zstd_compress([b"hello hello hello hello"], [17])
for name, inputs, lv in (("references L13", refs, 13), ("delta packs L17", deltas, 17)):
    for r in range(2):
        t0 = time.time(); out = dev.zstd_compress(inputs, [lv] * len(inputs)); dt = time.time() - t0
    tot = sum(map(len, inputs))
    print(f"{name}: {len(inputs)} frames, {tot/1e6:.1f} MB (avg {tot//len(inputs)} B) -> {sum(map(len,out))/1e6:.1f} MB in {dt*1e3:.0f} ms = {len(inputs)/dt:.0f} frames/s, {tot/dt/1e6:.1f} MB/s", flush=True)
