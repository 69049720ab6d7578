"""profile driver for the device residual coder: one delta-pack-like input, level 17"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, agc_b200, orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(0)
ref = rng.integers(0, 4, 30000).astype(np.uint8); z = orc.LZ(ref, 20); parts = []
while sum(map(len, parts)) < n:
    t = ref.copy(); m = rng.random(len(t)) < 0.01; t[m] = (t[m] + rng.integers(1, 4, int(m.sum()))) % 4
    parts.append(z.encode(t) + b"\xff")
raw = b"".join(parts)[:n]
if len(sys.argv) > 3 and sys.argv[3] == "raw":                     # raw-group pack: near-copies of one sequence, 1 byte per base
    ref = rng.integers(0, 4, 27000).astype(np.uint8); parts = []
    while sum(map(len, parts)) < n:
        t = ref.copy(); m = rng.random(len(t)) < 0.01; t[m] = (t[m] + rng.integers(1, 4, int(m.sum()))) % 4
        parts.append(bytes(t))
    raw = b"".join(parts)[:n]
dev = agc_b200.Device(k=21, min_match_len=20)
for r in range(reps):
    t0 = time.time(); out = dev.zstd_compress([raw], [17]); dt = time.time() - t0
    print(f"{n} bytes -> {len(out[0])} in {dt:.3f} s = {dt * 1e6 / n:.1f} us/B")
