import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import numpy as np, agc_b200
rng = np.random.default_rng(0)
dev = agc_b200.Device(k=21, min_match_len=20)
for n, lv in ((0, 19), (5, 19), (100, 19), (300, 17), (5000, 13)):
    raw = bytes(rng.integers(65, 69, n, dtype=np.uint8))
    if n == 5000: dev.zstd_compress([raw], [lv])
