"""TEST TOOL: extended differential fuzz of the diagonal-streaming LZ-diff encoder (agc_b200/csrc/lz_diag_core.cuh, host build
tests/lzd_host) against the C oracle's CLZDiff_V2::Encode.  usage: python tools/fuzz_lz_diag.py [n_cases] [first_seed] [LZD_DEFS]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from test_lzc_host import make_case, mutate
from test_lzd_host import build, enc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
L = build(sys.argv[3] if len(sys.argv) > 3 else "")
bad = 0
for s in range(s0, s0 + n):
    rng, mml, ref, t = make_case(s)
    if s % 5 == 0:                       # extra shapes: clustered mismatches, long insertions, text longer / shorter than the reference
        k = int(rng.integers(0, 4))
        if k == 0 and len(ref) > 2000:
            t = ref.copy(); a = int(rng.integers(0, len(t) - 600)); idx = a + np.sort(rng.choice(600, int(rng.integers(2, 60)), replace=False)); t[idx] = (t[idx] + 1) % 4
        elif k == 1:
            t = np.concatenate([mutate(rng, ref, 0.002), rng.integers(0, 4, int(rng.integers(1, 9000))).astype(np.uint8)])
        elif k == 2 and len(ref) > 400:
            t = mutate(rng, ref[:int(rng.integers(200, len(ref)))], 0.004, 1)
        else:
            t = mutate(rng, np.concatenate([ref[len(ref) // 3:], ref[:len(ref) // 3]]), 0.003)      # rotated: two diagonals
    z = orc.LZ(ref, mml)
    r, got = enc(L, t, ref, mml, z, int(rng.random() < 0.4), int(rng.integers(0, 70)))
    if got != z.encode(t):
        bad += 1
        print("MISMATCH seed", s, "mml", mml, "n", len(t), "m", len(ref), flush=True)
print(f"{n} cases from seed {s0}: {bad} mismatches")
