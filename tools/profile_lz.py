"""LZ-diff kernel micro-benchmark (used for ncu captures and for roofline numbers in DESIGN.md).
usage: python tools/profile_lz.py [n_segments] [seg_len] [p_snp] [n_groups] [reps]"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import agc_b200
import gen_data

n_seg = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
seg_len = int(sys.argv[2]) if len(sys.argv) > 2 else 60031
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.001
n_groups = int(sys.argv[4]) if len(sys.argv) > 4 else 64
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5

rng = np.random.default_rng(1)
LET = np.frombuffer(b"ACGT", np.uint8)
refs = [rng.integers(0, 4, seg_len, dtype=np.uint8) for _ in range(n_groups)]
contigs = [LET[r].tobytes() for r in refs]
for i in range(n_seg):
    contigs.append(LET[gen_data.substitute(rng, refs[i % n_groups], p)].tobytes())
dev = agc_b200.Device(k=31, min_match_len=20)
dev.set_splitters(np.zeros(0, np.uint64))
dev.scan_contigs(contigs)
dev.put_references([(g, 0, seg_len, False, 16 + g) for g in range(n_groups)])
reqs = [(n_groups + i, 0, seg_len, False, 16 + (i % n_groups)) for i in range(n_seg)]
arr = dev._reqs(reqs)
out = np.zeros(n_seg * (seg_len // 8 + 64), np.uint8)
offs = np.zeros(n_seg + 1, np.uint64)
best = None
for r in range(reps):
    t0 = time.time()
    dev.lz_encode_raw(arr, n_seg, out, offs)
    wall = time.time() - t0
    st = dev.stats()
    gbs = st.lz_alg_bytes / (st.last_lz_kernel_ms * 1e-3) / 1e9
    print(f"rep {r}: lz kernel {st.last_lz_kernel_ms:.3f} ms, alg bytes {st.lz_alg_bytes}, {gbs:.1f} GB/s algorithmic, "
          f"{n_seg * seg_len / (st.last_lz_kernel_ms * 1e-3) / 1e9:.1f} Gbase/s, call wall {wall * 1e3:.1f} ms, delta bytes {int(offs[-1])}")
