"""GPU tool: the LZ-diff encode kernels on a batch shaped like one HPP-scale device batch (N segments x 60 031 bases at a given SNP
rate against G reference segments): CUDA-event time of the launch, algorithmic GB/s, share handed to the sequential kernel.
usage: python tools/lz_hpp_bench.py [n_seg] [snp_rate] [n_groups]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import agc_b200, gen_data


def run(n_seg=4096, p=0.001, n_groups=64, seg_len=60031, reps=5, device=0, check=False, flush=False):
    rng = np.random.default_rng(1)
    LET = np.frombuffer(b"ACGT", np.uint8)
    refs = [rng.integers(0, 4, seg_len, dtype=np.uint8) for _ in range(n_groups)]
    texts = [gen_data.substitute(rng, refs[i % n_groups], p) for i in range(n_seg)]
    contigs = [LET[r].tobytes() for r in refs] + [LET[t].tobytes() for t in texts]
    dev = agc_b200.Device(k=31, min_match_len=20, device=device)
    try:
        dev.set_splitters(np.zeros(0, np.uint64))
        dev.scan_contigs(contigs)
        dev.put_references([(g, 0, seg_len, False, 16 + g) for g in range(n_groups)])
        arr = dev._reqs([(n_groups + i, 0, seg_len, False, 16 + (i % n_groups)) for i in range(n_seg)])
        out = np.zeros(n_seg * (seg_len // 8 + 64), np.uint8)
        offs = np.zeros(n_seg + 1, np.uint64)
        ms = []
        fl = None
        if flush:
            import torch
            fl = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{device}")
        for _ in range(reps):
            if fl is not None:
                fl.fill_(1); torch.cuda.synchronize()
            dev.lz_encode_raw(arr, n_seg, out, offs)
            st = dev.stats()
            ms.append(st.last_lz_kernel_ms)
        med = sorted(ms[1:])[len(ms[1:]) // 2]
        res = {"n_seg": n_seg, "seg_len": seg_len, "snp": p, "groups": n_groups, "kernel_ms": med, "all_ms": ms, "alg_bytes": int(st.lz_alg_bytes),
               "GBps": st.lz_alg_bytes / (med * 1e-3) / 1e9, "Gbase_per_s": n_seg * seg_len / (med * 1e-3) / 1e9,
               "chunk_segments": int(st.lz_chunk_segments), "sequential_segments": int(st.lz_sequential_segments)}
        if check:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import orc
            bad = 0
            for i in range(0, n_seg, max(1, n_seg // 64)):
                z = orc.LZ(refs[i % n_groups], 20)
                if z.encode(texts[i]) != out[int(offs[i]):int(offs[i + 1])].tobytes(): bad += 1
            res["checked_mismatches"] = bad
        return res
    finally:
        dev.close()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.001
    g = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    print(json.dumps(run(n, p, g, check=True)))
