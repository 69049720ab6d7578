#!/usr/bin/env python
"""bench.py -- input Gbp/s of the AGC compression hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

A "step" is one complete `agc create` of the workload through the reference-facing API of this repo
(agc_b200's CAGCCompressor mirror behind the C ABI): splitter determination on the reference sample, ingest + 2-bit
packing, splitter scan, hash-assign, reference index build, LZ-diff encoding of every segment, reference tuple packing,
residual coding (when the device coder is built) and archive assembly.
  value : whole-job throughput with the raw FASTA bodies already resident in HBM when the timed region starts
  e2e   : the same call with HOST (pinned) buffers: host->device copies of every input byte and device->host copies of
          all deltas / packed references inside the timed region
Workload (config.workload): BASELINE.json configs[1] = 1000 synthetic 30 kb viral genomes (1 % SNP from one random
reference), k=25, 1 GPU.  N>1 (torchrun): every rank compresses its own replica of the workload (weak scaling, no
data-path collective; see DESIGN.md "multi-GPU").

--impl reference : the UNMODIFIED reference binary (oracle/_ref/agc, built from /root/reference by oracle/Makefile.ref)
on the host cores, same files / flags / metric.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

K, MML, SEG, PACK = 25, 20, 60000, 50
N_SAMPLES, REF_LEN, P_SNP, SEED = 1000, 30000, 0.01, 1


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_workload(tmp):
    import gen_data
    files, _ = gen_data.viral(os.path.join(tmp, "data"), n_samples=N_SAMPLES, ref_len=REF_LEN, p=P_SNP, seed=SEED)
    total = (N_SAMPLES + 1) * REF_LEN
    return files, total


def read_bodies(files):
    """FASTA -> (sample names, per-contig sample index, contig ids, concatenated raw bodies, offsets)"""
    names, soc, ids, bodies = [], [], [], []
    for fn in files:
        nm = os.path.basename(fn)
        for suf in (".fa",):
            if nm.endswith(suf):
                nm = nm[:-len(suf)]
        names.append(nm)
        data = open(fn, "rb").read()
        for rec in data.split(b">")[1:]:
            nl = rec.index(b"\n")
            ids.append(rec[:nl].decode())
            bodies.append(rec[nl + 1:])
            soc.append(len(names) - 1)
    offs = np.zeros(len(bodies) + 1, np.uint64)
    offs[1:] = np.cumsum([len(b) for b in bodies])
    return names, np.array(soc, np.uint32), ids, b"".join(bodies), offs


class ClockSampler(threading.Thread):
    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.stop_flag, self.sm, self.reasons, self.max_sm = dev, False, [], set(), None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.sm.append(float(o[0])); self.max_sm = float(o[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), o[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "agc")
    tmp = tempfile.mkdtemp(prefix="agcbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        files, total = make_workload(tmp)
        cores = os.cpu_count() or 1
        lst = os.path.join(tmp, "list.txt")
        open(lst, "w").write("\n".join(files[1:]) + "\n")
        cmd = [ref_bin, "create", "-k", str(K), "-t", str(cores), "-o", os.path.join(tmp, "ref.agc"), "-i", lst, files[0]]
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        val = total / (ms * 1e-3) / 1e9
        line = {"impl": "reference", "metric": "input Gbp/s (agc create, bit-exact .agc)", "value": val, "unit": "Gbp/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": workload_config(), "gpu_launches": 0,
                "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": cores, "kind": "reference",
                                 "sample": f"full workload ({total} bases), oracle/_ref/agc create -t {cores}, files in tmpfs"},
                "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def workload_config():
    return {"workload": f"BASELINE configs[1]: {N_SAMPLES} synthetic {REF_LEN} b viral genomes + reference, {P_SNP:.0%} SNP, seed {SEED}, agc create -k {K} (l={MML} s={SEG} b={PACK})",
            "l2_policy": "inputs re-uploaded / outputs re-written every step; working set per step (raw FASTA 30.4 MB + packed + deltas) plus a 256 MiB L2 flush write between steps",
            "parallelism": "one process per GPU, replicas"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import agc_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: agc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = agc_b200.lib()
    vp = C.c_void_p
    L.agcgpu_compressor_create.restype = C.c_int
    L.agcgpu_compressor_create.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                           C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_char_p, C.POINTER(vp)]
    L.agcgpu_compressor_add_samples_memory.restype = C.c_int
    L.agcgpu_compressor_add_samples_memory.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_char_p),
                                                       C.c_uint32, vp, C.POINTER(C.c_uint64), C.c_int]
    L.agcgpu_compressor_set_discard_parts.restype = C.c_int; L.agcgpu_compressor_set_discard_parts.argtypes = [vp, C.c_int]
    L.agcgpu_compressor_close.restype = C.c_int; L.agcgpu_compressor_close.argtypes = [vp, C.c_uint32]
    L.agcgpu_compressor_last_error.restype = C.c_char_p; L.agcgpu_compressor_last_error.argtypes = [vp]
    L.agcgpu_compressor_ctx.restype = vp; L.agcgpu_compressor_ctx.argtypes = [vp]
    L.agcgpu_compressor_total_bases.restype = C.c_uint64; L.agcgpu_compressor_total_bases.argtypes = [vp]
    L.agcgpu_compressor_last_stats.restype = C.c_int; L.agcgpu_compressor_last_stats.argtypes = [C.POINTER(agc_b200.Stats)]

    tmp = tempfile.mkdtemp(prefix=f"agcbench{rank}_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        files, total = make_workload(tmp)
        names, soc, ids, raw, offs = read_bodies(files)
        n_ctg = len(ids)
        c_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
        c_ids = (C.c_char_p * n_ctg)(*[i.encode() for i in ids])
        raw_np = np.frombuffer(raw, np.uint8)
        pinned = torch.empty(len(raw_np) + 64, dtype=torch.uint8).pin_memory()
        pinned[:len(raw_np)] = torch.from_numpy(raw_np.copy())
        dev_raw = torch.empty(len(raw_np) + 64, dtype=torch.uint8, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        has_zstd = os.environ.get("AGC_BENCH_RESIDUAL", "auto")
        out_path = os.path.join(tmp, "out.agc")

        def one_step(resident):
            """returns (seconds, stats)"""
            flush.fill_(1)                      # L2 flush between timed iterations
            if resident:
                dev_raw[:len(raw_np)].copy_(pinned[:len(raw_np)])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            h = vp()
            rc = L.agcgpu_compressor_create(out_path.encode(), PACK, K, files[0].encode(), SEG, MML, 0, 0, 0, 1, 0.0, local_rank, None, C.byref(h))
            if rc:
                raise SystemExit("compressor_create failed: " + L.agcgpu_compressor_last_error(None).decode())
            if not RESIDUAL:
                L.agcgpu_compressor_set_discard_parts(h, 1)
            ptr = vp(dev_raw.data_ptr()) if resident else vp(pinned.data_ptr())
            rc = L.agcgpu_compressor_add_samples_memory(h, c_names, len(names), soc.ctypes.data_as(C.POINTER(C.c_uint32)), c_ids, n_ctg,
                                                        ptr, offs.ctypes.data_as(C.POINTER(C.c_uint64)), 1 if resident else 0)
            if rc:
                raise SystemExit("add_samples failed: " + L.agcgpu_compressor_last_error(h).decode())
            st = agc_b200.Stats()
            L.agcgpu_get_stats(L.agcgpu_compressor_ctx(h), C.byref(st))
            stats = {k: getattr(st, k) for k in ("kernel_launches", "h2d_bytes", "d2h_bytes", "lz_alg_bytes", "last_lz_kernel_ms", "last_scan_kernel_ms")}
            rc = L.agcgpu_compressor_close(h, 1)
            if rc:
                raise SystemExit("close failed: " + L.agcgpu_compressor_last_error(None).decode())
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            L.agcgpu_compressor_last_stats(C.byref(st))          # counters incl. the residual coder, which runs inside close
            for k in ("kernel_launches", "h2d_bytes", "d2h_bytes", "zstd_kernel_ms", "zstd_input_mb"):
                stats[k] = getattr(st, k)
            return dt, stats

        # is the device residual coder available?  (probe once; without it the step stops before zstd and says so)
        global RESIDUAL
        RESIDUAL = True
        if has_zstd == "0":
            RESIDUAL = False
        elif has_zstd == "auto":
            try:
                one_step(True)
            except SystemExit:
                RESIDUAL = False

        def timed(resident):
            for _ in range(args.warmup):
                one_step(resident)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            times, last = [], None
            for _ in range(args.steps):
                dt, last = one_step(resident)
                times.append(dt)
            torch.cuda.synchronize()
            t = torch.tensor([sum(times)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.barrier()
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()) / args.steps, last

        sampler = ClockSampler(local_rank)
        sampler.start()
        sec_res, st_res = timed(True)
        sec_e2e, st_e2e = timed(False)
        sampler.stop_flag = True
        sampler.join(timeout=2)

        if rank == 0:
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(peaks_path):
                peak = json.load(open(peaks_path))["hbm_gbs"]; peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
            else:
                peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
            lz_gbs = st_res["lz_alg_bytes"] / (st_res["last_lz_kernel_ms"] * 1e-3) / 1e9 if st_res["last_lz_kernel_ms"] else 0.0
            prof = os.path.join(ROOT, "profiles", "r01_lz_traffic.json")
            traffic = json.load(open(prof)).get("traffic_bytes_per_launch") if os.path.exists(prof) else None
            # the same LZ kernel on a batch shaped like one HPP-scale device batch (context for the roofline figure: the C2
            # launch above holds 16 MB of algorithmic bytes, less than a launch latency worth of HBM traffic)
            lz_hpp = lz_hpp_batch(local_rank, peak)
            # bounded CPU sample of the same workload: reference binary on the first 200 samples
            cpu = cpu_baseline(files, tmp)
            cfg = workload_config()
            cfg["stages"] = ("determine_splitters, ingest+2bit pack, splitter scan, hash-assign, LZ index, LZ-diff encode, ref tuple pack, "
                             + ("residual coder (zstd frames), archive write" if RESIDUAL else "archive part assembly -- residual coder (zstd, SURVEY a24) NOT yet in the step"))
            line = {"metric": "input Gbp/s (agc create" + (", bit-exact .agc)" if RESIDUAL else ", residual coder excluded)"),
                    "value": world * total / sec_res / 1e9, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": sec_res * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                    "data": "synthetic", "config": cfg,
                    "e2e": {"value": world * total / sec_e2e / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]),
                            "d2h_bytes_per_step": int(st_e2e["d2h_bytes"]), "ms_per_step": sec_e2e * 1e3},
                    "gpu_launches": int(st_res["kernel_launches"]) * args.steps,
                    "roofline": {"kernel": "k_lz_packed<0> (LZ-diff encode)", "bound": "hbm", "achieved": lz_gbs, "peak": peak, "unit": "GB/s",
                                 "frac": lz_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                                 "algorithmic_bytes_per_launch": int(st_res["lz_alg_bytes"]), "kernel_ms": st_res["last_lz_kernel_ms"]},
                    "lz_kernel_hpp_like_batch": lz_hpp,
                    "residual_coder": {"kernel": "k_zstd (bit-exact zstd frames; one CTA per part: parser warp + 15 warps of tree walks)", "ms_per_step": float(st_res["zstd_kernel_ms"]),
                                       "input_bytes_per_step": int(st_res["zstd_input_mb"] * 1e6),
                                       "share_of_step": float(st_res["zstd_kernel_ms"]) / (sec_res * 1e3),
                                       "note": "critical path of the step = the largest frame (1.35 MB raw-group pack, btopt): the optimal parse is sequential per frame; latency bound, not an HBM-roofline kernel (DESIGN.md section 5)"},
                    "cpu_baseline": cpu, "clocks": sampler.summary()}
            print(json.dumps(line))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


def lz_hpp_batch(device, peak):
    """k_lz_packed<0> on 4096 segments of 60 031 bases (0.1 % SNP, 64 reference segments): CUDA-event time of the launch"""
    import agc_b200
    import gen_data
    n_seg, seg_len, n_groups = 4096, 60031, 64
    rng = np.random.default_rng(1)
    LET = np.frombuffer(b"ACGT", np.uint8)
    refs = [rng.integers(0, 4, seg_len, dtype=np.uint8) for _ in range(n_groups)]
    contigs = [LET[r].tobytes() for r in refs] + [LET[gen_data.substitute(rng, refs[i % n_groups], 0.001)].tobytes() for i in range(n_seg)]
    dev = agc_b200.Device(k=31, min_match_len=20, device=device)
    try:
        dev.set_splitters(np.zeros(0, np.uint64))
        dev.scan_contigs(contigs)
        dev.put_references([(g, 0, seg_len, False, 16 + g) for g in range(n_groups)])
        arr = dev._reqs([(n_groups + i, 0, seg_len, False, 16 + (i % n_groups)) for i in range(n_seg)])
        out = np.zeros(n_seg * (seg_len // 8 + 64), np.uint8)
        offs = np.zeros(n_seg + 1, np.uint64)
        ms = []
        for _ in range(5):
            dev.lz_encode_raw(arr, n_seg, out, offs)
            st = dev.stats()
            ms.append(st.last_lz_kernel_ms)
        ms = sorted(ms[2:])[len(ms[2:]) // 2]
        gbs = st.lz_alg_bytes / (ms * 1e-3) / 1e9
        tp = os.path.join(ROOT, "profiles", "r01_lz_traffic_hpp.json")
        traffic = json.load(open(tp)).get("traffic_bytes_per_launch") if os.path.exists(tp) else None
        return {"workload": f"{n_seg} segments x {seg_len} bases, 0.1% SNP, {n_groups} reference segments (working set 125 MB ~ L2 size, 2 warm-up launches)",
                "kernel_ms": ms, "algorithmic_bytes_per_launch": int(st.lz_alg_bytes), "achieved": gbs, "unit": "GB/s", "frac": gbs / peak,
                "traffic": traffic}
    finally:
        dev.close()


def cpu_baseline(files, tmp):
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "agc")
    if not os.path.exists(ref_bin):
        return {"value": None, "unit": "Gbp/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/agc not built"}
    cores = os.cpu_count() or 1
    n = 200
    sub = files[:n + 1]
    bases = (n + 1) * REF_LEN
    lst = os.path.join(tmp, "cpu_list.txt")
    open(lst, "w").write("\n".join(sub[1:]) + "\n")
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        subprocess.check_call([ref_bin, "create", "-k", str(K), "-t", str(cores), "-o", os.path.join(tmp, "cpu.agc"), "-i", lst, sub[0]],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": bases / best / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "reference",
            "sample": f"first {n} samples + reference of the workload ({bases} bases), oracle/_ref/agc create -t {cores}, best of 2"}


RESIDUAL = True

if __name__ == "__main__":
    main()
