#!/usr/bin/env python
"""bench.py -- input Gbp/s of the AGC compression hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

A "step" is one complete `agc create` of the workload through the reference-facing API of this repo (agc_b200's
CAGCCompressor mirror behind the C ABI): splitter determination on the reference sample, ingest + 2-bit packing, splitter
scan, hash-assign, reference index build, LZ-diff encoding of every segment, reference tuple packing, residual coding
(bit-exact zstd frames) and archive assembly.  The archive is written in every step; its sha256 is compared IN THE RUN with
the archive the reference binary writes for the same files (cpu_baseline leg) -- a mismatch aborts the bench.
  value : whole-job throughput with the raw FASTA bodies already resident in HBM when the timed region starts
  e2e   : the same create from the FASTA FILES (tmpfs): file read + parse + host->device copies + device->host copies +
          archive write inside the timed region -- what `agc create` does, and what the reference arm is timed on
Workload (config.workload): BASELINE.json configs[2] = 64 synthetic 5 Mb bacterial genomes, adaptive mode (-a), k=29 -- the
largest configuration BASELINE.json gives to one GPU.  `other_workloads` carries configs[1] (1000 x 30 kb viral, k=25) and configs[3]
at full size (8 x 250 Mb human-like contigs + reference: 2.25 Gbases fit one B200) measured the same way with fewer steps;
`lz_kernel_hpp_like_batch` / `roofline.batch_*` the LZ-diff encode kernel on one HPP-scale device batch with the reference's own
Encode timed on one host core beside it; `residual_coder.cpu_zstd` the reference's libzstd over the workload's exact parts.
N>1 (torchrun): ONE create sharded over the N GPUs (strong scaling): every rank keeps the O(#segments) bookkeeping, the
per-base device work (LZ-diff encoding, residual coding) is split across the ranks and all-gathered over NCCL from C++
(agc_b200/csrc/comm.cu); rank 0 writes the archive, whose sha256 is checked like at N=1.

--impl reference : the UNMODIFIED reference binary (oracle/_ref/agc, built from /root/reference by oracle/Makefile.ref) on
the host cores, same files / flags / metric.
"""
import argparse
import ctypes as C
import hashlib
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

MML, SEG, PACK = 20, 60000, 50
WORKLOADS = {
    # name: (description, generator kwargs, k, adaptive)
    "c3": ("BASELINE configs[2]: 63 synthetic 5 Mb bacterial genomes (1% SNP, 20 indels, one novel 120 kb contig each) + 5 Mb reference, "
           "seed 2, agc create -a -k 29 (l=20 s=60000 b=50)", dict(kind="bacterial", seed=2, n_samples=63, ref_len=5_000_000), 29, 1),
    # not part of the default run: BASELINE configs[3] in shape at 1/5 of its size (python bench.py --workload c4s)
    "c4s": ("BASELINE configs[3] in shape, scaled: 4 synthetic 50 Mb human-chromosome-like contigs (0.1% SNP, an indel every 10 kb, repeats) + "
            "50 Mb reference, seed 3, agc create -k 31 (l=20 s=60000 b=50)", dict(kind="human", seed=3, n_samples=4, ref_len=50_000_000), 31, 0),
    "c4": ("BASELINE configs[3]: 8 synthetic 250 Mb human-chromosome-like contigs (0.1% SNP, an indel every 10 kb, repeats) + 250 Mb reference, "
           "seed 3, agc create -k 31 (l=20 s=60000 b=50)", dict(kind="human", seed=3, n_samples=8, ref_len=250_000_000), 31, 0),
    "c2": ("BASELINE configs[1]: 1000 synthetic 30 kb viral genomes (1% SNP) + reference, seed 1, agc create -k 25 (l=20 s=60000 b=50)",
           dict(kind="viral", seed=1, n_samples=1000, ref_len=30000), 25, 0),
}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_workload(tmp, name):
    import gen_data
    kw = dict(WORKLOADS[name][1]); kind = kw.pop("kind")
    d = os.path.join(tmp, "data_" + name)
    if kind == "viral":
        files, _ = gen_data.viral(d, n_samples=kw["n_samples"], ref_len=kw["ref_len"], p=0.01, seed=kw["seed"])
    elif kind == "human":
        files = gen_data.human_chromosome(d, seed=kw["seed"], n_samples=kw["n_samples"], ctg_len=kw["ref_len"], n_repeats=40 if kw["ref_len"] < 100_000_000 else 200)
    else:
        files = gen_data.bacterial_adaptive(d, seed=kw["seed"], n_samples=kw["n_samples"], ref_len=kw["ref_len"])
    return files, gen_data.total_bases(files)


def read_bodies(files):
    """FASTA -> (sample names, per-contig sample index, contig ids, concatenated raw bodies, offsets)"""
    names, soc, ids, bodies = [], [], [], []
    for fn in files:
        nm = os.path.basename(fn)
        if nm.endswith(".fa"):
            nm = nm[:-3]
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


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


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


def workload_config(name, world):
    return {"workload": WORKLOADS[name][0],
            "l2_policy": "inputs re-uploaded / outputs re-written every step and a 256 MiB L2 flush write between steps; working set per step (raw FASTA + packed + deltas) exceeds L2",
            "parallelism": "single GPU" if world == 1 else f"one create sharded over {world} GPUs (LZ-diff encoding and residual coding split across ranks, NCCL all-gather from C++; rank 0 writes the archive)"}


def ref_cmd(ref_bin, name, cores, out, lst, ref_fa):
    k, adaptive = WORKLOADS[name][2], WORKLOADS[name][3]
    return [ref_bin, "create"] + (["-a"] if adaptive else []) + ["-k", str(k), "-t", str(cores), "-o", out, "-i", lst, ref_fa]


def run_reference_once(files, tmp, name, tag):
    """the unmodified reference binary on the whole workload with every host core -> (seconds, sha256 of its archive)"""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "agc")
    cores = os.cpu_count() or 1
    lst = os.path.join(tmp, f"list_{name}.txt")
    open(lst, "w").write("\n".join(files[1:]) + "\n")
    out = os.path.join(tmp, f"ref_{name}_{tag}.agc")
    t0 = time.perf_counter()
    subprocess.check_call(ref_cmd(ref_bin, name, cores, out, lst, files[0]), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    return dt, sha256_file(out), cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    name = args.workload
    tmp = tempfile.mkdtemp(prefix="agcbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        files, total = make_workload(tmp, name)
        times, sha, cores = [], None, 1
        for it in range(args.warmup + args.steps):
            dt, sha, cores = run_reference_once(files, tmp, name, "arm")
            if it >= args.warmup:
                times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        val = total / (ms * 1e-3) / 1e9
        line = {"impl": "reference", "metric": "input Gbp/s (agc create, bit-exact .agc)", "value": val, "unit": "Gbp/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": workload_config(name, args.gpus), "gpu_launches": 0, "archive_sha256": sha,
                "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": cores, "kind": "reference",
                                 "sample": f"full workload ({total} bases), oracle/_ref/agc create -t {cores}, files in tmpfs"},
                "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


STAT_KEYS = ("kernel_launches", "h2d_bytes", "d2h_bytes", "zstd_kernel_ms", "zstd_input_mb", "lz_alg_bytes_total", "lz_kernel_ms_total",
             "scan_bytes_total", "scan_kernel_ms_total", "lz_encode_launches", "scan_launches", "lz_chunk_segments", "lz_sequential_segments",
             "zstd_wait_ms", "lz_diag_segments")


class Runner:
    """one workload through the CAGCCompressor facade of the C ABI"""

    def __init__(self, L, name, tmp, local_rank, rank, world):
        import torch
        import agc_b200
        self.torch, self.agc, self.L, self.name, self.rank, self.world, self.local_rank = torch, agc_b200, L, name, rank, world, local_rank
        self.k, self.adaptive = WORKLOADS[name][2], WORKLOADS[name][3]
        self.files, self.total = make_workload(tmp, name)
        names, self.soc, ids, raw, self.offs = read_bodies(self.files)
        self.n_ctg = len(ids)
        self.c_names = (C.c_char_p * len(names))(*[n.encode() for n in names]); self.n_names = len(names)
        self.c_ids = (C.c_char_p * self.n_ctg)(*[i.encode() for i in ids])
        raw_np = np.frombuffer(raw, np.uint8)
        self.raw_len = len(raw_np)
        self.pinned = torch.empty(self.raw_len + 64, dtype=torch.uint8).pin_memory()
        self.pinned[:self.raw_len] = torch.from_numpy(raw_np.copy())
        self.dev_raw = torch.empty(self.raw_len + 64, dtype=torch.uint8, device="cuda")
        self.c_files = (C.c_char_p * (len(self.files) - 1))(*[f.encode() for f in self.files[1:]])
        self.c_snames = (C.c_char_p * (len(self.files) - 1))(*[n.encode() for n in names[1:]])
        self.out_path = os.path.join(tmp, f"out_{name}_{rank}.agc")

    def step(self, resident, flush):
        """returns (seconds, stats); the archive is at self.out_path afterwards (rank 0)"""
        torch, L, vp = self.torch, self.L, C.c_void_p
        flush.fill_(1)                                   # L2 flush between timed iterations
        if resident:
            self.dev_raw[:self.raw_len].copy_(self.pinned[:self.raw_len])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = vp()
        rc = L.agcgpu_compressor_create(self.out_path.encode(), PACK, self.k, self.files[0].encode(), SEG, MML, 0, self.adaptive, 0, 1, 0.0,
                                        self.local_rank, None, C.byref(h))
        if rc:
            raise SystemExit("compressor_create failed: " + L.agcgpu_compressor_last_error(None).decode())
        if resident:
            # sample 0 (the reference file) is added by Create's caller in `agc create` as well: all samples go through here
            rc = L.agcgpu_compressor_add_samples_memory(h, self.c_names, self.n_names, self.soc.ctypes.data_as(C.POINTER(C.c_uint32)), self.c_ids,
                                                        self.n_ctg, vp(self.dev_raw.data_ptr()), self.offs.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
        else:
            all_files = (C.c_char_p * len(self.files))(*[f.encode() for f in self.files])
            all_names = (C.c_char_p * len(self.files))(*[self.c_names[i] for i in range(self.n_names)])
            rc = L.agcgpu_compressor_add_sample_files(h, all_names, all_files, len(self.files), 1)
        if rc:
            raise SystemExit("add_samples failed: " + L.agcgpu_compressor_last_error(h).decode())
        rc = L.agcgpu_compressor_close(h, 1)
        if rc:
            raise SystemExit("close failed: " + L.agcgpu_compressor_last_error(None).decode())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = self.agc.Stats()
        L.agcgpu_compressor_last_stats(C.byref(st))      # counters of the compressor just closed (incl. the residual coder in close)
        return dt, {k: getattr(st, k) for k in STAT_KEYS}

    def timed(self, resident, steps, warmup, flush, dist):
        torch = self.torch
        for _ in range(warmup):
            self.step(resident, flush)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        times, last = [], None
        for _ in range(steps):
            dt, last = self.step(resident, flush)
            times.append(dt)
        torch.cuda.synchronize()
        t = torch.tensor([sum(times)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, last


def bind(L):
    vp = C.c_void_p
    import agc_b200
    L.agcgpu_compressor_create.restype = C.c_int
    L.agcgpu_compressor_create.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                           C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_char_p, C.POINTER(vp)]
    L.agcgpu_compressor_add_samples_memory.restype = C.c_int
    L.agcgpu_compressor_add_samples_memory.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_char_p),
                                                       C.c_uint32, vp, C.POINTER(C.c_uint64), C.c_int]
    L.agcgpu_compressor_add_sample_files.restype = C.c_int
    L.agcgpu_compressor_add_sample_files.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32]
    L.agcgpu_compressor_close.restype = C.c_int; L.agcgpu_compressor_close.argtypes = [vp, C.c_uint32]
    L.agcgpu_compressor_last_error.restype = C.c_char_p; L.agcgpu_compressor_last_error.argtypes = [vp]
    L.agcgpu_compressor_last_stats.restype = C.c_int; L.agcgpu_compressor_last_stats.argtypes = [C.POINTER(agc_b200.Stats)]


def main():
    # stdout carries exactly ONE line (the JSON): whatever libraries print there (NCCL's version banner ...) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workload and the LZ micro-batch")
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
    bind(L)
    comm_info = None
    if world > 1:
        from agc_b200 import dist as agc_dist
        comm_info = agc_dist.install_exchange(local_rank)          # NCCL communicator inside libagcgpu (C++), bootstrap over torch.distributed

    tmp = tempfile.mkdtemp(prefix=f"agcbench{rank}_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        run = Runner(L, args.workload, tmp, local_rank, rank, world)
        sampler = ClockSampler(local_rank)
        sampler.start()
        sec_res, st_res = run.timed(True, args.steps, args.warmup, flush, dist)
        sha_res = sha256_file(run.out_path) if rank == 0 else None
        sec_e2e, st_e2e = run.timed(False, args.steps, args.warmup, flush, dist)
        sha_e2e = sha256_file(run.out_path) if rank == 0 else None
        sampler.stop_flag = True
        sampler.join(timeout=2)

        if rank == 0:
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(peaks_path):
                peak = json.load(open(peaks_path))["hbm_gbs"]; peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
            else:
                peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
            # the reference binary on the WHOLE workload with every host core: the CPU baseline and the bit-exactness oracle of the run
            ref_dt, ref_sha, cores = run_reference_once(run.files, tmp, args.workload, "cpu")
            if sha_res != ref_sha or sha_e2e != ref_sha:
                raise SystemExit(f"bench.py: archive differs from the reference's (ours {sha_res} / {sha_e2e}, reference {ref_sha}): no valid number")
            cpu = {"value": run.total / ref_dt / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "reference",
                   "sample": f"full workload ({run.total} bases), oracle/_ref/agc create -t {cores}, files in tmpfs, one run"}
            total = run.total
            lz_ms, lz_b = st_res["lz_kernel_ms_total"], st_res["lz_alg_bytes_total"]
            lz_gbs = lz_b / (lz_ms * 1e-3) / 1e9 if lz_ms else 0.0
            sc_ms, sc_b = st_res["scan_kernel_ms_total"], st_res["scan_bytes_total"]
            sc_gbs = sc_b / (sc_ms * 1e-3) / 1e9 if sc_ms else 0.0
            prof = os.path.join(ROOT, "profiles", "r02_lz_traffic.json")      # dram__bytes_read + dram__bytes_write of one ncu --set full capture
            traffic = json.load(open(prof)) if os.path.exists(prof) else {}
            line = {"metric": "input Gbp/s (agc create, bit-exact .agc)", "value": total / sec_res / 1e9, "unit": "Gbp/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_res * 1e3, "higher_is_better": True,
                    "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                    "config": workload_config(args.workload, world),
                    "e2e": {"value": total / sec_e2e / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]),
                            "d2h_bytes_per_step": int(st_e2e["d2h_bytes"]), "ms_per_step": sec_e2e * 1e3,
                            "includes": "FASTA file read + parse, H2D/D2H, archive write"},
                    "gpu_launches": int(st_res["kernel_launches"]) * args.steps,
                    "archive_sha256": sha_e2e, "reference_sha256": ref_sha, "bit_exact": True,
                    "roofline": {"kernel": "k_lz_diag (+ k_lzc_parse / k_lzc_stitch for segments above 128 kb): LZ-diff encode, all launches of the step", "bound": "hbm", "achieved": lz_gbs, "peak": peak,
                                 "unit": "GB/s", "frac": lz_gbs / peak, "traffic": traffic.get("step_traffic_bytes_per_launch"), "peak_source": peak_src,
                                 "algorithmic_bytes_per_step": int(lz_b), "kernel_ms_per_step": lz_ms, "launches_per_step": int(st_res["lz_encode_launches"]),
                                 "segments_diagonal_kernel": int(st_res["lz_diag_segments"]), "segments_chunk_parallel": int(st_res["lz_chunk_segments"]),
                                 "segments_sequential_kernel": int(st_res["lz_sequential_segments"]),
                                 "note": "-a mode: one device batch per sample, so a launch holds ~90 segments (2.7 MB algorithmic) and lasts as long as its slowest warp: launch-latency bound; the HBM figure of the kernel is lz_kernel_hpp_like_batch (segments_chunk_parallel also counts the cost-vector requests of the missing-middle decisions)"},
                    "scan_roofline": {"kernel": "k_scan (splitter scan over 2-bit packed contigs)", "bound": "hbm", "achieved": sc_gbs, "peak": peak, "unit": "GB/s",
                                      "frac": sc_gbs / peak, "algorithmic_bytes_per_step": int(sc_b), "kernel_ms_per_step": sc_ms, "launches_per_step": int(st_res["scan_launches"])},
                    "residual_coder": {"kernel": "k_zstd / k_zstd_narrow (bit-exact zstd frames, one CTA per part)", "ms_per_step": float(st_res["zstd_kernel_ms"]),
                                       "input_bytes_per_step": int(st_res["zstd_input_mb"] * 1e6), "host_wait_ms_per_step": float(st_res["zstd_wait_ms"]),
                                       "share_of_step": float(st_res["zstd_wait_ms"] or st_res["zstd_kernel_ms"]) / (sec_res * 1e3),
                                       "note": "batches are queued behind the pipeline (agcgpu_zstd_submit); host_wait_ms = the part that did not overlap"},
                    "cpu_baseline": cpu, "clocks": sampler.summary()}
            if comm_info:
                from agc_b200 import dist as agc_dist
                line["comm"] = agc_dist.comm_stats()          # counters after the run: collectives issued by the data path
            if not args.no_extra and world == 1:
                line["lz_kernel_hpp_like_batch"] = lz_hpp_batch(local_rank, peak, traffic)
                # the same kernel where it is not launch-latency bound (one HPP-scale device batch), next to the step's own figure
                hb = line["lz_kernel_hpp_like_batch"]
                line["roofline"].update({"batch_achieved": hb["achieved"], "batch_frac": hb["frac"], "batch_traffic": hb["traffic"],
                                         "batch_workload": hb["workload"]})
                # the other BASELINE configurations that fit one GPU, measured the same way with fewer steps (the headline stays C3):
                # configs[1] and configs[3] at full size (2.25 Gbases -- the HPP-like regime, where the device has frames and segments
                # by the thousand).  A failure here is reported, it does not take the headline line with it; a mismatch of an
                # archive still aborts.
                line["residual_coder"]["cpu_zstd"] = cpu_zstd_leg(run, tmp)
                line["other_workloads"] = {}
                for other, (st, wu) in (("c2", (2, 1)), ("c4", (1, 1))):
                    if other == args.workload:
                        continue
                    try:
                        r2 = Runner(L, other, tmp, local_rank, rank, world)
                        s_res, st2 = r2.timed(True, st, wu, flush, dist)
                        s_e2e, _ = r2.timed(False, st, wu, flush, dist)
                        sha2 = sha256_file(r2.out_path)
                        rdt, rsha, _ = run_reference_once(r2.files, tmp, other, "cpu")
                    except SystemExit:
                        raise
                    except Exception as e:                                   # out of host memory for the 2.3 GB of FASTA, a full tmpfs ...
                        line["other_workloads"][other] = {"workload": WORKLOADS[other][0], "error": repr(e)[:200]}
                        continue
                    if sha2 != rsha:
                        raise SystemExit(f"bench.py: {other} archive differs from the reference's")
                    line["other_workloads"][other] = {"workload": WORKLOADS[other][0], "value": r2.total / s_res / 1e9, "e2e": r2.total / s_e2e / 1e9,
                                                      "unit": "Gbp/s", "steps": st, "warmup": wu, "ms_per_step": s_res * 1e3, "e2e_ms_per_step": s_e2e * 1e3,
                                                      "reference_cpu": r2.total / rdt / 1e9, "reference_ms": rdt * 1e3,
                                                      "residual_coder_ms": float(st2["zstd_kernel_ms"]), "lz_kernel_ms": float(st2["lz_kernel_ms_total"]),
                                                      "bit_exact": True}
                    del r2
                    shutil.rmtree(os.path.join(tmp, "data_" + other), ignore_errors=True)
            print(json.dumps(line))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


def cpu_lz_encode_leg(n_texts=32, reps=4):
    """BASELINE.md section 3 per-stage CPU number: the reference's own CLZDiff_V2::Encode (oracle/_ref/liblzdiff_ref.so, index prepared
    once) on segments of the same shape as the device batch, one thread"""
    so = os.path.join(ROOT, "oracle", "_ref", "liblzdiff_ref.so")
    if not os.path.exists(so):
        return None
    import gen_data
    L = C.CDLL(so)
    if not hasattr(L, "ref_lz_encode_many_ns"):
        return None
    L.ref_lz_encode_many_ns.restype = C.c_long
    L.ref_lz_encode_many_ns.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(1)
    ref = rng.integers(0, 4, 60031, dtype=np.uint8)
    texts = [gen_data.substitute(rng, ref, 0.001) for _ in range(n_texts)]
    cat = np.concatenate(texts); offs = np.zeros(n_texts + 1, np.int64); offs[1:] = np.cumsum([len(t) for t in texts])
    ob = C.c_long()
    ns = L.ref_lz_encode_many_ns(ref.ctypes.data, len(ref), cat.ctypes.data, offs.ctypes.data, n_texts, MML, reps, C.byref(ob))
    bases = n_texts * reps * 60031
    return {"value": bases / ns, "unit": "Gbase/s", "cores": 1, "kind": "reference",
            "sample": f"CLZDiff_V2::Encode of {n_texts} x 60 031-base segments (0.1% SNP) x {reps} against one prepared reference, single thread"}


def cpu_zstd_leg(run, tmp, budget_s=15.0):
    """BASELINE.md section 3 per-stage CPU number for the residual coder: ZSTD_compress of the reference's vendored libzstd
    (oracle/_ref/libzstd_ref.so) over the EXACT parts of the workload -- one more create with the facade's part dump switched on
    (parts are written to a file instead of being coded), then every part at its level on one host thread, until `budget_s`."""
    try:
        import agc_parts
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libzstd_ref.so")):
            return None
        L, vp = run.L, C.c_void_p
        dump = os.path.join(tmp, "parts.dump")
        h = vp()
        if L.agcgpu_compressor_create(os.path.join(tmp, "dump_out.agc").encode(), PACK, run.k, run.files[0].encode(), SEG, MML, 0, run.adaptive, 0, 1, 0.0,
                                      run.local_rank, dump.encode(), C.byref(h)):
            return None
        all_files = (C.c_char_p * len(run.files))(*[f.encode() for f in run.files])
        all_names = (C.c_char_p * len(run.files))(*[run.c_names[i] for i in range(run.n_names)])
        rc = L.agcgpu_compressor_add_sample_files(h, all_names, all_files, len(run.files), 1)
        rc = L.agcgpu_compressor_close(h, 1) or rc
        if rc:
            return None
        _, recs = agc_parts.read_dump(dump)
        tasks = [(int(lv), raw) for r in recs if r["kind"] != 9 for lv, raw in r["tasks"] if int(lv) in (13, 17, 18, 19)]
        os.remove(dump)
        tasks.sort(key=lambda t: -len(t[1]))
        total = sum(len(t[1]) for t in tasks)
        done_b, done_n, largest_s = 0, 0, 0.0
        t0 = time.perf_counter()
        for lv, raw in tasks:
            t1 = time.perf_counter()
            agc_parts.zstd_compress(bytes(raw), lv)
            dt = time.perf_counter() - t1
            if done_n == 0:
                largest_s = dt
            done_b += len(raw); done_n += 1
            if time.perf_counter() - t0 > budget_s:
                break
        el = time.perf_counter() - t0
        return {"value": done_b / el / 1e6, "unit": "MB/s", "cores": 1, "kind": "reference",
                "sample": f"ZSTD_compress (vendored libzstd 1.5.5) of the workload's own parts at their levels, largest first: {done_n} of {len(tasks)} parts, "
                          f"{done_b} of {total} bytes in {el:.1f} s on one thread; the largest part ({len(tasks[0][1])} B, level {tasks[0][0]}) took {largest_s * 1e3:.0f} ms"}
    except Exception as e:
        return {"error": repr(e)[:200]}


def lz_hpp_batch(device, peak, traffic):
    """the LZ-diff encode kernel on 16384 segments of 60 031 bases (0.1 % SNP, 256 reference segments; 0.98 Gbase, one HPP-scale device
    batch): CUDA-event time of the launch (k_lz_diag), L2 flushed before every launch (working set 0.5 GB > L2 anyway)"""
    import lz_hpp_bench
    r = lz_hpp_bench.run(16384, 0.001, 256, reps=5, device=device, flush=True)
    t = traffic.get("hpp_traffic_bytes_per_launch")
    return {"workload": f"{r['n_seg']} segments x {r['seg_len']} bases, 0.1% SNP, {r['groups']} reference segments (one HPP-scale device batch; working set >> L2, L2 flushed between launches)",
            "kernel": "k_lz_diag", "kernel_ms": r["kernel_ms"], "algorithmic_bytes_per_launch": r["alg_bytes"], "achieved": r["GBps"], "unit": "GB/s", "frac": r["GBps"] / peak,
            "Gbase_per_s": r["Gbase_per_s"], "segments_sequential_kernel": r["sequential_segments"], "traffic": t,
            "traffic_source": traffic.get("source"), "cpu_lz_encode": cpu_lz_encode_leg()}


if __name__ == "__main__":
    main()
