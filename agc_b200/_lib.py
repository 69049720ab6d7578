"""ctypes bindings of libagcgpu.so (C ABI declared in include/agcgpu.h)."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)


class AgcGpuError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("kmer_length", C.c_uint32), ("min_match_len", C.c_uint32), ("segment_size", C.c_uint32),
                ("pack_cardinality", C.c_uint32), ("device", C.c_int32), ("flags", C.c_uint32)]


class Cut(C.Structure):
    _fields_ = [("contig", C.c_uint32), ("has_front", C.c_uint32), ("has_back", C.c_uint32), ("reserved", C.c_uint32),
                ("start", C.c_uint64), ("len", C.c_uint64), ("front_dir", C.c_uint64), ("front_rc", C.c_uint64),
                ("back_dir", C.c_uint64), ("back_rc", C.c_uint64)]


class SegReq(C.Structure):
    _fields_ = [("contig", C.c_uint32), ("is_rc", C.c_uint32), ("start", C.c_uint64), ("len", C.c_uint32),
                ("group_id", C.c_uint32), ("bound", C.c_uint32), ("reserved", C.c_uint32)]


class SplitReq(C.Structure):
    _fields_ = [("contig", C.c_uint32), ("len", C.c_uint32), ("start", C.c_uint64), ("group1", C.c_uint32), ("group2", C.c_uint32),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class Assign(C.Structure):
    _fields_ = [("key1", C.c_uint64), ("key2", C.c_uint64), ("group_id", C.c_int32), ("is_rc", C.c_uint32),
                ("klass", C.c_uint32), ("reserved", C.c_uint32)]


class FKmer(C.Structure):
    _fields_ = [("pos", C.c_uint64), ("kmer", C.c_uint64), ("is_dir_oriented", C.c_uint32), ("is_symmetric", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("device_bytes_in_use", C.c_uint64), ("lz_alg_bytes", C.c_uint64), ("last_lz_kernel_ms", C.c_float),
                ("last_scan_kernel_ms", C.c_float), ("zstd_kernel_ms", C.c_float), ("zstd_input_mb", C.c_float),
                ("lz_chunk_segments", C.c_uint64), ("lz_sequential_segments", C.c_uint64),
                ("lz_alg_bytes_total", C.c_uint64), ("scan_bytes_total", C.c_uint64), ("lz_kernel_ms_total", C.c_float),
                ("scan_kernel_ms_total", C.c_float), ("lz_encode_launches", C.c_uint32), ("scan_launches", C.c_uint32),
                ("zstd_wait_ms", C.c_float), ("lz_diag_segments", C.c_uint32)]


# every symbol include/agcgpu.h declares (tests/test_abi.py checks header <-> library <-> this list)
EXPORTED_SYMBOLS = [
    "agcgpu_create", "agcgpu_destroy", "agcgpu_last_error", "agcgpu_sync", "agcgpu_stream", "agcgpu_get_stats",
    "agcgpu_determine_splitters", "agcgpu_set_splitters", "agcgpu_scan_contigs", "agcgpu_scan_contigs_dev",
    "agcgpu_get_segment", "agcgpu_map_insert", "agcgpu_assign_cuts", "agcgpu_group_put_reference_batch",
    "agcgpu_group_put_reference", "agcgpu_group_get_index", "agcgpu_lz_encode_batch", "agcgpu_lz_estimate_batch",
    "agcgpu_lz_cost_vector", "agcgpu_lz_cost_split_batch", "agcgpu_debug_lz_chunk_records", "agcgpu_lz_encode_batch_sharded",
    "agcgpu_zstd_compress_batch_sharded", "agcgpu_comm_unique_id", "agcgpu_comm_init", "agcgpu_comm_destroy", "agcgpu_comm_get_stats",
    "agcgpu_comm_last_error", "agcgpu_comm_world", "agcgpu_comm_rank", "agcgpu_pack_ref_batch", "agcgpu_zstd_compress_batch",
    "agcgpu_find_new_splitters", "agcgpu_rescan_contigs", "agcgpu_filtered_kmers", "agcgpu_last_splitter_positions",
    "agcgpu_zstd_decompress_batch", "agcgpu_lz_decode_batch", "agcgpu_zstd_submit", "agcgpu_zstd_submit_parts", "agcgpu_zstd_collect", "agcgpu_host_alloc", "agcgpu_host_free",
]


def lib_path():
    return os.path.join(_HERE, "libagcgpu.so")


def lib():
    """Load libagcgpu.so; fails loudly when it has not been built (no fallback of any kind)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise AgcGpuError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). agc_b200 has no CPU fallback.")
    L = C.CDLL(p)
    vp = C.c_void_p
    L.agcgpu_create.restype = C.c_int; L.agcgpu_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.agcgpu_destroy.restype = None; L.agcgpu_destroy.argtypes = [vp]
    L.agcgpu_last_error.restype = C.c_char_p; L.agcgpu_last_error.argtypes = [vp]
    L.agcgpu_sync.restype = C.c_int; L.agcgpu_sync.argtypes = [vp]
    L.agcgpu_stream.restype = vp; L.agcgpu_stream.argtypes = [vp]
    L.agcgpu_get_stats.restype = C.c_int; L.agcgpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.agcgpu_determine_splitters.restype = C.c_int
    L.agcgpu_determine_splitters.argtypes = [vp, u8p, u64p, C.c_uint32, u64p, C.c_uint64, u64p]
    L.agcgpu_set_splitters.restype = C.c_int; L.agcgpu_set_splitters.argtypes = [vp, u64p, C.c_uint64]
    L.agcgpu_scan_contigs.restype = C.c_int
    L.agcgpu_scan_contigs.argtypes = [vp, u8p, u64p, C.c_uint32, u64p, C.POINTER(Cut), C.c_uint64, u64p]
    L.agcgpu_find_new_splitters.restype = C.c_int
    L.agcgpu_find_new_splitters.argtypes = [vp, u32p, C.c_uint32, u64p, C.c_uint64, u64p]
    L.agcgpu_lz_decode_batch.restype = C.c_int
    L.agcgpu_lz_decode_batch.argtypes = [vp, u32p, u8p, u64p, C.c_uint32, u8p, C.c_uint64, u64p]
    L.agcgpu_zstd_decompress_batch.restype = C.c_int
    L.agcgpu_zstd_decompress_batch.argtypes = [vp, u8p, u64p, C.c_uint32, u8p, C.c_uint64, u64p]
    L.agcgpu_filtered_kmers.restype = C.c_int
    L.agcgpu_filtered_kmers.argtypes = [vp, C.POINTER(SegReq), C.c_uint32, C.c_uint64, C.POINTER(FKmer), C.c_uint64, u64p]
    L.agcgpu_last_splitter_positions.restype = C.c_int
    L.agcgpu_last_splitter_positions.argtypes = [vp, u32p, u64p, u64p, u8p, C.c_uint64, u64p]
    L.agcgpu_rescan_contigs.restype = C.c_int; L.agcgpu_rescan_contigs.argtypes = [vp, C.POINTER(Cut), C.c_uint64, u64p]
    L.agcgpu_scan_contigs_dev.restype = C.c_int
    L.agcgpu_scan_contigs_dev.argtypes = [vp, vp, C.c_uint64, u64p, C.c_uint32, u64p, C.POINTER(Cut), C.c_uint64, u64p]
    L.agcgpu_get_segment.restype = C.c_int
    L.agcgpu_get_segment.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, u8p]
    L.agcgpu_map_insert.restype = C.c_int; L.agcgpu_map_insert.argtypes = [vp, u64p, u64p, i32p, C.c_uint64]
    L.agcgpu_assign_cuts.restype = C.c_int; L.agcgpu_assign_cuts.argtypes = [vp, C.POINTER(Cut), C.c_uint64, C.POINTER(Assign)]
    L.agcgpu_group_put_reference_batch.restype = C.c_int
    L.agcgpu_group_put_reference_batch.argtypes = [vp, C.POINTER(SegReq), C.c_uint32]
    L.agcgpu_group_put_reference.restype = C.c_int; L.agcgpu_group_put_reference.argtypes = [vp, C.c_uint32, u8p, C.c_uint32]
    L.agcgpu_group_get_index.restype = C.c_int; L.agcgpu_group_get_index.argtypes = [vp, C.c_uint32, u32p, C.c_uint64, u64p]
    L.agcgpu_lz_encode_batch.restype = C.c_int
    L.agcgpu_lz_encode_batch.argtypes = [vp, C.POINTER(SegReq), C.c_uint32, u8p, C.c_uint64, u64p]
    L.agcgpu_lz_estimate_batch.restype = C.c_int; L.agcgpu_lz_estimate_batch.argtypes = [vp, C.POINTER(SegReq), C.c_uint32, u32p]
    L.agcgpu_lz_cost_vector.restype = C.c_int; L.agcgpu_lz_cost_vector.argtypes = [vp, C.POINTER(SegReq), C.c_int, u32p]
    L.agcgpu_lz_cost_split_batch.restype = C.c_int; L.agcgpu_lz_cost_split_batch.argtypes = [vp, C.POINTER(SplitReq), C.c_uint32, u32p, u32p]
    L.agcgpu_pack_ref_batch.restype = C.c_int
    L.agcgpu_pack_ref_batch.argtypes = [vp, u32p, C.c_uint32, u8p, C.c_uint64, u64p, u8p]
    L.agcgpu_zstd_submit.restype = C.c_int; L.agcgpu_zstd_submit.argtypes = [vp, u8p, u64p, i32p, C.c_uint32]
    L.agcgpu_zstd_collect.restype = C.c_int; L.agcgpu_zstd_collect.argtypes = [vp, C.c_uint32, u8p, C.c_uint64, u64p]
    L.agcgpu_host_alloc.restype = vp; L.agcgpu_host_alloc.argtypes = [C.c_uint64, u64p]
    L.agcgpu_host_free.restype = None; L.agcgpu_host_free.argtypes = [vp, C.c_uint64]
    L.agcgpu_zstd_compress_batch.restype = C.c_int
    L.agcgpu_zstd_compress_batch.argtypes = [vp, u8p, u64p, i32p, C.c_uint32, u8p, C.c_uint64, u64p]
    _LIB = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


class Device:
    """One context = the device side of one CAGCCompressor (src/core/agc_compressor.h:540-764)."""

    def __init__(self, k=31, min_match_len=20, segment_size=60000, pack_cardinality=50, device=0, adaptive=False):
        self.L = lib()
        self.params = Params(k, min_match_len, segment_size, pack_cardinality, device, 1 if adaptive else 0)
        h = C.c_void_p()
        rc = self.L.agcgpu_create(C.byref(self.params), C.byref(h))
        if rc != 0:
            raise AgcGpuError(f"agcgpu_create failed ({rc}): {self.L.agcgpu_last_error(None).decode()}")
        self.h = h
        self.contig_len = np.zeros(0, np.uint64)

    def close(self):
        if getattr(self, "h", None):
            self.L.agcgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise AgcGpuError(f"libagcgpu error {rc}: {self.L.agcgpu_last_error(self.h).decode()}")

    # ---- helpers -------------------------------------------------------------------------------
    @staticmethod
    def _cat(contigs):
        offs = np.zeros(len(contigs) + 1, np.uint64)
        if contigs:
            offs[1:] = np.cumsum([len(c) for c in contigs])
        raw = np.frombuffer(b"".join(bytes(c) for c in contigs), np.uint8).copy() if offs[-1] else np.zeros(1, np.uint8)
        return raw, offs

    @staticmethod
    def _reqs(reqs):
        arr = (SegReq * max(len(reqs), 1))()
        for i, r in enumerate(reqs):
            contig, start, length, is_rc, group = r[:5]
            bound = r[5] if len(r) > 5 else 0xFFFFFFFF
            arr[i] = SegReq(contig, int(bool(is_rc)), start, length, group, bound, 0)
        return arr

    def stats(self):
        s = Stats()
        self._ck(self.L.agcgpu_get_stats(self.h, C.byref(s)))
        return s

    def stream(self):
        return self.L.agcgpu_stream(self.h)

    # ---- splitters -----------------------------------------------------------------------------
    def set_splitters(self, splitters):
        s = np.ascontiguousarray(splitters, np.uint64)
        self._ck(self.L.agcgpu_set_splitters(self.h, _p(s, u64p), len(s)))

    def determine_splitters(self, raw_contigs):
        raw, offs = self._cat(raw_contigs)
        cap = int(offs[-1]) // max(1, min(self.params.segment_size, 100)) + 4 * len(raw_contigs) + 64
        out = np.zeros(cap, np.uint64)
        n = C.c_uint64(0)
        self._ck(self.L.agcgpu_determine_splitters(self.h, _p(raw, u8p), _p(offs, u64p), len(raw_contigs), _p(out, u64p), cap, C.byref(n)))
        return out[:n.value].copy()

    # ---- scan ----------------------------------------------------------------------------------
    def scan_contigs(self, raw_contigs):
        """raw_contigs: list of bytes (FASTA bodies, newlines allowed). Returns list of Cut; contigs stay resident."""
        raw, offs = self._cat(raw_contigs)
        return self._scan(raw, offs, len(raw_contigs), None)

    def _scan(self, raw, offs, n, raw_dev):
        cap = int(offs[-1]) // 16 + 4 * n + 64
        while True:
            cuts = (Cut * cap)()
            lens = np.zeros(max(n, 1), np.uint64)
            nc = C.c_uint64(0)
            if raw_dev is None:
                rc = self.L.agcgpu_scan_contigs(self.h, _p(raw, u8p), _p(offs, u64p), n, _p(lens, u64p), cuts, cap, C.byref(nc))
            else:
                rc = self.L.agcgpu_scan_contigs_dev(self.h, raw_dev, int(offs[-1]), _p(offs, u64p), n, _p(lens, u64p), cuts, cap, C.byref(nc))
            if rc == -4 and nc.value > cap:
                cap = nc.value + 16
                continue
            self._ck(rc)
            break
        self.contig_len = lens[:n].copy()
        return [cuts[i] for i in range(nc.value)]

    def find_new_splitters(self, contigs):
        """-a mode: new splitters of the listed resident contigs (CAGCCompressor::find_new_splitters); sorted, unique"""
        c = np.ascontiguousarray(contigs, np.uint32)
        cap = int(sum(int(self.contig_len[i]) for i in c)) // max(1, self.params.segment_size) + 2 * len(c) + 16
        out = np.zeros(cap, np.uint64)
        n = C.c_uint64(0)
        self._ck(self.L.agcgpu_find_new_splitters(self.h, _p(c, u32p), len(c), _p(out, u64p), cap, C.byref(n)))
        return out[:n.value].copy()

    def filtered_kmers(self, ranges, threshold):
        """-f mode: k-mers of resident ranges [(contig, start, len)] that pass kmer_filter_t; list of lists of FKmer"""
        reqs = self._reqs([(c, s, l, False, 0) for c, s, l in ranges])
        offs = np.zeros(len(ranges) + 1, np.uint64)
        cap = 1 << 12
        while True:
            out = (FKmer * cap)()
            rc = self.L.agcgpu_filtered_kmers(self.h, reqs, len(ranges), int(threshold), out, cap, _p(offs, u64p))
            if rc == -4 and int(offs[-1]) > cap:
                cap = int(offs[-1]) + 16
                continue
            self._ck(rc)
            break
        return [[out[j] for j in range(int(offs[i]), int(offs[i + 1]))] for i in range(len(ranges))]

    def last_splitter_positions(self):
        """(contig, pos, kmer, is_last) of the splitters the last determine_splitters / find_new_splitters call found"""
        cap = 1 << 16
        c = np.zeros(cap, np.uint32); p = np.zeros(cap, np.uint64); k = np.zeros(cap, np.uint64); l = np.zeros(cap, np.uint8)
        n = C.c_uint64(0)
        self._ck(self.L.agcgpu_last_splitter_positions(self.h, _p(c, u32p), _p(p, u64p), _p(k, u64p), _p(l, u8p), cap, C.byref(n)))
        return [(int(c[i]), int(p[i]), int(k[i]), int(l[i])) for i in range(n.value)]

    def rescan_contigs(self):
        """scan of the resident batch again under the current splitter set (-a mode, hard_contigs stage)"""
        cap = int(self.contig_len.sum()) // 16 + 4 * len(self.contig_len) + 64
        cuts = (Cut * cap)()
        nc = C.c_uint64(0)
        self._ck(self.L.agcgpu_rescan_contigs(self.h, cuts, cap, C.byref(nc)))
        return [cuts[i] for i in range(nc.value)]

    def scan_contigs_dev(self, dev_ptr, offs):
        offs = np.ascontiguousarray(offs, np.uint64)
        return self._scan(None, offs, len(offs) - 1, C.c_void_p(dev_ptr))

    def get_segment(self, contig, start, length, is_rc=False):
        out = np.zeros(max(length, 1), np.uint8)
        self._ck(self.L.agcgpu_get_segment(self.h, contig, start, length, int(bool(is_rc)), _p(out, u8p)))
        return out[:length].copy()

    # ---- map / assign --------------------------------------------------------------------------
    def map_insert(self, keys1, keys2, groups):
        k1 = np.ascontiguousarray(keys1, np.uint64); k2 = np.ascontiguousarray(keys2, np.uint64)
        g = np.ascontiguousarray(groups, np.int32)
        self._ck(self.L.agcgpu_map_insert(self.h, _p(k1, u64p), _p(k2, u64p), _p(g, i32p), len(g)))

    def assign_cuts(self, cuts):
        arr = (Cut * max(len(cuts), 1))(*cuts)
        out = (Assign * max(len(cuts), 1))()
        self._ck(self.L.agcgpu_assign_cuts(self.h, arr, len(cuts), out))
        return [out[i] for i in range(len(cuts))]

    # ---- references / LZ -----------------------------------------------------------------------
    def put_references(self, reqs):
        """reqs: (contig, start, len, is_rc, group_id) of resident segments that become group references."""
        arr = self._reqs(reqs)
        self._ck(self.L.agcgpu_group_put_reference_batch(self.h, arr, len(reqs)))

    def put_reference_host(self, group_id, symbols):
        s = np.ascontiguousarray(symbols, np.uint8)
        buf = s if len(s) else np.zeros(1, np.uint8)
        self._ck(self.L.agcgpu_group_put_reference(self.h, group_id, _p(buf, u8p), len(s)))

    def get_index(self, group_id):
        n = C.c_uint64(0)
        self._ck(self.L.agcgpu_group_get_index(self.h, group_id, None, 0, C.byref(n)))
        out = np.zeros(n.value, np.uint32)
        self._ck(self.L.agcgpu_group_get_index(self.h, group_id, _p(out, u32p), n.value, C.byref(n)))
        return out

    def lz_encode(self, reqs):
        arr = self._reqs(reqs)
        cap = sum((r[2] * 3) // 2 + 32 for r in reqs) + 64
        out = np.zeros(cap, np.uint8)
        offs = np.zeros(len(reqs) + 1, np.uint64)
        self._ck(self.L.agcgpu_lz_encode_batch(self.h, arr, len(reqs), _p(out, u8p), cap, _p(offs, u64p)))
        return [out[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(reqs))]

    def lz_encode_raw(self, arr, n, out, offs):
        """Pre-marshalled variant for timing loops: arr = (SegReq*n), out/offs = numpy buffers."""
        self._ck(self.L.agcgpu_lz_encode_batch(self.h, arr, n, _p(out, u8p), out.size, _p(offs, u64p)))

    def lz_estimate(self, reqs):
        arr = self._reqs(reqs)
        out = np.zeros(max(len(reqs), 1), np.uint32)
        self._ck(self.L.agcgpu_lz_estimate_batch(self.h, arr, len(reqs), _p(out, u32p)))
        return out[:len(reqs)].copy()

    def lz_cost_vector(self, req, prefix_costs):
        arr = self._reqs([req])
        out = np.zeros(max(req[2], 1), np.uint32)
        self._ck(self.L.agcgpu_lz_cost_vector(self.h, arr, int(bool(prefix_costs)), _p(out, u32p)))
        return out[:req[2]].copy()

    def lz_cost_split(self, reqs):
        """reqs: (contig, start, len, group1, group2, flags) -> (best_pos[], best_sum[]) of agcgpu_lz_cost_split_batch"""
        arr = (SplitReq * max(len(reqs), 1))()
        for i, (c, st, ln, g1, g2, fl) in enumerate(reqs):
            arr[i].contig, arr[i].start, arr[i].len, arr[i].group1, arr[i].group2, arr[i].flags = c, st, ln, g1, g2, fl
        pos = np.zeros(max(len(reqs), 1), np.uint32); sm = np.zeros(max(len(reqs), 1), np.uint32)
        self._ck(self.L.agcgpu_lz_cost_split_batch(self.h, arr, len(reqs), _p(pos, u32p), _p(sm, u32p)))
        return pos[:len(reqs)].copy(), sm[:len(reqs)].copy()

    def zstd_compress(self, inputs, levels):
        """ZSTD_compressCCtx(level) of every input on the device -> list of frames"""
        offs = np.zeros(len(inputs) + 1, np.uint64)
        if inputs:
            offs[1:] = np.cumsum([len(x) for x in inputs])
        src = np.frombuffer(b"".join(inputs), np.uint8).copy() if offs[-1] else np.zeros(1, np.uint8)
        lv = np.ascontiguousarray(levels, np.int32)
        cap = int(offs[-1]) + int(offs[-1]) // 128 + 1024 * (len(inputs) + 1)
        dst = np.zeros(cap, np.uint8)
        doffs = np.zeros(len(inputs) + 1, np.uint64)
        self._ck(self.L.agcgpu_zstd_compress_batch(self.h, _p(src, u8p), _p(offs, u64p), _p(lv, i32p), len(inputs), _p(dst, u8p), cap, _p(doffs, u64p)))
        return [dst[int(doffs[i]):int(doffs[i + 1])].tobytes() for i in range(len(inputs))]

    def zstd_submit(self, inputs, levels):
        """queue a batch for the residual coder (agcgpu_zstd_submit): returns at once, the frames come from zstd_collect"""
        offs = np.zeros(len(inputs) + 1, np.uint64)
        if inputs:
            offs[1:] = np.cumsum([len(x) for x in inputs])
        src = np.frombuffer(b"".join(inputs), np.uint8).copy() if offs[-1] else np.zeros(1, np.uint8)
        lv = np.ascontiguousarray(levels, np.int32)
        self._ck(self.L.agcgpu_zstd_submit(self.h, _p(src, u8p), _p(offs, u64p), _p(lv, i32p), len(inputs)))
        self._submitted = getattr(self, "_submitted", []) + [len(x) for x in inputs]

    def zstd_collect(self):
        """frames of every batch submitted since the last collect, in submission order"""
        sizes = getattr(self, "_submitted", [])
        self._submitted = []
        n = len(sizes); tot = int(sum(sizes))
        cap = tot + tot // 128 + 1024 * (n + 1)
        dst = np.zeros(max(cap, 1), np.uint8)
        doffs = np.zeros(n + 1, np.uint64)
        self._ck(self.L.agcgpu_zstd_collect(self.h, n, _p(dst, u8p), cap, _p(doffs, u64p)))
        return [dst[int(doffs[i]):int(doffs[i + 1])].tobytes() for i in range(n)]

    def lz_decode(self, group_ids, deltas):
        """CLZDiff_V2::Decode of every delta against its group's resident reference -> list of symbol arrays"""
        g = np.ascontiguousarray(group_ids, np.uint32)
        offs = np.zeros(len(deltas) + 1, np.uint64)
        if deltas:
            offs[1:] = np.cumsum([len(x) for x in deltas])
        src = np.frombuffer(b"".join(deltas), np.uint8).copy() if offs[-1] else np.zeros(1, np.uint8)
        ooffs = np.zeros(len(deltas) + 1, np.uint64)
        out = np.zeros(1, np.uint8)
        rc = self.L.agcgpu_lz_decode_batch(self.h, _p(g, u32p), _p(src, u8p), _p(offs, u64p), len(deltas), _p(out, u8p), 0, _p(ooffs, u64p))
        if rc == -4:
            out = np.zeros(int(ooffs[-1]) + 1, np.uint8)
            rc = self.L.agcgpu_lz_decode_batch(self.h, _p(g, u32p), _p(src, u8p), _p(offs, u64p), len(deltas), _p(out, u8p), int(ooffs[-1]), _p(ooffs, u64p))
        self._ck(rc)
        return [out[int(ooffs[i]):int(ooffs[i + 1])].copy() for i in range(len(deltas))]

    def zstd_decompress(self, frames):
        """ZSTD_decompressDCtx of every frame on the device -> list of outputs (sizes come from the frame headers)"""
        offs = np.zeros(len(frames) + 1, np.uint64)
        if frames:
            offs[1:] = np.cumsum([len(x) for x in frames])
        src = np.frombuffer(b"".join(frames), np.uint8).copy() if offs[-1] else np.zeros(1, np.uint8)
        doffs = np.zeros(len(frames) + 1, np.uint64)
        dst = np.zeros(1, np.uint8)
        rc = self.L.agcgpu_zstd_decompress_batch(self.h, _p(src, u8p), _p(offs, u64p), len(frames), _p(dst, u8p), 0, _p(doffs, u64p))
        if rc == -4:                                     # sized by the first call
            dst = np.zeros(int(doffs[-1]) + 1, np.uint8)
            rc = self.L.agcgpu_zstd_decompress_batch(self.h, _p(src, u8p), _p(offs, u64p), len(frames), _p(dst, u8p), int(doffs[-1]), _p(doffs, u64p))
        self._ck(rc)
        return [dst[int(doffs[i]):int(doffs[i + 1])].tobytes() for i in range(len(frames))]

    def pack_refs(self, group_ids, cap):
        ids = np.ascontiguousarray(group_ids, np.uint32)
        out = np.zeros(cap + 64, np.uint8)
        offs = np.zeros(len(ids) + 1, np.uint64)
        use = np.zeros(max(len(ids), 1), np.uint8)
        self._ck(self.L.agcgpu_pack_ref_batch(self.h, _p(ids, u32p), len(ids), _p(out, u8p), out.size, _p(offs, u64p), _p(use, u8p)))
        return [(out[int(offs[i]):int(offs[i + 1])].tobytes(), bool(use[i])) for i in range(len(ids))]
