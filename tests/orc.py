"""ctypes bindings of the C oracle (oracle/agc_oracle.c) and of the reference shims (oracle/_ref/*.so).
TEST INFRASTRUCTURE ONLY -- the product package agc_b200 never imports this module."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORC = None
_REF = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def _p8(a):
    return a.ctypes.data_as(u8p)


class Cut(C.Structure):
    _fields_ = [("start", C.c_uint64), ("len", C.c_uint64),
                ("front_dir", C.c_uint64), ("front_rc", C.c_uint64),
                ("back_dir", C.c_uint64), ("back_rc", C.c_uint64),
                ("has_front", C.c_uint32), ("has_back", C.c_uint32)]


def oracle():
    global _ORC
    if _ORC is None:
        so = os.path.join(ROOT, "oracle", "libagc_oracle.so")
        src = os.path.join(ROOT, "oracle", "agc_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.orc_murmur64.restype = C.c_uint64; L.orc_murmur64.argtypes = [C.c_uint64]
        L.orc_murmur_pair64.restype = C.c_uint64; L.orc_murmur_pair64.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_preprocess.restype = C.c_uint64; L.orc_preprocess.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_reverse_complement.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_scan_contig.restype = C.c_uint64
        L.orc_scan_contig.argtypes = [u8p, C.c_uint64, C.c_uint32, u64p, C.c_uint64, C.POINTER(Cut)]
        L.orc_enumerate_kmers.restype = C.c_uint64; L.orc_enumerate_kmers.argtypes = [u8p, C.c_uint64, C.c_uint32, u64p]
        L.orc_determine_splitters.restype = C.c_uint64
        L.orc_determine_splitters.argtypes = [u8p, u64p, C.c_uint32, C.c_uint32, C.c_uint64, u64p, u64p, u64p]
        L.orc_lz_prepare.restype = C.c_void_p; L.orc_lz_prepare.argtypes = [u8p, C.c_uint32, C.c_uint32]
        L.orc_lz_free.argtypes = [C.c_void_p]
        L.orc_lz_ht_size.restype = C.c_uint64; L.orc_lz_ht_size.argtypes = [C.c_void_p]
        L.orc_lz_is_short.restype = C.c_int; L.orc_lz_is_short.argtypes = [C.c_void_p]
        L.orc_lz_get_ht.argtypes = [C.c_void_p, u32p]
        L.orc_lz_encode.restype = C.c_uint64; L.orc_lz_encode.argtypes = [C.c_void_p, u8p, C.c_uint32, C.POINTER(u8p)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_lz_estimate.restype = C.c_uint64; L.orc_lz_estimate.argtypes = [C.c_void_p, u8p, C.c_uint32, C.c_uint32]
        L.orc_lz_cost_vector.restype = C.c_uint64; L.orc_lz_cost_vector.argtypes = [C.c_void_p, u8p, C.c_uint32, C.c_int, u32p]
        L.orc_lz_decode.restype = C.c_uint64; L.orc_lz_decode.argtypes = [u8p, C.c_uint32, u8p, C.c_uint64, C.c_uint32, u8p]
        L.orc_find_new_splitters.restype = C.c_uint64
        L.orc_find_new_splitters.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint64, u64p, C.c_uint64, u64p]
        L.orc_filtered_kmers.restype = C.c_uint64
        L.orc_filtered_kmers.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint64, u64p, u64p, u8p]
        L.orc_find_splitters_pos.restype = C.c_uint64
        L.orc_find_splitters_pos.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint64, u64p, C.c_uint64, u64p, u64p, u8p]
        L.orc_bytes2tuples.restype = C.c_uint64; L.orc_bytes2tuples.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_ref_use_tuples.restype = C.c_int; L.orc_ref_use_tuples.argtypes = [u8p, C.c_uint64]
        _ORC = L
    return _ORC


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liblzdiff_ref.so"))


def ref():
    """The unmodified reference lz_diff.cpp behind oracle/ref_shim.cpp (built by oracle/Makefile.ref)."""
    global _REF
    if _REF is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblzdiff_ref.so"))
        L.ref_lz_encode.restype = C.c_long; L.ref_lz_encode.argtypes = [u8p, C.c_long, u8p, C.c_long, C.c_int, u8p, C.c_long]
        L.ref_lz_estimate.restype = C.c_long; L.ref_lz_estimate.argtypes = [u8p, C.c_long, u8p, C.c_long, C.c_int, C.c_uint]
        L.ref_lz_cost_vector.restype = C.c_long; L.ref_lz_cost_vector.argtypes = [u8p, C.c_long, u8p, C.c_long, C.c_int, C.c_int, u32p]
        L.ref_lz_decode.restype = C.c_long; L.ref_lz_decode.argtypes = [u8p, C.c_long, u8p, C.c_long, C.c_int, u8p, C.c_long]
        _REF = L
    return _REF


# ---- numpy-friendly wrappers -------------------------------------------------------------------
def preprocess(raw: bytes) -> np.ndarray:
    a = np.frombuffer(raw, dtype=np.uint8).copy()
    out = np.empty(len(a) + 1, np.uint8)
    n = oracle().orc_preprocess(_p8(a), len(a), _p8(out))
    return out[:n].copy()


def revcomp(codes: np.ndarray) -> np.ndarray:
    codes = np.ascontiguousarray(codes, np.uint8)
    out = np.empty(len(codes) + 1, np.uint8)
    oracle().orc_reverse_complement(_p8(codes), len(codes), _p8(out))
    return out[:len(codes)].copy()


def scan_contig(codes, k, splitters_sorted):
    codes = np.ascontiguousarray(codes, np.uint8)
    spl = np.ascontiguousarray(splitters_sorted, np.uint64)
    sp = spl.ctypes.data_as(u64p)
    n = oracle().orc_scan_contig(_p8(codes), len(codes), k, sp, len(spl), None)
    cuts = (Cut * max(int(n), 1))()
    oracle().orc_scan_contig(_p8(codes), len(codes), k, sp, len(spl), cuts)
    return [cuts[i] for i in range(n)]


def determine_splitters(contigs, k, segment_size):
    cat = np.concatenate([np.ascontiguousarray(c, np.uint8) for c in contigs]) if contigs else np.zeros(0, np.uint8)
    offs = np.zeros(len(contigs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(c) for c in contigs])
    out = np.empty(len(cat) + 2 * len(contigs) + 2, np.uint64)
    sing = np.empty(len(cat) + 2, np.uint64)
    ns = C.c_uint64(0)
    n = oracle().orc_determine_splitters(_p8(cat), offs.ctypes.data_as(u64p), len(contigs), k, segment_size,
                                         out.ctypes.data_as(u64p), sing.ctypes.data_as(u64p), C.byref(ns))
    return out[:n].copy(), sing[:ns.value].copy()


class LZ:
    def __init__(self, refcodes, mml):
        self.ref = np.ascontiguousarray(refcodes, np.uint8)
        self.mml = mml
        self.h = oracle().orc_lz_prepare(_p8(self.ref), len(self.ref), mml)

    def __del__(self):
        try:
            oracle().orc_lz_free(self.h)
        except Exception:
            pass

    def ht(self):
        n = oracle().orc_lz_ht_size(self.h)
        out = np.empty(n, np.uint32)
        oracle().orc_lz_get_ht(self.h, out.ctypes.data_as(u32p))
        return out, bool(oracle().orc_lz_is_short(self.h))

    def encode(self, text) -> bytes:
        text = np.ascontiguousarray(text, np.uint8)
        p = u8p()
        n = oracle().orc_lz_encode(self.h, _p8(text), len(text), C.byref(p))
        r = C.string_at(p, n) if n else b""
        oracle().orc_free(p)
        return r

    def estimate(self, text, bound=0xFFFFFFFF) -> int:
        text = np.ascontiguousarray(text, np.uint8)
        return int(oracle().orc_lz_estimate(self.h, _p8(text), len(text), bound))

    def cost_vector(self, text, prefix) -> np.ndarray:
        text = np.ascontiguousarray(text, np.uint8)
        out = np.empty(len(text) + 1, np.uint32)
        n = oracle().orc_lz_cost_vector(self.h, _p8(text), len(text), int(prefix), out.ctypes.data_as(u32p))
        return out[:n].copy()


def lz_decode(refcodes, enc: bytes, mml, n_hint):
    refcodes = np.ascontiguousarray(refcodes, np.uint8)
    e = np.frombuffer(enc, np.uint8).copy() if enc else np.zeros(1, np.uint8)
    out = np.empty(n_hint + len(refcodes) + 64, np.uint8)
    n = oracle().orc_lz_decode(_p8(refcodes), len(refcodes), _p8(e), len(enc), mml, _p8(out))
    return out[:n].copy()


def bytes2tuples(codes) -> bytes:
    codes = np.ascontiguousarray(codes, np.uint8)
    out = np.empty(len(codes) + 2, np.uint8)
    n = oracle().orc_bytes2tuples(_p8(codes), len(codes), _p8(out))
    return out[:n].tobytes()


def ref_use_tuples(codes) -> bool:
    codes = np.ascontiguousarray(codes, np.uint8)
    return bool(oracle().orc_ref_use_tuples(_p8(codes), len(codes)))


# ---- reference (unmodified lz_diff.cpp) wrappers -----------------------------------------------
def ref_encode(refcodes, text, mml) -> bytes:
    r = np.ascontiguousarray(refcodes, np.uint8); t = np.ascontiguousarray(text, np.uint8)
    cap = 2 * len(t) + 64
    out = np.empty(cap, np.uint8)
    n = ref().ref_lz_encode(_p8(r), len(r), _p8(t), len(t), mml, _p8(out), cap)
    return out[:n].tobytes()


def ref_estimate(refcodes, text, mml, bound=0xFFFFFFFF) -> int:
    r = np.ascontiguousarray(refcodes, np.uint8); t = np.ascontiguousarray(text, np.uint8)
    return int(ref().ref_lz_estimate(_p8(r), len(r), _p8(t), len(t), mml, bound))


def ref_cost_vector(refcodes, text, mml, prefix) -> np.ndarray:
    r = np.ascontiguousarray(refcodes, np.uint8); t = np.ascontiguousarray(text, np.uint8)
    out = np.empty(len(t) + 1, np.uint32)
    n = ref().ref_lz_cost_vector(_p8(r), len(r), _p8(t), len(t), mml, int(prefix), out.ctypes.data_as(u32p))
    return out[:n].copy()


def enumerate_kmers(codes, k):
    codes = np.ascontiguousarray(codes, np.uint8)
    out = np.empty(len(codes) + 1, np.uint64)
    n = oracle().orc_enumerate_kmers(_p8(codes), len(codes), k, out.ctypes.data_as(u64p))
    return out[:n].copy()


def find_new_splitters(codes, k, segment_size, ref_kmers_sorted):
    """-a mode: CAGCCompressor::find_new_splitters for one contig; ref_kmers_sorted = every k-mer of the reference sample"""
    codes = np.ascontiguousarray(codes, np.uint8)
    rk = np.ascontiguousarray(ref_kmers_sorted, np.uint64)
    out = np.empty(len(codes) + 2, np.uint64)
    n = oracle().orc_find_new_splitters(_p8(codes), len(codes), k, segment_size, rk.ctypes.data_as(u64p), len(rk), out.ctypes.data_as(u64p))
    return out[:n].copy()


def filtered_kmers(codes, k, thr):
    """-f mode: (pos, kmer, is_dir_oriented, is_symmetric) of the k-mers passing kmer_filter_t, in position order"""
    codes = np.ascontiguousarray(codes, np.uint8)
    pos = np.empty(len(codes) + 1, np.uint64); km = np.empty(len(codes) + 1, np.uint64); fl = np.empty(len(codes) + 1, np.uint8)
    n = oracle().orc_filtered_kmers(_p8(codes), len(codes), k, int(thr), pos.ctypes.data_as(u64p), km.ctypes.data_as(u64p), _p8(fl))
    return [(int(pos[i]), int(km[i]), int(fl[i] & 1), int(fl[i] >> 1)) for i in range(n)]


def find_splitters_pos(codes, k, segment_size, cand_sorted):
    """find_splitters_in_contig with positions: (pos, kmer, is_last) in the order they were found"""
    codes = np.ascontiguousarray(codes, np.uint8)
    cd = np.ascontiguousarray(cand_sorted, np.uint64)
    out = np.empty(len(codes) + 2, np.uint64); pos = np.empty(len(codes) + 2, np.uint64); last = np.empty(len(codes) + 2, np.uint8)
    n = oracle().orc_find_splitters_pos(_p8(codes), len(codes), k, segment_size, cd.ctypes.data_as(u64p), len(cd),
                                        out.ctypes.data_as(u64p), pos.ctypes.data_as(u64p), _p8(last))
    return [(int(pos[i]), int(out[i]), int(last[i])) for i in range(n)]
