"""GPU diagnostic: chunk records of the device build vs the host build of lz_chunk_core.cuh on one (reference, text) pair"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, subprocess
import agc_b200, orc
subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "lzc_host"), "-s"])
H = C.CDLL(os.path.join(ROOT, "tests", "lzc_host", "liblzc_host.so"))
u8p = C.POINTER(C.c_uint8); u32p = C.POINTER(C.c_uint32)
H.lzc_host_encode_rec.restype = C.c_long
H.lzc_host_encode_rec.argtypes = [u8p, C.c_uint, u8p, C.c_uint, u32p, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, u8p, C.c_uint, C.c_void_p, C.c_uint]
FIELDS = ["flags", "lit0", "first_p", "first_ts", "first_mp", "first_len", "bytes", "end_i", "end_np", "end_pred", "end_diag", "open_ts", "open_mp", "open_predb", "pad0", "pad1"]
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
mml = 20; m = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
ref = rng.integers(0, 4, m).astype(np.uint8)
t = ref.copy(); mk = rng.random(m) < 0.005; t[mk] = (t[mk] + 1) % 4
z = orc.LZ(ref, mml); exp = z.encode(t)
ht, short = z.ht()
cap = len(t) * 2 + 64; out = np.zeros(cap, np.uint8)
nch = (len(t) + 2047) // 2048
hrec = np.zeros((nch, 16), np.uint32)
r = H.lzc_host_encode_rec(t.ctypes.data_as(u8p), len(t), ref.ctypes.data_as(u8p), m, ht.ctypes.data_as(u32p), len(ht), int(short), mml, 0, 0,
                          out.ctypes.data_as(u8p), cap, hrec.ctypes.data_as(C.c_void_p), nch)
print("host:", r, out[:r].tobytes() == exp)
LET = np.frombuffer(b"ACGT", np.uint8)
dev = agc_b200.Device(k=21, min_match_len=mml)
dev.set_splitters(np.zeros(0, np.uint64))
dev.scan_contigs([LET[ref].tobytes(), LET[t].tobytes()])
dev.put_references([(0, 0, m, False, 16)])
enc = dev.lz_encode([(1, 0, len(t), False, 16)])[0]
print("device:", len(enc), enc == exp, "seq segs", dev.stats().lz_sequential_segments)
L = agc_b200.lib()
L.agcgpu_debug_lz_chunk_records.restype = C.c_int
L.agcgpu_debug_lz_chunk_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
drec = np.zeros((nch + 4, 16), np.uint32); nn = C.c_uint64(0)
rc = L.agcgpu_debug_lz_chunk_records(dev.h, drec.ctypes.data_as(C.c_void_p), drec.nbytes, C.byref(nn))
print("records on device:", nn.value, "rc", rc)
for k in range(nch):
    if not np.array_equal(hrec[k, :14], drec[k, :14]):
        print("chunk", k, "differs")
        for i, f in enumerate(FIELDS[:14]):
            if hrec[k, i] != drec[k, i]: print(f"   {f}: host {hrec[k, i]} device {drec[k, i]}")
print("expected head:", exp[:60]); print("device   head:", enc[:60])
