"""Worker of the world-size-N `create` tests (launched with torch.distributed.run): every rank runs the same facade calls; the
exchange step is torch.distributed's all-gather (gloo).  argv: <library .so> <out.agc> <device or -1> <flags...> -- <files...>
A first "file" that ends in .agc is the archive to extend: the ranks then run `append` instead of `create`."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist
from agc_b200 import dist as adist
import agc_b200

so, out, device = sys.argv[1], sys.argv[2], int(sys.argv[3])
rest = sys.argv[4:]
flags, files = rest[:rest.index("--")], rest[rest.index("--") + 1:]
opt = {"-k": 31, "-l": 20, "-s": 60000, "-b": 50}
i = 0
adaptive = conc = 0
frac = 0.0
while i < len(flags):
    if flags[i] == "-a":
        adaptive = 1
    elif flags[i] == "-c":
        conc = 1
    elif flags[i] == "-f":
        frac = min(float(flags[i + 1]), 0.05); i += 1
    else:
        opt[flags[i]] = int(flags[i + 1]); i += 1
    i += 1

# AGC_EXCHANGE=nccl: one GPU per rank, NCCL communicator inside the library, device-to-device all-gathers from C++;
# AGC_EXCHANGE=staged-cpu / default: the callback entry (agcgpu_set_exchange) over gloo (CPU suite)
mode = os.environ.get("AGC_EXCHANGE", "gloo")
if mode == "nccl":
    import torch
    device = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(device)
    dist.init_process_group("nccl")
    L = agc_b200.lib()                                   # the NCCL communicator lives inside libagcgpu (agc_b200/csrc/comm.cu)
    info = adist.install_exchange(device)
    rank, world = info["rank"], info["comm_nranks_seen"]
else:
    dist.init_process_group("gloo")
    L = C.CDLL(so)
    rank, world = adist.install_exchange_callback(L, device="cpu" if mode == "staged-cpu" else None)
vp = C.c_void_p
L.agcgpu_compressor_create.restype = C.c_int
L.agcgpu_compressor_create.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                       C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_char_p, C.POINTER(vp)]
L.agcgpu_compressor_add_sample_files.restype = C.c_int
L.agcgpu_compressor_add_sample_files.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32]
L.agcgpu_compressor_close.restype = C.c_int; L.agcgpu_compressor_close.argtypes = [vp, C.c_uint32]
L.agcgpu_compressor_last_error.restype = C.c_char_p; L.agcgpu_compressor_last_error.argtypes = [vp]
L.agcgpu_compressor_last_stats.restype = C.c_int; L.agcgpu_compressor_last_stats.argtypes = [C.POINTER(agc_b200.Stats)]
L.agcgpu_compressor_append.restype = C.c_int
L.agcgpu_compressor_append.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_double, C.c_int, C.POINTER(vp)]
h = vp()
if files[0].endswith(".agc"):
    rc = L.agcgpu_compressor_append(files[0].encode(), out.encode(), 0, 1, conc, adaptive, 1, frac, max(device, 0), C.byref(h))
    files = files[1:]
else:
    rc = L.agcgpu_compressor_create(out.encode(), opt["-b"], opt["-k"], files[0].encode(), opt["-s"], opt["-l"], conc, adaptive, 0, 1, frac,
                                    max(device, 0), None, C.byref(h))
assert rc == 0, L.agcgpu_compressor_last_error(None)
names = [os.path.splitext(os.path.basename(f))[0].encode() for f in files]
cn = (C.c_char_p * len(files))(*names); cf = (C.c_char_p * len(files))(*[f.encode() for f in files])
rc = L.agcgpu_compressor_add_sample_files(h, cn, cf, len(files), 1)
assert rc == 0, L.agcgpu_compressor_last_error(h)
rc = L.agcgpu_compressor_close(h, 1)
assert rc == 0, L.agcgpu_compressor_last_error(None)
st = agc_b200.Stats()
L.agcgpu_compressor_last_stats(C.byref(st))
dist.barrier()
if mode == "nccl":
    print(f"rank{rank} comm {adist.comm_stats()}", flush=True)
print(f"rank{rank}/{world} zstd_input_mb={st.zstd_input_mb:.6f}", flush=True)
dist.destroy_process_group()
