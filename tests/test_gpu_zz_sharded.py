"""SURVEY 8e on the device: two ranks (two processes, two contexts -- both on cuda:0, so the test also runs on a one-GPU box;
the exchange step goes over gloo) write ONE archive through libagcgpu.so, byte-identical to the reference's.  On a multi-GPU
box test_two_gpus_nccl runs one GPU per rank with the NCCL communicator inside the library (agc_b200.dist.install_exchange)."""
import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dist_gloo import _sharded_create
from test_host_pipeline import REF_AGC

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(REF_AGC), reason="reference binary not built (make -f oracle/Makefile.ref)")
@pytest.mark.parametrize("case", ["complex", "adaptive"])
def test_two_ranks_one_archive(tmp_path, case):
    import agc_b200
    a, b, mb = _sharded_create(str(tmp_path), case, 2, 29551 if case == "complex" else 29552, agc_b200.lib_path(), 0)
    assert a == b, f"{case}: archive of 2 ranks differs from the reference's ({len(a)} vs {len(b)} bytes)"
    total = sum(mb.values())
    assert total > 0 and all(v > 0.25 * total for v in mb.values()), mb


def _n_gpus():
    try:
        import subprocess
        return len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
    except Exception:
        return 0


@pytest.mark.skipif(not os.path.exists(REF_AGC), reason="reference binary not built (make -f oracle/Makefile.ref)")
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs (NCCL cannot put two ranks on one device)")
def test_two_gpus_nccl(tmp_path):
    """one GPU per rank, the exchange step = ncclAllGather between device buffers from C++ (agc_b200/csrc/comm.cu)"""
    import agc_b200
    a, b, mb = _sharded_create(str(tmp_path), "complex", 2, 29553, agc_b200.lib_path(), 0, exchange="nccl")
    assert a == b
    total = sum(mb.values())
    assert total > 0 and all(v > 0.25 * total for v in mb.values()), mb
