"""world_size-2 gloo test of the N>1 plumbing (sample sharding + timing reductions) -- runs on CPU."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_cover_all_samples_once():
    sys.path.insert(0, ROOT)
    from agc_b200 import dist as d
    for n in (1, 2, 7, 1001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                sh = d.shard_samples(n, r, world)
                assert sh[0] == 0
                seen += sh[1:]
            assert sorted(seen) == list(range(1, n))


def test_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import sys, os
        sys.path.insert(0, "@ROOT@")
        import torch.distributed as dist
        from agc_b200 import dist as d
        rank, local, world = d.env()
        dist.init_process_group("gloo")
        sh = d.shard_samples(11, rank, world)
        assert d.max_over_ranks(1.0 + rank) == 2.0
        assert d.sum_over_ranks(len(sh) - 1) == 10
        dist.barrier()
        dist.destroy_process_group()
        print("ok%d" % rank, flush=True)
    """).replace("@ROOT@", ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ok0" in out.stdout and "ok1" in out.stdout


def _sharded_create(tmp, case, world, port, so, device, append_after=0, exchange="gloo"):
    """one archive from `world` ranks (agcgpu_set_exchange + torch.distributed all-gather over gloo); returns (archive bytes,
    reference archive bytes, per-rank residual-coder input in MB).  append_after = n: the reference creates a base from the
    first n files and the ranks extend it (`append`) with the rest."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_host_pipeline import collection, REF_AGC, append_flags
    files, flags = collection(case, tmp)
    ref = os.path.join(tmp, "ref.agc"); out = os.path.join(tmp, "our.agc")
    if append_after:
        base = os.path.join(tmp, "base.agc")
        subprocess.check_call([REF_AGC, "create", "-t", "4", "-o", base] + flags + files[:append_after], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.check_call([REF_AGC, "append", "-t", "4", "-o", ref] + append_flags(flags) + [base] + files[append_after:], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        files = [base] + files[append_after:]
    else:
        subprocess.check_call([REF_AGC, "create", "-t", "4", "-o", ref] + flags + files, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_create_worker.py"), so, out, str(device)] + flags + ["--"] + files,
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, AGC_EXCHANGE=exchange))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    import re
    mb = {int(m.group(1)): float(m.group(2)) for m in re.finditer(r"rank(\d+)/\d+ zstd_input_mb=([0-9.]+)", r.stdout)}
    assert sorted(mb) == list(range(world)), r.stdout
    return open(out, "rb").read(), open(ref, "rb").read(), mb


def test_sharded_create_gloo(tmp_path):
    """SURVEY 8e on the CPU: 2 and 3 ranks (host objects + the oracle-backed stand-in of the device ABI, tests/mock) write ONE
    archive, byte-identical to the reference's, and the residual-coder work is really split between the ranks."""
    import pytest
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_host_pipeline import REF_AGC, MOCK_DIR
    if not os.path.exists(REF_AGC):
        pytest.skip("reference binary not built")
    subprocess.check_call(["make", "-C", MOCK_DIR, "-j4"], stdout=subprocess.DEVNULL)
    so = os.path.join(MOCK_DIR, "libagcgpu_mock.so")
    # the third run takes the staging path of install_exchange (the one NCCL uses on a GPU box) with CPU tensors
    for case, world, port, app, exch in (("complex", 2, 29541, 0, "gloo"), ("adaptive", 3, 29542, 0, "gloo"), ("fallback", 2, 29543, 4, "staged-cpu")):
        tmp = os.path.join(str(tmp_path), case); os.makedirs(tmp)
        a, b, mb = _sharded_create(tmp, case, world, port, so, -1, append_after=app, exchange=exch)
        assert a == b, f"{case}: archive of {world} ranks differs from the reference's ({len(a)} vs {len(b)} bytes)"
        total = sum(mb.values())
        assert total > 0 and all(v > 0.5 * total / world for v in mb.values()), mb     # every rank coded its share of the parts
