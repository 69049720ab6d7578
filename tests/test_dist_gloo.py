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
