"""CPU-only, world_size 2 over gloo: the multi-GPU plumbing of bench.py —
page sharding without overlap or gaps, barrier, MAX-over-ranks timing."""
import os
import socket
import subprocess
import sys
import textwrap

from unpaper_gpu_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        from unpaper_gpu_b200 import shard
        rank, world, local = shard.env_rank()
        dist.init_process_group("gloo")
        lo, hi = shard.shard_range(4096, rank, world)
        dist.barrier()
        t = shard.max_over_ranks([10.0 + rank, float(hi - lo)])
        n = shard.sum_over_ranks([float(hi - lo)])
        assert t[0] == 10.0 + world - 1, t
        assert n[0] == 4096.0, n
        if rank == 0:
            print("OK", lo, hi, t, n)
        dist.destroy_process_group()
    """))
    port = _free_port()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK 0 2048" in out.stdout
