"""CPU, world_size 2 over gloo: the multi-GPU path shards independent streams by rank with no data-path
collective; rank-local unit counts add up and the max-over-ranks reduction used by bench.py works."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from pycricodecs_b200 import sharding
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [10 + (i * 7) %% 13 for i in range(37)]
    lo, hi = sharding.shard_range(len(sizes), rank, world)
    mine = sum(sizes[lo:hi])
    total = sharding.all_sum(float(mine))
    worst = sharding.all_max(float(rank + 1) * 1.5)
    ids = sharding.stream_ids(8, rank)
    if rank == 0:
        print(json.dumps({"total": total, "worst": worst, "expect": float(sum(sizes)), "lo": lo, "hi": hi, "ids0": ids[0]}))
    dist.destroy_process_group()
""")


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["total"] == d["expect"] and d["worst"] == 3.0 and d["lo"] == 0 and d["hi"] == 19


def test_shard_ranges_cover_everything_once():
    from pycricodecs_b200 import sharding
    for n in (0, 1, 7, 8, 8192, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
