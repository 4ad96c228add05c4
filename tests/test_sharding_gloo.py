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


GATHER_WORKER = textwrap.dedent("""
    import os, sys, json, hashlib
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from pycricodecs_b200 import sharding
    rank, world, device = sharding.init("gloo")
    equal = %r
    n = 23
    rng = np.random.default_rng(5)
    lens = np.full(n, 40, np.uint64) if equal else rng.integers(1, 90, n).astype(np.uint64)
    if equal:
        n = 24; lens = np.full(n, 40, np.uint64)
    offsets = np.zeros(n + 1, np.uint64); np.cumsum(lens, out=offsets[1:])
    blob = rng.integers(0, 256, int(offsets[-1]), dtype=np.uint8)
    keys = np.arange(n, dtype=np.uint64) * 3 + 1

    # stand-in for the engine: stream i becomes its bytes + key, twice (sizes known up front, like header-derived sizes)
    def sizes_of(piece, poff, keys=None, **kw):
        return np.diff(poff.astype(np.int64)) * 2
    def compute(piece, poff, out, keys=None, out_offsets=None, **kw):
        p = piece.numpy(); at = 0
        for i in range(len(poff) - 1):
            x = (p[int(poff[i]):int(poff[i + 1])].astype(np.int64) + int(keys[i])) %% 256
            y = np.concatenate([x, x]).astype(np.uint8)
            out[at:at + len(y)] = torch.from_numpy(y); at += len(y)
        st = -(np.asarray(keys) %% 5 == 0).astype(np.int64)
        return st
    mine_only = blob.copy()                       # a rank may only look at its own pieces
    chunks = 3
    owned = np.zeros(n, bool)
    for k in range(chunks):
        lo, hi = sharding.piece_ranges(n, world, chunks)[k * world + rank]
        owned[lo:hi] = True
    for i in range(n):
        if not owned[i]: mine_only[int(offsets[i]):int(offsets[i + 1])] = 0
    out, ooff, status = sharding.sharded_batch(0, mine_only, offsets, chunks=chunks, compute=compute, sizes_of=sizes_of, keys=keys)
    want = []
    for i in range(n):
        x = (blob[int(offsets[i]):int(offsets[i + 1])].astype(np.int64) + int(keys[i])) %% 256
        want.append(np.concatenate([x, x]).astype(np.uint8))
    ok = bool((out.numpy() == np.concatenate(want)).all()) and list(np.diff(ooff.astype(np.int64))) == [len(w) for w in want]
    ok = ok and list(status) == [-(int(k) %% 5 == 0) for k in keys]
    part, _, _ = sharding.sharded_batch(0, mine_only, offsets, chunks=chunks, compute=compute, sizes_of=sizes_of, keys=keys, gather=False)
    for i in range(n):
        if owned[i]:
            ok = ok and bool((part.numpy()[int(ooff[i]):int(ooff[i + 1])] == want[i]).all())
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print(json.dumps({"ok": all(flags), "world": world}))
    dist.destroy_process_group()
""")


def _run_two_ranks(tmp_path, text):
    script = tmp_path / "worker.py"
    script.write_text(text)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_two_rank_gather_ragged_pieces(tmp_path):
    """sharded_batch over gloo with a CPU stand-in for the engine: ragged streams -> padded all-gather + compaction;
    every rank ends with the whole packed output, global offsets and statuses."""
    assert _run_two_ranks(tmp_path, GATHER_WORKER % (ROOT, False)) == {"ok": True, "world": 2}


def test_two_rank_gather_equal_pieces_in_place(tmp_path):
    """Equal-size pieces: the all-gather runs in place in the final blob."""
    assert _run_two_ranks(tmp_path, GATHER_WORKER % (ROOT, True)) == {"ok": True, "world": 2}


def test_shard_ranges_cover_everything_once():
    from pycricodecs_b200 import sharding
    for n in (0, 1, 7, 8, 8192, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
