"""GPU, two ranks over NCCL (needs two GPUs: `gpurun --gpus 2`; skipped on a one-GPU box): the sharded decode with the
all-gather gives every rank the bytes the one-GPU call gives, for equal-size streams (in-place gather) and ragged ones
(padded gather + compaction). On one GPU the same entry point runs with world size 1."""
import json
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json, hashlib
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from pycricodecs_b200 import _lib, engine, sharding, synth
    rank, world, device = sharding.init()
    ctx = engine.Context(device.index)
    KEY = 0xCF222F1FE0748978
    results = {}
    for name, lens in (("equal", [9000] * 24), ("ragged", [3000 + 517 * (i %% 7) for i in range(29)])):
        wavs = [synth.wav(700 + i, 2, n) for i, n in enumerate(lens)]
        hcas = engine.hca_encode_batch(wavs, quality=1, ctx=ctx)
        hcas = engine.hca_crypt_batch(hcas, True, keys=KEY, ctx=ctx)
        blob, offsets = engine.pack(hcas)
        keys = np.full(len(hcas), KEY, np.uint64)
        t = {}
        out, ooff, status = sharding.sharded_batch(_lib.JOB_HCA_DECODE, blob, offsets, ctx, keys=keys, chunks=3, timings=t)
        want = engine.hca_decode_batch(hcas, keys=KEY, ctx=ctx)
        host = out.cpu().numpy()
        ok = not status.any() and int(ooff[-1]) == sum(len(w) for w in want)
        ok = ok and all(host[int(ooff[i]):int(ooff[i + 1])].tobytes() == want[i] for i in range(len(want)))
        # ADX encode through the same path (other job kind, other sizes)
        wblob, woff = engine.pack(wavs)
        aout, aoff, ast = sharding.sharded_batch(_lib.JOB_ADX_ENCODE, wblob, woff, ctx, chunks=2)
        awant = engine.adx_encode_batch(wavs, ctx=ctx)
        ahost = aout.cpu().numpy()
        ok = ok and all(ahost[int(aoff[i]):int(aoff[i + 1])].tobytes() == awant[i] for i in range(len(awant)))
        results[name] = bool(ok)
    flags = [None] * world
    if world > 1:
        dist.all_gather_object(flags, results)
    else:
        flags = [results]
    if rank == 0:
        print(json.dumps({"world": world, "ranks": flags}))
    if world > 1:
        dist.destroy_process_group()
""")


def _run(tmp_path, nproc):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_sharded_batch_one_rank(tmp_path):
    d = _run(tmp_path, 1)
    assert d == {"world": 1, "ranks": [{"equal": True, "ragged": True}]}


def test_sharded_decode_two_ranks_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    d = _run(tmp_path, 2)
    assert d == {"world": 2, "ranks": [{"equal": True, "ragged": True}] * 2}
