#!/usr/bin/env python3
"""Throughput of the general decode kernels on synthetic HCA v3.0 streams (tests/helpers/hca3gen.py): 4096 stereo streams
of 94 frames (joint stereo, HFR, noise fill), resident in HBM, CUDA-event time of the kernels. Not part of bench.py's
contract (the headline workload is v2.0); prints one line for DESIGN.md.   python tools/bench_v3.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import hca3gen  # noqa: E402
from pycricodecs_b200 import _lib, engine  # noqa: E402


def main():
    distinct = [hca3gen.stream(seed=500 + i, frames=94, frame_size=2048, rate=48000) for i in range(32)]
    n = 4096
    streams = [distinct[i % len(distinct)] for i in range(n)]
    blob = np.frombuffer(b"".join(streams), np.uint8)
    offs = np.zeros(n + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in streams])
    ctx = engine.Context(0)
    job = engine.Job(ctx, _lib.JOB_HCA_DECODE, blob, offs)
    for _ in range(3):
        job.run()
    ms = 0.0
    steps = 10
    for _ in range(steps):
        job.run()
        ms += ctx.last_kernel_ms
    ms /= steps
    out, status = job.download(np.empty(job.out_bytes, np.uint8))
    assert int((status != 0).sum()) == 0
    import oracle
    want = oracle.port().hca_decode(streams[5])[1]
    oo = job.out_offsets
    got = bytes(out[int(oo[5]):int(oo[6])])
    print(f"v3.0 general path: {n} streams x 94 frames, {ms:.2f} ms per batch, {job.units / (ms * 1e-3) / 1e6:.1f} M frames/s, parity {got == want}")


if __name__ == "__main__":
    main()
