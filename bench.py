#!/usr/bin/env python3
"""Benchmark of the ADX / HCA hot path (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic 48 kHz stereo
streams (2 s each). Default workload = BASELINE.json configs[1]: 8192 x stereo
HCA v2.0 (keyless, quality High) decode on one GPU. `value` is frames/s with the
batch resident in HBM (kernels only, CUDA events on the engine's stream);
`e2e` is the same metric through the C-ABI batch call with pinned HOST buffers
(header parsing, planning, H2D, kernels, D2H all inside the timed region).

Weak scaling: with --gpus N every rank decodes its own 8192 streams (streams are
independent; there is no data-path collective), time = max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KEY = 0xCF222F1FE0748978
WORKLOADS = {
    # name: (kind, unit, description, algorithmic bytes per unit of the whole path)
    "hca_decode": "8192 x 48kHz stereo HCA v2.0 decode (keyless, High), 2 s streams",
    "hca_decrypt_decode": "encrypted HCA (type 56, key 0xCF222F1FE0748978) decrypt+decode, 8192 streams per GPU",
    "adx_encode": "8192 x 48kHz stereo ADX encode bitdepth=4 blocksize=18",
    "adx_decode": "8192 x 48kHz stereo ADX decode bitdepth=4 blocksize=18",
}


def rank_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ----------------------------------------------------------------- inputs
def _encode_unique(args):
    """Worker: synthesise stream `sid` and encode it with the CPU checker (input synthesis only)."""
    sid, what = args
    import oracle
    from pycricodecs_b200 import synth
    port = oracle.port()
    w = synth.wav(sid, 2)
    if what == "wav":
        return w
    if what == "adx":
        return port.adx_encode(w)[1]
    h = port.hca_encode(w, 1)[1]
    if what == "hca_enc":
        h = port.hca_crypt(h, 1, 56, KEY)[1]
    return h


def make_inputs(workload: str, streams: int, unique: int, rank: int):
    """Returns (list of unique stream bytes, tiling factor). Stream ids are offset per rank."""
    what = {"hca_decode": "hca", "hca_decrypt_decode": "hca_enc", "adx_encode": "wav", "adx_decode": "adx"}[workload]
    unique = min(unique, streams)
    ids = [(rank * streams + i, what) for i in range(unique)]
    if what == "wav":
        uniq = [_encode_unique(a) for a in ids]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32)) as pool:
            uniq = pool.map(_encode_unique, ids, chunksize=4)
    return uniq, (streams + unique - 1) // unique


def tile(uniq, streams):
    from pycricodecs_b200 import engine
    blob_u, off_u = engine.pack(uniq)
    reps = (streams + len(uniq) - 1) // len(uniq)
    sizes = np.diff(off_u)
    sizes_all = np.tile(sizes, reps)[:streams]
    offsets = np.zeros(streams + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(sizes_all, dtype=np.uint64)
    blob = np.tile(blob_u, reps)[: int(offsets[-1])] if streams % len(uniq) == 0 else np.concatenate(
        [blob_u] * (streams // len(uniq)) + [blob_u[: int(off_u[streams % len(uniq)])]])
    return np.ascontiguousarray(blob), offsets


# ------------------------------------------------------------ clock sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                smax = max(smax, float(p[2]))
                if t0 <= t <= t1:
                    sm.append(float(p[1]))
                    for nm, v in zip(names, p[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ ours
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(a):
    rank, local_rank, world = rank_info()
    import torch
    from pycricodecs_b200 import _lib, engine
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = engine.Context(local_rank)
    kind = {"hca_decode": _lib.JOB_HCA_DECODE, "hca_decrypt_decode": _lib.JOB_HCA_DECODE, "adx_encode": _lib.JOB_ADX_ENCODE,
            "adx_decode": _lib.JOB_ADX_DECODE}[a.workload]
    uniq, reps = make_inputs(a.workload, a.streams, a.unique, rank)
    blob, offsets = tile(uniq, a.streams)
    keys = np.full(a.streams, KEY, np.uint64) if a.workload == "hca_decrypt_decode" else None
    kw = dict(keys=keys)
    if kind == _lib.JOB_ADX_ENCODE:
        kw["adx"] = engine.adx_params()

    # pinned host buffers for the end-to-end leg
    pin_in = torch.empty(len(blob), dtype=torch.uint8, pin_memory=True)
    pin_in.numpy()[:] = blob
    job = engine.Job(ctx, kind, pin_in.numpy(), offsets, **kw)
    units = job.units
    out_bytes = job.out_bytes
    pin_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    launches0 = ctx.launches
    for _ in range(a.warmup):
        job.run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    launches1 = ctx.launches
    t0 = time.perf_counter()
    dev_ms = dom_ms = 0.0
    for _ in range(a.steps):
        job.run()
        dev_ms += ctx.last_kernel_ms
        dom_ms += ctx.last_dominant_ms
    barrier()
    t1 = time.perf_counter()
    launches = ctx.launches - launches1
    clocks = sampler.stop(t0, t1)
    out_check, status = job.download(pin_out.numpy())
    assert int((status != 0).sum()) == 0, "engine reported per-stream errors"
    # spot parity inside the bench: first unique stream against the CPU checker (not timed, not the product path)
    import oracle
    port = oracle.port()
    oo = job.out_offsets
    got0 = bytes(out_check[int(oo[0]):int(oo[1])])
    if a.workload == "adx_encode":
        want0 = port.adx_encode(uniq[0])[1]
    elif a.workload == "adx_decode":
        want0 = port.adx_decode(uniq[0])[1]
    else:
        want0 = port.hca_decode(uniq[0], KEY if keys is not None else 0)[1]
    parity = got0 == want0
    job.close()

    # ---- end to end through the C-ABI batch call, host buffers in, host buffers out
    L = _lib.lib()
    status_arr = np.zeros(a.streams, np.int32)
    out_off = np.ascontiguousarray(oo, dtype=np.uint64)
    adxp = engine.adx_params()

    def e2e_call():
        if kind == _lib.JOB_HCA_DECODE:
            rc = L.cri_hca_decode_batch(ctx.handle, pin_in.data_ptr(), offsets.ctypes.data, a.streams,
                                        keys.ctypes.data if keys is not None else None, None, pin_out.data_ptr(),
                                        out_off.ctypes.data, status_arr.ctypes.data)
        elif kind == _lib.JOB_ADX_ENCODE:
            rc = L.cri_adx_encode_batch(ctx.handle, pin_in.data_ptr(), offsets.ctypes.data, a.streams, ctypes.byref(adxp),
                                        pin_out.data_ptr(), out_off.ctypes.data, status_arr.ctypes.data)
        else:
            rc = L.cri_adx_decode_batch(ctx.handle, pin_in.data_ptr(), offsets.ctypes.data, a.streams, pin_out.data_ptr(),
                                        out_off.ctypes.data, status_arr.ctypes.data)
        ctx.check(rc)

    e2e_call()  # warm-up (allocator, page mapping)
    barrier()
    e0 = time.perf_counter()
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    for _ in range(e2e_steps):
        e2e_call()
    barrier()
    e2e_s = (time.perf_counter() - e0) / e2e_steps

    ms = dev_ms / a.steps
    wall_ms = (t1 - t0) * 1e3 / a.steps
    dom = dom_ms / a.steps
    if dist is not None:
        t = torch.tensor([ms, wall_ms, e2e_s, dom], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall_ms, e2e_s, dom = [float(x) for x in t.tolist()]
        u = torch.tensor([units], device="cuda", dtype=torch.float64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        total_units = float(u.item())
    else:
        total_units = float(units)

    if rank == 0:
        peak, peak_src = peaks()
        is_hca = a.workload.startswith("hca")
        frame_bytes = int(np.diff(offsets)[0] - 96) // 94 if is_hca else 18
        # algorithmic (compulsory) bytes per unit: compressed bytes + PCM16 bytes, once each
        path_bytes = (frame_bytes + 4096) if is_hca else 82
        # dominant kernel: HCA = IMDCT transform kernel (reads int16 spectra 4096 B + gains 1024 B, writes PCM 4096 B
        # per stereo frame); ADX = the single encode/decode kernel (18 B + 64 B per block)
        dom_bytes = 9216 if is_hca else 82
        per_rank_units = total_units / world
        achieved = per_rank_units * dom_bytes / (dom * 1e-3) / 1e9 if dom > 0 else None
        line = {
            "metric": "audio frames/sec (48kHz stereo) HCA-decode + ADX-encode at 1/2/4/8 B200 vs CPU ref",
            "value": total_units / (ms * 1e-3),
            "unit": "HCA frames/s (1024 samples x 2 ch)" if is_hca else "ADX blocks/s (32 samples x 1 ch)",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "wall_ms_per_step": wall_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (no FMA, bit-exact) + int16/u8 bitstream" if is_hca else "int32",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[a.workload], "name": a.workload, "streams_per_gpu": a.streams,
                       "unique_streams_per_gpu": len(uniq), "tiling": f"{len(uniq)} unique synthetic streams tiled x{reps}",
                       "stream_seconds": 2.0, "sample_rate": 48000, "channels": 2,
                       "input_synthesis": "pycricodecs_b200.synth PCM; HCA/ADX inputs encoded once, untimed, by the CPU checker",
                       "l2": "inputs and outputs exceed the 126 MB L2 (no flush needed)",
                       "in_bytes_per_gpu": int(offsets[-1]), "out_bytes_per_gpu": int(out_bytes)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "traffic": None,
                         "kernel": "hca_imdct_kernel" if is_hca else f"adx_{a.workload.split('_')[1]}_fast_kernel",
                         "kernel_ms": dom, "algorithmic_bytes_per_unit": dom_bytes, "peak_source": peak_src,
                         "whole_path": {"algorithmic_bytes_per_unit": path_bytes,
                                        "achieved_gbs": per_rank_units * path_bytes / (ms * 1e-3) / 1e9,
                                        "frac": per_rank_units * path_bytes / (ms * 1e-3) / 1e9 / peak}},
            "e2e": {"value": total_units / e2e_s, "unit": "HCA frames/s" if is_hca else "ADX blocks/s",
                    "h2d_bytes_per_step": int(offsets[-1]), "d2h_bytes_per_step": int(out_bytes), "ms_per_step": e2e_s * 1e3,
                    "api": "cri_*_batch (C-ABI), pinned host buffers, parse+plan+H2D+kernels+D2H"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_spot_check": bool(parity),
        }
        if not a.no_cpu and world >= 1:
            line["cpu_baseline"] = cpu_baseline(a.workload, uniq, keys is not None, a.cpu_seconds)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------- CPU baselines
def _cpu_one(fn_name, data, key):
    import oracle
    impl = oracle.ref() if oracle.have_ref() else oracle.port()
    if fn_name == "hca_decode":
        r = impl.hca_decode(data, key)
    elif fn_name == "adx_encode":
        r = impl.adx_encode(data)
    else:
        r = impl.adx_decode(data)
    return 1


def cpu_kind():
    import oracle
    return "reference" if oracle.have_ref() else "port"


def units_of(workload, data):
    if workload.startswith("hca"):
        return int.from_bytes(data[16:20], "big")          # frames
    if workload == "adx_encode":
        return (len(data) - 44) // 2 // 32                  # blocks over all channels
    return int.from_bytes(data[12:16], "big") * data[7] // 32


def cpu_baseline(workload, uniq, keyed, seconds):
    """Single-thread CPU baseline on a bounded sample (rank 0 only)."""
    fn = "hca_decode" if workload.startswith("hca") else workload
    key = KEY if keyed else 0
    t0 = time.perf_counter()
    n = units = 0
    while time.perf_counter() - t0 < seconds:
        d = uniq[n % len(uniq)]
        _cpu_one(fn, d, key)
        units += units_of(workload, d)
        n += 1
    dt = time.perf_counter() - t0
    is_hca = workload.startswith("hca")
    return {"value": units / dt, "unit": "HCA frames/s" if is_hca else "ADX blocks/s", "cores": 1, "kind": cpu_kind(),
            "sample": f"{n} streams of 2 s decoded/encoded back to back on one host thread in {dt:.1f} s"}


def _ref_worker(args):
    fn, data, key, reps = args
    for _ in range(reps):
        _cpu_one(fn, data, key)
    return reps


def run_reference(a):
    rank, _, world = rank_info()
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    uniq, _ = make_inputs(a.workload, a.streams, min(a.unique, 64), 0)
    fn = "hca_decode" if a.workload.startswith("hca") else a.workload
    key = KEY if a.workload == "hca_decrypt_decode" else 0
    per_step = cores * 24                                       # streams per step: a bounded sample of the workload
    tasks = [(fn, uniq[i % len(uniq)], key, 1) for i in range(per_step)]
    units_step = sum(units_of(a.workload, t[1]) for t in tasks)
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(1, a.warmup)):
            pool.map(_ref_worker, tasks, chunksize=4)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_ref_worker, tasks, chunksize=4)
        dt = (time.perf_counter() - t0) / a.steps
    is_hca = a.workload.startswith("hca")
    unit = "HCA frames/s (1024 samples x 2 ch)" if is_hca else "ADX blocks/s (32 samples x 1 ch)"
    v = units_step / dt
    print(json.dumps({
        "impl": "reference", "metric": "audio frames/sec (48kHz stereo) HCA-decode + ADX-encode at 1/2/4/8 B200 vs CPU ref",
        "value": v, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 scalar SSE2 (no FMA)" if is_hca else "int32",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[a.workload], "name": a.workload, "streams_per_step": per_step,
                   "note": "the reference's own CPU implementation (oracle/_ref = unmodified CriCodecs C++), one process per host core"},
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": cpu_kind(),
                         "sample": f"{per_step} streams of 2 s per step over a {cores}-process pool"},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hca_decode", choices=list(WORKLOADS))
    ap.add_argument("--streams", type=int, default=8192)
    ap.add_argument("--unique", type=int, default=512)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
