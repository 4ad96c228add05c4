#!/usr/bin/env python3
"""Benchmark of the ADX / HCA hot path (contract: task statement; layout of the numbers: DESIGN.md section Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--scaling weak|strong]

A "step" is one pass of the hot path over one batch of synthetic 48 kHz stereo streams (2 s each, all distinct).
Default workload = BASELINE.json configs[1]: 8192 x stereo HCA v2.0 (keyless, quality High) decode per GPU.

  value        frames/s with the batch resident in HBM (kernels only, CUDA events on the engine's stream)
  roofline     the WHOLE PATH against the measured HBM peak: compulsory bytes (compressed bytes + PCM16 bytes, once each,
               SURVEY.md section 8d) x units / step time; `kernels` lists every kernel of the step with its own figure
  e2e          the same metric through the C-ABI batch call with pinned HOST buffers (header parsing, planning, H2D,
               kernels, D2H inside the timed region), next to the bare pinned-copy time of the same bytes measured in the
               same process (`pcie_ceiling_ms`: what cudaMemcpyAsync alone needs)
  e2e_device   the same call with DEVICE buffers on the caller's stream (cri_*_batch_dev): no PCIe for the payload
  gather       N > 1: the sharded product path (pycricodecs_b200.sharding.sharded_batch) without and with the NCCL
               all-gather that reassembles the output blob on every rank
  adx_encode, hca_decrypt_decode, hca_encode
               companion objects of the default workload: BASELINE.json configs[2], configs[3] (8192 encrypted streams
               per GPU, weak) and configs[4] (4096 WAV -> HCA streams in total, STRONG scaling over the N GPUs)

Inputs are synthesised by the product itself: PCM from pycricodecs_b200.synth (torch on the GPU), HCA / ADX streams by
this repo's own GPU encoders (bit-exact with the reference, tests/test_hca_encode_gpu.py), untimed. Weak scaling: every
rank works on its own streams (streams are independent; no data-path collective), time = max over ranks.
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
QUALITY = 1          # HCA quality of the synthetic streams: 0 Highest, 1 High (BASELINE configs), 2 Middle, 3 Low (--quality)
METRIC = "audio frames/sec (48kHz stereo) HCA-decode + ADX-encode at 1/2/4/8 B200 vs CPU ref"
WORKLOADS = {
    "hca_decode": "8192 x 48kHz stereo HCA v2.0 decode (keyless, High), 2 s streams",
    "hca_decrypt_decode": "encrypted HCA (type 56, key 0xCF222F1FE0748978) decrypt+decode, 8192 streams per GPU",
    "hca_decrypt": "encrypted HCA (type 56, key 0xCF222F1FE0748978) decrypt only (HcaCrypt), 8192 streams per GPU",
    "hca_encode": "48kHz stereo WAV -> HCA v2.0 encode (High), 2 s streams",
    "adx_encode": "8192 x 48kHz stereo ADX encode bitdepth=4 blocksize=18",
    "adx_decode": "8192 x 48kHz stereo ADX decode bitdepth=4 blocksize=18",
}
# kernels of each workload's step; the decode step is two kernels (device time is split at an event between them)
KERNELS = {
    "hca_decode": ("hca_unpack_fast_kernel", "hca_imdct_fast_kernel"), "hca_decrypt_decode": ("hca_unpack_fast_kernel", "hca_imdct_fast_kernel"),
    "hca_decrypt": ("hca_crypt_lut_kernel",), "hca_encode": ("hca_encode_kernel",),
    "adx_encode": ("adx_encode_fast_kernel",), "adx_decode": ("adx_decode_fast_kernel",),
}


def rank_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def is_hca(w):
    return w.startswith("hca")


def unit_of(w):
    return "HCA frames/s (1024 samples x 2 ch)" if is_hca(w) else "ADX blocks/s (32 samples x 1 ch)"


def job_kind(w):
    from pycricodecs_b200 import _lib
    return {"hca_decode": _lib.JOB_HCA_DECODE, "hca_decrypt_decode": _lib.JOB_HCA_DECODE, "hca_decrypt": _lib.JOB_HCA_CRYPT,
            "hca_encode": _lib.JOB_HCA_ENCODE, "adx_encode": _lib.JOB_ADX_ENCODE, "adx_decode": _lib.JOB_ADX_DECODE}[w]


def job_kwargs(w, streams):
    from pycricodecs_b200 import _lib, engine
    kind = job_kind(w)
    kw = dict(keys=np.full(streams, KEY, np.uint64) if w in ("hca_decrypt_decode", "hca_decrypt") else None)
    if kind in (_lib.JOB_ADX_ENCODE, _lib.JOB_HCA_ENCODE):
        kw["adx"] = engine.adx_params()
    if kind == _lib.JOB_HCA_ENCODE:
        kw["quality"] = QUALITY
    if kind == _lib.JOB_HCA_CRYPT:
        kw.update(encrypt=0, ciph_type=0)
    return kw


# ----------------------------------------------------------------- inputs
def make_wav_blob(streams, first_id, device):
    """(pinned uint8 tensor, offsets): `streams` distinct synthetic stereo WAV images, stream ids first_id .. first_id + streams - 1."""
    import torch
    from pycricodecs_b200 import synth
    n, ch = synth.DEFAULT_SAMPLES, 2
    hdr = np.frombuffer(synth.wav_header(ch, n), np.uint8)
    size = len(hdr) + n * ch * 2
    blob = torch.empty(streams * size, dtype=torch.uint8, pin_memory=True)
    view = blob.numpy().reshape(streams, size)
    view[:, : len(hdr)] = hdr
    chunk = 128
    for s0 in range(0, streams, chunk):
        ids = range(first_id + s0, first_id + min(s0 + chunk, streams))
        pcm = synth.pcm_batch_torch(ids, ch, n, device=device).cpu().numpy()
        view[s0:s0 + len(pcm), len(hdr):] = pcm.reshape(len(pcm), -1).view(np.uint8)
    offsets = (np.arange(streams + 1, dtype=np.uint64) * np.uint64(size))
    return blob, offsets


def run_job_to_pinned(ctx, kind, blob_np, offsets, **kw):
    """Run one engine job and return (pinned uint8 tensor with the output blob, out_offsets)."""
    import torch
    from pycricodecs_b200 import engine
    with engine.Job(ctx, kind, blob_np, offsets, **kw) as job:
        job.run()
        out = torch.empty(max(job.out_bytes, 1), dtype=torch.uint8, pin_memory=True)
        _, status = job.download(out.numpy())
        assert int((status != 0).sum()) == 0, "input synthesis failed"
        return out, job.out_offsets.copy()


class Inputs:
    """Synthetic inputs of one rank: the WAV blob and, on demand, what this repo's own encoders make of it."""

    def __init__(self, ctx, streams, first_id, device):
        self.ctx, self.streams = ctx, streams
        self.wav, self.woff = make_wav_blob(streams, first_id, device)
        self._hca = self._enc = self._adx = None

    def wav_np(self):
        return self.wav.numpy()[: int(self.woff[-1])]

    def hca(self):
        from pycricodecs_b200 import _lib, engine
        if self._hca is None:
            self._hca = run_job_to_pinned(self.ctx, _lib.JOB_HCA_ENCODE, self.wav_np(), self.woff, quality=QUALITY, adx=engine.adx_params())
        return self._hca

    def encrypted(self):
        from pycricodecs_b200 import _lib
        if self._enc is None:
            hca, hoff = self.hca()
            keys = np.full(self.streams, KEY, np.uint64)
            self._enc = run_job_to_pinned(self.ctx, _lib.JOB_HCA_CRYPT, hca.numpy()[: int(hoff[-1])], hoff, keys=keys, encrypt=1, ciph_type=56)
        return self._enc

    def adx(self):
        from pycricodecs_b200 import _lib, engine
        if self._adx is None:
            self._adx = run_job_to_pinned(self.ctx, _lib.JOB_ADX_ENCODE, self.wav_np(), self.woff, adx=engine.adx_params())
        return self._adx

    def of(self, workload):
        """(pinned input blob, offsets) of `workload`."""
        if workload in ("adx_encode", "hca_encode"):
            return self.wav, self.woff
        if workload == "adx_decode":
            return self.adx()
        if workload == "hca_decode":
            return self.hca()
        return self.encrypted()


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
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, power, smax, reasons = [], [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                smax = max(smax, float(p[2]))
                if t0 <= t <= t1:
                    sm.append(float(p[1]))
                    power.append(float(p[3]))
                    for nm, v in zip(names, p[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def traffic_of(kernel, streams):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        if int(t.get("streams_per_launch", 0)) == int(streams) and kernel in t:
            return int(t[kernel]["bytes"])
    except (OSError, ValueError, KeyError):
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class Dist:
    """torch.distributed plumbing of the bench: barrier, MAX / SUM of a few floats."""

    def __init__(self, world, device):
        self.world, self.device, self.d = world, device, None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=device)
            self.d = dist

    def barrier(self):
        import torch
        torch.cuda.synchronize()
        if self.d is not None:
            self.d.barrier()

    def reduce(self, values, op):
        import torch
        if self.d is None:
            return [float(v) for v in values]
        t = torch.tensor(list(values), device="cuda", dtype=torch.float64)
        self.d.all_reduce(t, op=getattr(self.d.ReduceOp, op))
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.d is not None:
            self.d.destroy_process_group()


def time_resident(ctx, job, steps, warmup, dist):
    """Kernels only, batch resident in HBM: (ms per step, ms of the step's last kernel, wall ms per step, launches), max over ranks."""
    for _ in range(warmup):
        job.run()
    dist.barrier()
    l0 = ctx.launches
    t0 = time.perf_counter()
    dev = dom = 0.0
    for _ in range(steps):
        job.run()
        dev += ctx.last_kernel_ms
        dom += ctx.last_dominant_ms
    dist.barrier()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    launches = ctx.launches - l0
    ms, dom, wall = dist.reduce([dev / steps, dom / steps, wall], "MAX")
    return ms, dom, wall, launches, t0


def companion(ctx, name, blob_np, offsets, steps, dist, scaling="weak"):
    """One more workload of the same synthetic corpus, kernels only, resident: an extra object of the JSON line."""
    from pycricodecs_b200 import engine
    n = len(offsets) - 1
    with engine.Job(ctx, job_kind(name), blob_np, offsets, **job_kwargs(name, n)) as job:
        ms, _, _, _, _ = time_resident(ctx, job, steps, 3, dist)
        units = float(job.units)
        path = float(int(offsets[-1]) + job.out_bytes)
    units_all, = dist.reduce([units], "SUM")
    path_max, = dist.reduce([path], "MAX")
    hbm, _ = peaks()
    return {"workload": WORKLOADS[name], "name": name, "value": units_all / (ms * 1e-3), "unit": unit_of(name), "ms_per_step": ms,
            "scaling": scaling, "streams_per_gpu": n, "streams_total": int(dist.reduce([n], "SUM")[0]),
            "roofline_frac": path_max / (ms * 1e-3) / 1e9 / hbm, "kernel": KERNELS[name][-1] if len(KERNELS[name]) == 1 else "+".join(KERNELS[name])}


# ------------------------------------------------------------------ ours
def run_ours(a):
    rank, local_rank, world = rank_info()
    import torch
    from pycricodecs_b200 import _lib, engine, sharding
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = sharding.bind_to_gpu_numa(local_rank)          # before any page-locked allocation: node-local host buffers
    dist = Dist(world, device)
    ctx = engine.Context(local_rank)
    strong = a.scaling == "strong"
    from pycricodecs_b200.sharding import shard_range
    if strong:
        lo, hi = shard_range(a.streams, rank, world)       # a.streams = streams of the whole job
        streams, first_id = hi - lo, lo
    else:
        streams, first_id = a.streams, rank * a.streams    # a.streams = streams per GPU
    kind = job_kind(a.workload)
    inputs = Inputs(ctx, streams, first_id, device)
    pin_in, offsets = inputs.of(a.workload)
    in_bytes = int(offsets[-1])
    keyed = a.workload in ("hca_decrypt_decode", "hca_decrypt")
    kw = job_kwargs(a.workload, streams)
    keys = kw.get("keys")
    blob_np = pin_in.numpy()[:in_bytes]
    job = engine.Job(ctx, kind, blob_np, offsets, **kw)
    units, out_bytes = job.units, job.out_bytes
    pin_out = torch.empty(max(out_bytes, 1), dtype=torch.uint8, pin_memory=True)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ms, dom, wall_ms, launches, t0 = time_resident(ctx, job, a.steps, a.warmup, dist)
    t1 = time.perf_counter()
    if t1 - t0 < 0.6:                       # keep the GPU busy long enough for nvidia-smi to see the clocks under load (not timed)
        t_hold = time.perf_counter()
        while time.perf_counter() - t_hold < 0.6:
            job.run()
        t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    out_check, status = job.download(pin_out.numpy())
    assert int((status != 0).sum()) == 0, "engine reported per-stream errors"
    oo = job.out_offsets.copy()
    job.close()

    # spot parity inside the bench: stream 0 against the CPU checker (not timed, not on the product path): the compiled
    # reference where it is built, else its C restatement
    import oracle
    chk = oracle.ref() if oracle.have_ref() else oracle.port()
    got0 = bytes(out_check[int(oo[0]):int(oo[1])])
    in0 = bytes(blob_np[int(offsets[0]):int(offsets[1])])
    want0 = cpu_one(chk, a.workload, in0, KEY if keyed else 0)
    parity = got0 == want0

    # ---- end to end through the C-ABI batch call, host buffers in, host buffers out
    L = _lib.lib()
    status_arr = np.zeros(streams, np.int32)
    out_off = np.ascontiguousarray(oo, dtype=np.uint64)
    adxp = engine.adx_params()
    kp = keys.ctypes.data if keys is not None else None

    def batch_call(src, dst, dev_stream=None):
        """cri_*_batch (host pointers) or cri_*_batch_dev (device pointers + stream)."""
        sfx = () if dev_stream is None else (ctypes.c_void_p(dev_stream),)
        d = "" if dev_stream is None else "_dev"
        o, oo_, st = offsets.ctypes.data, out_off.ctypes.data, status_arr.ctypes.data
        if kind == _lib.JOB_HCA_DECODE:
            rc = getattr(L, "cri_hca_decode_batch" + d)(ctx.handle, src, o, streams, kp, None, dst, oo_, st, *sfx)
        elif kind == _lib.JOB_HCA_CRYPT:
            rc = getattr(L, "cri_hca_crypt_batch" + d)(ctx.handle, src, o, streams, 0, 0, kp, None, dst, st, *sfx)
        elif kind == _lib.JOB_HCA_ENCODE:
            rc = getattr(L, "cri_hca_encode_batch" + d)(ctx.handle, src, o, streams, QUALITY, 0, dst, oo_, st, *sfx)
        elif kind == _lib.JOB_ADX_ENCODE:
            rc = getattr(L, "cri_adx_encode_batch" + d)(ctx.handle, src, o, streams, ctypes.byref(adxp), dst, oo_, st, *sfx)
        else:
            rc = getattr(L, "cri_adx_decode_batch" + d)(ctx.handle, src, o, streams, dst, oo_, st, *sfx)
        ctx.check(rc)

    def timed_host(fn, reps):
        fn()                                 # warm-up (allocator, page mapping)
        dist.barrier()
        e0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dist.barrier()
        return dist.reduce([(time.perf_counter() - e0) / reps], "MAX")[0]

    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    e2e_s = timed_host(lambda: batch_call(pin_in.data_ptr(), pin_out.data_ptr()), e2e_steps)
    d_ref = pin_out[:out_bytes].to(device)                     # what the host-buffer call produced (pin_out is reused below)

    # the same bytes through bare cudaMemcpyAsync (both directions at once, pinned memory): the platform's ceiling for e2e
    d_in = torch.empty(max(in_bytes, 1), dtype=torch.uint8, device=device)
    d_out = torch.empty(max(out_bytes, 1), dtype=torch.uint8, device=device)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def bare_copies():
        with torch.cuda.stream(s_up):
            d_in[:in_bytes].copy_(pin_in[:in_bytes], non_blocking=True)
        with torch.cuda.stream(s_dn):
            pin_out[:out_bytes].copy_(d_out[:out_bytes], non_blocking=True)
        torch.cuda.synchronize()
    ceiling_s = timed_host(bare_copies, e2e_steps)

    # device-resident caller: the same batch call with device pointers on the caller's stream
    d_in[:in_bytes].copy_(pin_in[:in_bytes])
    cur = torch.cuda.current_stream().cuda_stream
    dev_s = timed_host(lambda: batch_call(d_in.data_ptr(), d_out.data_ptr(), cur), e2e_steps)
    dev_ok = bool(torch.equal(d_out[:out_bytes], d_ref))
    del d_ref

    # ---- N > 1: the sharded product path with and without the all-gather that reassembles the output on every rank
    gather = None
    if world > 1 and a.workload in ("hca_decode", "hca_decrypt_decode", "adx_decode", "adx_encode", "hca_encode") and not a.no_gather:
        chunks = 4
        n_all = streams * world
        sizes_in = np.diff(offsets.astype(np.int64))
        g_off = np.zeros(n_all + 1, np.uint64)             # every rank holds equally sized streams: global layout by tiling
        np.cumsum(np.tile(sizes_in, world).astype(np.uint64), out=g_off[1:])
        pieces = sharding.piece_ranges(n_all, world, chunks)
        per = streams // chunks
        local_of = {pieces[k * world + rank]: (k * per, (k + 1) * per if k < chunks - 1 else streams) for k in range(chunks)}

        class PieceBlob:                                     # sharded_batch only asks for this rank's pieces
            device = d_in.device

            def __call__(self, lo, hi):
                l0, l1 = local_of[(lo, hi)]
                return d_in[int(offsets[l0]):int(offsets[l1])]
        res = {}
        gk = dict(keys=np.full(n_all, KEY, np.uint64)) if keyed else {}
        skw = {k: v for k, v in kw.items() if k != "keys"}
        for label, do_gather in (("compute_only", False), ("with_gather", True)):
            tot = 0.0
            for it in range(5):                              # two untimed passes (allocator, NCCL channels), three timed
                t = {}
                out_t, _, st = sharding.sharded_batch(kind, PieceBlob(), g_off, ctx, gather=do_gather, chunks=chunks, timings=t, **skw, **gk)
                del out_t
                if it >= 2:
                    tot += dist.reduce([t["total_ms"]], "MAX")[0] / 3
                if it == 4 and rank == 0 and os.environ.get("CRI_GATHER_TRACE"):
                    print(f"[gather trace] {label}: compute_ms {t.get('compute_ms'):.3f} total_ms {t.get('total_ms'):.3f} chunks (compute end, gather start, gather end) {t.get('chunks')}", file=sys.stderr)
            res[label] = {"ms_per_step": tot, "value": units * world / (tot * 1e-3)}
        gathered_bytes = int(out_bytes) * world
        extra_ms = max(res["with_gather"]["ms_per_step"] - res["compute_only"]["ms_per_step"], 0.0)
        gather = {"api": "sharding.sharded_batch (cri_*_batch_dev per piece, NCCL all_gather_into_tensor in place, 4 chunks on a second stream)",
                  **res, "gathered_bytes_per_rank": gathered_bytes, "received_bytes_per_rank": gathered_bytes * (world - 1) // world,
                  "gather_extra_ms": extra_ms,
                  "note": "per step: header fetch + planning + kernels for this rank's four pieces (device-resident input), then the in-place all-gather of "
                          "every chunk; compute_only leaves the other ranks' pieces unfilled"}
        torch.cuda.empty_cache()
    del d_in, d_out

    # ---- companions of the default workload (BASELINE.json names HCA decode + ADX encode; configs[3] and configs[4])
    comp = {}
    if a.workload == "hca_decode" and not a.no_companion and not strong:
        comp["adx_encode"] = companion(ctx, "adx_encode", inputs.wav_np(), inputs.woff, a.steps, dist)
        enc, eoff = inputs.encrypted()
        comp["hca_decrypt_decode"] = companion(ctx, "hca_decrypt_decode", enc.numpy()[: int(eoff[-1])], eoff, a.steps, dist)
        # configs[4]: 4096 streams in total, split over the ranks (strong scaling)
        lo, hi = shard_range(min(4096, streams * world), rank, world)
        m = hi - lo
        comp["hca_encode"] = companion(ctx, "hca_encode", inputs.wav.numpy()[: int(inputs.woff[m])], inputs.woff[: m + 1], a.steps, dist, "strong")

    total_units, = dist.reduce([units], "SUM")
    if rank == 0:
        peak, peak_src = peaks()
        per_rank_units = total_units / world
        # algorithmic (compulsory) bytes per unit of the whole path = bytes in + bytes out, once each
        path_bytes = (in_bytes + out_bytes) / per_rank_units
        names = KERNELS[a.workload]
        kernels = []
        if len(names) == 2 and dom > 0:
            # decode = unpack (frame bytes in, 2 x 8 x 128 fp32 spectra out) + transform (spectra in, PCM16 out)
            parts = ((names[0], ms - dom, in_bytes / per_rank_units + 8192), (names[1], dom, 12288.0))
        else:
            parts = ((names[-1], dom if dom > 0 else ms, path_bytes),)
        for nm, kms, kb in parts:
            ach = per_rank_units * kb / (kms * 1e-3) / 1e9
            kernels.append({"kernel": nm, "kernel_ms": kms, "share_of_step": kms / ms, "algorithmic_bytes_per_unit": kb,
                            "achieved": ach, "frac": ach / peak, "traffic": traffic_of(nm, streams)})
        if len(names) == 2:
            xf = kernels[1]
            xf["read_stream"] = {"bytes_per_unit": 8192, "achieved": per_rank_units * 8192 / (xf["kernel_ms"] * 1e-3) / 1e9,
                                 "frac": per_rank_units * 8192 / (xf["kernel_ms"] * 1e-3) / 1e9 / peak,
                                 "note": "BASELINE.md section 3: the transform kernel's HBM-read stream (fp32 spectra) against the measured peak; target >= 0.40"}
        top = max(kernels, key=lambda k: k["kernel_ms"])
        achieved = per_rank_units * path_bytes / (ms * 1e-3) / 1e9
        traffics = [k["traffic"] for k in kernels]
        hca = is_hca(a.workload)
        line = {
            "metric": METRIC, "value": total_units / (ms * 1e-3), "unit": unit_of(a.workload),
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "wall_ms_per_step": wall_ms,
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f32 (no FMA, bit-exact) + int16/u8 bitstream" if hca else "int32",
            "data": "synthetic",
            "config": config_of(a, streams, in_bytes, int(out_bytes)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": sum(traffics) if all(t is not None for t in traffics) else None,
                         "scope": "whole path: compulsory bytes in + out per unit x units / step time (all kernels of the step)",
                         "algorithmic_bytes_per_unit": path_bytes, "kernel": top["kernel"], "kernel_ms": top["kernel_ms"],
                         "peak_source": peak_src, "kernels": kernels},
            "e2e": {"value": total_units / e2e_s, "unit": unit_of(a.workload), "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(out_bytes),
                    "ms_per_step": e2e_s * 1e3, "api": "cri_*_batch (C-ABI), pinned host buffers, parse+plan+H2D+kernels+D2H",
                    "pcie_ceiling_ms": ceiling_s * 1e3, "frac_of_pcie_ceiling": ceiling_s / e2e_s,
                    "pcie_ceiling_note": "the same H2D + D2H bytes through bare cudaMemcpyAsync on two streams, pinned memory, all ranks at once",
                    "numa": numa},
            "e2e_device": {"value": total_units / dev_s, "unit": unit_of(a.workload), "ms_per_step": dev_s * 1e3, "matches_host_path": dev_ok,
                           "api": "cri_*_batch_dev (C-ABI), device buffers on the caller's stream: header fetch + plan + kernels, no PCIe payload"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_spot_check": bool(parity),
            "parity_checker": "reference" if oracle.have_ref() else "port",
        }
        if gather is not None:
            line["gather"] = gather
        if len(names) == 2:
            # the transform is bounded by fp32 issue before HBM -- 16 transforms x 3968 separately rounded
            # fp32 operations per stereo frame (no FMA: the reference rounds every product and sum) against one fp32
            # instruction per lane and clock
            sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
            ops = per_rank_units * 16 * 3968 / (kernels[1]["kernel_ms"] * 1e-3)
            peak_ops = sm * 128 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
            line["roofline"]["fp32_issue"] = {"achieved_tops": ops / 1e12, "peak_tops": peak_ops / 1e12, "frac": ops / peak_ops,
                                              "kernel": names[1],
                                              "note": "separately rounded fp32 mul/add per second vs SMs x 128 lanes x SM clock"}
        line.update(comp)
        if not a.no_cpu:
            sample = [bytes(blob_np[int(offsets[i]):int(offsets[i + 1])]) for i in range(min(64, streams))]
            line["cpu_baseline"] = cpu_baseline(a.workload, sample, keyed, a.cpu_seconds)
        emit_line(line)
    dist.close()


def config_of(a, streams, in_bytes=None, out_bytes=None):
    hca = is_hca(a.workload)
    wl = WORKLOADS[a.workload] if QUALITY == 1 or not hca else WORKLOADS[a.workload].replace("High", ["Highest", "High", "Middle", "Low"][QUALITY])
    c = {"workload": wl, "name": a.workload, "streams_per_gpu": streams, "unique_streams_per_gpu": streams,
         "stream_seconds": 2.0, "sample_rate": 48000, "channels": 2}
    if in_bytes is not None:
        c.update({"input_synthesis": "pycricodecs_b200.synth PCM; compressed inputs made once, untimed, by this repo's own GPU encoders",
                  "l2": "inputs and outputs exceed the 126 MB L2 (no flush needed)" if in_bytes + out_bytes > 3e8 else "working set near L2 size",
                  "in_bytes_per_gpu": in_bytes, "out_bytes_per_gpu": out_bytes})
    return c


# ----------------------------------------------------------- CPU baselines
def cpu_one(impl, workload, data, key):
    """One stream through the CPU checker; returns the output bytes."""
    if workload in ("hca_decode", "hca_decrypt_decode"):
        r = impl.hca_decode(data, key)
    elif workload == "hca_decrypt":
        r = impl.hca_crypt(data, 0, 0, key)
    elif workload == "hca_encode":
        r = impl.hca_encode(data, QUALITY)
    elif workload == "adx_encode":
        r = impl.adx_encode(data)
    else:
        r = impl.adx_decode(data)
    return bytes(r[1]) if isinstance(r, tuple) else bytes(r)


def _cpu_one(workload, data, key):
    import oracle
    cpu_one(oracle.ref() if oracle.have_ref() else oracle.port(), workload, data, key)


def cpu_kind():
    import oracle
    return "reference" if oracle.have_ref() else "port"


def units_of(workload, data):
    if workload == "hca_encode":
        return ((len(data) - 44) // 4 + 128 + 1023) // 1024   # frames = ceil((samples + delay) / 1024)
    if is_hca(workload):
        return int.from_bytes(data[16:20], "big")             # frame count from the fmt chunk
    if workload == "adx_encode":
        return (len(data) - 44) // 2 // 32                    # blocks over all channels
    return int.from_bytes(data[12:16], "big") * data[7] // 32


def cpu_baseline(workload, sample, keyed, seconds):
    """Single-thread CPU baseline on a bounded sample of the same workload (rank 0 only)."""
    key = KEY if keyed else 0
    _cpu_one(workload, sample[0], key)
    t0 = time.perf_counter()
    n = units = 0
    while time.perf_counter() - t0 < seconds:
        d = sample[n % len(sample)]
        _cpu_one(workload, d, key)
        units += units_of(workload, d)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": units / dt, "unit": "HCA frames/s" if is_hca(workload) else "ADX blocks/s", "cores": 1, "kind": cpu_kind(),
            "sample": f"{n} of the workload's 2 s streams processed back to back on one host thread in {dt:.1f} s"}


_REF_DATA = None


def _ref_worker(args):
    workload, idx, key = args
    _cpu_one(workload, _REF_DATA[idx], key)
    return 1


def reference_inputs(workload, count):
    """Inputs for the CPU arm, made by the CPU implementation itself (no GPU needed)."""
    import oracle
    from pycricodecs_b200 import synth
    impl = oracle.ref() if oracle.have_ref() else oracle.port()
    out = []
    for sid in range(count):
        w = synth.wav(sid, 2)
        if workload in ("adx_encode", "hca_encode"):
            out.append(w)
        elif workload == "adx_decode":
            out.append(impl.adx_encode(w)[1])
        else:
            h = impl.hca_encode(w, QUALITY)[1]
            if workload in ("hca_decrypt_decode", "hca_decrypt"):
                h = bytes(impl.hca_crypt(h, 1, 56, KEY)) if cpu_kind() == "reference" else impl.hca_crypt(h, 1, 56, KEY)[1]
            out.append(h)
    return out


def run_reference(a):
    """The reference's own CPU implementation of the path on all host cores, on our arm's config: every step is the stated
    per-GPU batch (--streams streams, 32 distinct ones tiled), one process per core (the reference holds the GIL)."""
    global _REF_DATA
    rank, _, world = rank_info()
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    _REF_DATA = reference_inputs(a.workload, 32)
    key = KEY if a.workload in ("hca_decrypt_decode", "hca_decrypt") else 0
    per_step = a.ref_streams or a.streams
    tasks = [(a.workload, i % len(_REF_DATA), key) for i in range(per_step)]
    units_step = sum(units_of(a.workload, _REF_DATA[t[1]]) for t in tasks)
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(1, a.warmup)):
            pool.map(_ref_worker, tasks, chunksize=8)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_ref_worker, tasks, chunksize=8)
        dt = (time.perf_counter() - t0) / a.steps
    v = units_step / dt
    unit = unit_of(a.workload)
    emit_line({
        "impl": "reference", "metric": METRIC, "value": v, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 scalar SSE2 (no FMA)" if is_hca(a.workload) else "int32", "data": "synthetic",
        "config": config_of(a, a.streams),
        "note": ("the reference's own CPU implementation (oracle/_ref = unmodified CriCodecs C++ where built, else the C port), one process "
                 f"per host core; a step = {per_step} streams (32 distinct, tiled) = one GPU's batch at any N (CPU throughput does not depend on N)"),
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": cpu_kind(),
                         "sample": f"{per_step} streams of 2 s per step over a {cores}-process pool"},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_JSON_OUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: native libraries write banners to file descriptor 1 (NCCL prints its version
    there at the first collective), so the descriptor is pointed at stderr for the run and the line goes to a saved copy."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hca_decode", choices=list(WORKLOADS))
    ap.add_argument("--streams", type=int, default=8192, help="streams per GPU (weak scaling) or of the whole job (--scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ref-streams", type=int, default=0, help="reference arm: streams per step (default: --streams)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--no-companion", action="store_true", help="skip the companion measurements of the default workload")
    ap.add_argument("--quality", type=int, default=1, choices=[0, 1, 2, 3], help="HCA quality of the synthetic streams (default High)")
    a = ap.parse_args()
    global QUALITY
    QUALITY = a.quality
    if a.impl == "reference":
        run_reference(a)
    else:
        a.warmup = max(a.warmup, 3)
        run_ours(a)


if __name__ == "__main__":
    main()
