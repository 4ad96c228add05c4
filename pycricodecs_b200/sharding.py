"""Multi-GPU execution of a batch: one process per GPU (torchrun), streams sharded, optional NCCL all-gather.

Streams are independent, so the data path needs no exchange while it computes: the batch is cut
into `world x chunks` consecutive *pieces* of whole streams (`shard_range`); piece p = k * world + r is rank r's k-th chunk.
Each rank runs its pieces through the device-pointer entry points of the C-ABI (`engine.batch_device`, payload resident
in HBM). With `gather=True` the output shards are reassembled on EVERY rank with one all-gather per chunk
(`torch.distributed.all_gather_into_tensor`, NCCL over NVLink/NVSwitch) issued on a second CUDA stream, so the gather of
chunk k overlaps the kernels of chunk k + 1. Because piece k * world + r is rank r's k-th chunk, the gathered pieces of a
chunk are consecutive in the global stream order: with equal piece sizes (equal-length streams, the benchmark's case)
the collective writes straight into the final packed blob; otherwise pieces are padded to the chunk's largest one and
compacted afterwards with device-to-device copies.

    rank, world, device = sharding.init()                       # under torchrun
    out, out_offsets, status = sharding.sharded_batch(_lib.JOB_HCA_DECODE, blob, offsets)   # same result on all ranks

`compute=` swaps the per-piece engine call for a stand-in, which is how the world-size-2 gloo test exercises the piece
assignment / padding / compaction logic on CPU tensors.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of `n` streams for `rank` of `world`."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def stream_ids(per_rank: int, rank: int):
    """Weak scaling: every rank synthesises its own `per_rank` stream ids."""
    return list(range(rank * per_rank, (rank + 1) * per_rank))


def piece_ranges(n: int, world: int, chunks: int):
    """[(lo, hi)] for the world * chunks pieces; piece k * world + r belongs to rank r, chunk k."""
    return [shard_range(n, p, world * chunks) for p in range(world * chunks)]


def init(backend: Optional[str] = None):
    """Join the torchrun job: (rank, world, device). NCCL on GPUs, gloo otherwise; 127.0.0.1 rendezvous by default."""
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if cuda else "gloo")
        dist.init_process_group(backend, rank=rank, world_size=world, **({"device_id": device} if backend == "nccl" else {}))
    return rank, world, device


def bind_to_gpu_numa(device_index: int) -> Optional[str]:
    """Pin this process to the CPUs of the GPU's NUMA node (page-locked buffers allocated afterwards are node-local, which is
    what keeps eight ranks' host copies off the inter-socket link). Returns a note, or None when the topology is unknown."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return f"node {node} cpus {cpus}"
    except Exception:
        return None


def sharded_batch(kind: int, blob, offsets: np.ndarray, ctx=None, *, gather: bool = True, chunks: int = 4, group=None,
                  compute: Optional[Callable] = None, sizes_of: Optional[Callable] = None, timings: Optional[dict] = None, **kw):
    """Run one batch over all ranks of the process group.

    `blob` is the WHOLE batch (host numpy uint8 array, or a torch tensor on any device; a rank only touches the bytes of
    its own pieces) or a callable `blob(lo, hi)` that returns the device tensor with streams lo .. hi - 1 of one of this
    rank's pieces; `offsets` are the n + 1 stream offsets. Per-stream keyword arrays (`keys`, `subkeys`) are global and are
    sliced per piece.

    Returns (out, out_offsets, status); `out_offsets` (n + 1) and `status` (n) are global:
      gather=True   `out` = the packed output blob of ALL streams (torch uint8 on this rank's device), identical on every rank;
      gather=False  same buffer and layout, but only this rank's pieces are filled in: outputs stay sharded.

    Output sizes are known from the headers before anything is decoded, so every piece is written at its final place in the
    packed blob and the all-gather of a chunk runs IN PLACE (send buffer = this rank's slice of the receive buffer).
    """
    import torch
    import torch.distributed as dist
    from . import engine

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    chunks = max(1, min(chunks, max(1, n // max(world, 1))))
    pieces = piece_ranges(n, world, chunks)
    if compute is None:
        ctx = ctx or engine.default_context()
        device = torch.device("cuda", ctx.device)

        def compute(piece_blob, piece_offsets, out, out_offsets=None, **pkw):
            return engine.batch_device(kind, piece_blob, piece_offsets, ctx, out=out, out_offsets=out_offsets, **pkw)[2]

        def sizes_of(piece_blob, piece_offsets, **pkw):
            return engine.sizes_device(kind, piece_blob, piece_offsets, ctx, **pkw)[0]
    else:
        device = blob.device if isinstance(blob, torch.Tensor) else torch.device("cpu")
    on_gpu = device.type == "cuda"
    comm_stream = torch.cuda.Stream(device) if on_gpu and world > 1 and gather else None
    ev0 = ev1 = ev2 = None
    if timings is not None and on_gpu:
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ev0.record()

    def piece_kw(lo, hi):
        pkw = dict(kw)
        for name in ("keys", "subkeys"):
            if pkw.get(name) is not None and not np.isscalar(pkw[name]):
                pkw[name] = np.asarray(pkw[name])[lo:hi]
        return pkw

    inputs = {}

    def piece_input(lo, hi):
        if (lo, hi) not in inputs:
            b0, b1 = int(offsets[lo]), int(offsets[hi])
            if callable(blob):                                   # the caller hands out pieces itself (e.g. already in HBM)
                inputs[(lo, hi)] = blob(lo, hi)
            elif isinstance(blob, torch.Tensor):
                t = blob[b0:b1]
                inputs[(lo, hi)] = t if t.device == device else t.to(device, non_blocking=True)
            else:
                inputs[(lo, hi)] = torch.from_numpy(np.ascontiguousarray(blob[b0:b1])).to(device, non_blocking=True)
        return inputs[(lo, hi)]

    mine = [pieces[k * world + rank] for k in range(chunks)]
    # ---- output layout: per-stream sizes from the headers of this rank's pieces, summed over ranks
    sizes = np.zeros(n, np.int64)
    for lo, hi in mine:
        if hi > lo:
            sizes[lo:hi] = np.asarray(sizes_of(piece_input(lo, hi), offsets[lo:hi + 1] - offsets[lo], **piece_kw(lo, hi)), np.int64)
    if world > 1:
        t = torch.from_numpy(sizes).to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        sizes = t.cpu().numpy()
    out_offsets = np.zeros(n + 1, np.uint64)
    np.cumsum(sizes.astype(np.uint64), out=out_offsets[1:])
    total = int(out_offsets[-1])
    final = torch.empty(max(total, 1), dtype=torch.uint8, device=device)
    if comm_stream is not None:
        comm_stream.wait_stream(torch.cuda.current_stream(device))

    # ---- per chunk: this rank's piece straight into its place, then the chunk's all-gather on the second stream
    status = np.zeros(n, np.int64)
    for k, (lo, hi) in enumerate(mine):
        o0, o1 = int(out_offsets[lo]), int(out_offsets[hi])
        if hi > lo:
            status[lo:hi] = compute(piece_input(lo, hi), offsets[lo:hi + 1] - offsets[lo], final[o0:max(o1, o0 + 1)],
                                    out_offsets=out_offsets[lo:hi + 1] - out_offsets[lo], **piece_kw(lo, hi))
        inputs.pop((lo, hi), None)
        if world == 1 or not gather:
            continue
        ps = range(k * world, (k + 1) * world)
        pbytes = [int(out_offsets[pieces[p][1]] - out_offsets[pieces[p][0]]) for p in ps]
        first = int(out_offsets[pieces[k * world][0]])
        marks = timings.setdefault("chunks", []) if timings is not None and ev0 is not None else None
        if comm_stream is not None:
            done = torch.cuda.Event(enable_timing=marks is not None)
            done.record()
        with (torch.cuda.stream(comm_stream) if comm_stream is not None else _Null()):
            if comm_stream is not None:
                comm_stream.wait_event(done)
            if marks is not None and comm_stream is not None:    # (compute end, gather start, gather end) of this chunk
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                marks.append((done, g0, g1))
            if len(set(pbytes)) == 1:
                if pbytes[0]:
                    dist.all_gather_into_tensor(final[first:first + world * pbytes[0]], final[o0:o1], group=group)
            else:
                width = (max(pbytes) + 15) // 16 * 16
                padded = torch.zeros(width, dtype=torch.uint8, device=device)
                padded[: o1 - o0] = final[o0:o1]
                staging = torch.empty(world * width, dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(staging, padded, group=group)
                at = first
                for r, nb in enumerate(pbytes):
                    if r != rank:
                        final[at:at + nb] = staging[r * width:r * width + nb]
                    at += nb
            if marks is not None and comm_stream is not None:
                marks[-1][2].record()
    if ev1 is not None:
        ev1.record()
    if comm_stream is not None:
        torch.cuda.current_stream(device).wait_stream(comm_stream)
    if world > 1:
        t = torch.from_numpy(status).to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        status = t.cpu().numpy()
    if timings is not None and ev0 is not None:
        ev2.record()
        torch.cuda.synchronize(device)
        timings.update(compute_ms=ev0.elapsed_time(ev1), total_ms=ev0.elapsed_time(ev2))
        if "chunks" in timings:          # ms from the start of the call: when each chunk's compute ended, its gather ran
            timings["chunks"] = [tuple(round(ev0.elapsed_time(e), 3) for e in m) for m in timings["chunks"]]
    return final[:total], out_offsets, status.astype(np.int32)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _reduce(value: float, op_name: str) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def all_max(value: float) -> float:
    return _reduce(value, "MAX")


def all_sum(value: float) -> float:
    return _reduce(value, "SUM")


# ------------------------------------------------------------------ torchrun entry point
def main(argv=None):
    """python -m torch.distributed.run --nproc-per-node N -m pycricodecs_b200.sharding KIND FILES... [--out-dir DIR]

    KIND: hca_decode | hca_encode | adx_decode | adx_encode. Every rank reads the files of its own pieces, the outputs are
    gathered over NCCL and rank 0 writes them to --out-dir (same base names, new extension)."""
    import argparse
    from . import _lib
    ap = argparse.ArgumentParser(prog="pycricodecs_b200.sharding")
    ap.add_argument("kind", choices=["hca_decode", "hca_encode", "adx_decode", "adx_encode"])
    ap.add_argument("files", nargs="+")
    ap.add_argument("--out-dir", default=".")
    ap.add_argument("--key", type=lambda s: int(s, 0), default=0)
    ap.add_argument("--quality", type=int, default=1)
    a = ap.parse_args(argv)
    rank, world, device = init()
    kind = {"hca_decode": _lib.JOB_HCA_DECODE, "hca_encode": _lib.JOB_HCA_ENCODE, "adx_decode": _lib.JOB_ADX_DECODE,
            "adx_encode": _lib.JOB_ADX_ENCODE}[a.kind]
    ext = {"hca_decode": ".wav", "adx_decode": ".wav", "hca_encode": ".hca", "adx_encode": ".adx"}[a.kind]
    lens = np.array([os.path.getsize(f) for f in a.files], np.uint64)
    offsets = np.zeros(len(a.files) + 1, np.uint64)
    np.cumsum(lens, out=offsets[1:])
    blob = np.zeros(int(offsets[-1]), np.uint8)                    # only this rank's pieces are read from disk
    chunks = 4
    chunks = max(1, min(chunks, max(1, len(a.files) // world)))
    for k in range(chunks):
        lo, hi = piece_ranges(len(a.files), world, chunks)[k * world + rank]
        for i in range(lo, hi):
            blob[int(offsets[i]):int(offsets[i + 1])] = np.fromfile(a.files[i], np.uint8)
    from . import engine
    kw = {}
    if kind == _lib.JOB_HCA_DECODE and a.key:
        kw["keys"] = np.full(len(a.files), a.key, np.uint64)
    if kind == _lib.JOB_HCA_ENCODE:
        kw["quality"] = a.quality
    out, ooff, status = sharded_batch(kind, blob, offsets, engine.Context(device.index or 0), chunks=chunks, **kw)
    if rank == 0:
        os.makedirs(a.out_dir, exist_ok=True)
        host = out.cpu().numpy()
        for i, f in enumerate(a.files):
            if status[i] == 0:
                host[int(ooff[i]):int(ooff[i + 1])].tofile(os.path.join(a.out_dir, os.path.splitext(os.path.basename(f))[0] + ext))
            else:
                print(f"{f}: status {status[i]} ({engine.strerror(int(status[i]))})")
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
