#!/usr/bin/env python3
"""Config 5 of BASELINE.json: a batch of CGPS-like 2048 x 2048 u8 slices, segmenting + merging, sharded
data-parallel over the GPUs of one box (one process per GPU, no data-path collective).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_batch.py \
        [--slices 1024] [--chunk 64] [--reps 2]

Every rank takes slices/N slices and runs them in chunks of `chunk` slices (one device plan per chunk
shape, reused).  Per chunk, inside the timed region: H2D of the slices from pinned host memory (double
buffered on a copy stream, so that chunk k + 1 crosses the link while chunk k is in the kernels), seed
finding on the device, the segmenting run, the merging run, D2H of the per-slice lake counts.  Rank 0
prints one JSON line: total time = max over ranks; device_ms = the kernels alone (CUDA events).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fieldgen  # noqa: E402
from wsb200_loader import load  # noqa: E402

LEVELS, S = 255, 2048


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slices", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic slices per rank (tiled to a chunk)")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ws = load()
    ctx = ws.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    per_rank = args.slices // world
    chunk = min(args.chunk, per_rank)
    nchunks = per_rank // chunk
    base = np.stack([fieldgen.cgps_like(S, S, seed=rank * args.distinct + i) for i in range(args.distinct)])
    host = torch.from_numpy(np.concatenate([base] * ((chunk + args.distinct - 1) // args.distinct))[:chunk]).pin_memory()
    plan = ws.Plan(ctx, chunk, S, S)
    d_img = [torch.empty((chunk, S, S), dtype=torch.uint8, device="cuda") for _ in range(2)]  # double buffer
    d_off = torch.zeros(chunk + 1, dtype=torch.int32, device="cuda")
    cap = int(0.12 * chunk * S * S)
    d_seeds = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
    counts_h = torch.empty((chunk, 256), dtype=torch.int32).pin_memory()
    copy_stream = torch.cuda.Stream(device=local)
    ready = [torch.cuda.Event() for _ in range(2)]     # slices of buffer b are on the device
    done = [torch.cuda.Event() for _ in range(2)]      # the kernels have finished with buffer b

    class _Dev:
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<i4", "data": (int(ptr), False), "version": 2}

    def upload(b):
        """H2D of the next chunk on the copy stream: overlaps the kernels of the current one."""
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[b])
            d_img[b].copy_(host, non_blocking=True)
            ready[b].record(copy_stream)

    def compute(b):
        stream.wait_event(ready[b])
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n = plan.find_local_minima(d_img[b].data_ptr(), d_seeds.data_ptr(), cap, d_off.data_ptr())
        plan.run(0, 254, d_img[b].data_ptr(), d_seeds.data_ptr(), d_off.data_ptr(), n)
        plan.run(1, 254, d_img[b].data_ptr(), d_seeds.data_ptr(), d_off.data_ptr(), n)
        e1.record(stream)
        done[b].record(stream)
        with torch.cuda.stream(stream):
            counts_h.copy_(torch.as_tensor(_Dev(plan.lake_counts_ptr, (chunk, 256)), device="cuda"), non_blocking=True)
        stream.synchronize()
        return n, e0.elapsed_time(e1)

    def run_all():
        dev_ms, seeds = 0.0, 0
        upload(0)
        for c in range(nchunks):
            if c + 1 < nchunks:
                upload((c + 1) & 1)
            n, ms = compute(c & 1)
            dev_ms += ms
            seeds += n
        return dev_ms, seeds

    for b in range(2):
        done[b].record(stream)
    run_all()  # warm-up
    best = None
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev_ms, seeds = run_all()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t = torch.tensor([wall, dev_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if best is None or t[0].item() < best[0]:
            best = (t[0].item(), t[1].item(), seeds)
    if rank == 0:
        total = nchunks * chunk * world
        pxl = 2 * total * S * S * LEVELS
        print(json.dumps({
            "config": f"{total} CGPS-like 2048x2048 u8 slices, segmenting + merging (lake counts), {world} GPU(s), "
                      f"{chunk} slices per plan, slices sharded over ranks (no collective)",
            "n_gpus": world, "slices": total, "wall_ms": 1e3 * best[0], "device_ms_max_rank": best[1],
            "Mpx_levels_per_s_e2e": pxl / best[0] / 1e6, "Mpx_levels_per_s_device": pxl / (best[1] * 1e-3) / 1e6,
            "h2d_bytes": total * S * S, "d2h_bytes": total * 1024, "seeds_rank0": best[2],
            "lakes_slice0_levels_0_127_254": [int(counts_h[0, 0]), int(counts_h[0, 127]), int(counts_h[0, 254])]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
