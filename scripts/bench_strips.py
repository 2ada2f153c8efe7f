#!/usr/bin/env python3
"""Config 4 of BASELINE.json: merging watershed of ONE S x S field cut into row strips, one strip per
GPU, boundary rows exchanged with NCCL send/recv over NVLink (torchrun, one process per GPU).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_strips.py \
        [--size 16384] [--field uniform|smooth] [--reps 3] [--check]

Prints one JSON line on rank 0: time of the whole strip solve (max over ranks, CUDA-synchronised wall
clock), rounds of the two exchange loops, lakes per level at a few levels, and with --check the
comparison against a single-GPU solve of the whole field on rank 0.
"""
import argparse
import importlib
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--field", default="uniform", choices=["uniform", "smooth"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    ws = load()
    st = importlib.import_module("rustronomy_watershed_b200.strips")
    ctx = ws.Context(local)
    S = args.size
    img = fieldgen.uniform(S, S, 0) if args.field == "uniform" else fieldgen.smooth(S, S, 16.0, 0)
    parts = st.partition_rows(S, world)
    g = st.StripGeometry(rank, world, S, parts[rank])
    lo, hi = g.local_rows
    strip = st.CudaStrip(ws, ctx, g, torch.from_numpy(img[lo:hi].copy()).cuda())
    comm = st.DistComm()
    comm.set_device(torch.device("cuda", local))
    times, res = [], None
    for rep in range(args.reps + 1):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = st.solve([strip], comm, st.MERGING, 254)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep:
            times.append(float(dt.item()))
    out = {"config": f"{S}x{S} u8 {args.field}, merging, {world} row strips (one per GPU), NCCL halo exchange",
           "n_gpus": world, "ms": 1e3 * min(times), "ms_all": [1e3 * t for t in times],
           "Mpx_levels_per_s": S * S * 255 / min(times) / 1e6, "flood_exchange_rounds": res.flood_rounds,
           "label_exchange_rounds": res.label_rounds, "seeds": res.nseeds_total, "forest_edges": res.edges_total,
           "lakes_0_64_127_191_254": [int(res.lake_counts[i]) for i in (0, 64, 127, 191, 254)],
           "halo_bytes_per_round_per_neighbour": S * 4,
           "phase_ms_rank0_last_rep": {k: round(1e3 * v, 3) for k, v in (res.phase_s or {}).items()}}
    if args.check:
        lab = torch.from_numpy(strip.owned_labels().astype(np.int32)).cuda()
        gathered = [torch.empty((parts[r][1] - parts[r][0], S), dtype=torch.int32, device="cuda") for r in range(world)]
        dist.all_gather(gathered, lab)
        if rank == 0:
            whole = torch.cat(gathered).cpu().numpy().astype(np.uint32)
            seg = ws.TransformBuilder.default().set_device(local).build_merging()
            seeds = seg.find_local_minima(img)
            ref_lab, _ = seg.transform_compact(img, seeds)
            ref_lakes, _ = seg.lake_counts(img, seeds)
            out["check_labels_equal_single_gpu"] = bool(np.array_equal(whole, ref_lab))
            out["check_lakes_equal_single_gpu"] = bool(np.array_equal(res.lake_counts, ref_lakes))
    if rank == 0:
        print(json.dumps(out))
    strip.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
