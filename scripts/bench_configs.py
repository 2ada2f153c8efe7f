#!/usr/bin/env python3
"""Time the five BASELINE.json configs through the reference-facing API (host buffers, wall clock)
on ONE GPU, with the oracle timed beside the small ones.  Writes one JSON object to stdout.

    python scripts/bench_configs.py [--slices 16] [--skip-history] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fieldgen  # noqa: E402
from wsb200_loader import load  # noqa: E402

LEVELS = 255


def timed(fn, reps):
    fn()                                  # warm-up (allocations, first-touch)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slices", type=int, default=16, help="slices of config 5 on this GPU (128 = one GPU's share)")
    ap.add_argument("--skip-history", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true", help="also time the oracle on configs 1 and 2")
    args = ap.parse_args()
    ws = load()
    seg = ws.TransformBuilder.default().build_segmenting()
    mrg = ws.TransformBuilder.default().build_merging()
    res = {}

    # ---- configs 1 and 2: 512^2 uniform (README example / per-level lake counts) --------------------
    img = fieldgen.uniform(512, 512, 0)
    t_min, seeds = timed(lambda: seg.find_local_minima(img), args.reps)
    t1, _ = timed(lambda: seg.transform(img, seeds), args.reps)
    t2, (lakes, _) = timed(lambda: mrg.lake_counts(img, seeds), args.reps)
    t2l, _ = timed(lambda: mrg.transform_to_list(img, seeds), 1)
    px = img.size
    res["c1_segmenting_512_uniform"] = {"ms": 1e3 * t1, "find_local_minima_ms": 1e3 * t_min, "seeds": len(seeds),
                                        "Mpx_levels_per_s": px * LEVELS / t1 / 1e6}
    res["c2_merging_512_uniform"] = {"lake_counts_ms": 1e3 * t2, "transform_to_list_ms": 1e3 * t2l,
                                     "lakes_0_127_254": [int(lakes[0]), int(lakes[127]), int(lakes[254])],
                                     "Mpx_levels_per_s": px * LEVELS / t2 / 1e6}
    if args.cpu:
        from oracle import oracle as orc
        t0 = time.perf_counter()
        orc.transform(orc.SEGMENTING, img, seeds)
        tc1 = time.perf_counter() - t0
        t0 = time.perf_counter()
        orc.transform(orc.MERGING, img, seeds, fast_closure=False, hook=lambda l, c: None)
        tc2 = time.perf_counter() - t0
        res["c1_segmenting_512_uniform"]["oracle_ms"] = 1e3 * tc1
        res["c2_merging_512_uniform"]["oracle_ms"] = 1e3 * tc2
        res["oracle_threads"] = orc.num_threads()

    # ---- config 3: 4096^2 smoothed field, transform_history at every level -----------------------
    img = fieldgen.smooth(4096, 4096, 8.0, 0)
    seeds = seg.find_local_minima(img)
    tc, (lab, lvl) = timed(lambda: seg.transform_compact(img, seeds), args.reps)
    entry = {"seeds": len(seeds), "compact_ms": 1e3 * tc, "compact_bytes": int(lab.nbytes + lvl.nbytes)}
    if not args.skip_history:
        import ctypes as C
        N = ws._native
        ctx = ws.default_context()
        hist = np.empty((LEVELS, 4096, 4096), dtype=np.uint64)          # 34 GB, caller-owned like the Vec
        hist[:] = 0                                                     # touch the pages once
        lv = np.empty(LEVELS, dtype=np.uint8)
        cfg, view, s = seg._cfg(), N.image_view(img), N.seeds_array(seeds)

        def run_hist():
            ctx.check(ctx.lib.ws_transform_history(ctx.handle, C.byref(cfg), C.byref(view), s.ctypes.data,
                                                   s.shape[0], lv.ctypes.data, hist.ctypes.data))
        th, _ = timed(run_hist, 1)
        ok = bool(np.array_equal(hist[254], lab.astype(np.uint64))
                  and np.array_equal(hist[100], np.where(lvl <= 100, lab, 0).astype(np.uint64)))
        entry.update({"history_ms": 1e3 * th, "history_bytes": int(hist.nbytes),
                      "history_GBps": hist.nbytes / th / 1e9, "history_consistent_with_compact": ok,
                      "Mpx_levels_per_s": img.size * LEVELS / th / 1e6})
        del hist
    res["c3_segmenting_history_4096_smooth8"] = entry

    # ---- config 5: batch of CGPS-like 2048^2 slices, segmenting + merging ------------------------
    n = args.slices
    imgs = np.stack([fieldgen.cgps_like(2048, 2048, seed=i) for i in range(n)])
    tmin, (bseeds, boff) = timed(lambda: mrg.find_local_minima_batch(imgs), 1)
    tb, (labels, counts) = timed(lambda: mrg.transform_batch(imgs, bseeds, boff, want_labels=True,
                                                              want_lake_counts=True), args.reps)
    tbc, _ = timed(lambda: mrg.transform_batch(imgs, bseeds, boff, want_labels=False, want_lake_counts=True),
                   args.reps)
    res["c5_batch_2048_cgps"] = {"slices": n, "seeds_total": int(boff[-1]), "find_local_minima_ms": 1e3 * tmin,
                                 "labels_and_counts_ms": 1e3 * tb, "counts_only_ms": 1e3 * tbc,
                                 "Mpx_levels_per_s_both_transforms": 2 * imgs.size * LEVELS / tb / 1e6,
                                 "lakes_first_slice_0_127_254": [int(counts[0][0]), int(counts[0][127]), int(counts[0][254])]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
