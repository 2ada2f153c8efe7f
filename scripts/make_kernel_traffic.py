#!/usr/bin/env python3
"""profiles/kernel_traffic.json from the ncu summaries of one capture stage:

    python scripts/make_kernel_traffic.py <stage, e.g. r02_e> <commit the capture ran at> [field_size, default uniform_16384]

reads profiles/<stage>_<kernel>_<field>16k_ncu_summary.txt (scripts/ncu_summary.py output) for every kernel of
the pipeline and records DRAM bytes per launch, the launch time under ncu and the file each number came from.
bench.py reports them as roofline.kernels[].dram_bytes_per_launch next to the live CUDA-event times."""
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"fill_state_kernel": "fill_state", "fill_rows_kernel": "fill_state", "seeds_scan_kernel": "seeds_scan", "seed_init_kernel": "seed_init", "flood_kernel": "flood",
         "label_tile_kernel": "label_tile", "rim_jump_kernel": "rim_jump", "label_finish_kernel": "label_finish",
         "merge_reduce_kernel": "merge_reduce", "forest_init_kernel": "forest_init",
         "forest_boruvka_kernel": "forest_boruvka"}
UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def metric(text, name):
    m = re.search(rf"^\s*{re.escape(name)}\s+([0-9.,]+)\s+(\S+)", text, re.M)
    if not m:
        return None
    return float(m.group(1).replace(",", "")) * UNITS.get(m.group(2), 1)


def main():
    stage, commit = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else "uniform_16384"
    field = key.split("_")[0]
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    entry = out.setdefault(key, {})
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{stage}_*_{field}16k_ncu_summary.txt"))):
        kern = os.path.basename(f)[len(stage) + 1:].split(f"_{field}16k")[0]
        if kern not in NAMES:
            continue
        t = open(f).read()
        r, w, d = metric(t, "dram__bytes_read.sum"), metric(t, "dram__bytes_write.sum"), metric(t, "gpu__time_duration.sum")
        if r is None or w is None:
            continue
        entry[NAMES[kern]] = {"dram_bytes": int(r + w), "dram_read_bytes": int(r), "dram_write_bytes": int(w),
                              "ms_under_ncu": d, "file": "profiles/" + os.path.basename(f), "commit": commit}
    out["captured_at"] = f"stage {stage}, commit {commit}: ncu --set full --clock-control none, one launch per kernel"
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
