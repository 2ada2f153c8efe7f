"""Per-tile statistics of merge_reduce (diagnostic).  Needs a library built with -DWS_MERGE_STATS, e.g.
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -DWS_MERGE_STATS \\
         -o scripts/libws_stats.so rustronomy-watershed_b200/csrc/{kernels,flood,labels,merge,engine}.cu
    WS_B200_LIB=scripts/libws_stats.so python scripts/merge_stats.py 8192 uniform"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import fieldgen
from wsb200_loader import load
ws = load()
S = int(sys.argv[1]); kind = sys.argv[2]
img = fieldgen.uniform(S, S, 0) if kind == "uniform" else fieldgen.smooth(S, S, 16.0, 0) if kind == "smooth" else fieldgen.cgps_like(S, S, 0)
ctx = ws.default_context()
plan = ws.Plan(ctx, 1, S, S)
d_img = torch.from_numpy(img).cuda()
off = torch.zeros(2, dtype=torch.int32, device="cuda")
n = plan.find_local_minima(d_img.data_ptr(), 0, 0, off.data_ptr())
seeds = torch.empty((max(n, 1), 2), dtype=torch.int32, device="cuda")
plan.find_local_minima(d_img.data_ptr(), seeds.data_ptr(), n, off.data_ptr())
plan.run(1, 254, d_img.data_ptr(), seeds.data_ptr(), off.data_ptr(), n)
print(kind, S, "seeds", n, plan.stats(), plan.phase_ms())
