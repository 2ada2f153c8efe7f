import sys, importlib, numpy as np, torch
sys.path.insert(0, '.')
import fieldgen
from wsb200_loader import load
ws = load(); st = importlib.import_module("rustronomy_watershed_b200.strips")
S = int(sys.argv[1]); n = int(sys.argv[2])
img = fieldgen.uniform(S, S, 0)
ctx = ws.default_context()
m = ws.TransformBuilder.default().build_merging()
seeds = m.find_local_minima(img)
ref_lakes, _ = m.lake_counts(img, seeds)
parts = st.partition_rows(S, n)
strips = []
for sid in range(n):
    g = st.StripGeometry(sid, n, S, parts[sid]); lo, hi = g.local_rows
    strips.append(st.CudaStrip(ws, ctx, g, torch.from_numpy(img[lo:hi].copy()).cuda()))
res = st.solve(strips, st.LocalComm(n), st.MERGING, 254)
print("edges", res.edges_total, "seeds", res.nseeds_total, len(seeds))
print("lakes equal", np.array_equal(res.lake_counts, ref_lakes), res.lake_counts[[0, 64, 127, 254]], ref_lakes[[0, 64, 127, 254]])
# edge sanity
abs_, ws_ = [], []
for s in strips:
    ab, w, nd = s.edges(); abs_.append(ab); ws_.append(w)
ab = torch.cat(abs_).cpu().numpy().astype(np.int64); w = torch.cat(ws_).cpu().numpy()
print("max id", ab.max(), "min id", ab.min(), "ncolours", res.nseeds_total, "self loops", int((ab[:, 0] == ab[:, 1]).sum()))
# host Kruskal on the gathered edges
order = np.argsort(w, kind="stable")
parent = np.arange(res.nseeds_total)
def find(x):
    while parent[x] != x:
        parent[x] = parent[parent[x]]; x = parent[x]
    return x
cnt = res.nseeds_total; unions = np.zeros(256, np.int64)
if S <= 2048:
    for i in order:
        a, b = find(ab[i, 0]), find(ab[i, 1])
        if a != b:
            parent[max(a, b)] = min(a, b); unions[w[i]] += 1
    host = res.nseeds_total - np.cumsum(unions)[:255]
    print("host kruskal equals ref", np.array_equal(host.astype(np.uint64), ref_lakes), host[[0, 64, 127, 254]])
