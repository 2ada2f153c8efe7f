"""Full-size checks (BASELINE.json sizes) through size-independent properties -- the oracle cannot run
a 16384^2 field, but the result is characterised completely by properties that torch can verify on the GPU:

  * arrival times: T is THE fixed point of  T(p) = max(A(p), 1 + min over 4-neighbours T(q))  with seeds at 0
    (the fixed point is unique and equals the reference's nested loops;
    tests/test_oracle_semantics.py checks that statement against the literal loop on small fields);
  * labels: every coloured non-seed pixel carries the label of its parent = the first neighbour in the order
    down, right, left, up with a smaller arrival time (lib.rs:190, 245); seeds carry index + 1;
  * merging: the per-level lake counts (tile-contracted fast path: FINAL edges only counted) equal the number
    of distinct representatives in the per-level snapshots (merge tree built from all edges by the global
    level-ordered union-find) -- two different algorithms -- and never increase with the level.
"""
import numpy as np
import pytest
import torch

import fieldgen
from conftest import big_field
from wsb200_loader import load

pytestmark = pytest.mark.gpu

INF = 0xFF000000


class _Dev:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(ptr, shape, typestr):
    return torch.as_tensor(_Dev(ptr, shape, typestr), device="cuda")


def _u32(ptr, shape):
    return _view(ptr, shape, "<i4").to(torch.int64) & 0xFFFFFFFF


def _shift(x, dr, dc, fill):
    """x[r + dr, c + dc] with `fill` outside."""
    out = torch.full_like(x, fill)
    R, C = x.shape
    rs, re = max(0, -dr), min(R, R - dr)
    cs, ce = max(0, -dc), min(C, C - dc)
    out[rs:re, cs:ce] = x[rs + dr:re + dr, cs + dc:ce + dc]
    return out


def _run(kind, S, field):
    ws = load()
    img = big_field(field, S)
    ctx = ws.default_context()
    plan = ws.Plan(ctx, 1, S, S)
    d_img = torch.from_numpy(img).cuda()
    off = torch.zeros(2, dtype=torch.int32, device="cuda")
    n = plan.find_local_minima(d_img.data_ptr(), 0, 0, off.data_ptr())
    seeds = torch.empty((max(n, 1), 2), dtype=torch.int32, device="cuda")
    plan.find_local_minima(d_img.data_ptr(), seeds.data_ptr(), n, off.data_ptr())
    plan.run(kind, 254, d_img.data_ptr(), seeds.data_ptr(), off.data_ptr(), n)
    return ws, ctx, plan, d_img, seeds, n


@pytest.mark.parametrize("S,field", [(16384, "uniform"), (8192, "smooth")])
def test_arrival_times_and_labels_full_size(S, field):
    ws, ctx, plan, d_img, seeds, n = _run(0, S, field)
    T = _u32(plan.arrival_times_ptr, (S, S))
    img = d_img.to(torch.int64)
    interior = torch.zeros((S, S), dtype=torch.bool, device="cuda")
    interior[1:-1, 1:-1] = True
    A = torch.where(interior & (img <= 254), (img << 24) | 1, torch.full_like(img, INF))
    is_seed = torch.zeros((S, S), dtype=torch.bool, device="cuda")
    sr, sc = seeds[:n, 0].long(), seeds[:n, 1].long()
    is_seed[sr, sc] = True
    m = torch.minimum(torch.minimum(_shift(T, 1, 0, INF), _shift(T, -1, 0, INF)),
                      torch.minimum(_shift(T, 0, 1, INF), _shift(T, 0, -1, INF)))
    never = (A >= INF) | (m >= INF)
    expect = torch.maximum(A, m + 1)
    ok = torch.where(is_seed, T == 0, torch.where(never, T >= INF, T == expect))
    assert bool(ok.all()), f"{int((~ok).sum())} pixels violate the fixed-point equation"
    assert int((T == 0).sum()) == n
    del m, never, expect, ok, A

    lab = _u32(plan.labels_ptr, (S, S))
    assert bool((lab >> 31 == 1).all()), "unresolved label words"
    lab = lab & 0x7FFFFFFF
    lvl = _view(plan.levels_ptr, (S, S), "|u1").to(torch.int64)
    assert bool((lvl == torch.where(T >= INF, torch.full_like(T, 255), T >> 24)).all())
    coloured = T < INF
    assert bool((lab[~coloured] == 0).all())
    assert bool((lab[sr, sc] == torch.arange(1, n + 1, device="cuda")).all())
    parent_lab = _shift(lab, -1, 0, 0)                                              # up (last choice)
    parent_lab = torch.where(_shift(T, 0, -1, INF) < T, _shift(lab, 0, -1, 0), parent_lab)   # left
    parent_lab = torch.where(_shift(T, 0, 1, INF) < T, _shift(lab, 0, 1, 0), parent_lab)     # right
    parent_lab = torch.where(_shift(T, 1, 0, INF) < T, _shift(lab, 1, 0, 0), parent_lab)     # down (first choice)
    inner = coloured & ~is_seed
    assert bool((lab[inner] == parent_lab[inner]).all())
    assert bool((lab[inner] > 0).all())
    plan.close()


@pytest.mark.parametrize("S,field", [(16384, "uniform"), (8192, "smooth")])
def test_lake_counts_against_merge_tree_full_size(S, field):
    ws, ctx, plan, d_img, seeds, n = _run(1, S, field)
    counts = _view(plan.lake_counts_ptr, (256,), "<i4").to(torch.int64).cpu().numpy()
    assert counts[254] >= 1 and counts[255] == 0
    assert np.all(np.diff(counts[:255]) <= 0), "lake counts must not increase with the water level"
    assert counts[0] <= n
    out = torch.empty((S, S), dtype=torch.int64, device="cuda")
    for level in (0, 37, 128, 254):
        plan.snapshot(1, 0, level, out.data_ptr())     # builds the merge tree on the first call
        torch.cuda.synchronize()
        u = torch.unique(out)
        lakes = int(u.numel()) - int((u == 0).any())
        assert lakes == counts[level], (level, lakes, int(counts[level]))
        del u
    plan.close()


def test_one_long_dependency_chain():
    """A serpentine corridor: the whole flood is ONE chain of ~half a million dependent hops through ~8000 tile
    visits, almost every CTA idles all the time (the worklist's idle / termination path), and the answer is
    known in closed form: the hop count along the corridor."""
    S = 1024
    ws = load()
    img = fieldgen.maze(S, S)
    ctx = ws.default_context()
    plan = ws.Plan(ctx, 1, S, S)
    d_img = torch.from_numpy(img).cuda()
    off = torch.tensor([0, 1], dtype=torch.int32, device="cuda")
    seeds = torch.tensor([[1, 1]], dtype=torch.int32, device="cuda")
    plan.run(0, 254, d_img.data_ptr(), seeds.data_ptr(), off.data_ptr(), 1)
    T = _u32(plan.arrival_times_ptr, (S, S)).cpu().numpy()
    lab = (_u32(plan.labels_ptr, (S, S)) & 0x7FFFFFFF).cpu().numpy()
    # walk the corridor on the host: the k-th pixel after the seed floods in pass k of level 0
    open_px = (img == 0)
    open_px[0, :] = open_px[-1, :] = False
    open_px[:, 0] = open_px[:, -1] = False
    expect = np.full((S, S), INF, np.int64)
    r, c, k, prev = 1, 1, 0, None
    while True:
        expect[r, c] = k
        nxt = [(rr, cc) for rr, cc in ((r + 1, c), (r, c + 1), (r, c - 1), (r - 1, c))
               if open_px[rr, cc] and (rr, cc) != prev and expect[rr, cc] == INF]
        if not nxt:
            break
        prev, (r, c), k = (r, c), nxt[0], k + 1
    assert k > 400000
    assert np.array_equal(np.minimum(T, INF), expect)
    assert np.array_equal(lab, (expect < INF).astype(np.int64))
    plan.close()
