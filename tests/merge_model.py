"""CPU model of csrc/merge.cu (merge_reduce_kernel + the global level-ordered union): the same rounds, the same
rules (who picks, which side of a mutual pick moves, which picks are FINAL, identities through FINAL moves),
written with numpy so that the ALGORITHM can be checked against the oracle without a GPU
(tests/test_merge_model.py).  Test infrastructure only.

    lakes(L) = colours on the canvas - FINAL edges with level <= L - successful unions of DEFERRED edges <= L
"""
from __future__ import annotations

import numpy as np

TILE_W, TILE_H = 64, 32
NONE = np.iinfo(np.int64).max


def _find_roots(parent: np.ndarray) -> np.ndarray:
    p = parent.copy()
    while True:
        pp = p[p]
        if np.array_equal(pp, p):
            return p
        p = pp


def tile_boruvka(lab: np.ndarray, lvl: np.ndarray, r0: int, c0: int, rows: int, cols: int, contract: bool = True):
    """One tile.  lab / lvl: whole images.  Returns (final_edges, deferred_edges, rounds): arrays [n][3] of
    (colour a, colour b, level); DEFERRED edges are between identities."""
    r1, c1 = min(r0 + TILE_H + 1, rows), min(c0 + TILE_W + 1, cols)
    L = lab[r0:r1, c0:c1].astype(np.int64)
    V = lvl[r0:r1, c0:c1].astype(np.int64)
    H, W = L.shape
    colours, inv = np.unique(L, return_inverse=True)
    inv = inv.reshape(H, W)
    n = len(colours)
    rr, cc = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    rim = (rr == TILE_H) | (cc == TILE_W) | ((rr == 0) & (r0 > 0)) | ((cc == 0) & (c0 > 0))
    open_ = np.zeros(n, bool)
    if contract:
        open_[np.unique(inv[rim & (L != 0)])] = True
    else:
        open_[:] = True
    gr, gc = rr + r0, cc + c0
    centre = (gr >= 1) & (gr <= rows - 2) & (gc >= 1) & (gc <= cols - 2)
    intile = (rr < TILE_H) & (cc < TILE_W)
    E = []
    a, b = inv[:, :-1], inv[:, 1:]
    ok = intile[:, :-1] & (L[:, :-1] != 0) & (L[:, 1:] != 0) & (a != b) & (centre[:, :-1] | centre[:, 1:])
    E.append(np.stack([a[ok], b[ok], np.maximum(V[:, :-1], V[:, 1:])[ok]], 1))
    a, b = inv[:-1, :], inv[1:, :]
    ok = intile[:-1, :] & (L[:-1, :] != 0) & (L[1:, :] != 0) & (a != b) & (centre[:-1, :] | centre[1:, :])
    E.append(np.stack([a[ok], b[ok], np.maximum(V[:-1, :], V[1:, :])[ok]], 1))
    E = np.concatenate(E)
    empty = np.zeros((0, 3), np.int64)
    if len(E) == 0:
        return empty, empty, 0
    u, v, w = E[:, 0].copy(), E[:, 1].copy(), E[:, 2].copy()
    parent = np.arange(n)
    comp = np.arange(n)
    link = np.arange(n)              # identity links (FINAL moves of closed roots)
    fin = np.zeros(n, bool)
    moved_edge = np.full(n, -1)      # index into E of the edge an id went along
    live = np.arange(len(E))
    rounds = 0
    while True:
        rounds += 1
        # offers
        cu, cv = comp[u[live]], comp[v[live]]
        alive = cu != cv
        live, cu, cv = live[alive], cu[alive], cv[alive]
        if len(live) == 0:
            break
        key = w[live] * (1 << 32) + np.arange(len(live))
        best = np.full(n, NONE)
        np.minimum.at(best, cu, key)
        np.minimum.at(best, cv, key)
        # hooks (all decisions on the state at the start of the round)
        new_parent = parent.copy()
        for i in np.where(best != NONE)[0]:
            j = best[i] & 0xFFFFFFFF
            e = live[j]
            a_, b_ = comp[u[e]], comp[v[e]]
            other = b_ if a_ == i else a_
            mutual = best[other] == best[i]
            ci, co = not open_[i], not open_[other]
            if mutual and ((i < other) if ci == co else (not ci)):
                continue
            new_parent[i] = other
            moved_edge[i] = e
            if ci or (mutual and co):
                fin[i] = True
                if ci:
                    link[i] = v[e] if a_ == i else u[e]
        parent = new_parent
        # flatten + openness to the roots
        roots = _find_roots(parent)
        np.logical_or.at(open_, roots, open_.copy())
        parent = roots.copy()
        comp = roots
    ident = _find_roots(link)
    moved = np.where(moved_edge >= 0)[0]
    F, D = [], []
    for i in moved:
        e = moved_edge[i]
        if fin[i]:
            F.append((colours[u[e]], colours[v[e]], w[e]))
        else:
            ia, ib = ident[u[e]], ident[v[e]]
            assert ia != ib, "a DEFERRED edge inside one identity"
            D.append((colours[ia], colours[ib], w[e]))
    return (np.array(F, np.int64).reshape(-1, 3), np.array(D, np.int64).reshape(-1, 3), rounds)


def reduce_image(lab: np.ndarray, lvl: np.ndarray, contract: bool = True):
    rows, cols = lab.shape
    Fs, Ds, rounds = [], [], []
    for r0 in range(0, rows, TILE_H):
        for c0 in range(0, cols, TILE_W):
            F, D, r = tile_boruvka(lab, lvl, r0, c0, rows, cols, contract)
            Fs.append(F)
            Ds.append(D)
            rounds.append(r)
    return np.concatenate(Fs), np.concatenate(Ds), rounds


class _UF:
    def __init__(self, n):
        self.p = list(range(n))

    def find(self, x):
        while self.p[x] != x:
            self.p[x] = self.p[self.p[x]]
            x = self.p[x]
        return x

    def union(self, a, b):
        a, b = self.find(a), self.find(b)
        if a == b:
            return False
        if a < b:
            a, b = b, a
        self.p[a] = b
        return True


def lake_counts(lab: np.ndarray, lvl: np.ndarray, nseeds: int, lmax: int = 254, contract: bool = True) -> np.ndarray:
    """Lakes per level 0..=lmax the way the engine computes them."""
    F, D, _ = reduce_image(lab, lvl, contract)
    ndistinct = len(np.unique(lab[lab != 0]))
    fin_hist = np.bincount(F[:, 2], minlength=256) if len(F) else np.zeros(256, np.int64)
    uf = _UF(nseeds + 1)
    unions = np.zeros(256, np.int64)
    for a, b, w in D[np.argsort(D[:, 2], kind="stable")] if len(D) else []:
        if uf.union(int(a), int(b)):
            unions[w] += 1
    return ndistinct - np.cumsum(fin_hist + unions)[: lmax + 1]


def forest_rounds(edges: np.ndarray, n_nodes: int, open_: np.ndarray = None):
    """csrc/forest.cu: Boruvka rounds over an arbitrary edge list [n][3] = (a, b, level) with the same rules as the
    tile kernel.  Returns (FINAL picks [k][3], DEFERRED picks between identities [m][3], rounds)."""
    open_ = np.zeros(n_nodes, bool) if open_ is None else open_.copy()
    parent = np.arange(n_nodes)
    link = np.arange(n_nodes)
    orig = edges[:, :2].copy()
    cur = edges[:, :2].copy()          # current roots of the ends
    w = edges[:, 2].copy()
    F, D, rounds = [], [], 0
    alive = cur[:, 0] != cur[:, 1]
    orig, cur, w = orig[alive], cur[alive], w[alive]
    while len(cur):
        rounds += 1
        key = w * (1 << 32) + np.arange(len(cur))
        best = np.full(n_nodes, NONE)
        np.minimum.at(best, cur[:, 0], key)
        np.minimum.at(best, cur[:, 1], key)
        new_parent = parent.copy()
        for i in range(len(cur)):
            ca, cb = cur[i]
            pa, pb = best[ca] == key[i], best[cb] == key[i]
            if not (pa or pb):
                continue
            cla, clb = not open_[ca], not open_[cb]
            if pa and pb:
                a_moves = (ca > cb) if cla == clb else cla
            else:
                a_moves = pa
            mover, other = (ca, cb) if a_moves else (cb, ca)
            mover_closed = cla if a_moves else clb
            fin = mover_closed or (pa and pb and (cla or clb))
            new_parent[mover] = other
            if fin:
                if mover_closed:
                    link[mover] = orig[i][1] if a_moves else orig[i][0]
                F.append((orig[i][0], orig[i][1], w[i]))
            else:
                D.append((orig[i][0], orig[i][1], w[i]))
        parent = new_parent
        roots = _find_roots(parent)
        np.logical_or.at(open_, roots, open_.copy())
        parent = roots.copy()
        cur = roots[cur]
        alive = cur[:, 0] != cur[:, 1]
        orig, cur, w = orig[alive], cur[alive], w[alive]
    ident = _find_roots(link)
    D = np.array(D, np.int64).reshape(-1, 3)
    if len(D):
        D[:, 0], D[:, 1] = ident[D[:, 0]], ident[D[:, 1]]
    return np.array(F, np.int64).reshape(-1, 3), D, rounds


def lake_counts_forest(lab: np.ndarray, lvl: np.ndarray, nseeds: int, lmax: int = 254) -> np.ndarray:
    """Lakes per level with the forest rounds instead of the level-ordered union (what the engine runs)."""
    F, D, _ = reduce_image(lab, lvl, True)
    ndistinct = len(np.unique(lab[lab != 0]))
    F2, D2, rounds = forest_rounds(D, nseeds + 1)
    assert len(D2) == 0, "every node is closed: no DEFERRED pick"
    hist = np.bincount(np.concatenate([F[:, 2], F2[:, 2]]).astype(np.int64), minlength=256)
    return ndistinct - np.cumsum(hist)[: lmax + 1]


def partitions(lab: np.ndarray, lvl: np.ndarray, nseeds: int, levels, contract: bool = True):
    """Merging label images at `levels` from the merge tree built of ALL emitted edges (FINAL ones between the
    basins themselves, DEFERRED ones between identities), representative = smallest colour."""
    F, D, _ = reduce_image(lab, lvl, contract)
    E = np.concatenate([F, D])
    E = E[np.argsort(E[:, 2], kind="stable")] if len(E) else E
    out = {}
    uf = _UF(nseeds + 1)
    k = 0
    for L in sorted(levels):
        while k < len(E) and E[k, 2] <= L:
            uf.union(int(E[k, 0]), int(E[k, 1]))
            k += 1
        rep = np.array([uf.find(c) for c in range(nseeds + 1)], np.int64)
        out[L] = np.where(lvl <= L, rep[lab.astype(np.int64)], 0)
    return out
