"""Independent statement of the hot path's semantics
in pure Python, used to cross-check the C oracle on small images:

  arrival time  T(p) = max(A_p, 1 + min_q T(q)),  A_p = (img[p] << 24) | 1,
                seeds T = 0, border / img > Lmax never flood;
  segmenting    label(p) = label(first q in order down,right,left,up with T(q) < T(p));
  merging       partition at level L = components of {lvl <= L} under
                4-adjacency with at least one interior endpoint.
"""
import heapq

import numpy as np

INF = 0xFF000000
NB = ((1, 0), (0, 1), (0, -1), (-1, 0))        # lib.rs:190


def arrival_times(img, seeds_rc, lmax=254):
    H, W = img.shape
    T = np.full((H, W), INF, np.int64)
    heap = []
    for r, c in seeds_rc:
        T[r, c] = 0
    for r, c in {(int(r), int(c)) for r, c in seeds_rc}:
        heapq.heappush(heap, (0, r, c))
    while heap:
        t, r, c = heapq.heappop(heap)
        if t != T[r, c]:
            continue
        for dr, dc in NB:
            rr, cc = r + dr, c + dc
            if not (1 <= rr < H - 1 and 1 <= cc < W - 1):
                continue                               # only window centres flood (lib.rs:220)
            if img[rr, cc] > lmax or T[rr, cc] == 0:
                continue
            cand = max(t + 1, (int(img[rr, cc]) << 24) | 1)
            if cand < T[rr, cc]:
                T[rr, cc] = cand
                heapq.heappush(heap, (cand, rr, cc))
    return T


def segment_labels(T, seeds_rc):
    H, W = T.shape
    lab = np.zeros((H, W), np.int64)
    for i, (r, c) in enumerate(seeds_rc):
        lab[r, c] = i + 1
    order = np.argsort(T, axis=None, kind="stable")
    for flat in order:
        r, c = divmod(int(flat), W)
        t = T[r, c]
        if t >= INF:
            break
        if t == 0:
            continue
        for dr, dc in NB:
            rr, cc = r + dr, c + dc
            if 0 <= rr < H and 0 <= cc < W and T[rr, cc] < t:
                lab[r, c] = lab[rr, cc]
                break
    return lab


def merging_partition(T, level):
    """Component id image (0 = uncoloured) at `level`; ids are arbitrary."""
    H, W = T.shape
    col = (T < INF) & ((T >> 24) <= level)
    parent = np.arange(H * W)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for r in range(H):
        for c in range(W):
            if not col[r, c]:
                continue
            for dr, dc in ((1, 0), (0, 1)):
                rr, cc = r + dr, c + dc
                if rr >= H or cc >= W or not col[rr, cc]:
                    continue
                p_int = 1 <= r < H - 1 and 1 <= c < W - 1
                q_int = 1 <= rr < H - 1 and 1 <= cc < W - 1
                if not (p_int or q_int):
                    continue
                a, b = find(r * W + c), find(rr * W + cc)
                if a != b:
                    parent[max(a, b)] = min(a, b)
    out = np.zeros((H, W), np.int64)
    for r in range(H):
        for c in range(W):
            if col[r, c]:
                out[r, c] = find(r * W + c) + 1
    return out
