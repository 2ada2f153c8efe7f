"""A numpy stand-in for CudaStrip (tests only): the same per-strip operations, stated directly from
the semantics of DESIGN.md section 2, so that the strip protocol (strips.solve + the communicators) can
be exercised on CPU with gloo.  Small images only."""
import numpy as np
import torch

INF = np.uint32(0xFF000000)
RES = np.uint32(0x80000000)


def _as_i32(a_u32):
    return torch.from_numpy(np.ascontiguousarray(a_u32).view(np.int32).copy())


def _as_u32(t):
    return t.numpy().view(np.uint32)


class NumpyStrip:
    def __init__(self, geom, local_img):
        self.geom = geom
        self.img = np.ascontiguousarray(local_img, dtype=np.uint8)
        self.rows, self.cols = self.img.shape
        im = self.img.astype(np.int32)
        seeds = []
        for r in range(1, self.rows - 1):
            for c in range(1, self.cols - 1):
                nb = im[r - 1:r + 2, c - 1:c + 2].copy()
                t = nb[1, 1]
                nb[1, 1] = -1
                if (nb < t).all():
                    seeds.append((r, c))
        self.seeds = seeds
        self.nseeds = len(seeds)

    # rows of the plan that the strip owns
    @property
    def r0(self):
        return 1 if self.geom.halo_top else 0

    @property
    def r1(self):
        return self.rows - 1 - (1 if self.geom.halo_bottom else 0)

    def _relax(self):
        T, A = self.T, self.A
        while True:
            P = np.pad(T, 1, constant_values=INF).astype(np.uint64)
            m = np.minimum(np.minimum(P[2:, 1:-1], P[:-2, 1:-1]), np.minimum(P[1:-1, 2:], P[1:-1, :-2]))
            new = np.minimum(T.astype(np.uint64), np.maximum(A.astype(np.uint64), m + 1)).astype(np.uint32)
            if np.array_equal(new, T):
                return
            T[:] = new

    def begin(self, kind, lmax, colour_base):
        self.kind, self.lmax, self.base = kind, lmax, colour_base
        img = self.img.astype(np.uint32)
        flood = np.zeros(img.shape, bool)
        flood[1:-1, 1:-1] = True                      # local border rows: field border or halo
        flood &= img <= lmax
        self.A = np.where(flood, (img << 24) | 1, INF).astype(np.uint32)
        self.T = np.full(img.shape, INF, np.uint32)
        self.seedlab = np.zeros(img.shape, np.uint32)
        for i, (r, c) in enumerate(self.seeds):
            self.T[r, c] = 0
            self.seedlab[r, c] = colour_base + i + 1
        self._relax()

    def export_times(self):
        top = _as_i32(self.T[self.r0]) if self.geom.halo_top else None
        bot = _as_i32(self.T[self.r1]) if self.geom.halo_bottom else None
        return top, bot

    def import_times(self, top, bottom):
        changed = False
        for row, buf in ((0, top), (self.rows - 1, bottom)):
            if buf is None:
                continue
            v = _as_u32(buf)
            lower = v < self.T[row]
            if lower.any():
                self.T[row] = np.minimum(self.T[row], v)
                changed = True
        if changed:
            self._relax()
        return changed

    def _resolve(self):
        T, lab = self.T, self.lab
        order = np.argsort(T, axis=None, kind="stable")
        for flat in order:
            r, c = divmod(int(flat), self.cols)
            if T[r, c] >= INF:
                break
            if lab[r, c] & RES or self.halo[r]:
                continue
            pr, pc = self.parent[r][c]
            if lab[pr, pc] & RES:
                lab[r, c] = lab[pr, pc]

    def labels(self):
        T = self.T
        self.halo = np.zeros(self.rows, bool)
        if self.geom.halo_top:
            self.halo[0] = True
        if self.geom.halo_bottom:
            self.halo[-1] = True
        self.lvl = np.where(T >= INF, 255, T >> 24).astype(np.uint8)
        self.lab = np.zeros(T.shape, np.uint32)                 # 0 = pending
        self.lab[T >= INF] = RES
        own_seed = (self.seedlab != 0)
        self.lab[own_seed] = RES | self.seedlab[own_seed]
        self.parent = [[None] * self.cols for _ in range(self.rows)]
        for r in range(self.rows):
            if self.halo[r]:
                self.lab[r][T[r] < INF] = 0                     # owned by the neighbour: pending
                continue
            for c in range(self.cols):
                t = T[r, c]
                if t >= INF or t == 0:
                    continue
                for dr, dc in ((1, 0), (0, 1), (0, -1), (-1, 0)):
                    if T[r + dr, c + dc] < t:
                        self.parent[r][c] = (r + dr, c + dc)
                        break
        self._resolve()

    def export_labels(self):
        top = _as_i32(self.lab[self.r0]) if self.geom.halo_top else None
        bot = _as_i32(self.lab[self.r1]) if self.geom.halo_bottom else None
        return top, bot

    def import_labels(self, top, bottom):
        for row, buf in ((0, top), (self.rows - 1, bottom)):
            if buf is None:
                continue
            v = _as_u32(buf)
            take = ((v & RES) != 0) & ((self.lab[row] & RES) == 0)
            self.lab[row][take] = v[take]
        self._resolve()
        return int(((self.lab[self.r0:self.r1 + 1] & RES) == 0).sum())

    def edges(self):
        g = self.geom
        lab = (self.lab & ~RES).astype(np.int64)
        off = g.local_rows[0]

        def centre(r, c):
            return 1 <= r + off <= g.global_rows - 2 and 1 <= c <= self.cols - 2
        ab, w = [], []
        for r in range(self.r0, self.r1 + 1):
            for c in range(self.cols):
                a = lab[r, c]
                if a == 0:
                    continue
                for rr, cc in ((r, c + 1), (r + 1, c)):
                    if rr >= self.rows or cc >= self.cols:
                        continue
                    b = lab[rr, cc]
                    if b == 0 or b == a or not (centre(r, c) or centre(rr, cc)):
                        continue
                    ab.append((a - 1, b - 1))
                    w.append(max(int(self.lvl[r, c]), int(self.lvl[rr, cc])))
        nd = sum(1 for i, (r, c) in enumerate(self.seeds) if lab[r, c] == self.base + i + 1)
        return (torch.tensor(ab, dtype=torch.int32).reshape(-1, 2), torch.tensor(w, dtype=torch.uint8), nd)

    def union(self, ab, w, ncolours, ndistinct, lmax):
        parent = list(range(ncolours))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        counts = np.zeros(256, np.uint32)
        n = ndistinct
        edges = sorted(zip(w.tolist(), ab.tolist())) if ab is not None and len(w) else []
        k = 0
        for l in range(lmax + 1):
            while k < len(edges) and edges[k][0] == l:
                a, b = find(edges[k][1][0]), find(edges[k][1][1])
                if a != b:
                    parent[max(a, b)] = min(a, b)
                    n -= 1
                k += 1
            counts[l] = n
        return counts

    def owned_labels(self):
        return (self.lab & ~RES)[self.r0:self.r1 + 1]

    def owned_levels(self):
        return self.lvl[self.r0:self.r1 + 1]
