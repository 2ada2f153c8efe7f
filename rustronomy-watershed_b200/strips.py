"""Row-strip decomposition of ONE large field over several GPUs.

The arrival-time fixed point does not depend on how the field is cut, so a
strip solve is bit-exact with the single-GPU solve.  Protocol, per strip (= one ws_plan over the
strip's rows plus one halo row per neighbour):

  1. seeds of the owned rows; colour base = number of seeds in the strips above (all-gather);
  2. local flood to a fixed point, then repeat { send first / last owned row of arrival times to the
     neighbours (NCCL send/recv over NVLink), min-merge the received rows into the halo rows, re-flood
     from the tiles that saw a lower value } until no halo value changed anywhere (all-reduce max);
  3. labels: parent pointers + pointer jumping inside the strip, halo pixels pending; repeat
     { exchange boundary-row labels, jump again } until no owned pixel is pending (all-reduce sum);
  4. merging: every strip reduces the basin edges of its tiles to a spanning forest; the forests are
     all-gathered and one Kruskal over the union gives the lakes per level (boundary union-merge).

`solve()` is written against two small interfaces so the same driver runs (a) one strip per rank under
torch.distributed (NCCL on GPUs, gloo in the CPU tests), and (b) several strips in one process.

On the GPU the protocol runs WITHOUT a host round trip per exchange round (`solve_async`): everything -- the
library's kernels, the row copies, NCCL -- is enqueued on the library's stream, the "did anything change" /
"is anything pending" answers accumulate in device words, and the host looks at them once per batch of rounds
(rounds after the fixed point change nothing and cost a few tens of microseconds).  Step 4 exchanges no edge
lists: every strip reduces its basin graph to the number of certain forest edges per level plus the few forest
edges between basins on its boundary rows (csrc/forest.cu), and one all-gather of those packets (a few
hundred KB per strip) feeds a last round of the same reduction on every rank.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

SEGMENTING, MERGING = 0, 1


def partition_rows(global_rows: int, n_strips: int) -> List[Tuple[int, int]]:
    """Contiguous, nearly equal owned row ranges [r0, r1) of the strips."""
    if n_strips < 1 or global_rows < n_strips:
        raise ValueError("need at least one row per strip")
    base, extra = divmod(global_rows, n_strips)
    out, r = [], 0
    for s in range(n_strips):
        h = base + (1 if s < extra else 0)
        out.append((r, r + h))
        r += h
    return out


@dataclass
class StripGeometry:
    sid: int
    n_strips: int
    global_rows: int
    own: Tuple[int, int]            # owned rows [r0, r1) of the field

    @property
    def halo_top(self) -> bool:
        return self.sid > 0

    @property
    def halo_bottom(self) -> bool:
        return self.sid < self.n_strips - 1

    @property
    def local_rows(self) -> Tuple[int, int]:     # rows of the field held by the plan, incl. halos
        return self.own[0] - (1 if self.halo_top else 0), self.own[1] + (1 if self.halo_bottom else 0)


# --------------------------------------------------------------------------------------------
# communication
# --------------------------------------------------------------------------------------------

class LocalComm:
    """All strips live in this process (one GPU, or a test)."""

    def __init__(self, n_strips: int):
        self.n_strips = n_strips
        self.local_ids = list(range(n_strips))

    def exchange(self, outgoing: Dict[Tuple[int, str], torch.Tensor]) -> Dict[Tuple[int, str], torch.Tensor]:
        inc = {}
        for (sid, side), buf in outgoing.items():
            if side == "top" and sid > 0:
                inc[(sid - 1, "bottom")] = buf.clone()
            if side == "bottom" and sid < self.n_strips - 1:
                inc[(sid + 1, "top")] = buf.clone()
        return inc

    def allreduce_max(self, x: int) -> int:
        return x

    def allreduce_sum(self, x: int) -> int:
        return x

    def allgather_ints(self, local: Dict[int, int]) -> List[int]:
        return [local[s] for s in range(self.n_strips)]

    def allreduce_sum_vec(self, v: np.ndarray) -> np.ndarray:
        return v

    def allgather_edges(self, ab: List[torch.Tensor], w: List[torch.Tensor]):
        return torch.cat(ab) if ab else None, torch.cat(w) if w else None

    # -- asynchronous protocol: nothing below waits for the device -------------------------------------
    def exchange_rows(self, strips: Dict[int, "CudaStrip"], kind: str):
        """send_{top,bottom} of every strip -> recv_{bottom,top} of its neighbour (device copies, stream ordered)."""
        for sid, s in strips.items():
            snd = s.rows_buf[kind]
            if sid > 0:
                strips[sid - 1].rows_buf[kind]["recv_bottom"].copy_(snd["send_top"], non_blocking=True)
            if sid < self.n_strips - 1:
                strips[sid + 1].rows_buf[kind]["recv_top"].copy_(snd["send_bottom"], non_blocking=True)

    def reduce_flag(self, t: torch.Tensor, op: str) -> int:
        return int(t.item())                      # the only wait of a batch of rounds

    def allgather_packets(self, packets: List[torch.Tensor]) -> torch.Tensor:
        return torch.cat(packets)


class DistComm:
    """One strip per rank of a torch.distributed process group (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.n_strips = dist.get_rank(group), dist.get_world_size(group)
        self.local_ids = [self.rank]

    def exchange(self, outgoing):
        dist, r, n = self.dist, self.rank, self.n_strips
        ops, inc = [], {}
        if r > 0:
            inc[(r, "top")] = torch.empty_like(outgoing[(r, "top")])
            ops.append(dist.P2POp(dist.isend, outgoing[(r, "top")], r - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, inc[(r, "top")], r - 1, self.group))
        if r < n - 1:
            inc[(r, "bottom")] = torch.empty_like(outgoing[(r, "bottom")])
            ops.append(dist.P2POp(dist.isend, outgoing[(r, "bottom")], r + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, inc[(r, "bottom")], r + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            if next(iter(inc.values())).is_cuda:
                torch.cuda.synchronize()
        return inc

    def _reduce(self, x: int, op, device) -> int:
        t = torch.tensor([x], dtype=torch.int64, device=device)
        self.dist.all_reduce(t, op=op, group=self.group)
        return int(t.item())

    def allreduce_max(self, x: int) -> int:
        return self._reduce(x, self.dist.ReduceOp.MAX, self._dev)

    def allreduce_sum(self, x: int) -> int:
        return self._reduce(x, self.dist.ReduceOp.SUM, self._dev)

    def allreduce_sum_vec(self, v: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.int64)).to(self._dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    _dev = "cpu"

    def set_device(self, device):
        self._dev = device

    def allgather_ints(self, local):
        t = torch.tensor([local[self.rank]], dtype=torch.int64, device=self._dev)
        out = [torch.zeros_like(t) for _ in range(self.n_strips)]
        self.dist.all_gather(out, t, group=self.group)
        return [int(v.item()) for v in out]

    # -- asynchronous protocol ---------------------------------------------------------------------------
    def exchange_rows(self, strips, kind: str):
        dist, r, n = self.dist, self.rank, self.n_strips
        b = strips[r].rows_buf[kind]
        ops = []
        if r > 0:
            ops.append(dist.P2POp(dist.isend, b["send_top"], r - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, b["recv_top"], r - 1, self.group))
        if r < n - 1:
            ops.append(dist.P2POp(dist.isend, b["send_bottom"], r + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, b["recv_bottom"], r + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()                        # orders the current stream behind the transfer; the host goes on

    def reduce_flag(self, t: torch.Tensor, op: str) -> int:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def allgather_packets(self, packets):
        out = torch.empty(self.n_strips * packets[0].numel(), dtype=torch.uint8, device=packets[0].device)
        self.dist.all_gather_into_tensor(out, packets[0], group=self.group)
        return out

    def allgather_edges(self, ab, w):
        """Variable-length all-gather: sizes first, then padded buffers."""
        a, b = ab[0], w[0]
        sizes = self.allgather_ints({self.rank: int(a.shape[0])})
        m = max(max(sizes), 1)
        pa = torch.zeros((m, 2), dtype=a.dtype, device=a.device)
        pw = torch.zeros((m,), dtype=b.dtype, device=b.device)
        pa[: a.shape[0]] = a
        pw[: b.shape[0]] = b
        ga = [torch.empty_like(pa) for _ in range(self.n_strips)]
        gw = [torch.empty_like(pw) for _ in range(self.n_strips)]
        self.dist.all_gather(ga, pa, group=self.group)
        self.dist.all_gather(gw, pw, group=self.group)
        return (torch.cat([g[:n] for g, n in zip(ga, sizes)]), torch.cat([g[:n] for g, n in zip(gw, sizes)]))


# --------------------------------------------------------------------------------------------
# the CUDA backend: one ws_plan per strip
# --------------------------------------------------------------------------------------------

class CudaStrip:
    """A strip on the GPU of `ctx`: owns the plan, the local image and the exchange buffers."""

    def __init__(self, ws, ctx, geom: StripGeometry, local_img: torch.Tensor):
        assert local_img.is_cuda and local_img.dtype == torch.uint8 and local_img.is_contiguous()
        self.ws, self.ctx, self.geom, self.img = ws, ctx, geom, local_img
        self.rows, self.cols = int(local_img.shape[0]), int(local_img.shape[1])
        lr = geom.local_rows
        assert self.rows == lr[1] - lr[0]
        self.dev = local_img.device
        self.plan = ws.Plan(ctx, 1, self.rows, self.cols)
        self.off = torch.zeros(2, dtype=torch.int32, device=self.dev)
        self.nseeds = self.plan.find_local_minima(self.img.data_ptr(), 0, 0, self.off.data_ptr())
        self.seeds = torch.empty((max(self.nseeds, 1), 2), dtype=torch.int32, device=self.dev)
        if self.nseeds:
            self.plan.find_local_minima(self.img.data_ptr(), self.seeds.data_ptr(), self.nseeds, self.off.data_ptr())
        self.colour_base = 0

        # the asynchronous protocol: everything runs on the library's stream, row buffers are reused
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=self.dev)
        with torch.cuda.stream(self.stream):
            self.rows_buf = {k: {n: torch.empty(self.cols, dtype=torch.int32, device=self.dev)
                                 for n in ("send_top", "send_bottom", "recv_top", "recv_bottom")}
                             for k in ("times", "labels")}
            self.packet = torch.empty(self.plan.strip_packet_bytes(), dtype=torch.uint8, device=self.dev)

    def _row(self):
        return torch.empty(self.cols, dtype=torch.int32, device=self.dev)

    def _ptr(self, kind: str, name: str, present: bool) -> int:
        return self.rows_buf[kind][name].data_ptr() if present else 0

    def begin_async(self, kind: int, lmax: int, colour_base: int):
        g = self.geom
        self.colour_base = colour_base
        self.plan.strip_begin_async(kind, lmax, g.global_rows, g.local_rows[0], g.halo_top, g.halo_bottom, colour_base,
                                    self.img.data_ptr(), self.seeds.data_ptr(), self.nseeds)

    def export_async(self, kind: str):
        g = self.geom
        fn = self.plan.strip_export_times_async if kind == "times" else self.plan.strip_export_labels_async
        fn(self._ptr(kind, "send_top", g.halo_top), self._ptr(kind, "send_bottom", g.halo_bottom))

    def import_async(self, kind: str, acc: torch.Tensor):
        g = self.geom
        fn = self.plan.strip_import_times_async if kind == "times" else self.plan.strip_import_labels_async
        fn(self._ptr(kind, "recv_top", g.halo_top), self._ptr(kind, "recv_bottom", g.halo_bottom), acc.data_ptr())

    def labels_async(self):
        self.plan.strip_labels_async()

    def forest_async(self, ncolours_total: int) -> torch.Tensor:
        self.plan.strip_forest(ncolours_total, self.packet.data_ptr())
        return self.packet

    def lakes_from_packets(self, packets: torch.Tensor, n_packets: int, ncolours_total: int, lmax: int) -> np.ndarray:
        self.plan.forest_packets(packets.data_ptr(), n_packets, ncolours_total, lmax)
        counts = self.ctx.d2h(self.plan.lake_counts_ptr, (256,), np.uint32)      # (waits for the stream)
        self.plan.strip_check()
        return counts

    def check(self):
        self.plan.strip_check()

    def begin(self, kind: int, lmax: int, colour_base: int):
        g = self.geom
        self.colour_base = colour_base
        self.plan.strip_begin(kind, lmax, g.global_rows, g.local_rows[0], g.halo_top, g.halo_bottom, colour_base,
                              self.img.data_ptr(), self.seeds.data_ptr(), self.nseeds)

    def export_times(self):
        top = self._row() if self.geom.halo_top else None
        bot = self._row() if self.geom.halo_bottom else None
        self.plan.strip_export_times(top.data_ptr() if top is not None else 0, bot.data_ptr() if bot is not None else 0)
        return top, bot

    def import_times(self, top, bottom) -> bool:
        torch.cuda.current_stream(self.dev).synchronize()      # rows received / cloned on torch's stream
        return self.plan.strip_import_times(top.data_ptr() if top is not None else 0,
                                            bottom.data_ptr() if bottom is not None else 0)

    def labels(self):
        self.plan.strip_labels()

    def export_labels(self):
        top = self._row() if self.geom.halo_top else None
        bot = self._row() if self.geom.halo_bottom else None
        self.plan.strip_export_labels(top.data_ptr() if top is not None else 0,
                                      bot.data_ptr() if bot is not None else 0)
        return top, bot

    def import_labels(self, top, bottom) -> int:
        torch.cuda.current_stream(self.dev).synchronize()
        return self.plan.strip_import_labels(top.data_ptr() if top is not None else 0,
                                             bottom.data_ptr() if bottom is not None else 0)

    def edges(self):
        d_ab, d_w, n, nd = self.plan.strip_edges()
        ab = torch.empty((n, 2), dtype=torch.int32, device=self.dev)
        w = torch.empty((n,), dtype=torch.uint8, device=self.dev)
        if n:
            self.ctx.d2d(ab.data_ptr(), d_ab, n * 8)
            self.ctx.d2d(w.data_ptr(), d_w, n)
        # FINAL edges (bit 31 of the second colour: negative as int32) are certain forest edges inside this
        # strip: they are counted per level here and never leave the GPU; only the DEFERRED ones are gathered
        fin = ab[:, 1] < 0
        self.fin_hist = torch.bincount(w[fin].to(torch.int64), minlength=256).cpu().numpy().astype(np.int64)
        keep = ~fin
        return ab[keep].contiguous(), w[keep].contiguous(), nd

    def union(self, ab, w, ncolours: int, ndistinct: int, lmax: int) -> np.ndarray:
        n = int(ab.shape[0]) if ab is not None else 0
        # ab / w were produced on torch's stream (all-gather, cat); the library runs on its own stream
        torch.cuda.current_stream(self.dev).synchronize()
        self.plan.union_edges(ab.data_ptr() if n else 0, w.data_ptr() if n else 0, n, ncolours, ndistinct, lmax)
        return self.ctx.d2h(self.plan.lake_counts_ptr, (256,), np.uint32)

    # results of the owned rows, on the host
    def owned_labels(self) -> np.ndarray:
        lab = self.ctx.d2h(self.plan.labels_ptr, (self.rows, self.cols), np.uint32) & 0x7FFFFFFF
        return lab[(1 if self.geom.halo_top else 0): self.rows - (1 if self.geom.halo_bottom else 0)]

    def owned_levels(self) -> np.ndarray:
        lvl = self.ctx.d2h(self.plan.levels_ptr, (self.rows, self.cols), np.uint8)
        return lvl[(1 if self.geom.halo_top else 0): self.rows - (1 if self.geom.halo_bottom else 0)]

    def close(self):
        self.plan.close()


# --------------------------------------------------------------------------------------------
# the driver
# --------------------------------------------------------------------------------------------

@dataclass
class StripResult:
    lake_counts: Optional[np.ndarray]     # merging: lakes at levels 0..=max
    flood_rounds: int
    label_rounds: int
    nseeds_total: int
    edges_total: int
    phase_s: Optional[dict] = None        # wall seconds per phase of the protocol (this process)


def solve_async(strips: Sequence, comm, kind: int = MERGING, max_water_level: int = 254, max_rounds: int = 100000,
                first_batch: int = 4, batch: int = 2) -> StripResult:
    """The protocol on GPUs with the exchange rounds kept in flight: the host waits once per batch of rounds."""
    import time
    by_id = {s.geom.sid: s for s in strips}
    assert sorted(by_id) == sorted(comm.local_ids)
    any_strip = next(iter(by_id.values()))
    phase_s, t_last = {}, time.perf_counter()

    def lap(name):
        nonlocal t_last
        now = time.perf_counter()
        phase_s[name] = phase_s.get(name, 0.0) + now - t_last
        t_last = now

    counts = comm.allgather_ints({sid: s.nseeds for sid, s in by_id.items()})
    bases = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ncolours = int(bases[-1])
    lap("seed counts (all-gather)")
    with torch.cuda.stream(any_strip.stream):
        acc = torch.zeros(1, dtype=torch.int32, device=any_strip.dev)
        for sid, s in by_id.items():
            s.begin_async(kind, max_water_level, int(bases[sid]))

        def rounds_until(kind_name: str, op: str) -> int:
            """Exchange rounds in batches; stops when the LAST round of a batch left `acc` at zero everywhere."""
            done, size = 0, first_batch
            while True:
                for r in range(size):
                    if r == size - 1:
                        acc.zero_()                      # only the batch's last round decides
                    for s in by_id.values():
                        s.export_async(kind_name)
                    comm.exchange_rows(by_id, kind_name)
                    for s in by_id.values():
                        s.import_async(kind_name, acc)
                done += size
                if comm.reduce_flag(acc, op) == 0:
                    return done
                if done >= max_rounds:
                    raise RuntimeError("strip exchange did not converge")
                size = batch

        flood_rounds = rounds_until("times", "max")
        lap("arrival-time exchange rounds")
        for s in by_id.values():
            s.labels_async()
        label_rounds = rounds_until("labels", "sum")
        # Every OWNED pixel is resolved now, but a halo row still shows what its owner exported at the start of
        # the last round; the edges towards the halo rows need the final words: one more exchange.
        for s in by_id.values():
            s.export_async("labels")
        comm.exchange_rows(by_id, "labels")
        for s in by_id.values():
            s.import_async("labels", acc)
        label_rounds += 1
        for s in by_id.values():
            s.plan.strip_labels_finish_async()       # the label plane itself, once
        lap("labels + label exchange rounds")

        lake_counts, edges_total = None, 0
        if kind == MERGING:
            packets = [by_id[sid].forest_async(ncolours) for sid in sorted(by_id)]
            gathered = comm.allgather_packets(packets)
            first = by_id[sorted(by_id)[0]]
            counts256 = first.lakes_from_packets(gathered, comm.n_strips, ncolours, max_water_level)
            lake_counts = counts256[: max_water_level + 1].astype(np.uint64)
            hdr = gathered.view(torch.int32)[: 4].cpu().numpy() if comm.n_strips == 1 else None
            edges_total = int(hdr[0]) if hdr is not None else 0
            lap("strip forests, packet all-gather, boundary forest")
        for s in by_id.values():
            s.check()
    return StripResult(lake_counts, flood_rounds, label_rounds, ncolours, edges_total, phase_s)


def solve(strips: Sequence, comm, kind: int = MERGING, max_water_level: int = 254, max_rounds: int = 100000) -> StripResult:
    """Run the protocol over the strips this process holds (`strips[i]` has geometry sid = comm.local_ids[i])."""
    if strips and all(hasattr(s, "begin_async") for s in strips) and hasattr(comm, "exchange_rows"):
        return solve_async(strips, comm, kind, max_water_level, max_rounds)
    return solve_sync(strips, comm, kind, max_water_level, max_rounds)


def solve_sync(strips: Sequence, comm, kind: int = MERGING, max_water_level: int = 254, max_rounds: int = 100000) -> StripResult:
    """The protocol with a host decision after every exchange round (the CPU test backend; the first GPU version)."""
    import time
    by_id = {s.geom.sid: s for s in strips}
    assert sorted(by_id) == sorted(comm.local_ids)
    phase_s, t_last = {}, time.perf_counter()

    def lap(name):
        nonlocal t_last
        now = time.perf_counter()
        phase_s[name] = phase_s.get(name, 0.0) + now - t_last
        t_last = now
    counts = comm.allgather_ints({sid: s.nseeds for sid, s in by_id.items()})
    bases = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    for sid, s in by_id.items():
        s.begin(kind, max_water_level, int(bases[sid]))
    lap("begin: state, seeds, first flood")

    def exchange(exporter, importer, reducer, stop_when):
        rounds = 0
        while True:
            out = {}
            for sid, s in by_id.items():
                top, bot = exporter(s)
                if top is not None:
                    out[(sid, "top")] = top
                if bot is not None:
                    out[(sid, "bottom")] = bot
            inc = comm.exchange(out)
            acc = 0
            for sid, s in by_id.items():
                acc = reducer(acc, importer(s, inc.get((sid, "top")), inc.get((sid, "bottom"))))
            rounds += 1
            if stop_when(acc):
                return rounds
            if rounds >= max_rounds:
                raise RuntimeError("strip exchange did not converge")

    flood_rounds = exchange(lambda s: s.export_times(), lambda s, t, b: int(s.import_times(t, b)),
                            max, lambda acc: comm.allreduce_max(acc) == 0)
    lap("arrival-time exchange rounds")
    for s in by_id.values():
        s.labels()
    lap("labels")
    label_rounds = exchange(lambda s: s.export_labels(), lambda s, t, b: s.import_labels(t, b),
                            lambda a, b: a + b, lambda acc: comm.allreduce_sum(acc) == 0)
    # Every OWNED pixel is resolved now, but a halo row still shows what its owner exported at the start
    # of the last round; the edges towards the halo rows need the final words: one more exchange.
    label_rounds += exchange(lambda s: s.export_labels(), lambda s, t, b: s.import_labels(t, b),
                             lambda a, b: a + b, lambda acc: True)
    lap("label exchange rounds")

    lake_counts, edges_total = None, 0
    if kind == MERGING:
        abs_, ws_, nd = [], [], 0
        for sid in sorted(by_id):
            ab, w, n = by_id[sid].edges()
            abs_.append(ab)
            ws_.append(w)
            nd += n
        nd = comm.allreduce_sum(nd)
        ab_all, w_all = comm.allgather_edges(abs_, ws_)
        edges_total = int(ab_all.shape[0]) if ab_all is not None else 0
        first = by_id[sorted(by_id)[0]]
        lake_counts = first.union(ab_all, w_all, int(bases[-1]), nd, max_water_level)[: max_water_level + 1].astype(np.int64)
        # forest edges that were contracted inside a strip's tiles: counted, not unioned
        fin = np.zeros(256, np.int64)
        for sid in by_id:
            fin += getattr(by_id[sid], "fin_hist", np.zeros(256, np.int64))
        fin = comm.allreduce_sum_vec(fin)
        lake_counts = (lake_counts - np.cumsum(fin)[: max_water_level + 1]).astype(np.uint64)
        edges_total += int(fin.sum())
        lap("forest edges, gather, union")
    return StripResult(lake_counts, flood_rounds, label_rounds, int(bases[-1]), edges_total, phase_s)
