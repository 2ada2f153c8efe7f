"""Row strips over NCCL: two ranks (one process per GPU, torch.multiprocessing.spawn), halo rows by NCCL
send/recv on the library's stream, strip packets by all-gather -- labels, levels and lakes per level must equal
the single-GPU transform of the whole field bit for bit.  Skipped on a box with one GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, field: str, rows: int, cols: int, out_dir: str):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import importlib
    import torch.distributed as dist
    import fieldgen
    from wsb200_loader import load
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    ws = load()
    st = importlib.import_module("rustronomy_watershed_b200.strips")
    img = fieldgen.uniform(rows, cols, 5) if field == "uniform" else fieldgen.smooth(rows, cols, 6.0, 5)
    ctx = ws.Context(rank)
    parts = st.partition_rows(rows, world)
    g = st.StripGeometry(rank, world, rows, parts[rank])
    lo, hi = g.local_rows
    strip = st.CudaStrip(ws, ctx, g, torch.from_numpy(img[lo:hi].copy()).cuda())
    comm = st.DistComm()
    comm.set_device(torch.device("cuda", rank))
    res = None
    for _ in range(2):                      # twice: the second run reuses every buffer
        res = st.solve([strip], comm, st.MERGING, 254)
    np.save(os.path.join(out_dir, f"lab{rank}.npy"), strip.owned_labels())
    np.save(os.path.join(out_dir, f"lvl{rank}.npy"), strip.owned_levels())
    if rank == 0:
        np.save(os.path.join(out_dir, "lakes.npy"), res.lake_counts)
        np.save(os.path.join(out_dir, "rounds.npy"), np.array([res.flood_rounds, res.label_rounds]))
    strip.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("field", ["uniform", "smooth"])
def test_two_ranks_nccl_match_single_gpu(tmp_path, field):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from wsb200_loader import load
    import fieldgen
    rows, cols = 1500, 2048
    mp.spawn(_worker, args=(2, _free_port(), field, rows, cols, str(tmp_path)), nprocs=2, join=True)
    ws = load()
    img = fieldgen.uniform(rows, cols, 5) if field == "uniform" else fieldgen.smooth(rows, cols, 6.0, 5)
    t = ws.TransformBuilder.default().build_merging()
    seeds = t.find_local_minima(img)
    lab, lvl = t.transform_compact(img, seeds)
    lakes, _ = t.lake_counts(img, seeds)
    got_lab = np.concatenate([np.load(tmp_path / f"lab{r}.npy") for r in range(2)])
    got_lvl = np.concatenate([np.load(tmp_path / f"lvl{r}.npy") for r in range(2)])
    assert np.array_equal(got_lvl, lvl)
    assert np.array_equal(got_lab, lab)
    assert np.array_equal(np.load(tmp_path / "lakes.npy"), lakes)
