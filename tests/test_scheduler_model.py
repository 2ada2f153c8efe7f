"""The design claim behind the asynchronous flood: the arrival times are the unique
fixed point of the relaxation, so ANY tile schedule -- FIFO sweeps, strict level order, the asynchronous
level-bucketed worklist the kernel uses -- must give the same T, and that T is the reference loop's
(level, pass).  Checked here on the CPU with the scheduler model that was used to choose the kernel's policy
(scripts/sim/flood_sim.c) against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts", "sim"))

import fieldgen  # noqa: E402


@pytest.fixture(scope="module")
def sim():
    import run_sim
    return run_sim


@pytest.mark.parametrize("name", ["smooth", "uniform", "plateaus"])
def test_every_schedule_reaches_the_reference_arrival_times(sim, oracle, name):
    img = {"smooth": lambda: fieldgen.smooth(160, 224, 5.0, 11), "uniform": lambda: fieldgen.uniform(130, 200, 12),
           "plateaus": lambda: fieldgen.plateaus(128, 192, 5, 3.0, 13)}[name]()
    seeds = sim.maxima(img)
    ref = oracle.transform(oracle.SEGMENTING, img, seeds.astype(np.uint64))
    exp = (ref.lvl.astype(np.uint32) << 24) | ref.hop
    exp[ref.lvl == 255] = 0xFF000000
    acts = {}
    for policy, delta in ((0, 0), (1, 0), (3, 2), (3, 5)):   # FIFO sweeps, level-exact, async buckets of 4 / 32 levels
        T, st, _, _ = sim.run(img, seeds, 64, 32, policy, delta, ncta=8)
        assert np.array_equal(np.minimum(T, 0xFF000000), exp), (policy, delta)
        acts[(policy, delta)] = st[1]
    # the point of the priority order: fewer tile activations than level-blind FIFO sweeps on a smooth field
    if name == "smooth":
        assert acts[(3, 2)] <= acts[(0, 0)]
