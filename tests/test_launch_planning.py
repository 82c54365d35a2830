"""Host-side launch planning of the library (no GPU): the CTA plan of eval_kernel and the track
split of pp_scan_kernel, through the C-ABI hooks f1l_debug_eval_plan / f1l_debug_pp_parts."""
import ctypes as C
import math

import pytest

from f1tenth_planning_b200 import _lib

SM = 148   # B200


def eval_plan(n_cand, S, M, generator=0, sm=SM):
    out = (C.c_int32 * 4)()
    rc = _lib.lib().f1l_debug_eval_plan(n_cand, S, M, sm, generator, out, 4)
    assert rc == 0, rc
    return dict(nw=out[0], chunk=out[1], ctas_per_scenario=out[2], item=out[3])


def test_eval_plan_of_the_baseline_configs():
    # C4: batches of the default 4x7 goal grid -> one 4-warp CTA per scenario whose warps pull the
    # seven items of four candidates (shared Newton solve); measured faster than one 7-warp CTA
    p = eval_plan(28, 100000, 100)
    assert p == dict(nw=4, chunk=28, ctas_per_scenario=1, item=4), p
    # ... also for the chunks of the host-buffer pipeline (8192 scenarios): the throughput-regime
    # rule needs eight waves of the four resident 4-warp CTAs per SM, 4736 scenarios
    assert eval_plan(28, 8192, 100) == p
    assert eval_plan(28, 4736, 100) == p
    assert eval_plan(28, 4735, 100)["nw"] != 4 or eval_plan(28, 4735, 100)["ctas_per_scenario"] != 1
    # the G1-clothoid generator has no shared solve: the wave model decides, never items of four
    assert eval_plan(28, 100000, 100, generator=1)["item"] == 1
    # ... a few scenarios (less than a wave) are a latency problem: one candidate per warp
    p = eval_plan(28, 64, 100)
    assert p == dict(nw=7, chunk=7, ctas_per_scenario=4, item=1), p
    # C1: a single query is a latency problem -> one candidate per warp, four CTAs
    p = eval_plan(28, 1, 100)
    assert p == dict(nw=7, chunk=7, ctas_per_scenario=4, item=1), p
    # C3: 4096 candidates of one query -> about one wave of 7-warp CTAs (4 resident per SM)
    p = eval_plan(4096, 1, 100)
    assert p["nw"] == 7 and p["ctas_per_scenario"] <= 4 * SM
    assert p["ctas_per_scenario"] > 3 * SM
    # C5: M = 200 runs the 8-warp shapes only (the 72-register builds of those spill)
    p = eval_plan(65536, 1, 200)
    assert p["nw"] == 8


@pytest.mark.parametrize("n_cand", [1, 3, 7, 27, 28, 29, 64, 100, 1000, 4096, 65536])
@pytest.mark.parametrize("S", [1, 5, 1000])
@pytest.mark.parametrize("M", [16, 100, 200])
def test_eval_plan_invariants(n_cand, S, M):
    for gen in (0, 1):
        p = eval_plan(n_cand, S, M, gen)
        assert p["nw"] in (4, 7, 8)
        assert p["chunk"] % p["nw"] == 0 and p["chunk"] >= p["nw"]
        # the chunks cover every candidate, and no CTA is empty
        assert p["ctas_per_scenario"] * p["chunk"] >= n_cand
        assert (p["ctas_per_scenario"] - 1) * p["chunk"] < n_cand
        # the shared Newton solve needs the cubic generator and four candidates per warp
        assert p["item"] == (4 if gen == 0 and p["chunk"] >= 4 * p["nw"] else 1)


def test_eval_plan_rejects_bad_arguments():
    out = (C.c_int32 * 4)()
    L = _lib.lib()
    assert L.f1l_debug_eval_plan(0, 1, 100, SM, 0, out, 4) < 0
    assert L.f1l_debug_eval_plan(28, 1, 1, SM, 0, out, 4) < 0
    assert L.f1l_debug_eval_plan(28, 1, 100, SM, 0, out, 3) < 0
    assert L.f1l_debug_pp_parts(0, 2000, 3552) < 0
    assert L.f1l_debug_pp_parts(10, 1, 3552) < 0


def wave_efficiency(n_poses, n_wpts, slots, parts):
    tasks = math.ceil(n_poses / 128) * parts
    return tasks / (math.ceil(tasks / slots) * slots) if tasks > slots else 1.0


@pytest.mark.parametrize("n_poses", [1, 100, 4096, 100000, 10**6])
@pytest.mark.parametrize("n_wpts", [2, 3, 33, 200, 2000, 20000])
@pytest.mark.parametrize("slots", [24 * SM, 25 * SM, 16 * SM])
def test_pp_parts_invariants(n_poses, n_wpts, slots):
    parts = _lib.lib().f1l_debug_pp_parts(n_poses, n_wpts, slots)
    nblk = (n_wpts - 1 + 31) // 32
    assert 1 <= parts <= max(1, min(24, nblk))      # no task without a 32-segment block
    if nblk >= 6:
        assert parts >= 6
    if nblk >= 24 and math.ceil(n_poses / 128) * 6 > slots:
        # more than a wave of work: whole waves, never worse than the fixed 8-part split
        assert wave_efficiency(n_poses, n_wpts, slots, parts) >= 0.9
        assert wave_efficiency(n_poses, n_wpts, slots, parts) >= \
            wave_efficiency(n_poses, n_wpts, slots, 8) - 0.05


def test_pp_parts_of_baseline_config_2():
    # 10^5 poses on a 2000-waypoint track: 782 groups; the split fills whole waves
    for per_sm in (24, 25):
        parts = _lib.lib().f1l_debug_pp_parts(100000, 2000, per_sm * SM)
        assert wave_efficiency(100000, 2000, per_sm * SM, parts) > 0.93, parts
    # a small batch spreads over more parts (more SMs busy) than a large one needs
    assert _lib.lib().f1l_debug_pp_parts(256, 2000, 24 * SM) >= 16
