"""Closed-loop rollout (reference lattice_planner.py:204-214 and the loop of
examples/control/pure_pursuit.py:47-57): a kinematic bicycle driven by LatticePlanner.plan on the
Spielberg raceline + map from the example's start pose, the previous path carried from call to
call, the float64 oracle stepping in lockstep on the same poses."""
import os

import numpy as np
import pytest

from f1tenth_planning_b200 import LatticePlanner, PurePursuitPlanner
from oracle import c_oracle as co
from tests import helpers as H

pytestmark = pytest.mark.gpu

N_STEPS = 400
DT = 0.01          # the gym's integration step
WHEELBASE = 0.33


def _spielberg(golden_spielberg):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "maps.npz"))
    w = int(g["spielberg_shape"][1])
    occ = np.unpackbits(g["spielberg_bits"], axis=1)[:, :w]
    return golden_spielberg["waypoints"], (occ, tuple(g["spielberg_origin"]), float(g["spielberg_res"]))


@pytest.mark.parametrize("start", ["example_start_pose", "hairpin"])
def test_closed_loop_rollout_matches_oracle(golden_spielberg, start):
    wp, grid = _spielberg(golden_spielberg)
    la, wd = np.linspace(0.8, 3.0, 8), np.linspace(-0.6, 0.6, 9)
    pl = LatticePlanner(wheelbase=WHEELBASE, waypoints=wp)
    pl.set_map(*grid)
    pl.set_goal_grid(la, wd)
    cfg = H.oracle_config_from_engine(pl.engine)
    world = co.World_(wp, la, wd, grid=grid[0], grid_origin=grid[1], grid_res=grid[2])
    tracker = PurePursuitPlanner(wheelbase=WHEELBASE)

    if start == "example_start_pose":
        x, y, th, v = 0.0, -0.84, 3.40, 0.0      # examples/control/pure_pursuit.py:47
        v_cap, dev_max = 100.0, 0.5
    else:   # 12 m before the tightest corner of the raceline (kappa = 0.45 rad/m), 5 m/s
        i0 = int(np.argmax(np.abs(wp[:, 4]))) - 60
        x, y, th, v = float(wp[i0, 0]), float(wp[i0, 1]), float(wp[i0, 3]), 5.0
        v_cap, dev_max = 5.0, 0.8
    counts = {"steps": 0, "same_idx": 0, "idx_diff:cost_gap<1e-5": 0, "idx_diff:margin_flag": 0,
              "steer_err_max": 0.0, "map_tracker_steer_err_max": 0.0, "infeasible_steps": 0}
    flag_counts = {}
    path, seen_idx = [], set()
    for k in range(N_STEPS):
        d = pl.plan_detailed(x, y, th, v, want_states=False, want_map=True)
        o = co.plan(cfg, world, np.array([x, y, th, v]), None, want_states=True)
        H.classify_flags(d.flags, o, cfg, counts=flag_counts)
        same = (d.flags & 0xF) == (o["flags"] & 0xF)
        fin = np.isfinite(d.costs) & np.isfinite(o["costs"])
        assert np.array_equal(np.isfinite(d.costs)[same], np.isfinite(o["costs"])[same])
        assert H.close(d.costs[fin], o["costs"][fin]).all(), k
        assert d.no_feasible == o["no_feasible"] or not same.all()
        counts["steps"] += 1
        seen_idx.add(int(d.best_idx))
        counts["infeasible_steps"] += int(d.no_feasible)
        if d.best_idx == o["best_idx"]:
            counts["same_idx"] += 1
            if not d.no_feasible:
                err = abs(d.steer - o["steer"])
                counts["steer_err_max"] = max(counts["steer_err_max"], err)
                assert err < 1e-4 + 1e-4 * abs(o["steer"]), (k, d.steer, o["steer"])
                assert d.speed == o["speed"]      # the raceline speed itself (float64)
                sc = H.traj_scale(o["best_traj"])[0]
                assert (np.abs(d.best_traj_map[:, :2] - o["best_traj_map"][:, :2])
                        <= H.ABS + H.REL * max(sc[0], sc[1])).all(), k
        else:
            gap = abs(float(o["costs"][d.best_idx]) - float(o["costs"][o["best_idx"]]))
            if gap < 1e-5:
                counts["idx_diff:cost_gap<1e-5"] += 1
            else:
                assert not same[[d.best_idx, o["best_idx"]]].all(), ("argmin differs", k, gap)
                counts["idx_diff:margin_flag"] += 1
        if k % 10 == 0 and not d.no_feasible and d.tracker_found:
            # B.8: the map-frame [X, Y, v, Theta] trajectory is what a map-frame tracker consumes --
            # pure pursuit from the MAP pose on it gives the steer the planner's own
            # (vehicle-frame) tracker produced
            st2, sp2 = tracker.plan(x, y, th, pl.engine.config.tracker_lookahead,
                                    waypoints=np.ascontiguousarray(d.best_traj_map[:, :3]))
            counts["map_tracker_steer_err_max"] = max(counts["map_tracker_steer_err_max"], abs(st2 - d.steer))
            assert abs(st2 - d.steer) < 1e-5 and sp2 == d.speed, (k, st2, d.steer)
        # teacher-forced carry: the oracle's next similarity term uses the device's previous path
        world.set_prev(d.best_traj[:, 2])
        # kinematic bicycle (single-track, rear axle), speed first-order towards the command
        steer = float(np.clip(d.steer, -0.4189, 0.4189))
        v += float(np.clip(min(d.speed, v_cap) - v, -9.51 * DT, 9.51 * DT))
        x += v * np.cos(th) * DT
        y += v * np.sin(th) * DT
        th += v / WHEELBASE * np.tan(steer) * DT
        path.append((x, y))
    path = np.array(path)
    # the car went somewhere and stayed on the raceline (within half a track width)
    assert np.hypot(*(path[-1] - path[0])) > 5.0
    dev = [co.nearest_point(p, wp[:, :2])[1] for p in path[::20]]
    assert max(dev) < dev_max, max(dev)
    counts.update({"flag:" + k_: v_ for k_, v_ in flag_counts.items()})
    counts["raceline_deviation_max_m"] = float(max(dev))
    counts["distinct_best_idx"] = len(seen_idx)
    H.record_parity("closed_loop_spielberg_400_steps_" + start, counts)
    assert counts["same_idx"] + counts["idx_diff:cost_gap<1e-5"] + counts["idx_diff:margin_flag"] == N_STEPS
    assert counts["same_idx"] >= 0.97 * N_STEPS, counts
    assert start != "hairpin" or len(seen_idx) >= 3
    assert counts["infeasible_steps"] == 0, counts
