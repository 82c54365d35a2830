"""The reference-facing class on the GPU against the oracle: LatticePlanner.plan / plan_detailed /
plan_batch, the plug-in hooks (add_sample_function, add_cost_function + cost_weights,
add_selection_function) and the module-level sample_lookahead_square
(reference lattice_planner.py:57-128, 174-214, 223-260), plus regression tests for the round-1
advisor findings (stale CUDA graph after a buffer grows, in-place raceline edits, the plug-in
tracker's frame)."""
import numpy as np
import pytest

from f1tenth_planning_b200 import LatticePlanner, synth
from f1tenth_planning_b200 import lattice_planner as lp
from oracle import c_oracle as co
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _planner(track, grid, la=None, wd=None, **cfg):
    pl = LatticePlanner(waypoints=track, **cfg)
    if grid is not None:
        pl.set_map(*grid)
    if la is not None:
        pl.set_goal_grid(la, wd)
    return pl


def _oracle(pl, track, grid, la, wd):
    cfg = H.oracle_config_from_engine(pl.engine)
    kw = dict(grid=grid[0], grid_origin=grid[1], grid_res=grid[2]) if grid is not None else {}
    return cfg, co.World_(track, la, wd, **kw)      # the oracle's own float64 LUT


def test_plan_matches_oracle_default_grid(ellipse, corridor):
    """C1 through the class: plan() 3-tuple, .last detail, previous path carried across calls."""
    la, wd = synth.goal_grid(1)
    pl = _planner(ellipse, corridor, window=0, kappa_max=0.0)
    cfg, world = _oracle(pl, ellipse, corridor, la, wd)
    poses, opp, n_opp = synth.scenario_batch(ellipse, 6, 1, 1001)
    for s in range(6):
        steer, speed, traj = pl.plan(*poses[s], opponent_poses=opp[s, :n_opp[s]])
        d = pl.last
        o = co.plan(cfg, world, poses[s], opp[s, :n_opp[s]], want_states=True)
        H.compare_plan(d, o, cfg, record="planner_c1_plan_%d" % s)
        # the 3-tuple of lattice_planner.py:214: (steer, speed, [M,4] x, y, theta, |kappa|)
        assert steer == d.steer and speed == d.speed
        assert traj.shape == (100, 4) and traj.dtype == np.float64
        assert np.array_equal(traj[:, :3], d.best_traj[:, :3].astype(np.float64))
        assert np.array_equal(traj[:, 3], np.abs(d.best_traj[:, 3].astype(np.float64)))
        world.set_prev(d.best_traj[:, 2])     # teacher-forced previous path for the next call


def test_plan_detailed_and_batch_match_oracle(ellipse, corridor):
    la, wd = np.linspace(0.6, 3.2, 10), np.linspace(-1.0, 1.0, 11)
    pl = _planner(ellipse, corridor, la, wd, kappa_max=2.5)
    cfg, world = _oracle(pl, ellipse, corridor, la, wd)
    poses, opp, n_opp = synth.scenario_batch(ellipse, 64, 6, 31)
    for s in range(3):
        d = pl.plan_detailed(*poses[s], opponent_poses=opp[s, :n_opp[s]], want_states=True, want_map=True)
        pl.engine.set_prev_path(None)
        o = co.plan(cfg, world, poses[s], opp[s, :n_opp[s]], want_states=True)
        H.compare_plan(d, o, cfg, record="planner_plan_detailed_%d" % s)
        assert d.best_traj_map.shape == (100, 4)
    b = pl.plan_batch(poses, opp, n_opp, want_flags=True)
    counts = H.compare_batch(b, np.arange(64), poses, opp, n_opp, cfg, world)
    H.record_parity("planner_plan_batch", counts)
    ob = co.plan_batch(cfg, world, poses, opp, n_opp, n_threads=co.max_threads())
    ok = (b.best_idx == ob["best_idx"]) & np.isfinite(ob["best_cost"])
    assert ok.sum() > 30
    assert H.close(b.best_traj[ok], ob["best_traj"][ok], scale=H.traj_scale(ob["best_traj"][ok])).all()
    assert H.close(b.steer_speed[ok], ob["steer_speed"][ok], 1e-4, 1e-4).all()


def test_sample_lookahead_square_and_sample_plugin(ellipse, corridor):
    """module-level sampler (reference :223-260, repaired per DESIGN.md) == the oracle's B.1 goals;
    registered as the sample_func it reproduces the built-in sampler's plan."""
    la, wd = [0.4, 0.6, 0.8, 1.0], np.linspace(-1.0, 1.0, 7)
    pl = _planner(ellipse, corridor, la, wd, window=0, kappa_max=0.0)
    cfg, world = _oracle(pl, ellipse, corridor, la, wd)
    poses, opp, n_opp = synth.scenario_batch(ellipse, 4, 2, 77)
    for s in range(4):
        goals = lp.sample_lookahead_square(*poses[s], ellipse)      # defaults = la, wd
        o = co.plan(cfg, world, poses[s], opp[s, :n_opp[s]], want_states=True)
        assert goals.shape == (28, 3) and goals.dtype == np.float64
        assert H.close(goals, o["goals"], 1e-5, 1e-5).all()
    seen = []

    def sampler(px, py, pth, v, wpts):
        seen.append((px, py, pth, v))
        return lp.sample_lookahead_square(px, py, pth, v, wpts, la, wd)
    pl2 = _planner(ellipse, corridor, window=0, kappa_max=0.0)
    pl2.add_sample_function(sampler)
    for s in range(4):
        d_builtin = pl.plan_detailed(*poses[s], opponent_poses=opp[s, :n_opp[s]], want_states=True)
        d_plugin = pl2.plan_detailed(*poses[s], opponent_poses=opp[s, :n_opp[s]], want_states=True)
        pl.engine.set_prev_path(None)
        pl2.engine.set_prev_path(None)
        assert seen[-1] == tuple(poses[s])
        # explicit goals are float32 copies of the sampler's float32 goals: the same spirals
        assert d_plugin.best_idx == d_builtin.best_idx
        assert H.close(d_plugin.costs[np.isfinite(d_builtin.costs)],
                       d_builtin.costs[np.isfinite(d_builtin.costs)]).all()
        assert np.array_equal(np.isfinite(d_plugin.costs), np.isfinite(d_builtin.costs))
        o = co.plan(cfg, world, poses[s], opp[s, :n_opp[s]],
                    goals=lp.sample_lookahead_square(*poses[s], ellipse, la, wd), want_states=True)
        H.compare_plan(d_plugin, o, cfg)


def test_custom_selection_equals_builtin(ellipse, corridor):
    """A registered selection_func that is argmin in disguise must give the default plan: same
    index, trajectory, steer and speed (the chosen candidate is tracked on the device like the
    built-in winner -- vehicle frame, raceline speed, configured wheelbase)."""
    la, wd = np.linspace(0.6, 3.0, 8), np.linspace(-1.0, 1.0, 9)
    poses, opp, n_opp = synth.scenario_batch(ellipse, 5, 4, 5)
    for literal in (0, 1):
        ref = _planner(ellipse, corridor, la, wd, kappa_max=0.0, literal_tracker=literal, wheelbase=0.3)
        cus = _planner(ellipse, corridor, la, wd, kappa_max=0.0, literal_tracker=literal, wheelbase=0.3)
        cus.add_selection_function(lambda costs: int(np.argmin(np.asarray(costs))))
        for s in range(5):
            a = ref.plan_detailed(*poses[s], opponent_poses=opp[s, :n_opp[s]], want_map=True)
            b = cus.plan_detailed(*poses[s], opponent_poses=opp[s, :n_opp[s]], want_map=True)
            assert b.best_idx == a.best_idx and np.float32(b.best_cost) == np.float32(a.best_cost)
            assert b.steer == a.steer and b.speed == a.speed
            assert b.tracker_found == a.tracker_found and b.no_feasible == a.no_feasible
            assert np.array_equal(b.best_traj, a.best_traj)
            assert np.array_equal(b.best_traj_map, a.best_traj_map)
            # previous path carried identically: the next call's similarity terms agree (checked by
            # the equality of the next iteration's costs)
            assert np.array_equal(b.costs, a.costs)
    # a selection that picks something else is tracked as that candidate
    far = _planner(ellipse, corridor, la, wd, kappa_max=0.0)
    far.add_selection_function(lambda costs: int(np.nonzero(np.isfinite(costs))[0][-1]))
    d = far.plan_detailed(*poses[0], opponent_poses=opp[0, :n_opp[0]], want_states=True)
    last = int(np.nonzero(np.isfinite(d.costs))[0][-1])
    assert d.best_idx == last and np.array_equal(d.best_traj, d.states[last])
    assert d.best_cost == pytest.approx(float(d.costs[last]))


def test_custom_cost_functions_and_weights(ellipse, corridor):
    """registered cost_funcs + cost_weights (reference eval, :130-156) over the GPU trajectories,
    against the same user functions over the oracle's trajectories."""
    la, wd = np.linspace(0.8, 3.0, 6), np.linspace(-0.9, 0.9, 7)
    wide = synth.corridor_grid(half_width=2.5)       # walls beyond the outer goals: most candidates free
    pl = _planner(ellipse, wide, la, wd, kappa_max=0.0)
    cfg, world = _oracle(pl, ellipse, wide, la, wd)

    def end_offset(traj):        # user cost 1: lateral offset of the end point
        return abs(traj[-1, 1])

    def bending(traj):           # user cost 2: integral of |kappa| (column 3 is unsigned)
        assert (traj[:, 3] >= 0).all()
        return float(np.sum(traj[:, 3]))
    pl.add_cost_function([end_offset, bending])
    pl.cost_weights = [0.25, 0.75]
    poses, opp, n_opp = synth.scenario_batch(ellipse, 5, 4, 9)
    n_both = 0
    for s in range(5):
        steer, speed, traj = pl.plan(*poses[s], opponent_poses=opp[s, :n_opp[s]])
        d = pl.last
        o = co.plan(cfg, world, poses[s], opp[s, :n_opp[s]], want_states=True)
        ost = o["states"].copy()
        ost[:, :, 3] = np.abs(ost[:, :, 3])
        ocost = np.array([0.25 * end_offset(t) + 0.75 * bending(t) for t in ost])
        ocost[~np.isfinite(o["costs"])] = np.inf      # invalid / collided stay infeasible
        gfin, ofin = np.isfinite(d.costs), np.isfinite(ocost)
        counts = H.classify_flags(d.flags, o, cfg)
        same = (d.flags & 0xF) == (o["flags"] & 0xF)
        assert np.array_equal(gfin[same], ofin[same])
        both = gfin & ofin
        n_both += int(both.sum())
        assert H.close(d.costs[both], ocost[both]).all()
        oi = int(np.argmin(ocost))
        if d.best_idx != oi:
            assert abs(ocost[d.best_idx] - ocost[oi]) < 1e-5 or not same[[d.best_idx, oi]].all(), counts
        else:
            assert H.close(d.best_traj, o["states"][oi], scale=H.traj_scale(o["states"][oi])).all()
            # tracker on the chosen trajectory: vehicle frame, raceline speed at the goal centre
            wp = np.column_stack([o["states"][oi][:, 0], o["states"][oi][:, 1],
                                  np.full(100, speed)])
            ot = co.pure_pursuit_batch(wp, np.zeros((1, 3)), cfg.tracker_lookahead,
                                       wheelbase=cfg.wheelbase, max_reacquire=cfg.max_reacquire)
            assert abs(steer - ot["actuation"][0, 0]) < 1e-4 + 1e-4 * abs(ot["actuation"][0, 0])
            assert speed > 0 and d.tracker_found
    assert n_both >= 40, n_both
    with pytest.raises(ValueError):       # reference :145-146
        pl.cost_weights = [0.5, 0.75]
        pl.plan(*poses[0])


def test_graph_survives_buffer_growth(ellipse, corridor):
    """advisor r1: a buffer that generate() / plan_goals() grow must not leave the captured
    single-query graph pointing at freed memory."""
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, kappa_max=0.0)
    pose, opp = H.scenario(ellipse, 3, 2)
    a = eng.plan(pose, opp, update_prev=False, want_states=True)
    gx, gy = np.meshgrid(np.linspace(0.8, 3.5, 40), np.linspace(-1.0, 1.0, 41), indexing="ij")
    goals = np.stack([gx.ravel(), gy.ravel(), 0.1 * gy.ravel()], axis=1)
    eng.generate(goals)                                             # grows q_states / q_goals
    eng.plan_goals(pose, goals, opp, update_prev=False, detail=False)   # grows q_detail
    b = eng.plan(pose, opp, update_prev=False, want_states=True)
    assert np.array_equal(a.states, b.states) and np.array_equal(a.costs, b.costs)
    assert a.best_idx == b.best_idx and np.array_equal(a.best_traj, b.best_traj)
    o = co.plan(cfg, world, pose, opp, want_states=True)
    H.compare_plan(b, o, cfg)


def test_in_place_raceline_edit_is_detected(ellipse):
    """advisor r1: editing rows of the SAME waypoint array -- here only rows that a stride-16
    subsample never sees -- must reach the device."""
    wp = ellipse.copy()
    pl = LatticePlanner(waypoints=wp, kappa_max=0.0)
    pose = np.array([wp[100, 0], wp[100, 1], wp[100, 3], 4.0])
    pl.plan(*pose)
    before = pl.last
    mask = np.arange(wp.shape[0]) % 16 != 0
    wp[mask, 2] *= 0.5              # speed profile of the rows in between
    wp[mask, 1] += 0.05             # and a 5 cm lateral shift
    steer1, speed1, _ = pl.plan(*pose)
    after = pl.last
    assert not np.array_equal(after.costs, before.costs)
    fresh = LatticePlanner(waypoints=wp.copy(), kappa_max=0.0)
    steer2, speed2, _ = fresh.plan(*pose)
    assert (steer1, speed1) == (steer2, speed2)
    # (the first call stored a previous path; the fresh planner has none: compare without it)
    pl.engine.set_prev_path(None)
    pl.plan(*pose)
    assert np.array_equal(pl.last.costs, fresh.last.costs)


def test_select_candidate_needs_a_current_query(ellipse):
    from f1tenth_planning_b200.engine import Engine, F1LError
    eng = Engine()
    eng.set_track(ellipse)
    eng.set_goal_grid(*synth.goal_grid(1))
    with pytest.raises(F1LError):
        eng.select_candidate(0)                       # no query yet
    pose, opp = H.scenario(ellipse, 1, 1)
    d = eng.plan(pose, opp, update_prev=False)
    t = eng.select_candidate(d.best_idx, d.best_cost, update_prev=False)
    assert t.steer == d.steer and np.array_equal(t.best_traj, d.best_traj)
    with pytest.raises(F1LError):
        eng.select_candidate(28)                      # out of range
    eng.configure(kappa_max=1.0)
    with pytest.raises(F1LError):
        eng.select_candidate(0)                       # configuration changed since the query


def test_intersect_point_start_parameter_is_validated(ellipse):
    from f1tenth_planning_b200 import utils
    xy = ellipse[:, :2]
    n = xy.shape[0]
    p = xy[10] + np.array([0.05, 0.0])
    assert utils.intersect_point(p, 0.8, xy, 10.0, True)[0] is not None
    assert utils.intersect_point(p, 0.8, xy, n - 0.5, True)[0] is not None   # last start index
    for bad in (-1.0, float(n), 2.0 * n, float("nan"), float("inf")):
        with pytest.raises(IndexError):
            utils.intersect_point(p, 0.8, xy, bad, True)
