"""BASELINE.json's full sizes through size-independent properties plus oracle samples:
config 4 (10^5 scenarios x 28 candidates) and config 5 (65 536 candidates x 200 samples)."""
import numpy as np
import pytest

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_config4_hundred_thousand_scenarios(ellipse, corridor):
    la, wd = synth.goal_grid(4)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, kappa_max=0.0)
    S, K = 100000, 8
    poses, opp, n_opp = synth.scenario_batch(ellipse, S, K, 1004)
    b = eng.plan_batch(poses, opp, n_opp, want_flags=True)
    # argmin semantics at full size: first minimum of the cost vector, +inf rows -> index 0
    am = np.argmin(b.costs, axis=1)
    assert np.array_equal(b.best_idx, am.astype(np.int32))
    assert np.array_equal(b.best_cost, b.costs[np.arange(S), am])
    valid = (b.flags & 1) != 0
    hit = (b.flags & 6) != 0
    assert np.array_equal(np.isfinite(b.costs), valid & ~hit)     # cost finite <=> valid and free
    assert 0.85 < valid.mean() <= 1.0 and 0.02 < hit[valid].mean() < 0.5
    feasible = np.isfinite(b.best_cost)
    assert np.isfinite(b.best_traj[feasible]).all()
    # trajectories start at the vehicle origin and end at the selected goal
    assert np.abs(b.best_traj[feasible][:, 0, :3]).max() == 0.0
    # scenario sharding: any block evaluated alone gives the same answers (weak-scaling shards)
    for lo, hi in ((0, 12500), (50000, 62500), (87500, 100000)):
        part = eng.plan_batch(poses[lo:hi], opp[lo:hi], n_opp[lo:hi], want_flags=True)
        assert np.array_equal(part.best_idx, b.best_idx[lo:hi])
        assert np.array_equal(part.costs, b.costs[lo:hi])
        assert np.array_equal(part.best_traj, b.best_traj[lo:hi])
    # idempotence
    b2 = eng.plan_batch(poses, opp, n_opp, want_flags=True)
    assert np.array_equal(b2.costs, b.costs) and np.array_equal(b2.best_traj, b.best_traj)
    # oracle on a 256-scenario sample
    sub = np.random.default_rng(0).choice(S, 256, replace=False)
    o = co.plan_batch(cfg, world, poses[sub], opp[sub], n_opp[sub], n_threads=co.max_threads())
    fin = np.isfinite(o["costs"]) & np.isfinite(b.costs[sub])
    assert (np.isfinite(o["costs"]) != np.isfinite(b.costs[sub])).mean() < 0.005
    assert H.close(b.costs[sub][fin], o["costs"][fin]).all()
    agree = b.best_idx[sub] == o["best_idx"]
    for k in np.nonzero(~agree)[0]:
        assert abs(float(o["costs"][k, b.best_idx[sub][k]]) - float(o["best_cost"][k])) < 1e-5
    assert agree.mean() > 0.98


def test_config5_dense_sweep(ellipse, corridor):
    la, wd = synth.goal_grid(5)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, n_samples=200)
    pose, opp = H.scenario(ellipse, 1005, 8)
    d = eng.plan(pose, opp, update_prev=False)
    C = 65536
    assert d.costs.shape == (C,)
    assert d.best_idx == int(np.argmin(d.costs)) and d.best_cost == d.costs[d.best_idx]
    valid = (d.flags & 1) != 0
    assert np.array_equal(np.isfinite(d.costs), valid & ((d.flags & 6) == 0))
    assert valid.sum() > 20000
    # neighbouring goals give neighbouring costs (the cost field is smooth where it is finite)
    grid = d.costs.reshape(256, 256)
    both = np.isfinite(grid[:, 1:]) & np.isfinite(grid[:, :-1])
    assert np.median(np.abs(np.diff(grid, axis=1))[both]) < 5e-3
    # candidate shards reassemble to the full sweep
    best = (np.inf, C)
    for g in range(8):
        lo, hi = g * C // 8, (g + 1) * C // 8
        part = eng.plan(pose, opp, update_prev=False, shard=(lo, hi))
        assert np.array_equal(part.costs[lo:hi], d.costs[lo:hi])
        best = min(best, (float(part.best_cost), int(part.best_idx)))
    assert best[1] == d.best_idx
    # oracle on two slices of the sweep
    for lo, hi in ((30000, 30512), (60000, 60256)):
        o = co.plan(cfg, world, pose, opp, c_begin=lo, c_end=hi, want_states=False)
        fin = np.isfinite(o["costs"][lo:hi]) & np.isfinite(d.costs[lo:hi])
        assert (np.isfinite(o["costs"][lo:hi]) != np.isfinite(d.costs[lo:hi])).mean() < 0.01
        assert H.close(d.costs[lo:hi][fin], o["costs"][lo:hi][fin]).all()


def test_config2_full_size_properties(ellipse):
    """10^5 poses on the 2k-waypoint track: permutation invariance and block consistency"""
    from f1tenth_planning_b200.engine import Engine
    eng = Engine()
    eng.set_track(ellipse)
    poses, _ = synth.random_poses(ellipse, 100000, np.random.default_rng(1002))
    poses = np.ascontiguousarray(poses[:, :3])
    r = eng.pure_pursuit_batch(poses, 0.8)
    perm = np.random.default_rng(1).permutation(100000)
    rp = eng.pure_pursuit_batch(poses[perm], 0.8)
    assert np.array_equal(rp.nearest_i, r.nearest_i[perm])
    assert np.array_equal(rp.actuation, r.actuation[perm])
    part = eng.pure_pursuit_batch(poses[40000:40100], 0.8)
    assert np.array_equal(part.nearest, r.nearest[40000:40100])
    assert (r.status > 0).all()
