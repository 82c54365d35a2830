"""Randomised configuration sweep of the lattice pipeline against the oracle (through the C-ABI):
sample counts that hit every compiled kernel shape, windows, goal grids, vehicle sizes, weights,
end-curvature and tracker modes, tracks of different scale, grids with different resolution."""
import numpy as np
import pytest

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([300, 801, 2000]))
    a = float(rng.uniform(15, 80))
    b = float(rng.uniform(8, 0.6 * a))
    track = synth.ellipse_track(n=n, a=a, b=b, speed=float(rng.uniform(2, 9)))
    if rng.random() < 0.5:
        track = track[::-1].copy()                      # clockwise track
        track[:, 3] = np.mod(track[:, 3] + np.pi, 2 * np.pi)
    m = int(rng.choice([24, 50, 100, 104, 137, 200, 256]))
    cfg = dict(
        n_samples=m, n_newton=int(rng.choice([6, 8, 12])),
        window=int(rng.choice([0, 33, 64, 128, 300])),
        n_shift=int(rng.integers(0, 4)), n_cull=int(rng.integers(0, 6)),
        kappa_max=float(rng.choice([0.0, 1.35, 3.0])),
        car_length=float(rng.uniform(0.4, 0.7)), car_width=float(rng.uniform(0.2, 0.4)),
        use_goal_kappa=int(rng.integers(0, 2)), literal_tracker=int(rng.integers(0, 2)),
        tracker_lookahead=float(rng.uniform(0.4, 1.2)),
        prune_window=int(rng.integers(0, 2)),
    )
    w = rng.uniform(0.05, 1.0, 5)
    cfg["weights"] = list(w / w.sum())
    nl, nw = int(rng.integers(2, 10)), int(rng.integers(3, 12))
    la = np.sort(rng.uniform(0.5, 3.8, nl))
    wd = np.linspace(-rng.uniform(0.4, 1.3), rng.uniform(0.4, 1.3), nw)
    grid = None
    if rng.random() < 0.7:
        grid = synth.corridor_grid(a=a, b=b, half_width=float(rng.uniform(0.8, 1.6)),
                                   res=float(rng.choice([0.04, 0.05, 0.08])), margin=3.0)
    k = int(rng.integers(0, 7))
    return track, la, wd, grid, cfg, k, rng


@pytest.mark.parametrize("seed", list(range(100, 116)))
def test_random_configuration(seed):
    track, la, wd, grid, cfg, k, rng = _case(seed)
    eng, ocfg, world = H.make_pair(track, la, wd, grid=grid, **cfg)
    for q in range(2):
        poses, opp, n_opp = synth.scenario_batch(track, 1, max(k, 1), int(rng.integers(1 << 30)))
        o_in = opp[0, :k] if k else None
        prev = None
        if q == 1:   # second query sees the first one's best trajectory as prev_path
            prev = d.best_traj[:, 2].copy()
            eng.set_prev_path(prev)
            world.set_prev(prev)
        d = eng.plan(poses[0], o_in, update_prev=False, want_states=True)
        o = co.plan(ocfg, world, poses[0], o_in, want_states=True)
        H.compare_plan(d, o, ocfg)
        # the same query cut into W row-interleaved shards (f1l_plan_rows) and into W contiguous
        # blocks (f1l_plan_shard): per-candidate results at the same global indices, bit for bit
        W = int(rng.integers(2, 5))
        nL, nW = len(la), len(wd)
        for mode in ("rows", "blocks"):
            seen = np.zeros(nL * nW, bool)
            best = []
            for r in range(min(W, nL)):
                if mode == "rows":
                    p = eng.plan(poses[0], o_in, update_prev=False, rows=(r, min(W, nL)))
                    mine = np.zeros((nL, nW), bool)
                    mine[r::min(W, nL)] = True
                    mine = mine.ravel()
                else:
                    from f1tenth_planning_b200 import sharding
                    lo, hi = sharding.block(nL * nW, r, min(W, nL))
                    p = eng.plan(poses[0], o_in, update_prev=False, shard=(lo, hi))
                    mine = np.zeros(nL * nW, bool)
                    mine[lo:hi] = True
                assert np.array_equal(p.costs[mine], d.costs[mine]), (mode, r)
                assert np.array_equal(p.flags[mine], d.flags[mine]), (mode, r)
                assert np.isinf(p.costs[~mine]).all()
                best.append((float(p.best_cost), int(p.best_idx)))
                seen |= mine
            assert seen.all()
            if np.isfinite(d.costs).any():
                assert min(best)[1] == d.best_idx, (mode, best, d.best_idx)
