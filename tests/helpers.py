"""Parity helpers shared by the GPU tests, __graft_entry__.smoke() and tools/: build a matching
(oracle world, CUDA engine) pair and compare one plan against the oracle with the tolerance
classes BASELINE.json's north_star states:
    - spiral states and costs within 1e-4 relative (+1e-5 absolute floor),
    - collision flags and selected index bit-exact except where the decision margin is tiny
      (cost gap < 1e-5; collision margin < 1e-4 m; validity margin on kappa_max / tolerance).
"""
import numpy as np

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co

REL = 1e-4
ABS = 1e-5

CFG_FIELDS = ["n_samples", "n_newton", "window", "n_shift", "n_cull", "literal_tracker",
              "use_goal_kappa", "generator", "kappa_max", "car_length", "car_width", "converge_tol",
              "tracker_lookahead", "wheelbase", "max_reacquire"]


def oracle_config_from_engine(eng):
    c = eng.config
    kw = {k: getattr(c, k) for k in CFG_FIELDS}
    kw["weights"] = [c.weights[i] for i in range(5)]
    return co.default_config(**kw)


def make_pair(track, lookaheads, widths, grid=None, device=None, use_device_lut=True, **config):
    """(engine, oracle_cfg, oracle_world).  With use_device_lut the oracle seeds Newton from the
    LUT the device built (teacher-forced seed); otherwise from its own float64 LUT."""
    from f1tenth_planning_b200.engine import Engine
    eng = Engine(device=device, **config)
    eng.set_track(track)
    eng.set_goal_grid(lookaheads, widths)
    kw = {}
    if grid is not None:
        occ, origin, res = grid
        eng.set_grid(occ, origin, res)
        kw = dict(grid=occ, grid_origin=origin, grid_res=res)
    if use_device_lut:
        lut, ranges = eng.get_lut()
        kw.update(lut=lut, lut_ranges=ranges)
    world = co.World_(track, lookaheads, widths, **kw)
    return eng, oracle_config_from_engine(eng), world


def close(a, b, rel=REL, abs_=ABS, scale=None):
    """|a - b| <= abs + rel * max(|b|, scale): 1e-4 relative with a 1e-5 absolute floor; `scale`
    (broadcastable) makes the relative part refer to the magnitude of the whole column of a
    trajectory rather than to a value that happens to cross zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ref = np.abs(b) if scale is None else np.maximum(np.abs(b), scale)
    return np.abs(a - b) <= abs_ + rel * ref


def traj_scale(states):
    """per-trajectory, per-column magnitude [..., 1, 4] of a [..., M, 4] state array"""
    return np.abs(np.asarray(states, dtype=np.float64)).max(axis=-2, keepdims=True)


def compare_plan(d, o, cfg, kappa_max=None, verbose=False):
    """d: engine.PlanDetail (GPU), o: oracle plan dict.  Returns a dict of statistics and raises
    AssertionError on a parity violation."""
    stats = {}
    C = o["costs"].shape[0]
    gf = d.flags.astype(np.int32)
    of = o["flags"].astype(np.int32)
    # goals
    ok_centre = (of & co.FLAG_NO_CENTRE) == 0
    assert ((gf & co.FLAG_NO_CENTRE) == (of & co.FLAG_NO_CENTRE)).all(), "centre flags differ"
    assert close(d.goals[ok_centre], o["goals"][ok_centre], 1e-5, 1e-5).all(), "goals differ"
    # validity: equal except at the decision margin
    gv, ov = (gf & 1) != 0, (of & 1) != 0
    mism = np.nonzero(gv != ov)[0]
    stats["valid_mismatch"] = int(mism.size)
    stats["n_valid"] = int(ov.sum())
    km = cfg.kappa_max if kappa_max is None else kappa_max
    for c in mism:
        # explainable only by max|kappa| ~ kappa_max or an endpoint error ~ tolerance
        maxk_o = np.abs(o["states"][c, :, 3]).max() if "states" in o else np.nan
        g = o["goals"][c]
        tol = cfg.converge_tol * max(1.0, float(np.linalg.norm(g)))
        end_err = np.abs(o["states"][c, -1, :3] - g).max() if "states" in o else np.nan
        near_k = km > 0 and abs(maxk_o - km) < 1e-3 * km
        near_t = abs(end_err - tol) < 0.5 * tol or not np.isfinite(end_err)
        assert near_k or near_t, ("validity flag differs away from the margin", c, maxk_o, end_err)
    both = gv & ov
    stats["n_both_valid"] = int(both.sum())
    if both.any():
        # spiral parameters and states
        pe = np.abs(d.params[both, :3] - o["params"][both, :3]) / (ABS / REL + np.abs(o["params"][both, :3]))
        stats["param_rel_max"] = float(pe.max())
        assert close(d.params[both, :3], o["params"][both, :3]).all(), ("params differ", pe.max())
        if d.states is not None and "states" in o:
            sc = traj_scale(o["states"][both])
            se = np.abs(d.states[both] - o["states"][both]) / (ABS / REL + np.maximum(np.abs(o["states"][both]), sc))
            stats["state_rel_max"] = float(se.max())
            assert close(d.states[both], o["states"][both], scale=sc).all(), ("states differ", se.max())
        te = np.abs(d.terms[both] - o["terms"][both]) / (ABS / REL + np.abs(o["terms"][both]))
        stats["term_rel_max"] = [float(x) for x in te.max(axis=0)]
        assert close(d.terms[both], o["terms"][both]).all(), ("cost terms differ", te.max(axis=0))
    # collision flags: bit-exact except within 1e-4 m of the decision boundary
    for bit, col, name in ((co.FLAG_COLLIDE_OPP, 0, "opp"), (co.FLAG_COLLIDE_MAP, 1, "map")):
        mm = np.nonzero(both & ((gf & bit) != (of & bit)))[0]
        stats["collide_%s_mismatch" % name] = int(mm.size)
        stats["collide_%s_count" % name] = int(((of & bit) != 0).sum())
        for c in mm:
            assert o["margins"][c, col] < 1e-4, ("collision flag differs away from the boundary",
                                                 name, c, o["margins"][c, col])
    # costs
    fin = np.isfinite(o["costs"]) & np.isfinite(d.costs)
    stats["n_finite"] = int(fin.sum())
    if fin.any():
        assert close(d.costs[fin], o["costs"][fin]).all(), "total costs differ"
    # argmin: same index, or the cost gap is below 1e-5
    stats["best_idx"] = (int(d.best_idx), int(o["best_idx"]))
    if d.best_idx != o["best_idx"]:
        if np.isfinite(o["best_cost"]):
            gap = abs(float(d.costs[d.best_idx]) - float(o["costs"][o["best_idx"]]))
            cross = abs(float(o["costs"][d.best_idx]) - float(o["costs"][o["best_idx"]]))
            assert gap < 1e-5 or cross < 1e-5, ("argmin differs with a cost gap", gap, cross)
        else:
            raise AssertionError("argmin differs on an infeasible query")
    else:
        # best trajectory and tracker output
        if np.isfinite(o["best_cost"]):
            assert close(d.best_traj, o["best_traj"], scale=traj_scale(o["best_traj"])).all(), \
                "best trajectory differs"
            assert abs(d.steer - o["steer"]) < 1e-4 + 1e-4 * abs(o["steer"]), ("steer", d.steer, o["steer"])
            assert abs(d.speed - o["speed"]) < 1e-4 + 1e-4 * abs(o["speed"]), ("speed", d.speed, o["speed"])
    assert d.no_feasible == o["no_feasible"]
    if verbose:
        print(stats)
    return stats


def scenario(track, seed, k=1):
    poses, opp, n_opp = synth.scenario_batch(track, 1, k, seed)
    return poses[0], opp[0]
