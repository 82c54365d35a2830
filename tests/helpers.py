"""Parity helpers shared by the GPU tests, __graft_entry__.smoke() and tools/: build a matching
(oracle world, CUDA engine) pair and compare one plan against the oracle with the tolerance
classes BASELINE.json's north_star states:
    - spiral states and costs within 1e-4 relative (+1e-5 absolute floor),
    - collision flags and selected index bit-exact except where the decision margin is tiny
      (cost gap < 1e-5; collision margin < 1e-4 m; validity margin on kappa_max / tolerance).
"""
import json
import os

import numpy as np

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co

REL = 1e-4
ABS = 1e-5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARITY_LOG = os.environ.get("F1L_PARITY_LOG") or os.path.join(ROOT, "gpurun_out", "parity_counts.json")


def record_parity(name, stats):
    """Observed mismatch counts of a parity test -> one JSON file (merged by test name).  The file
    written by the GPU run is committed as profiles/r<round>_parity.json."""
    try:
        os.makedirs(os.path.dirname(PARITY_LOG), exist_ok=True)
        try:
            with open(PARITY_LOG) as f:
                data = json.load(f)
        except Exception:
            data = {}
        data[name] = stats
        with open(PARITY_LOG, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True, default=float)
    except OSError:
        pass

CFG_FIELDS = ["n_samples", "n_newton", "window", "n_shift", "n_cull", "literal_tracker",
              "use_goal_kappa", "generator", "collision_mode", "kappa_max", "car_length", "car_width", "converge_tol",
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


def compare_plan(d, o, cfg, kappa_max=None, verbose=False, record=None):
    """d: engine.PlanDetail (GPU), o: oracle plan dict.  Returns a dict of statistics and raises
    AssertionError on a parity violation.  record: name under which the counts go to the parity
    log (record_parity)."""
    stats = {"candidates": int(o["costs"].shape[0])}
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
    for c in mism:
        # explainable only by max|kappa| ~ kappa_max or an endpoint error ~ tolerance
        if "states" in o:
            cls = validity_class(o, c, cfg, kappa_max)
            assert cls is not None, ("validity flag differs away from the margin", int(c))
            stats["valid:" + cls] = stats.get("valid:" + cls, 0) + 1
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
            if getattr(d, "best_traj_map", None) is not None and "best_traj_map" in o:
                # (X, Y, v, Theta): positions to the vehicle-frame tolerance, speed exact
                dm, om = d.best_traj_map, o["best_traj_map"]
                sc = traj_scale(o["best_traj"])[0]
                assert (np.abs(dm[:, :2] - om[:, :2]) <= ABS + REL * max(sc[0], sc[1])).all(), "map-frame xy"
                assert np.array_equal(dm[:, 2], om[:, 2]), "map-frame speed column"
                assert (np.abs(dm[:, 3] - om[:, 3]) <= ABS + REL * sc[2]).all(), "map-frame heading"
    assert d.no_feasible == o["no_feasible"]
    if verbose:
        print(stats)
    if record:
        record_parity(record, stats)
    return stats


def validity_class(o, c, cfg, kappa_max=None):
    """Why may the validity flag of candidate c differ between the FP32 device path and the
    float64 oracle?  Only at a decision margin: max|kappa| within 0.1 % of kappa_max, or the
    endpoint error within 50 % of the tolerance (it IS the quantity thresholded, and FP32
    quadrature moves it by ~1e-5 m against a tolerance of 1e-4 m), or a non-finite solution.
    Returns the class name or None (unexplained)."""
    km = cfg.kappa_max if kappa_max is None else kappa_max
    st = o["states"][c]
    g = o["goals"][c]
    tol = cfg.converge_tol * max(1.0, float(np.linalg.norm(g)))
    maxk = np.abs(st[:, 3]).max()
    end_err = np.abs(st[-1, :3] - g).max()
    if km > 0 and abs(maxk - km) < 1e-3 * km:
        return "kappa_margin"
    if not np.isfinite(end_err):
        return "non_finite"
    if abs(end_err - tol) < 0.5 * tol:
        return "tolerance_margin"
    return None


def classify_flags(gf, o, cfg, kappa_max=None, counts=None):
    """Every candidate whose flags (valid, collide_opp, collide_map, no_centre) differ between the
    device (gf [C] uint8) and the oracle plan dict o (with states and margins) must fall into a
    margin class; raises otherwise.  Returns / updates the counts per class."""
    counts = {} if counts is None else counts
    gf = gf.astype(np.int32) & 0xF
    of = o["flags"].astype(np.int32) & 0xF
    assert ((gf & co.FLAG_NO_CENTRE) == (of & co.FLAG_NO_CENTRE)).all(), "centre flags differ"
    gv, ov = (gf & 1) != 0, (of & 1) != 0
    for c in np.nonzero(gv != ov)[0]:
        cls = validity_class(o, c, cfg, kappa_max)
        assert cls is not None, ("validity flag differs away from the margin", int(c))
        counts["valid:" + cls] = counts.get("valid:" + cls, 0) + 1
    both = gv & ov
    for bit, col, name in ((co.FLAG_COLLIDE_OPP, 0, "opp"), (co.FLAG_COLLIDE_MAP, 1, "map")):
        for c in np.nonzero(both & ((gf & bit) != (of & bit)))[0]:
            assert o["margins"][c, col] < 1e-4, ("collision flag differs away from the boundary",
                                                 name, int(c), float(o["margins"][c, col]))
            key = "collide_%s:boundary<1e-4m" % name
            counts[key] = counts.get(key, 0) + 1
    return counts


def compare_batch(b, sub, poses, opp, n_opp, cfg, world, kappa_max=None):
    """b: engine.BatchPlan of the whole batch (with flags); sub: scenario indices checked against
    the oracle.  Every flag / finiteness / argmin mismatch is classified (margin classes of
    classify_flags; argmin: cost gap < 1e-5); anything else raises.  Returns the counts."""
    o = co.plan_batch(cfg, world, poses[sub], opp[sub] if opp is not None else None,
                      n_opp[sub] if n_opp is not None else None, n_threads=co.max_threads())
    gfl = b.flags[sub].astype(np.int32) & 0xF
    ofl = o["flags"].astype(np.int32) & 0xF
    counts = {"scenarios": int(len(sub)), "candidates": int(gfl.size),
              "flag_mismatch_candidates": int((gfl != ofl).sum()),
              "argmin_mismatch": 0, "argmin:cost_gap<1e-5": 0}
    fin = np.isfinite(o["costs"]) & np.isfinite(b.costs[sub])
    assert close(b.costs[sub][fin], o["costs"][fin]).all(), "total costs differ"
    err = np.abs(b.costs[sub][fin] - o["costs"][fin]) / (ABS / REL + np.abs(o["costs"][fin]))
    counts["cost_rel_err_max"] = float(err.max()) if err.size else 0.0
    bad = np.nonzero((gfl != ofl).any(axis=1))[0]
    for k in bad:   # re-run the single-query oracle for the margins / states of this scenario
        s = sub[k]
        n = int(n_opp[s]) if n_opp is not None else (opp.shape[1] if opp is not None else 0)
        os_ = co.plan(cfg, world, poses[s], opp[s, :n] if n else None, want_states=True)
        assert np.array_equal(os_["flags"], o["flags"][k])
        classify_flags(b.flags[s], os_, cfg, kappa_max, counts)
    # a flag mismatch changes finiteness; everywhere else finiteness must agree
    same_flags = gfl == ofl
    assert (np.isfinite(o["costs"]) == np.isfinite(b.costs[sub]))[same_flags].all()
    for k in np.nonzero(b.best_idx[sub] != o["best_idx"])[0]:
        counts["argmin_mismatch"] += 1
        gi, oi = int(b.best_idx[sub][k]), int(o["best_idx"][k])
        gap = abs(float(o["costs"][k, gi]) - float(o["costs"][k, oi]))
        gap2 = abs(float(b.costs[sub][k, gi]) - float(o["costs"][k, oi]))
        if gap < 1e-5 or gap2 < 1e-5:
            counts["argmin:cost_gap<1e-5"] += 1
        else:
            # the only other legitimate cause: a classified flag mismatch on either winner
            assert (gfl[k, gi] != ofl[k, gi]) or (gfl[k, oi] != ofl[k, oi]), \
                ("argmin differs with a cost gap", int(sub[k]), gi, oi, gap)
            counts["argmin:margin_flag_on_winner"] = counts.get("argmin:margin_flag_on_winner", 0) + 1
    return counts


def scenario(track, seed, k=1):
    poses, opp, n_opp = synth.scenario_batch(track, 1, k, seed)
    return poses[0], opp[0]
