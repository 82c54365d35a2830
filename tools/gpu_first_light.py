"""Diagnostics for a first run on the GPU box: prints parity statistics instead of asserting."""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f1tenth_planning_b200 import synth  # noqa: E402
from f1tenth_planning_b200.engine import Engine  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    tr = synth.ellipse_track()
    t0 = time.time()
    eng = Engine()
    print("engine create %.3fs" % (time.time() - t0))
    print("peaks (fp32 TFLOP/s, mufu Gop/s):", eng.measure_peaks())
    lut, ranges = eng.get_lut()
    ref = co.lut_build()
    print("lut flag agree", (lut[..., 3] == ref[..., 3]).mean(), "converged frac", lut[..., 3].mean(),
          "maxdiff", np.abs(lut - ref)[(lut[..., 3] == 1) & (ref[..., 3] == 1)].max())
    # K1
    eng.set_track(tr)
    rng = np.random.default_rng(0)
    poses, _ = synth.random_poses(tr, 2000, rng)
    r = eng.pure_pursuit_batch(poses[:, :3], 0.8)
    o = co.pure_pursuit_batch(tr, poses[:, :3], 0.8, n_threads=8)
    print("K1 idx agree", (r.nearest_i == o["nearest_i"]).mean(), "nearest maxdiff",
          np.abs(r.nearest - o["nearest"]).max(), "act maxdiff", np.abs(r.actuation - o["actuation"]).max(),
          "status agree", (r.status == o["status"]).mean())
    # lattice
    for name, cfgid, kw, k in (("C1", 1, dict(window=0, kappa_max=0.0), 1), ("C3", 3, dict(), 8)):
        la, wd = synth.goal_grid(cfgid)
        grid = synth.corridor_grid() if cfgid == 3 else None
        eng2, cfg, world = H.make_pair(tr, la, wd, grid=grid, **kw)
        pose, opp = H.scenario(tr, 1000 + cfgid, k)
        d = eng2.plan(pose, opp, update_prev=False, want_states=True)
        ob = co.plan(cfg, world, pose, opp, want_states=True)
        gv, ov = (d.flags & 1) != 0, (ob["flags"] & 1) != 0
        both = gv & ov
        print(name, "valid gpu/oracle/both", gv.sum(), ov.sum(), both.sum(), "flags equal",
              ((d.flags & 15) == ob["flags"]).mean(), "newton passes", (d.flags >> 4).mean())
        if both.any():
            print("  goals maxdiff", np.abs(d.goals - ob["goals"]).max())
            print("  params maxrel", (np.abs(d.params[both, :3] - ob["params"][both, :3]) /
                                      (0.1 + np.abs(ob["params"][both, :3]))).max())
            print("  states maxabs", np.abs(d.states[both] - ob["states"][both]).max(axis=(0, 1)))
            print("  terms maxrel", (np.abs(d.terms[both] - ob["terms"][both]) /
                                     (0.1 + np.abs(ob["terms"][both]))).max(axis=0))
            fin = np.isfinite(d.costs) & np.isfinite(ob["costs"])
            print("  finite gpu/oracle", np.isfinite(d.costs).sum(), np.isfinite(ob["costs"]).sum(),
                  "cost maxrel", (np.abs(d.costs[fin] - ob["costs"][fin]) / np.abs(ob["costs"][fin])).max()
                  if fin.any() else None)
        print("  best", d.best_idx, ob["best_idx"], d.best_cost, ob["best_cost"], "steer", d.steer,
              ob["steer"], "speed", d.speed, ob["speed"])
        try:
            print("  compare:", H.compare_plan(d, ob, cfg))
        except AssertionError as e:
            print("  compare FAILED:", e)
        eng2.set_timing(True)
        for _ in range(3):
            eng2.plan(pose, opp, update_prev=False, detail=False)
        ts = []
        for _ in range(50):
            t = time.perf_counter()
            eng2.plan(pose, opp, update_prev=False, detail=False)
            ts.append(time.perf_counter() - t)
        print("  plan() p50 %.1f us, kernel ms (sample, eval, select) %s" %
              (1e6 * np.median(ts), eng2.last_kernel_ms()))
    # batch throughput quick look
    la, wd = synth.goal_grid(4)
    eng3, cfg, world = H.make_pair(tr, la, wd, grid=synth.corridor_grid())
    S = 20000
    poses, opp, n_opp = synth.scenario_batch(tr, S, 8, 4)
    eng3.plan_batch(poses, opp, n_opp)
    t = time.perf_counter()
    eng3.plan_batch(poses, opp, n_opp)
    dt = time.perf_counter() - t
    print("batch e2e: %d scenarios x %d cands in %.1f ms -> %.3e cand/s" %
          (S, eng3.n_candidates, dt * 1e3, S * eng3.n_candidates / dt))


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
