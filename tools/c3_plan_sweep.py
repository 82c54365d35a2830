"""Config 3 (4096 candidates x 100 samples, 8 opponents, grid): eval_kernel time and plan() p50 for
several CTA plans (F1L_EVAL_PLAN tuning hook)."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from f1tenth_planning_b200 import synth
    from f1tenth_planning_b200.engine import Engine
    track = synth.ellipse_track()
    la, wd = synth.goal_grid(3)
    eng = Engine(n_samples=100, window=128)
    eng.set_track(track)
    eng.set_grid(*synth.corridor_grid())
    eng.set_goal_grid(la, wd)
    poses, opp, n_opp = synth.scenario_batch(track, 16, 8, 1003)
    for i in range(5):
        eng.plan(poses[i], opp[i], update_prev=False, detail=False)
    ts = []
    for i in range(300):
        t = time.perf_counter()
        eng.plan(poses[i % 16], opp[i % 16], update_prev=False, detail=False)
        ts.append(time.perf_counter() - t)
    eng.set_timing(True)
    for i in range(40):
        eng.plan(poses[i % 16], opp[i % 16], update_prev=False, detail=False)
    sm, ev, se, n = eng.mean_kernel_ms()
    print("plan %-8s p50 %.1f us | eval %.1f us (sample %.1f select %.1f) %s"
          % (os.environ.get("F1L_EVAL_PLAN", "auto"), 1e6 * np.percentile(ts, 50), 1e3 * ev, 1e3 * sm, 1e3 * se,
             eng.last_eval_shape()))
else:
    for plan in (None, "7,7", "7,14", "8,8", "8,16", "4,4", "4,8"):
        env = dict(os.environ)
        if plan:
            env["F1L_EVAL_PLAN"] = plan
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
