"""Config-5 shard (8192 candidates of the 65 536 x 200 query, rows r, r + 8, ...): eval_kernel time
for several CTA plans (F1L_EVAL_PLAN tuning hook) and the valid / colliding fractions."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from f1tenth_planning_b200 import synth
    from f1tenth_planning_b200.engine import Engine
    track = synth.ellipse_track()
    la, wd = synth.goal_grid(5)
    eng = Engine(n_samples=200, window=128)
    eng.set_track(track)
    eng.set_grid(*synth.corridor_grid())
    eng.set_goal_grid(la, wd)
    poses, opp, n_opp = synth.scenario_batch(track, 16, 8, 1005)
    d = eng.plan(poses[0], opp[0], update_prev=False, rows=(0, 8))
    mine = np.zeros((256, 256), bool); mine[0::8] = True
    f = d.flags[mine.ravel()]
    eng.set_timing(True)
    for i in range(40):
        eng.plan(poses[i % 16], opp[i % 16], update_prev=False, detail=False, rows=(0, 8))
    sm, ev, se, n = eng.mean_kernel_ms()
    print("plan %-8s eval %.1f us (sample %.1f select %.1f) shape %s | valid %.3f collide %.3f"
          % (os.environ.get("F1L_EVAL_PLAN", "auto"), 1e3 * ev, 1e3 * sm, 1e3 * se, eng.last_eval_shape()["name"],
             ((f & 1) != 0).mean(), ((f & 6) != 0).mean()))
else:
    for plan in (None, "8,8", "8,16", "8,24", "8,32", "8,64", "4,4", "4,8", "4,16"):
        env = dict(os.environ)
        if plan:
            env["F1L_EVAL_PLAN"] = plan
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
