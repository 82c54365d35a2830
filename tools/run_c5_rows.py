"""Config 5 (65 536 candidates x 200 samples) as ONE rank of a row-interleaved 8-way shard sees it:
per-kernel device times (sampler / eval / select) and wall-clock p50 of plan(rows=(r, W)), next to
the unsharded query.  No peers needed: without attach_peers the call returns the shard's winner."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f1tenth_planning_b200 import synth  # noqa: E402
from f1tenth_planning_b200.engine import Engine  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
track = synth.ellipse_track()
la, wd = synth.goal_grid(5)
eng = Engine(n_samples=200, window=128, prune_window=int(os.environ.get("PRUNE", "0")))
eng.set_track(track)
eng.set_grid(*synth.corridor_grid())
eng.set_goal_grid(la, wd)
poses, opp, n_opp = synth.scenario_batch(track, 16, 8, 1005)
for rows in (None, (0, W), (W - 1, W)):
    kw = {} if rows is None else {"rows": rows}
    eng.set_timing(False)
    for i in range(5):
        eng.plan(poses[i], opp[i], update_prev=False, detail=False, **kw)
    ts = []
    for i in range(200):
        t = time.perf_counter()
        eng.plan(poses[i % 16], opp[i % 16], update_prev=False, detail=False, **kw)
        ts.append(time.perf_counter() - t)
    eng.set_timing(True)
    for i in range(32):
        eng.plan(poses[i % 16], opp[i % 16], update_prev=False, detail=False, **kw)
    sm, ev, se, n = eng.mean_kernel_ms()
    print("rows=%s: wall p50 %.1f us p10 %.1f us | kernels (plain launches): sample %.1f eval %.1f select %.1f us | %s"
          % (rows, 1e6 * np.percentile(ts, 50), 1e6 * np.percentile(ts, 10), 1e3 * sm, 1e3 * ev, 1e3 * se,
             eng.last_eval_shape()))
