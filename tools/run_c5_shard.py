"""A few rank-local calls of the config-5 row shard (rows r, r + 8, ...) with plain launches -- used
under ncu to profile sample_kernel / select_kernel on the latency path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f1tenth_planning_b200 import synth  # noqa: E402
from f1tenth_planning_b200.engine import Engine  # noqa: E402

track = synth.ellipse_track()
la, wd = synth.goal_grid(5)
eng = Engine(n_samples=200, window=128, prune_window=int(os.environ.get("PRUNE", "0")))
eng.set_graph(False)
eng.set_track(track)
eng.set_grid(*synth.corridor_grid())
eng.set_goal_grid(la, wd)
poses, opp, n_opp = synth.scenario_batch(track, 8, 8, 1005)
for i in range(8):
    d = eng.plan(poses[i], opp[i], update_prev=True, detail=False, rows=(0, 8))
print(d.best_idx, d.best_cost)
