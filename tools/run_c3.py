"""A few config-3 single queries (4096 candidates) -- used under ncu to profile the latency path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f1tenth_planning_b200 import synth  # noqa: E402
from f1tenth_planning_b200.engine import Engine  # noqa: E402

track = synth.ellipse_track()
la, wd = synth.goal_grid(3)
eng = Engine(n_samples=100, window=128)
eng.set_graph(False)
eng.set_track(track)
eng.set_grid(*synth.corridor_grid())
eng.set_goal_grid(la, wd)
poses, opp, n_opp = synth.scenario_batch(track, 8, 8, 1003)
for i in range(8):
    d = eng.plan(poses[i], opp[i], update_prev=True, detail=False)
print(d.best_idx, d.best_cost)
