"""Config 1 (one query, default 4x7 goal grid, one opponent): where the plan() latency goes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f1tenth_planning_b200 import LatticePlanner, synth  # noqa: E402
from f1tenth_planning_b200.engine import Engine  # noqa: E402

track = synth.ellipse_track()
grid = synth.corridor_grid()
la, wd = synth.goal_grid(1)
p1, o1, _ = synth.scenario_batch(track, 64, 1, 1001)


def p50(f, n=1000):
    for i in range(10):
        f(i)
    ts = []
    for i in range(n):
        t = time.perf_counter()
        f(i)
        ts.append(time.perf_counter() - t)
    return 1e6 * float(np.percentile(ts, 50))


eng = Engine(n_samples=100, window=128)
eng.set_track(track)
eng.set_grid(*grid)
eng.set_goal_grid(la, wd)
print("engine.plan detail=False p50 %.1f us" % p50(lambda i: eng.plan(p1[i % 64], o1[i % 64], detail=False)))
print("engine.plan detail=True  p50 %.1f us" % p50(lambda i: eng.plan(p1[i % 64], o1[i % 64], detail=True)))
eng.set_timing(True)
for i in range(20):
    eng.plan(p1[i], o1[i], detail=False)
print("kernels (sample, eval, select) us:", [round(1e3 * x, 1) for x in eng.mean_kernel_ms()[:3]])
eng.set_timing(False)
eng.set_graph(False)
print("engine.plan detail=False, no graph p50 %.1f us" % p50(lambda i: eng.plan(p1[i % 64], o1[i % 64], detail=False)))
pl = LatticePlanner(waypoints=track, n_samples=100, window=128)
pl.set_map(*grid)
print("LatticePlanner.plan p50 %.1f us" % p50(lambda i: pl.plan(*p1[i % 64], opponent_poses=o1[i % 64])))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for i in range(2000):
    pl.plan(*p1[i % 64], opponent_poses=o1[i % 64])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
