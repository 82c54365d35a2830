import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from f1tenth_planning_b200 import synth
from f1tenth_planning_b200.engine import Engine
track = synth.ellipse_track()
la, wd = synth.goal_grid(5)
eng = Engine(n_samples=200, window=128)
eng.set_track(track); eng.set_grid(*synth.corridor_grid()); eng.set_goal_grid(la, wd)
poses, opp, n_opp = synth.scenario_batch(track, 16, 8, 1005)
x = torch.zeros(1 << 20, device="cuda")
def run(mode):
    ts = []
    for i in range(5 + 100):
        t0 = time.perf_counter() + 2e-3
        if mode == "warm":
            while time.perf_counter() < t0 - 150e-6:
                for _ in range(8): x.add_(1.0)
            torch.cuda.synchronize()
        while time.perf_counter() < t0:
            pass
        t = time.perf_counter()
        eng.plan(poses[i % 16], opp[i % 16], update_prev=False, detail=False, rows=(0, 8))
        ts.append(time.perf_counter() - t)
    return 1e6 * np.percentile(ts[5:], 50)
for mode in ("back2back", "idle2ms", "warm", "idle2ms", "warm"):
    if mode == "back2back":
        ts = []
        for i in range(105):
            t = time.perf_counter(); eng.plan(poses[i % 16], opp[i % 16], update_prev=False, detail=False, rows=(0, 8)); ts.append(time.perf_counter() - t)
        print(mode, 1e6 * np.percentile(ts[5:], 50))
    else:
        print(mode, run(mode))
