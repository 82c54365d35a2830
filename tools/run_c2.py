"""Config 2 (10^5 poses x 2000-waypoint track, batched nearest_point + pure pursuit), device
resident -- used under ncu to profile pp_batch_kernel, and to time it alone."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f1tenth_planning_b200 import synth  # noqa: E402
from f1tenth_planning_b200.engine import Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
track = synth.ellipse_track()
eng = Engine()
eng.set_track(track)
pp, _ = synth.random_poses(track, B, np.random.default_rng(1002))
dev = torch.device("cuda", 0)
tp = torch.from_numpy(np.ascontiguousarray(pp[:, :3])).to(dev)
near = torch.empty(B, 4, dtype=torch.float64, device=dev)
ni = torch.empty(B, dtype=torch.int32, device=dev)
look = torch.empty(B, 4, dtype=torch.float64, device=dev)
li = torch.empty(B, dtype=torch.int32, device=dev)
act = torch.empty(B, 2, dtype=torch.float64, device=dev)
stt = torch.empty(B, dtype=torch.int32, device=dev)
for _ in range(3):
    eng.pure_pursuit_batch_dev(tp, 0.8, near, ni, look, li, act, stt)
torch.cuda.synchronize(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng.pure_pursuit_batch_dev(tp, 0.8, near, ni, look, li, act, stt)
e1.record()
torch.cuda.synchronize(dev)
ms = e0.elapsed_time(e1) / 20
print("c2: %d poses, %.4f ms, %.3e poses/s, %.2f algorithmic TFLOP/s" %
      (B, ms, B / (ms * 1e-3), B * (17 * (track.shape[0] - 1) + 60) / (ms * 1e-3) / 1e12))
