"""Mints golden vectors by running the UNMODIFIED reference functions from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/pp_spielberg.npz, pp_ellipse.npz, misc.npz.  Inputs (waypoints, poses) are
stored beside the outputs so the tests need nothing else.
"""
import os
import sys
import types
import warnings

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from f1tenth_planning.utils.utils import (nearest_point, intersect_point, get_actuation,  # noqa: E402
                                          get_rotation_matrix, pi_2_pi)
from f1tenth_planning.control.pure_pursuit.pure_pursuit import PurePursuitPlanner  # noqa: E402

from f1tenth_planning_b200 import synth  # noqa: E402


def run_pp(wp, poses, L):
    xy = wp[:, 0:2]
    b = poses.shape[0]
    nearest = np.zeros((b, 4))
    nearest_i = np.zeros(b, np.int32)
    look = np.zeros((b, 4))
    look_i = np.zeros(b, np.int32)
    act = np.zeros((b, 2))
    planner = PurePursuitPlanner(waypoints=wp)
    for k in range(b):
        pos = np.array([poses[k, 0], poses[k, 1]], dtype=np.float64)
        proj, dist, t, i = nearest_point(pos, xy)
        nearest[k] = (proj[0], proj[1], dist, t)
        nearest_i[k] = i
        p, i2, t2 = intersect_point(pos, float(L), xy, float(i + t), wrap=True)
        if i2 is not None:
            look[k] = (p[0], p[1], t2, 1.0)
            look_i[k] = i2
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            steer, speed = planner.plan(float(poses[k, 0]), float(poses[k, 1]), float(poses[k, 2]),
                                        float(L))
        act[k] = (steer, speed)
    return dict(nearest=nearest, nearest_i=nearest_i, lookahead=look, lookahead_i=look_i,
                actuation=act)


def main():
    rng = np.random.default_rng(20261017)
    # ---- Spielberg (reference fixture examples/control/Spielberg_raceline.csv) ----
    wp = np.loadtxt(os.path.join(REF, "examples/control/Spielberg_raceline.csv"), delimiter=";")
    kat = np.array([[0.0, -0.84, 3.40], [-20.0, -6.0, 3.4], [5.0, 5.0, 0.0]])
    near, _ = synth.random_poses(wp, 400, rng)
    seam_idx = np.concatenate([np.arange(0, 6), np.arange(wp.shape[0] - 8, wp.shape[0])])
    seam = np.stack([wp[seam_idx, 0] + rng.normal(0, 0.2, seam_idx.size),
                     wp[seam_idx, 1] + rng.normal(0, 0.2, seam_idx.size),
                     wp[seam_idx, 3] + rng.normal(0, 0.1, seam_idx.size)], axis=1)
    far = np.stack([rng.uniform(-90, 40, 40), rng.uniform(-40, 30, 40), rng.uniform(-3, 3, 40)], 1)
    onvert = np.stack([wp[5:25, 0], wp[5:25, 1], wp[5:25, 3]], 1)  # exactly on waypoints (ties)
    poses = np.concatenate([kat, near[:, :3], seam, far, onvert])
    out = run_pp(wp, poses, 0.8)
    np.savez_compressed(os.path.join(HERE, "pp_spielberg.npz"), waypoints=wp, poses=poses,
                        lookahead_distance=0.8, **out)

    # ---- synthetic ellipse ----
    tr = synth.ellipse_track()
    near, _ = synth.random_poses(tr, 500, rng)
    seam_idx = np.concatenate([np.arange(0, 6), np.arange(tr.shape[0] - 8, tr.shape[0])])
    seam = np.stack([tr[seam_idx, 0] + rng.normal(0, 0.2, seam_idx.size),
                     tr[seam_idx, 1] + rng.normal(0, 0.2, seam_idx.size),
                     tr[seam_idx, 3] + rng.normal(0, 0.1, seam_idx.size)], axis=1)
    poses = np.concatenate([near[:, :3], seam])
    out = run_pp(tr, poses, 0.8)
    np.savez_compressed(os.path.join(HERE, "pp_ellipse.npz"), poses=poses, lookahead_distance=0.8,
                        **out)

    # ---- intersect_point direct calls (start parameter / wrap variants), misc helpers ----
    xy = wp[:, 0:2]
    q_pts, q_t, q_wrap, q_r, q_out, q_i = [], [], [], [], [], []
    for _ in range(300):
        i = int(rng.integers(0, wp.shape[0] - 1))
        pt = xy[i] + rng.normal(0, 0.3, 2)
        t0 = float(np.clip(i + rng.uniform(-3, 1), 0, wp.shape[0] - 1.001))
        wrap = bool(rng.integers(0, 2))
        r = float(rng.choice([0.4, 0.6, 0.8, 1.0, 2.5]))
        p, i2, t2 = intersect_point(pt.astype(np.float64), r, xy, t0, wrap=wrap)
        q_pts.append(pt); q_t.append(t0); q_wrap.append(wrap); q_r.append(r)
        q_out.append((p[0], p[1], t2, 1.0) if i2 is not None else (0.0, 0.0, 0.0, 0.0))
        q_i.append(i2 if i2 is not None else 0)
    act_in = np.stack([rng.uniform(-3, 3, 64), rng.uniform(-2, 2, 64), rng.uniform(-2, 2, 64),
                       rng.uniform(0, 8, 64), rng.uniform(-1, 1, 64), rng.uniform(-1, 1, 64),
                       rng.uniform(0.3, 2.0, 64)], axis=1)
    act_in[0] = (0.3, 1.0, 0.5, 4.0, 0.2, 0.1, 0.8)
    act_in[1] = (0.0, 1.0, 0.0, 3.0, 0.0, 0.0, 0.8)  # |wy| < 1e-6 branch
    act_out = np.array([get_actuation(r[0], np.array([r[1], r[2], r[3]]), np.array([r[4], r[5]]),
                                      r[6], 0.33) for r in act_in])
    angles = np.array([3.5, -3.5, 0.3, -0.3, 3.2, -3.2])
    # cost helper / select that do run in the reference: needs a pyclothoids stub
    sys.modules["pyclothoids"] = types.SimpleNamespace(Clothoid=object)
    from f1tenth_planning.planning.lattice_planner.lattice_planner import (LatticePlanner,
                                                                           get_length_cost)
    np.savez_compressed(
        os.path.join(HERE, "misc.npz"),
        ip_points=np.array(q_pts), ip_t=np.array(q_t), ip_wrap=np.array(q_wrap),
        ip_radius=np.array(q_r), ip_out=np.array(q_out), ip_i=np.array(q_i, np.int32),
        act_in=act_in, act_out=act_out, angles=angles,
        pi_2_pi=np.array([pi_2_pi(a) for a in angles]),
        rot_03=get_rotation_matrix(0.3),
        length_cost=get_length_cost(np.array([[2.0, 0.0], [4.0, 0.0]])),
        select=np.array([LatticePlanner().select([3.0, 1.0, 1.0, 2.0])]))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
