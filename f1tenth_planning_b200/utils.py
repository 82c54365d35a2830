"""Drop-in for f1tenth_planning/utils/utils.py (pure-pursuit and geometry utilities), backed by
the CUDA library.  Same names, argument meaning and return shapes as the reference:

    nearest_point(point, trajectory)                       utils/utils.py:37-67
    intersect_point(point, radius, trajectory, t, wrap)    utils/utils.py:69-151
    get_actuation(pose_theta, lookahead_point, position,
                  lookahead_distance, wheelbase)           utils/utils.py:153-161
    get_rotation_matrix(theta)                             utils/utils.py:271-274
    pi_2_pi(angle)                                         utils/utils.py:276-283

The free functions take the trajectory on every call like the reference does; the device copy is
cached on the array's content so repeated calls on the same track upload once.
"""
import math

import numpy as np

from .engine import Engine, fingerprint

_engines = {}      # device -> Engine used by the free functions
_track_key = {}    # device -> key of the trajectory currently uploaded


def _engine_for(trajectory, device=None):
    traj = np.ascontiguousarray(trajectory, dtype=np.float64)
    if traj.ndim != 2 or traj.shape[1] < 2:
        raise ValueError("trajectory must be (N, 2+)")
    eng = _engines.get(device)
    if eng is None:
        eng = Engine(device=device)
        _engines[device] = eng
    key = fingerprint(traj)
    if _track_key.get(device) != key:
        eng.set_track(traj)
        _track_key[device] = key
    return eng


def nearest_point(point, trajectory, device=None):
    """Nearest point on the open piecewise-linear trajectory.

    Returns (nearest_point (2,), nearest_dist, t, i) exactly like utils/utils.py:37-67."""
    eng = _engine_for(trajectory, device)
    pose = np.array([[float(point[0]), float(point[1]), 0.0]])
    r = eng.pure_pursuit_batch(pose, 0.0)
    return r.nearest[0, 0:2].copy(), float(r.nearest[0, 2]), float(r.nearest[0, 3]), int(r.nearest_i[0])


def intersect_point(point, radius, trajectory, t=0.0, wrap=False, device=None):
    """First intersection of the circle (point, radius) with the trajectory starting at parameter
    t.  Returns (p (2,) | None, i | None, t | None) like utils/utils.py:69-151."""
    eng = _engine_for(trajectory, device)
    if not (0.0 <= float(t) < eng.n_waypoints):
        # the reference indexes trajectory[int(t)] unchecked (numba: no bounds check)
        raise IndexError("start parameter t=%r outside [0, %d)" % (t, eng.n_waypoints))
    out, out_i = eng.intersect_point_batch(np.array([[float(point[0]), float(point[1])]]),
                                           np.array([float(t)]), radius, wrap)
    if out[0, 3] == 0.0:
        return None, None, None
    return out[0, 0:2].copy(), int(out_i[0]), float(out[0, 2])


def get_actuation(pose_theta, lookahead_point, position, lookahead_distance, wheelbase,
                  device=None):
    """Pure-pursuit steering law, returns (speed, steering_angle) like utils/utils.py:153-161."""
    eng = _engines.get(device)
    if eng is None:
        eng = Engine(device=device)
        _engines[device] = eng
    row = np.array([[float(pose_theta), float(lookahead_point[0]), float(lookahead_point[1]),
                     float(lookahead_point[2]), float(position[0]), float(position[1]),
                     float(lookahead_distance)]])
    out = eng.get_actuation_batch(row, wheelbase)
    return float(out[0, 0]), float(out[0, 1])


def get_rotation_matrix(theta):
    """utils/utils.py:271-274 (a 2x2 host constant; nothing to accelerate)."""
    c, s = np.cos(theta), np.sin(theta)
    return np.ascontiguousarray(np.array([[c, -s], [s, c]]))


def pi_2_pi(angle):
    """utils/utils.py:276-283"""
    if angle > math.pi:
        return angle - 2.0 * math.pi
    if angle < -math.pi:
        return angle + 2.0 * math.pi
    return angle
