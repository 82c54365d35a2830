"""Drop-in for f1tenth_planning/control/pure_pursuit/pure_pursuit.py:37-122, backed by the
batched CUDA kernel (K1).  `plan` keeps the reference signature and return order
(steering_angle, speed); `plan_batch` is the additive batched form (BASELINE config 2)."""
import warnings

import numpy as np

from .engine import Engine, fingerprint


class PurePursuitPlanner():
    """Pure pursuit tracking controller (Coulter 1992).  All poses in the map frame.

    Args:
        wheelbase (float): pure_pursuit.py:51
        waypoints (numpy.ndarray [N x m], m >= 3): columns x, y, velocity[, heading, ...]
        device (int, optional): CUDA device index (default: $LOCAL_RANK or 0)
    """

    def __init__(self, wheelbase=0.33, waypoints=None, device=None):
        self.max_reacquire = 20.
        self.wheelbase = wheelbase
        self.waypoints = waypoints
        self._device = device
        self._engine = None
        self._key = None

    def _sync(self):
        if self._engine is None:
            self._engine = Engine(device=self._device, wheelbase=float(self.wheelbase),
                                  max_reacquire=float(self.max_reacquire))
        cfg = self._engine.config
        if cfg.wheelbase != self.wheelbase or cfg.max_reacquire != self.max_reacquire:
            self._engine.configure(wheelbase=float(self.wheelbase),
                                   max_reacquire=float(self.max_reacquire))
        w = np.ascontiguousarray(self.waypoints, dtype=np.float64)
        key = fingerprint(w)
        if key != self._key:
            self._engine.set_track(w)
            self._key = key
        return self._engine

    def _check(self, waypoints):
        if waypoints is not None:
            if waypoints.shape[1] < 3 or len(waypoints.shape) != 2:
                raise ValueError('Waypoints needs to be a (Nxm), m >= 3, numpy array!')
            self.waypoints = waypoints
        else:
            if self.waypoints is None:
                raise ValueError('Please set waypoints to track during planner instantiation or when calling plan()')

    def plan(self, pose_x, pose_y, pose_theta, lookahead_distance, waypoints=None):
        """Returns (steering_angle, speed) -- pure_pursuit.py:85-122."""
        self._check(waypoints)
        eng = self._sync()
        r = eng.pure_pursuit_batch(np.array([[float(pose_x), float(pose_y), float(pose_theta)]]),
                                   lookahead_distance)
        if r.status[0] == 0:
            warnings.warn('Cannot find lookahead point, stopping...')
            return 0.0, 0.0
        return float(r.actuation[0, 0]), float(r.actuation[0, 1])

    def plan_batch(self, poses, lookahead_distance, waypoints=None):
        """poses [B,3] (x, y, theta) -> PurePursuitBatch (nearest, nearest_i, lookahead,
        lookahead_i, actuation [B,2] = (steer, speed), status)."""
        self._check(waypoints)
        return self._sync().pure_pursuit_batch(poses, lookahead_distance)
