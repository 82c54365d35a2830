"""Drop-in for f1tenth_planning/control/stanley/stanley.py:37-139 (front-wheel feedback
controller, Hoffmann et al. 2007), backed by the batched CUDA nearest-point kernel (K1 in
front-axle mode).  `plan` keeps the reference signature and return order (steering_angle, speed);
`plan_batch` is the additive batched form.  LQR's `calc_control_points`
(control/lqr/lqr.py:60-102) is the same front-axle computation: use `front_axle_errors`."""

import numpy as np

from .engine import Engine, fingerprint


class StanleyPlanner():
    """Args: wheelbase (float), waypoints (numpy.ndarray [N x m], m >= 4: x, y, velocity, heading)"""

    def __init__(self, wheelbase=0.33, waypoints=None, device=None):
        self.wheelbase = wheelbase
        self.waypoints = waypoints
        self._device = device
        self._engine = None
        self._key = None

    def _sync(self, waypoints=None):
        if waypoints is not None:
            if waypoints.shape[1] < 4 or len(waypoints.shape) != 2:
                raise ValueError('Waypoints needs to be a (Nxm), m >= 4, numpy array!')
            self.waypoints = waypoints
        elif self.waypoints is None:
            raise ValueError('Please set waypoints to track during planner instantiation or when calling plan()')
        if self._engine is None:
            self._engine = Engine(device=self._device)
        w = np.ascontiguousarray(self.waypoints, dtype=np.float64)
        key = fingerprint(w)
        if key != self._key:
            self._engine.set_track(w)
            self._key = key
        return self._engine

    def front_axle_errors(self, vehicle_states, k_path=5., waypoints=None):
        """[B,4] states -> (front [B,6] = theta_e, ef, theta_raceline, kappa_ref, goal_velocity,
        delta; target_index [B])"""
        return self._sync(waypoints).front_axle_batch(vehicle_states, self.wheelbase, k_path)

    def calc_theta_and_ef(self, vehicle_state, waypoints):
        """stanley.py:57-85 -> (theta_e, ef, target_index, goal_velocity)"""
        f, i = self.front_axle_errors(np.asarray(vehicle_state, dtype=np.float64)[None, :4], 0.0, waypoints)
        return float(f[0, 0]), np.array([f[0, 1]]), int(i[0]), float(f[0, 4])

    def controller(self, vehicle_state, waypoints, k_path):
        """stanley.py:87-112 -> (delta, goal_velocity)"""
        f, _ = self.front_axle_errors(np.asarray(vehicle_state, dtype=np.float64)[None, :4], k_path, waypoints)
        return float(f[0, 5]), float(f[0, 4])

    def plan(self, pose_x, pose_y, pose_theta, velocity, k_path=5., waypoints=None):
        """stanley.py:114-139 -> (steering_angle, speed)"""
        self._sync(waypoints)
        return self.controller(np.array([pose_x, pose_y, pose_theta, velocity]), None, k_path)

    def plan_batch(self, vehicle_states, k_path=5., waypoints=None):
        """[B,4] (x, y, theta, velocity) -> (steering_angle [B], speed [B])"""
        f, _ = self.front_axle_errors(vehicle_states, k_path, waypoints)
        return f[:, 5].copy(), f[:, 4].copy()
