"""The data-parallel step of f1tenth_planning/control/lqr/lqr.py: `calc_control_points`
(:60-102) -- front-axle nearest point, cross-track and heading errors, target heading / curvature /
velocity -- backed by the batched CUDA nearest-point kernel (K1 in front-axle mode), for one state
or a batch.  The rest of the reference's LQRPlanner (the discrete Riccati iteration and the 4x4
feedback law, :104-190) is sequential small-matrix algebra per query, not a data-parallel path
(DESIGN 10): `controller` / `plan` say so instead of silently computing on the
host."""
import numpy as np

from .stanley import StanleyPlanner


class LQRPlanner():
    """Args: wheelbase (float), waypoints (numpy.ndarray [N x 5]: x, y, velocity, heading, curvature)"""

    def __init__(self, wheelbase=0.33, waypoints=None, device=None):
        self.wheelbase = wheelbase
        self.waypoints = waypoints
        self.vehicle_control_e_cog = 0       # lqr.py:57-58
        self.vehicle_control_theta_e = 0
        self._front = StanleyPlanner(wheelbase=wheelbase, waypoints=waypoints, device=device)

    def _check(self, waypoints):
        if waypoints is not None:
            if len(waypoints.shape) != 2 or waypoints.shape[1] < 5:
                raise ValueError('Waypoints needs to be a (Nxm), m >= 5, numpy array!')
            self.waypoints = waypoints
        elif self.waypoints is None:
            raise ValueError('Please set waypoints to track during planner instantiation or when calling plan()')
        self._front.wheelbase = self.wheelbase
        return self.waypoints

    def calc_control_points_batch(self, vehicle_states, waypoints=None):
        """[B,4] (x, y, heading, velocity) -> [B,5] (theta_e, e_cog, theta_raceline, kappa_ref,
        goal_velocity) and the target indices [B]"""
        w = self._check(waypoints)
        f, idx = self._front.front_axle_errors(vehicle_states, 0.0, w)
        return f[:, :5].copy(), idx

    def calc_control_points(self, vehicle_state, waypoints=None):
        """lqr.py:60-102 -> (theta_e, e_cog, theta_raceline, kappa_ref, goal_velocity)"""
        f, _ = self.calc_control_points_batch(np.asarray(vehicle_state, dtype=np.float64)[None, :4], waypoints)
        self.vehicle_control_e_cog = float(f[0, 1])          # :97-98
        self.vehicle_control_theta_e = float(f[0, 0])
        return float(f[0, 0]), float(f[0, 1]), float(f[0, 2]), float(f[0, 3]), float(f[0, 4])

    def controller(self, *args, **kwargs):
        raise NotImplementedError('the Riccati iteration of lqr.py:104-190 is outside the accelerated '
                                  'path; use calc_control_points() for its inputs')

    plan = controller
