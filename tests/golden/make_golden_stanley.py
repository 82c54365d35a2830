"""Golden vectors for the front-axle consumers of nearest_point, minted from the UNMODIFIED
reference controllers (control/stanley/stanley.py, control/lqr/lqr.py).  Build container only:
    python tests/golden/make_golden_stanley.py   ->  tests/golden/stanley.npz
"""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from f1tenth_planning.control.stanley.stanley import StanleyPlanner  # noqa: E402
from f1tenth_planning.control.lqr.lqr import LQRPlanner  # noqa: E402

from f1tenth_planning_b200 import synth  # noqa: E402


def main():
    rng = np.random.default_rng(77)
    g = np.load(os.path.join(HERE, "pp_spielberg.npz"))
    wp = g["waypoints"]
    # stanley expects columns [x, y, velocity, heading]; the Spielberg csv is x, y, v, psi, kappa
    states, _ = synth.random_poses(wp, 300, rng)
    states = np.concatenate([states, [[0.0, -0.84, 3.40, 4.0], [5.0, 5.0, 0.0, 2.0]]])
    st = StanleyPlanner(waypoints=wp)
    lq = LQRPlanner(waypoints=wp)
    out = np.zeros((states.shape[0], 6))
    idx = np.zeros(states.shape[0], np.int32)
    plan = np.zeros((states.shape[0], 2))
    for k, s in enumerate(states):
        th_e, ef, ti, gv = st.calc_theta_and_ef(s, wp)
        delta, gv2 = st.controller(s, wp, 5.0)
        l_th_e, l_ef, l_thr, l_kap, l_gv = lq.calc_control_points(s, wp)
        assert abs(l_th_e - th_e) < 1e-15 and abs(l_ef - ef[0]) < 1e-15
        out[k] = (th_e, ef[0], l_thr, l_kap, gv, delta)
        idx[k] = ti
        plan[k] = st.plan(s[0], s[1], s[2], s[3], 5.0)
    np.savez_compressed(os.path.join(HERE, "stanley.npz"), states=states, front=out, target_index=idx,
                        plan=plan, k_path=5.0, wheelbase=0.33)
    print("stanley golden vectors written")


if __name__ == "__main__":
    main()
