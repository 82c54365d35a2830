"""Host-side mirror of the reference interface: names, argument meaning and error behaviour
(lattice_planner.py:113-172, pure_pursuit.py:100-106) -- no GPU needed for these paths."""
import numpy as np
import pytest

from f1tenth_planning_b200 import LatticePlanner, PurePursuitPlanner, flops, utils
from f1tenth_planning_b200 import lattice_planner as lp


def test_reference_error_conventions():
    p = LatticePlanner()
    with pytest.raises(NotImplementedError):          # lattice_planner.py:123-124
        p.sample(0.0, 0.0, 0.0, 0.0, None)
    with pytest.raises(NotImplementedError):          # :141-142
        p.eval([], [])
    p.add_cost_function([lambda t: 1.0, lambda t: 2.0])
    p.add_cost_function(lambda t: 3.0)
    assert len(p.cost_funcs) == 3
    with pytest.raises(ValueError):                   # :143-144
        p.eval([np.zeros((3, 4))], [0.5, 0.5])
    with pytest.raises(ValueError):                   # :145-146
        p.eval([np.zeros((3, 4))], [0.5, 0.5, 0.5])
    assert p.eval([np.zeros((3, 4))] * 2, [0.5, 0.25, 0.25]) == [1.75, 1.75]
    assert p.select([3.0, 1.0, 1.0, 2.0]) == 1        # :169-171 first minimum
    p.add_selection_function(lambda c: len(c) - 1)
    assert p.select([3.0, 1.0, 1.0, 2.0]) == 3
    with pytest.raises(ValueError):
        LatticePlanner().plan(0.0, 0.0, 0.0, 1.0)     # no waypoints
    with pytest.raises(ValueError):                   # pure_pursuit.py:105-106
        PurePursuitPlanner().plan(0.0, 0.0, 0.0, 0.8)
    with pytest.raises(ValueError):                   # :100-102
        PurePursuitPlanner().plan(0.0, 0.0, 0.0, 0.8, waypoints=np.zeros((4, 2)))


def test_sampler_plugin_signature():
    p = LatticePlanner()
    seen = {}

    def sampler(px, py, pth, v, wpts):
        seen["args"] = (px, py, pth, v, wpts)
        return np.zeros((3, 3))
    p.add_sample_function(sampler)
    g = p.sample(1.0, 2.0, 3.0, 4.0, "wp")
    assert g.shape == (3, 3) and seen["args"] == (1.0, 2.0, 3.0, 4.0, "wp")


def test_cost_helpers_follow_reference_formulas():
    rng = np.random.default_rng(0)
    traj = rng.normal(size=(3 * lp.NUM_STEPS, 4))
    prev = rng.normal(size=(lp.NUM_STEPS, 4))
    assert lp.get_length_cost(np.array([[2.0, 0], [4.0, 0]])).tolist() == [0.5, 0.25]
    mk = lp.get_max_curvature(traj, 3)
    mean = lp.get_mean_curvature(traj, 3)
    sim = lp.get_similarity_cost(traj, prev, 3)
    for i in range(3):
        blk = traj[i * 100:(i + 1) * 100]
        assert mk[i] == np.abs(blk[:, 3]).max()
        np.testing.assert_allclose(mean[i], np.abs(blk[:, 3]).mean())
        np.testing.assert_allclose(sim[i], np.sum((blk[:-15, 2] - prev[5:-10, 2]) ** 2))


def test_geometry_helpers():
    np.testing.assert_allclose(utils.get_rotation_matrix(0.3),
                               [[0.95533649, -0.29552021], [0.29552021, 0.95533649]], atol=1e-8)
    assert utils.pi_2_pi(3.5) == -2.7831853071795862 and utils.pi_2_pi(-3.5) == 2.7831853071795862
    assert utils.pi_2_pi(0.3) == 0.3


def test_flop_model_matches_survey():
    assert abs(flops.candidate_flops(M=100, W=128, K=8) - 256e3) < 2e3    # ~256 kFLOP
    assert abs(flops.candidate_flops(M=200, W=128, K=8) - 500e3) < 2e3    # C5: ~500 kFLOP
    assert abs(flops.pose_flops(2000) - 34e3) < 1e2                        # C2: ~34 kFLOP
    assert flops.candidate_flops(full=False) < 0.1 * flops.candidate_flops(full=True)


def test_lqr_mirror_error_conventions():
    """LQRPlanner keeps lqr.py's waypoint checks; the Riccati part is not silently emulated."""
    from f1tenth_planning_b200 import LQRPlanner
    with pytest.raises(ValueError):
        LQRPlanner().calc_control_points(np.zeros(4))
    with pytest.raises(ValueError):
        LQRPlanner().calc_control_points(np.zeros(4), waypoints=np.zeros((5, 4)))
    with pytest.raises(NotImplementedError):
        LQRPlanner(waypoints=np.zeros((5, 5))).plan(0.0, 0.0, 0.0, 1.0)


def test_fingerprint_sees_every_element():
    """raceline change detection (advisor r1): any single edited element of the array changes the
    fingerprint -- no subsampling"""
    from f1tenth_planning_b200.engine import fingerprint
    rng = np.random.default_rng(0)
    a = rng.normal(size=(2000, 5))
    k0 = fingerprint(a)
    assert fingerprint(a.copy()) == k0 and fingerprint(np.asfortranarray(a)) == k0
    for i, j in [(5, 2), (1999, 4), (0, 0), (17, 3), (1001, 1)]:
        b = a.copy()
        b[i, j] = np.nextafter(b[i, j], np.inf)          # one ulp
        assert fingerprint(b) != k0
    b = a.copy()
    b[[3, 4]] = b[[4, 3]]                                 # two rows swapped: same multiset
    assert fingerprint(b) != k0
    assert fingerprint(a[:1999]) != k0 and fingerprint(a.reshape(5, 2000))[0] != k0[0]
