"""K1 parity (through the C-ABI): batched nearest_point + pure pursuit vs the golden vectors minted
from the reference and vs the oracle at BASELINE config 2 sizes."""
import numpy as np
import pytest

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _engine(track):
    from f1tenth_planning_b200.engine import Engine
    eng = Engine()
    eng.set_track(track)
    return eng


def _check(r, g, n_wp):
    """r: PurePursuitBatch (GPU), g: dict of reference outputs."""
    gi = g["nearest_i"]
    same = r.nearest_i == gi
    # (k, t=1) == (k+1, t=0): a vertex tie may resolve to the neighbouring segment
    for k in np.nonzero(~same)[0]:
        assert abs(int(r.nearest_i[k]) - int(gi[k])) == 1, (k, r.nearest_i[k], gi[k])
        assert abs(r.nearest[k, 2] - g["nearest"][k, 2]) < 1e-9
    assert (~same).mean() < 0.02
    np.testing.assert_allclose(r.nearest[same], g["nearest"][same], rtol=1e-9, atol=1e-9)
    m = (g["nearest"][:, 2] < 0.8) & same
    found = g["lookahead"][:, 3] > 0
    assert (r.lookahead[m, 3] > 0).tolist() == found[m].tolist()
    mf = m & found
    assert (r.lookahead_i[mf] == g["lookahead_i"][mf]).all()
    np.testing.assert_allclose(r.lookahead[mf, :3], g["lookahead"][mf, :3], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r.actuation[same], g["actuation"][same], rtol=1e-9, atol=1e-9)


def test_golden_spielberg(golden_spielberg):
    g = golden_spielberg
    eng = _engine(g["waypoints"])
    r = eng.pure_pursuit_batch(g["poses"], float(g["lookahead_distance"]))
    _check(r, g, g["waypoints"].shape[0])
    # the three known-answer poses of SURVEY appendix C
    assert r.nearest_i[:3].tolist() == [1690, 103, 1659]
    assert r.lookahead_i[0] == 3 and r.lookahead_i[1] == 106
    np.testing.assert_allclose(r.actuation[:3, 0], [-0.00035935558090650324, 0.22560092622792158,
                                                    -1.3435444691165728], rtol=1e-9)
    assert r.status[:3].tolist() == [1, 1, 2]


def test_golden_ellipse(golden_ellipse, ellipse):
    eng = _engine(ellipse)
    r = eng.pure_pursuit_batch(golden_ellipse["poses"], 0.8)
    _check(r, golden_ellipse, ellipse.shape[0])


def test_config2_sample_vs_oracle(ellipse):
    """10^5 poses on the 2k-waypoint track; a 4000-pose sample against the oracle, the rest by
    properties (projection lies on its segment, distance consistent, t in [0,1])."""
    rng = np.random.default_rng(1002)
    poses, _ = synth.random_poses(ellipse, 100000, rng)
    poses = poses[:, :3].copy()
    eng = _engine(ellipse)
    r = eng.pure_pursuit_batch(poses, 0.8)
    sub = rng.choice(100000, 4000, replace=False)
    o = co.pure_pursuit_batch(ellipse, poses[sub], 0.8, n_threads=co.max_threads())
    same = r.nearest_i[sub] == o["nearest_i"]
    assert (~same).mean() < 0.02
    # every index mismatch is a vertex tie, (k, t=1) == (k+1, t=0): neighbouring segments, the
    # same distance and the same projection -- a wrong block from the FP32 scan at a
    # hairpin would differ by more than one segment or in distance
    ties = np.nonzero(~same)[0]
    for k in ties:
        gi, oi = int(r.nearest_i[sub][k]), int(o["nearest_i"][k])
        assert abs(gi - oi) == 1, (k, gi, oi)
        assert abs(r.nearest[sub][k, 2] - o["nearest"][k, 2]) < 1e-9
        np.testing.assert_allclose(r.nearest[sub][k, :2], o["nearest"][k, :2], rtol=0, atol=1e-9)
        lo_t = r.nearest[sub][k, 3] if gi < oi else o["nearest"][k, 3]
        hi_t = o["nearest"][k, 3] if gi < oi else r.nearest[sub][k, 3]
        assert lo_t > 1.0 - 1e-7 and hi_t < 1e-7, (k, lo_t, hi_t)
    from tests import helpers as H
    H.record_parity("c2_nearest_index", {"poses": 4000, "index_mismatch": int(ties.size),
                                         "index_mismatch:vertex_tie": int(ties.size)})
    np.testing.assert_allclose(r.nearest[sub][same], o["nearest"][same], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r.actuation[sub][same], o["actuation"][same], rtol=1e-9, atol=1e-9)
    assert (r.status[sub][same] == o["status"][same]).all()
    # properties at full size
    i = r.nearest_i
    a, b = ellipse[i, :2], ellipse[i + 1, :2]
    t = r.nearest[:, 3]
    assert ((t >= 0) & (t <= 1)).all()
    proj = a + t[:, None] * (b - a)
    np.testing.assert_allclose(proj, r.nearest[:, :2], rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.hypot(*(poses[:, :2] - proj).T), r.nearest[:, 2], rtol=0, atol=1e-9)
    # no vertex of the track is closer than the reported nearest distance
    d_vert = np.min(np.hypot(poses[:2000, None, 0] - ellipse[None, :, 0],
                             poses[:2000, None, 1] - ellipse[None, :, 1]), axis=1)
    assert (r.nearest[:2000, 2] <= d_vert + 1e-9).all()


def test_free_functions_match_reference(golden_spielberg, golden_misc):
    from f1tenth_planning_b200 import utils
    g, m = golden_spielberg, golden_misc
    xy = g["waypoints"][:, :2]
    for k in (0, 1, 2, 10, 50):
        p, d, t, i = utils.nearest_point(g["poses"][k, :2], xy)
        assert i == g["nearest_i"][k]
        np.testing.assert_allclose([p[0], p[1], d, t], g["nearest"][k], rtol=1e-9, atol=1e-9)
    for k in range(0, 300, 7):
        p, i, t = utils.intersect_point(m["ip_points"][k], m["ip_radius"][k], xy, m["ip_t"][k],
                                        bool(m["ip_wrap"][k]))
        if m["ip_out"][k, 3] == 0:
            assert p is None and i is None and t is None
        else:
            assert i == m["ip_i"][k]
            np.testing.assert_allclose([p[0], p[1], t], m["ip_out"][k, :3], rtol=1e-9, atol=1e-9)
    for k in range(m["act_in"].shape[0]):
        r = m["act_in"][k]
        sp, st = utils.get_actuation(r[0], r[1:4], r[4:6], r[6], 0.33)
        np.testing.assert_allclose([sp, st], m["act_out"][k], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(utils.get_rotation_matrix(0.3), m["rot_03"])
    assert [utils.pi_2_pi(a) for a in m["angles"]] == m["pi_2_pi"].tolist()


def test_planner_class_api(golden_spielberg):
    from f1tenth_planning_b200 import PurePursuitPlanner
    g = golden_spielberg
    pl = PurePursuitPlanner(waypoints=g["waypoints"])
    for k in (0, 1, 2):
        steer, speed = pl.plan(*g["poses"][k], 0.8)
        np.testing.assert_allclose([steer, speed], g["actuation"][k], rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError):
        PurePursuitPlanner().plan(0.0, 0.0, 0.0, 0.8)
    with pytest.raises(ValueError):
        pl.plan(0.0, 0.0, 0.0, 0.8, waypoints=np.zeros((5, 2)))
    far = PurePursuitPlanner(waypoints=g["waypoints"])
    with pytest.warns(UserWarning):
        assert far.plan(500.0, 500.0, 0.0, 0.8) == (0.0, 0.0)


def test_stanley_front_axle_matches_reference(golden_spielberg):
    """front-axle mode of K1 vs the golden vectors of the reference Stanley / LQR controllers"""
    import os
    from f1tenth_planning_b200 import StanleyPlanner
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stanley.npz"))
    wp = golden_spielberg["waypoints"]
    pl = StanleyPlanner(waypoints=wp)
    front, idx = pl.front_axle_errors(g["states"], k_path=5.0)
    same = idx == g["target_index"]
    assert (~same).mean() < 0.02
    np.testing.assert_allclose(front[same], g["front"][same], rtol=1e-9, atol=1e-9)
    steer, speed = pl.plan_batch(g["states"], 5.0)
    np.testing.assert_allclose(np.stack([steer, speed], 1)[same], g["plan"][same], rtol=1e-9, atol=1e-9)
    s = g["states"][0]
    d, v = pl.plan(s[0], s[1], s[2], s[3], 5.0)
    np.testing.assert_allclose([d, v], g["plan"][0], rtol=1e-9, atol=1e-9)
    th_e, ef, ti, gv = pl.calc_theta_and_ef(s, wp)
    assert ti == g["target_index"][0] and abs(th_e - g["front"][0, 0]) < 1e-9
    with pytest.raises(ValueError):
        StanleyPlanner().plan(0.0, 0.0, 0.0, 1.0)
    with pytest.raises(ValueError):
        pl.plan(0.0, 0.0, 0.0, 1.0, waypoints=np.zeros((5, 3)))


@pytest.mark.parametrize("n_wp", [2, 3, 9, 32, 33, 34, 65, 255, 257, 1025])
@pytest.mark.parametrize("n_poses", [1, 31, 129, 300])
def test_scan_shapes_small_tracks_and_ragged_batches(n_wp, n_poses):
    """K1's scan splits the track's 32-segment blocks over 6..24 one-warp tasks per group of 128
    poses and packs four poses per lane: tracks with fewer blocks than parts, partial last blocks
    and batches that are not a multiple of 128 poses (or of the finish kernel's 32-pose warps) must
    give the oracle's answer."""
    from f1tenth_planning_b200.engine import Engine
    rng = np.random.default_rng(n_wp * 1000 + n_poses)
    phi = np.sort(rng.uniform(0.0, 1.5 * np.pi, n_wp))     # open arc, irregular spacing
    track = np.stack([12.0 * np.cos(phi), 7.0 * np.sin(phi), np.full(n_wp, 3.0),
                      np.zeros(n_wp), np.zeros(n_wp)], axis=1)
    eng = Engine()
    eng.set_track(track)
    poses = np.stack([rng.uniform(-14, 14, n_poses), rng.uniform(-9, 9, n_poses),
                      rng.uniform(-np.pi, np.pi, n_poses)], axis=1)
    r = eng.pure_pursuit_batch(poses, 0.9)
    o = co.pure_pursuit_batch(track, poses, 0.9)
    same = r.nearest_i == o["nearest_i"]
    # an FP32 scan may pick a neighbour when two segments tie to ~1e-7 m; the distance still agrees
    np.testing.assert_allclose(r.nearest[:, 2], o["nearest"][:, 2], rtol=1e-9, atol=1e-9)
    assert same.mean() >= 0.97
    np.testing.assert_allclose(r.nearest[same], o["nearest"][same], rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(r.status[same], o["status"][same])
    np.testing.assert_allclose(r.actuation[same], o["actuation"][same], rtol=1e-9, atol=1e-9)


def test_lqr_control_points_match_reference(golden_spielberg):
    """LQRPlanner.calc_control_points (lqr.py:60-102) on K1's front-axle mode vs the golden vectors
    minted from the reference controllers (the same five quantities Stanley's step shares)."""
    import os
    from f1tenth_planning_b200 import LQRPlanner
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stanley.npz"))
    wp = golden_spielberg["waypoints"]
    pl = LQRPlanner(waypoints=wp)
    cp, idx = pl.calc_control_points_batch(g["states"])
    same = idx == g["target_index"]
    assert (~same).mean() < 0.02
    np.testing.assert_allclose(cp[same], g["front"][same][:, :5], rtol=1e-9, atol=1e-9)
    one = pl.calc_control_points(g["states"][0], wp)
    np.testing.assert_allclose(one, g["front"][0, :5], rtol=1e-9, atol=1e-9)
    assert pl.vehicle_control_theta_e == one[0] and pl.vehicle_control_e_cog == one[1]
    with pytest.raises(NotImplementedError):
        pl.plan(0.0, 0.0, 0.0, 1.0)
    with pytest.raises(ValueError):
        LQRPlanner().calc_control_points(np.zeros(4))
    with pytest.raises(ValueError):
        pl.calc_control_points(np.zeros(4), waypoints=np.zeros((5, 4)))


def _oracle_check(eng, track, n, seed):
    poses, _ = synth.random_poses(track, n, np.random.default_rng(seed))
    r = eng.pure_pursuit_batch(poses[:, :3], 0.8)
    o = co.pure_pursuit_batch(track, poses[:, :3], 0.8)
    _check(r, o, track.shape[0])


def test_constant_table_ownership_and_fallback():
    """The scan's constant-memory copy of the line form belongs to the engine that uploaded its
    track last (one table per device); an older engine scans its global-memory copy, a track too
    long for the table never uses it, and a re-upload takes the table back.  All of them must give
    the oracle's answers -- including a last 32-segment block that is partial (padded entries)."""
    t_a = synth.ellipse_track(n=2000)            # 1999 segments: partial last block
    t_b = synth.ellipse_track(n=1500, a=50.0, b=30.0)
    t_c = synth.ellipse_track(n=3000)            # beyond the table: global-memory scan
    eng_a, eng_b = _engine(t_a), _engine(t_b)    # B owns the table now
    _oracle_check(eng_a, t_a, 1500, 1)           # A: global path
    _oracle_check(eng_b, t_b, 1500, 2)           # B: constant path
    eng_c = _engine(t_c)                         # too long: the table keeps B's entries
    _oracle_check(eng_c, t_c, 1500, 3)
    _oracle_check(eng_b, t_b, 700, 4)
    eng_a.set_track(t_a)                         # A takes the table back
    _oracle_check(eng_a, t_a, 1500, 5)
    _oracle_check(eng_b, t_b, 700, 6)            # B: global path, same answers


def test_constant_table_taken_over_by_another_thread():
    """One host thread keeps scanning its track while another keeps uploading a different one on
    the same device (each upload takes the constant-memory table over): every batch of the first
    thread must still be its own track's answer, whichever copy of the table it was scanned on."""
    import threading
    t_a = synth.ellipse_track(n=2000)
    t_b = synth.ellipse_track(n=1200, a=30.0, b=20.0)
    eng_a, eng_b = _engine(t_a), _engine(t_b)
    poses, _ = synth.random_poses(t_a, 4096, np.random.default_rng(11))
    eng_a.set_track(t_a)
    ref = eng_a.pure_pursuit_batch(poses[:, :3], 0.8)
    stop = threading.Event()
    errors = []

    def uploader():
        try:
            while not stop.is_set():
                eng_b.set_track(t_b)
        except Exception as e:   # pragma: no cover
            errors.append(e)

    th = threading.Thread(target=uploader)
    th.start()
    try:
        for k in range(60):
            if k % 7 == 0:
                eng_a.set_track(t_a)   # take the table back now and then
            r = eng_a.pure_pursuit_batch(poses[:, :3], 0.8)
            assert np.array_equal(r.nearest_i, ref.nearest_i)
            assert np.array_equal(r.nearest, ref.nearest)
            assert np.array_equal(r.actuation, ref.actuation)
    finally:
        stop.set()
        th.join()
    assert not errors
