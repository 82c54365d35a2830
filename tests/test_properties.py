"""Property tests (hypothesis) of the oracle: the invariants the domain offers, independent of
any stored vector."""
import numpy as np
from hypothesis import given, settings, strategies as st

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co

TRACK = synth.ellipse_track(n=400, a=20.0, b=9.0)
XY = np.ascontiguousarray(TRACK[:, :2])
LUT = co.lut_build()
XS, YS, TS = np.linspace(0.2, 4.0, 20), np.linspace(-2, 2, 21), np.linspace(-np.pi / 2, np.pi / 2, 9)


@settings(max_examples=60, deadline=None)
@given(st.floats(0.6, 3.8), st.floats(-1.0, 1.0), st.floats(-0.7, 0.7))
def test_spiral_from_lut_seed_hits_the_goal(gx, fy, gth):
    gy = fy * min(1.2, 0.6 * gx)          # the fan of goals a raceline-following sampler produces
    i = int(np.clip(np.floor((gx - 0.2) / 0.2 + 0.5), 0, 19))
    j = int(np.clip(np.floor((gy + 2.0) / 0.2 + 0.5), 0, 20))
    k = int(np.clip(np.floor((gth + np.pi / 2) / (np.pi / 8) + 0.5), 0, 8))
    q, stt = co.spiral((gx, gy, gth), seed=LUT[i, j, k, :3].astype(np.float64), n_newton=8, m=100)
    assert np.isfinite(stt).all() and q[2] > 0
    assert np.abs(stt[-1, :3] - (gx, gy, gth)).max() < 1e-5
    assert q[2] >= np.hypot(gx, gy) - 1e-9          # arc length is at least the chord
    # arc samples are equally spaced in arc length: consecutive points <= h apart
    h = q[2] / 99
    assert (np.hypot(np.diff(stt[:, 0]), np.diff(stt[:, 1])) <= h * (1 + 1e-9)).all()


@settings(max_examples=60, deadline=None)
@given(st.floats(-25, 25), st.floats(-12, 12))
def test_nearest_point_is_the_nearest(px, py):
    proj, dist, t, i = co.nearest_point([px, py], XY)
    assert 0 <= i < XY.shape[0] - 1 and 0.0 <= t <= 1.0
    a, b = XY[i], XY[i + 1]
    np.testing.assert_allclose(proj, a + t * (b - a), atol=1e-12)
    np.testing.assert_allclose(dist, np.hypot(px - proj[0], py - proj[1]), atol=1e-12)
    assert dist <= np.hypot(XY[:, 0] - px, XY[:, 1] - py).min() + 1e-12
    # dense sampling of every segment never finds anything closer
    tt = np.linspace(0, 1, 21)[None, :, None]
    pts = XY[:-1, None, :] + tt * (XY[1:, None, :] - XY[:-1, None, :])
    assert dist <= np.hypot(pts[..., 0] - px, pts[..., 1] - py).min() + 1e-9


@settings(max_examples=80, deadline=None)
@given(st.integers(0, 398), st.floats(0, 0.999), st.floats(-0.5, 0.5), st.floats(0.3, 3.0), st.booleans())
def test_intersect_point_is_the_first_crossing(seg, frac, lat, radius, wrap):
    a, b = XY[seg], XY[seg + 1]
    n = np.array([-(b - a)[1], (b - a)[0]]) / np.hypot(*(b - a))
    p = a + frac * (b - a) + lat * n
    q, i, t = co.intersect_point(p, radius, XY, seg + frac, wrap)
    if q is None:
        return
    assert 0.0 <= t <= 1.0
    N = XY.shape[0]
    s, e = XY[i % N], XY[(i + 1) % N] + 1e-6            # the reference's shifted segment end
    np.testing.assert_allclose(q, s + t * (e - s), atol=1e-12)
    assert abs(np.hypot(*(q - p)) - radius) < 1e-7      # on the circle
    if 0 <= i and i >= seg:                             # forward hit: nothing earlier crosses
        for k in range(seg + 1, i):
            ss, ee = XY[k], XY[k + 1] + 1e-6
            d = np.hypot(*(ss + np.linspace(0, 1, 50)[:, None] * (ee - ss) - p).T)
            assert ((d - radius) > 0).all() or ((d - radius) < 0).all()


@settings(max_examples=40, deadline=None)
@given(st.lists(st.sampled_from([0.5, 1.0, 2.0, float("inf")]), min_size=2, max_size=12))
def test_argmin_takes_the_first_minimum(costs):
    c = np.array(costs)
    first = int(np.argmin(c))
    assert c[first] == c.min() and (c[:first] > c[first]).all()


def test_seam_wraps_like_the_reference(golden_spielberg):
    """on Spielberg the last waypoint repeats the first: a pose just before the seam finds its
    lookahead point in the wrap loop"""
    wp = golden_spielberg["waypoints"]
    xy = wp[:, :2]
    p, d, t, i = co.nearest_point([0.0, -0.84], xy)
    assert i == xy.shape[0] - 2
    q, i2, t2 = co.intersect_point([0.0, -0.84], 0.8, xy, i + t, wrap=True)
    assert i2 == 3 and co.intersect_point([0.0, -0.84], 0.8, xy, i + t, wrap=False)[0] is None
