"""Properties of the oracle's own stages (the ones with no reference code: cubic spiral, costs,
collision, argmin) -- checked against independent numerics (scipy) so that the oracle is not only
self-consistent."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co


def _kappa(q, u, p0=0.0, p3=0.0):
    p1, p2 = q[0], q[1]
    b1 = (-11 * p0 + 18 * p1 - 9 * p2 + 2 * p3) / 2
    b2 = (18 * p0 - 45 * p1 + 36 * p2 - 9 * p3) / 2
    b3 = (-9 * p0 + 27 * p1 - 27 * p2 + 9 * p3) / 2
    return p0 + b1 * u + b2 * u ** 2 + b3 * u ** 3


@pytest.mark.parametrize("goal", [(1.0, 0.2, 0.1), (2.5, -0.8, -0.4), (3.5, 1.2, 0.6), (0.8, 0.0, 0.0)])
def test_spiral_hits_goal_and_matches_ode(goal):
    q, st = co.spiral(goal, n_newton=12, m=100)
    np.testing.assert_allclose(st[-1, :3], goal, atol=2e-5)       # endpoint hits the goal
    assert st[0].tolist() == [0, 0, 0, 0]
    # knots: kappa(1/3) = p1, kappa(2/3) = p2, kappa(0) = kappa(1) = 0
    assert abs(_kappa(q, 1 / 3) - q[0]) < 1e-12 and abs(_kappa(q, 2 / 3) - q[1]) < 1e-12
    # independent integration of x' = cos th, y' = sin th, th' = kappa(s / s_f)
    sf = q[2]
    sol = solve_ivp(lambda s, z: [np.cos(z[2]), np.sin(z[2]), _kappa(q, s / sf)], [0, sf], [0, 0, 0],
                    t_eval=np.linspace(0, sf, 100), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(st[:, 0], sol.y[0], atol=5e-7)
    np.testing.assert_allclose(st[:, 1], sol.y[1], atol=5e-7)
    np.testing.assert_allclose(st[:, 2], sol.y[2], atol=1e-8)
    np.testing.assert_allclose(st[:, 3], _kappa(q, np.linspace(0, 1, 100)), atol=1e-12)


def test_lut_cells_solve_their_own_goal():
    lut = co.lut_build()
    assert lut.shape == (20, 21, 9, 4)
    assert lut[..., 3].mean() > 0.9
    xs, ys, ts = np.linspace(0.2, 4.0, 20), np.linspace(-2, 2, 21), np.linspace(-np.pi / 2, np.pi / 2, 9)
    rng = np.random.default_rng(0)
    for _ in range(40):
        i, j, k = rng.integers(20), rng.integers(21), rng.integers(9)
        if lut[i, j, k, 3] != 1:
            continue
        q, st = co.spiral((xs[i], ys[j], ts[k]), seed=lut[i, j, k, :3].astype(np.float64), n_newton=0)
        np.testing.assert_allclose(st[-1, :3], (xs[i], ys[j], ts[k]), atol=1e-4)


def _world(track, cid, **kw):
    la, wd = synth.goal_grid(cid)
    return co.World_(track, la, wd, **kw)


def test_sampler_goals_are_vehicle_frame_offsets_of_raceline_points(ellipse):
    w = _world(ellipse, 1)
    cfg = co.default_config(window=0, kappa_max=0.0)
    pose = np.array([ellipse[100, 0] + 0.1, ellipse[100, 1] - 0.2, ellipse[100, 3] + 0.05, 4.0])
    o = co.plan(cfg, w, pose)
    la, wd = synth.goal_grid(1)
    g = o["goals"].reshape(4, 7, 3)
    c, s = np.cos(pose[2]), np.sin(pose[2])
    gm = np.stack([pose[0] + c * g[..., 0] - s * g[..., 1], pose[1] + s * g[..., 0] + c * g[..., 1]], -1)
    for j in range(4):
        centre = gm[j, 3]                               # width 0
        d = np.hypot(*(ellipse[:, :2] - centre).T)
        assert d.min() < 1e-9                           # the centre is a raceline waypoint
        r = np.hypot(*(centre - pose[:2]))
        assert abs(r - la[j]) < 0.3                     # the segment-start waypoint near the circle
        lat = np.hypot(*(gm[j] - centre).T)
        np.testing.assert_allclose(lat, np.abs(wd), atol=1e-9)   # offsets along the normal
    assert o["n_candidates"] == 28 and (o["flags"] & 1).all()


def test_cost_terms_against_numpy(ellipse, corridor):
    occ, origin, res = corridor
    w = _world(ellipse, 1, grid=occ, grid_origin=origin, grid_res=res)
    cfg = co.default_config(window=0, kappa_max=0.0)
    pose = np.array([ellipse[700, 0], ellipse[700, 1] + 0.15, ellipse[700, 3] - 0.03, 5.0])
    prev = np.linspace(0, 0.3, 100).astype(np.float32)
    w.set_prev(prev)
    o = co.plan(cfg, w, pose, want_states=True)
    st, terms = o["states"], o["terms"]
    np.testing.assert_allclose(terms[:, 0], 1.0 / o["params"][:, 2], rtol=1e-14)
    np.testing.assert_allclose(terms[:, 1], np.abs(st[:, :, 3]).max(axis=1), rtol=1e-14)
    np.testing.assert_allclose(terms[:, 2], np.abs(st[:, :, 3]).mean(axis=1), rtol=1e-13)
    sim = ((st[:, :85, 2] - prev[None, 5:90].astype(np.float64)) ** 2).sum(axis=1)
    np.testing.assert_allclose(terms[:, 3], sim, rtol=1e-12)
    # raceline deviation with W = N-1 is the reference nearest_point distance of every sample
    c, s = np.cos(pose[2]), np.sin(pose[2])
    for cand in (0, 13, 27):
        X = pose[0] + c * st[cand, :, 0] - s * st[cand, :, 1]
        Y = pose[1] + s * st[cand, :, 0] + c * st[cand, :, 1]
        d = [co.nearest_point([x, y], ellipse[:, :2])[1] for x, y in zip(X, Y)]
        np.testing.assert_allclose(terms[cand, 4], np.mean(d), rtol=1e-12)
    fin = np.isfinite(o["costs"])
    wts = np.array([cfg.weights[i] for i in range(5)])
    np.testing.assert_allclose(o["costs"][fin], terms[fin] @ wts, rtol=1e-13)
    assert o["best_idx"] == int(np.argmin(o["costs"]))       # first minimum


def test_collision_against_shapely_free_geometry(ellipse):
    """SAT against a brute-force point-sampling overlap test of the two rectangles."""
    w = _world(ellipse, 1)
    cfg = co.default_config(window=0, kappa_max=0.0)
    pose = np.array([ellipse[300, 0], ellipse[300, 1], ellipse[300, 3], 5.0])
    psi = ellipse[300, 3]
    rng = np.random.default_rng(3)
    checked = 0
    for _ in range(30):
        ahead, lat, dth = rng.uniform(0.3, 1.4), rng.uniform(-0.6, 0.6), rng.uniform(-0.5, 0.5)
        opp = np.array([[pose[0] + ahead * np.cos(psi) - lat * np.sin(psi),
                         pose[1] + ahead * np.sin(psi) + lat * np.cos(psi), psi + dth]])
        o = co.plan(cfg, w, pose, opp, want_states=True)
        c, s = np.cos(pose[2]), np.sin(pose[2])
        for cand in (3, 10, 24):
            st = o["states"][cand]
            hit = False
            for x, y, th in st[:, :3]:
                X, Y, TH = pose[0] + c * x - s * y, pose[1] + s * x + c * y, th + pose[2]
                hit |= _rect_overlap((X, Y, TH), opp[0])
            margin = o["margins"][cand, 0]
            if margin > 5e-3:   # the dense point test resolves ~2 mm
                assert bool(o["flags"][cand] & 2) == hit, (cand, margin)
                checked += 1
    assert checked > 40


def _rect_overlap(a, b, hl=0.29, hw=0.155, n=160):
    """dense point sampling of rectangle a's area tested against rectangle b (and vice versa)"""
    def pts(r):
        u, v = np.meshgrid(np.linspace(-hl, hl, n), np.linspace(-hw, hw, n // 2))
        c, s = np.cos(r[2]), np.sin(r[2])
        return r[0] + c * u - s * v, r[1] + s * u + c * v

    def inside(px, py, r):
        c, s = np.cos(r[2]), np.sin(r[2])
        dx, dy = px - r[0], py - r[1]
        return (np.abs(c * dx + s * dy) < hl) & (np.abs(-s * dx + c * dy) < hw)
    return bool(inside(*pts(a), b).any() or inside(*pts(b), a).any())


def test_grid_probe_convention(ellipse):
    """cell (row, col) covers [ox + col*res, ox + (col+1)*res); out of bounds is occupied."""
    occ = np.zeros((2000, 4000), np.uint8)
    origin, res = (-100.0, -50.0), 0.05
    w = _world(ellipse, 1, grid=occ, grid_origin=origin, grid_res=res)
    cfg = co.default_config(window=0, kappa_max=0.0)
    pose = np.array([ellipse[0, 0], ellipse[0, 1], ellipse[0, 3], 5.0])
    o = co.plan(cfg, w, pose)
    assert not (o["flags"] & 4).any()
    # occupy the cell under the car's centre at the start pose -> every candidate collides
    col, row = int(np.floor((pose[0] - origin[0]) / res)), int(np.floor((pose[1] - origin[1]) / res))
    occ2 = occ.copy(); occ2[row, col] = 1
    w2 = _world(ellipse, 1, grid=occ2, grid_origin=origin, grid_res=res)
    o2 = co.plan(cfg, w2, pose)
    valid = (o2["flags"] & 1) != 0
    assert ((o2["flags"][valid] & 4) != 0).all() and o2["no_feasible"]
    assert o2["best_idx"] == 0
    # a grid that does not cover the car at all: out of bounds = occupied
    w3 = _world(ellipse, 1, grid=np.zeros((10, 10), np.uint8), grid_origin=(0.0, 0.0), grid_res=res)
    o3 = co.plan(cfg, w3, pose)
    assert ((o3["flags"][(o3["flags"] & 1) != 0] & 4) != 0).all()


def test_shard_concatenation_equals_unsharded(ellipse, corridor):
    """multi-GPU partitioning logic on the CPU: candidate shards and scenario shards."""
    occ, origin, res = corridor
    la, wd = np.linspace(0.6, 3.0, 12), np.linspace(-1, 1, 10)
    w = co.World_(ellipse, la, wd, grid=occ, grid_origin=origin, grid_res=res)
    cfg = co.default_config()
    poses, opp, n_opp = synth.scenario_batch(ellipse, 6, 4, 2)
    full = co.plan(cfg, w, poses[0], opp[0, :n_opp[0]])
    C = 120
    best = (np.inf, C)
    for g in range(4):
        part = co.plan(cfg, w, poses[0], opp[0, :n_opp[0]], c_begin=g * 30, c_end=(g + 1) * 30)
        assert np.array_equal(part["costs"][g * 30:(g + 1) * 30], full["costs"][g * 30:(g + 1) * 30])
        best = min(best, (part["best_cost"], part["best_idx"]))
    assert best[1] == full["best_idx"]
    whole = co.plan_batch(cfg, w, poses, opp, n_opp)
    halves = [co.plan_batch(cfg, w, poses[i::2], opp[i::2], n_opp[i::2]) for i in range(2)]
    for i in range(2):
        assert np.array_equal(whole["best_idx"][i::2], halves[i]["best_idx"])
        assert np.array_equal(whole["costs"][i::2], halves[i]["costs"])


@pytest.mark.parametrize("goal", [(1.0, 0.2, 0.1), (2.5, -0.8, -0.4), (3.5, 1.2, 0.9), (1.0, 1.0, 0.0),
                                  (2.0, 0.0, 0.0), (0.5, -0.6, -1.2), (3.0, 0.5, -0.3)])
def test_g1_clothoid_against_ode(goal):
    """G1 Hermite clothoid (what lattice_planner.py:196 asks pyclothoids for): linear curvature,
    reaches the goal pose, matches an independent ODE integration."""
    kdl, st, ok = co.clothoid(goal, n_newton=10, m=100)
    assert ok
    k0, dk, L = kdl
    np.testing.assert_allclose(st[-1, :3], goal, atol=5e-6)
    np.testing.assert_allclose(st[:, 3], k0 + dk * np.linspace(0, L, 100), atol=1e-12)   # linear kappa
    sol = solve_ivp(lambda s, z: [np.cos(z[2]), np.sin(z[2]), k0 + dk * s], [0, L], [0, 0, 0],
                    t_eval=np.linspace(0, L, 100), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(st[:, 0], sol.y[0], atol=1e-6)
    np.testing.assert_allclose(st[:, 1], sol.y[1], atol=1e-6)
    np.testing.assert_allclose(st[:, 2], sol.y[2], atol=1e-9)
    # the reference's own example goal, G1Hermite(0,0,0,1,1,0) (test_pyclothoids.py:23): symmetric
    if goal == (1.0, 1.0, 0.0):
        np.testing.assert_allclose(st[:, 3], -st[::-1, 3], atol=1e-9)


@pytest.mark.parametrize("R", [0.8, 2.0, 5.0, 40.0])
@pytest.mark.parametrize("phi", [0.05, 0.3, 1.0, np.pi / 2, -0.7, -1.4])
def test_g1_clothoid_known_answers_circular_arcs(R, phi):
    """Analytic known answers of G1 Hermite interpolation (the boundary lattice_planner.py:196
    crosses into pyclothoids): the pose reached after turning by phi on a circle of radius R is
    joined by that arc -- kappa = sign(phi)/R, no curvature rate, length R |phi|."""
    goal = (R * np.sin(abs(phi)), np.sign(phi) * R * (1 - np.cos(phi)), phi)
    kdl, st, ok = co.clothoid(goal, n_newton=10, m=100)
    assert ok
    # tolerances: the oracle integrates with composite Simpson on Q = 32 intervals
    np.testing.assert_allclose(kdl[0], np.sign(phi) / R, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(kdl[1], 0.0, atol=1e-6 / R ** 2)
    np.testing.assert_allclose(kdl[2], R * abs(phi), rtol=1e-7)
    s = np.linspace(0, R * abs(phi), 100)
    np.testing.assert_allclose(st[:, 0], R * np.sin(s / R), atol=1e-6 * max(R, 1))
    np.testing.assert_allclose(st[:, 1], np.sign(phi) * R * (1 - np.cos(s / R)), atol=1e-6 * max(R, 1))
    np.testing.assert_allclose(st[:, 2], np.sign(phi) * s / R, atol=1e-6)


@pytest.mark.parametrize("L", [0.2, 1.0, 4.0])
def test_g1_clothoid_known_answer_straight_line(L):
    kdl, st, ok = co.clothoid((L, 0.0, 0.0), n_newton=10, m=50)
    assert ok
    np.testing.assert_allclose(kdl, [0.0, 0.0, L], atol=1e-12)
    np.testing.assert_allclose(st[:, 0], np.linspace(0, L, 50), atol=1e-12)
    assert np.abs(st[:, 1:]).max() < 1e-12


def test_g1_clothoid_mirror_symmetry():
    """Reflecting the goal about the x axis reflects the clothoid (y, theta, kappa change sign)."""
    for g in [(1.0, 1.0, 0.0), (2.5, -0.8, -0.4), (3.0, 0.5, -0.3), (0.5, -0.6, -1.2)]:
        k1, s1, ok1 = co.clothoid(g, n_newton=10, m=64)
        k2, s2, ok2 = co.clothoid((g[0], -g[1], -g[2]), n_newton=10, m=64)
        assert ok1 and ok2
        np.testing.assert_allclose(k2, [-k1[0], -k1[1], k1[2]], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(s2, s1 * np.array([1.0, -1.0, -1.0, -1.0]), atol=1e-9)


@pytest.mark.parametrize("R", [1.5, 4.0, 30.0])
@pytest.mark.parametrize("phi", [0.1, 0.6, 1.2, -0.5])
def test_spiral_known_answers_circular_arcs(R, phi):
    """Analytic known answers of the cubic-spiral boundary-value problem: with the
    end curvatures pinned to 1/R the spiral joining the two poses of a circular arc is the arc --
    every knot equals 1/R and s_f = R |phi|."""
    k = np.sign(phi) / R
    goal = (R * np.sin(abs(phi)), np.sign(phi) * R * (1 - np.cos(phi)), phi)
    q, st = co.spiral(goal, p0=k, p3=k, n_newton=12, seed=[0.0, 0.0, np.hypot(goal[0], goal[1])], m=100)
    # tolerances: composite Simpson on Q = 32 intervals inside the Newton residual
    np.testing.assert_allclose(q[:2], [k, k], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(q[2], R * abs(phi), rtol=1e-7)
    s = np.linspace(0, R * abs(phi), 100)
    np.testing.assert_allclose(st[:, 0], R * np.sin(s / R), atol=1e-6 * max(R, 1))
    np.testing.assert_allclose(st[:, 1], np.sign(phi) * R * (1 - np.cos(s / R)), atol=1e-6 * max(R, 1))
    np.testing.assert_allclose(st[:, 3], k, rtol=1e-5)


def test_spiral_known_answer_straight_line_and_mirror():
    q, st = co.spiral((2.0, 0.0, 0.0), n_newton=8, m=40)
    np.testing.assert_allclose(q, [0.0, 0.0, 2.0], atol=1e-12)
    np.testing.assert_allclose(st[:, 0], np.linspace(0, 2.0, 40), atol=1e-12)
    assert np.abs(st[:, 1:]).max() < 1e-12
    for g in [(1.0, 0.2, 0.1), (2.5, -0.8, -0.4), (3.5, 1.2, 0.6)]:
        q1, s1 = co.spiral(g, n_newton=12, m=64)
        q2, s2 = co.spiral((g[0], -g[1], -g[2]), n_newton=12, m=64)
        np.testing.assert_allclose(q2, [-q1[0], -q1[1], q1[2]], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(s2, s1 * np.array([1.0, -1.0, -1.0, -1.0]), atol=1e-9)


def _clothoid_closed_form(k0, dk, s):
    """(x, y) of the clothoid theta(t) = k0 t + dk t^2 / 2 from the origin, at arc lengths s, in
    closed form through the Fresnel integrals S, C (scipy.special.fresnel; their argument
    convention is int cos(pi t^2 / 2)).  Completing the square: theta = a (t + b)^2 - a b^2 with
    a = dk / 2, b = k0 / dk."""
    from scipy.special import fresnel
    s = np.asarray(s, dtype=np.float64)
    if abs(dk) < 1e-14:
        if abs(k0) < 1e-14:
            return s.copy(), np.zeros_like(s)
        return np.sin(k0 * s) / k0, (1.0 - np.cos(k0 * s)) / k0
    a, b = 0.5 * dk, k0 / dk
    sg = 1.0 if a > 0 else -1.0
    scale = np.sqrt(np.pi / (2.0 * abs(a)))
    S1, C1 = fresnel((s + b) / scale)
    S0, C0 = fresnel(b / scale)
    Ic, Is = scale * (C1 - C0), sg * scale * (S1 - S0)     # int cos / sin of a (t + b)^2
    ph = -a * b * b
    return np.cos(ph) * Ic - np.sin(ph) * Is, np.sin(ph) * Ic + np.cos(ph) * Is


@pytest.mark.parametrize("goal", [(1.0, 1.0, 0.0), (2.0, 0.5, 0.3), (3.0, -1.0, -0.4), (0.8, 0.6, 1.2),
                                  (4.0, 0.0, 0.0), (2.5, 1.5, 1.0), (1.5, -1.2, -1.3), (3.5, 0.2, -0.2)])
def test_g1_clothoid_pinned_to_fresnel_closed_form(goal):
    """pyclothoids' Clothoid.G1Hermite evaluates the clothoid through Fresnel
    integrals; the oracle's generator=1 path is pinned to that closed form at 1e-12 -- both the
    solved (kappa0, dkappa, L), whose exact endpoint must be the goal, and every sampled state."""
    kdl, st, ok = co.clothoid(goal, n_newton=12, m=100)
    assert ok
    k0, dk, L = kdl
    s = np.linspace(0.0, L, 100)
    x, y = _clothoid_closed_form(k0, dk, s)
    scale = max(1.0, L)
    assert np.abs(st[:, 0] - x).max() < 1e-12 * scale and np.abs(st[:, 1] - y).max() < 1e-12 * scale
    np.testing.assert_allclose(st[:, 2], k0 * s + 0.5 * dk * s * s, rtol=0, atol=1e-13)
    # the G1 interpolation conditions, evaluated in closed form
    assert abs(x[-1] - goal[0]) < 1e-12 * scale and abs(y[-1] - goal[1]) < 1e-12 * scale
    end_th = k0 * L + 0.5 * dk * L * L
    assert abs(np.angle(np.exp(1j * (end_th - goal[2])))) < 1e-12
