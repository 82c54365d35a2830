"""ctypes binding of the C oracle (oracle/c/f1o.c).  TEST INFRASTRUCTURE ONLY.

Mirrors the shapes of the product API so that parity tests read side by side.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

N_TERMS = 5
FLAG_VALID, FLAG_COLLIDE_OPP, FLAG_COLLIDE_MAP, FLAG_NO_CENTRE = 1, 2, 4, 8

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class Config(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_newton", C.c_int32), ("window", C.c_int32),
                ("n_shift", C.c_int32), ("n_cull", C.c_int32), ("literal_tracker", C.c_int32),
                ("use_goal_kappa", C.c_int32), ("generator", C.c_int32),
                ("prune_window", C.c_int32), ("collision_mode", C.c_int32),
                ("weights", C.c_double * N_TERMS), ("kappa_max", C.c_double),
                ("car_length", C.c_double), ("car_width", C.c_double),
                ("converge_tol", C.c_double), ("tracker_lookahead", C.c_double),
                ("wheelbase", C.c_double), ("max_reacquire", C.c_double)]


class World(C.Structure):
    _fields_ = [("wpts", _dp), ("n", C.c_int32), ("ncols", C.c_int32), ("grid", _bp),
                ("gh", C.c_int32), ("gw", C.c_int32), ("gox", C.c_double), ("goy", C.c_double),
                ("gres", C.c_double), ("lut", _fp), ("lut_dims", C.c_int32 * 3),
                ("pad0", C.c_int32), ("lut_ranges", C.c_double * 6), ("lookaheads", _dp),
                ("widths", _dp), ("n_lookaheads", C.c_int32), ("n_widths", C.c_int32),
                ("prev_theta", _fp), ("edt2", C.POINTER(C.c_uint16))]


class Result(C.Structure):
    _fields_ = [("steer", C.c_double), ("speed", C.c_double), ("best_idx", C.c_int32),
                ("no_feasible", C.c_int32), ("tracker_found", C.c_int32),
                ("n_candidates", C.c_int32), ("best_cost", C.c_double), ("best_traj", _dp),
                ("costs", _dp), ("terms", _dp), ("flags", _bp), ("goals", _dp), ("params", _dp),
                ("states", _dp), ("margins", _dp), ("best_traj_map", _dp)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        L.f1o_nearest_point.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _ip]
        L.f1o_nearest_point.restype = None
        L.f1o_intersect_point.argtypes = [_dp, C.c_double, _dp, C.c_int, C.c_int, C.c_double,
                                          C.c_int, _dp, _ip, _dp]
        L.f1o_intersect_point.restype = C.c_int
        L.f1o_get_actuation.argtypes = [C.c_double, _dp, _dp, C.c_double, C.c_double, _dp]
        L.f1o_get_actuation.restype = None
        L.f1o_pure_pursuit_batch.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_double,
                                             C.c_double, C.c_double, _dp, _ip, _dp, _ip, _dp, _ip,
                                             C.c_int]
        L.f1o_pure_pursuit_batch.restype = None
        L.f1o_lut_build.argtypes = [_ip, _dp, _fp, C.c_int]
        L.f1o_lut_build.restype = None
        L.f1o_spiral_solve.argtypes = [_dp, C.c_double, C.c_double, C.c_int, _dp]
        L.f1o_spiral_solve.restype = None
        L.f1o_spiral_sample.argtypes = [_dp, C.c_double, C.c_double, C.c_int, _dp]
        L.f1o_spiral_sample.restype = None
        L.f1o_plan.argtypes = [C.POINTER(Config), C.POINTER(World), _dp, _dp, C.c_int, _dp,
                               C.c_int, C.c_int, C.c_int, C.POINTER(Result)]
        L.f1o_plan.restype = C.c_int
        L.f1o_plan_batch.argtypes = [C.POINTER(Config), C.POINTER(World), _dp, _dp, _ip, C.c_int,
                                     C.c_int, _ip, _dp, _dp, _dp, _bp, _dp, C.c_int]
        L.f1o_plan_batch.restype = C.c_int64
        L.f1o_default_config.argtypes = [C.POINTER(Config)]
        L.f1o_default_config.restype = None
        L.f1o_max_threads.restype = C.c_int
        L.f1o_clothoid_g1.argtypes = [_dp, C.c_int, _dp]
        L.f1o_clothoid_g1.restype = C.c_int
        L.f1o_clothoid_sample.argtypes = [_dp, C.c_int, _dp]
        L.f1o_clothoid_sample.restype = None
        L.f1o_front_axle.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_double, C.c_double, _dp, _ip]
        L.f1o_front_axle.restype = None
        L.f1o_collide_f32.argtypes = [_fp, _fp, C.c_int, C.c_int, _fp, C.c_int, _fp, _ip, _bp,
                                      C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _bp]
        L.f1o_collide_f32.restype = None
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_config(**kw):
    cfg = Config()
    lib().f1o_default_config(C.byref(cfg))
    for k, v in kw.items():
        if k == "weights":
            for i, w in enumerate(v):
                cfg.weights[i] = float(w)
        else:
            setattr(cfg, k, v)
    return cfg


LUT_DIMS = (20, 21, 9)
LUT_RANGES = (0.2, 4.0, -2.0, 2.0, -np.pi / 2, np.pi / 2)
_lut_cache = {}


def lut_build(dims=LUT_DIMS, ranges=LUT_RANGES, n_threads=None):
    key = (tuple(dims), tuple(float(r) for r in ranges))
    if key not in _lut_cache:
        d = np.asarray(dims, dtype=np.int32)
        r = np.asarray(ranges, dtype=np.float64)
        out = np.zeros(tuple(dims) + (4,), dtype=np.float32)
        lib().f1o_lut_build(d.ctypes.data_as(_ip), _d(r), out.ctypes.data_as(_fp),
                            n_threads or lib().f1o_max_threads())
        _lut_cache[key] = out
    return _lut_cache[key]


def nearest_point(point, trajectory):
    """utils/utils.py:37-67 -> (proj(2,), dist, t, i)"""
    point = _as_f64(point)
    traj = np.asarray(trajectory, dtype=np.float64)
    base = _as_f64(traj)
    proj = np.zeros(2)
    dist = C.c_double()
    t = C.c_double()
    i = C.c_int32()
    lib().f1o_nearest_point(_d(point), _d(base), base.shape[0], base.shape[1], _d(proj),
                            C.byref(dist), C.byref(t), C.byref(i))
    return proj, dist.value, t.value, i.value


def intersect_point(point, radius, trajectory, t=0.0, wrap=False):
    """utils/utils.py:69-151 -> (p(2,)|None, i|None, t|None)"""
    point = _as_f64(point)
    base = _as_f64(trajectory)
    p = np.zeros(2)
    oi = C.c_int32()
    ot = C.c_double()
    found = lib().f1o_intersect_point(_d(point), float(radius), _d(base), base.shape[0],
                                      base.shape[1], float(t), int(bool(wrap)), _d(p),
                                      C.byref(oi), C.byref(ot))
    if not found:
        return None, None, None
    return p, oi.value, ot.value


def get_actuation(pose_theta, lookahead_point, position, lookahead_distance, wheelbase):
    """utils/utils.py:153-161 -> (speed, steer)"""
    lp = _as_f64(lookahead_point)
    pos = _as_f64(position)
    out = np.zeros(2)
    lib().f1o_get_actuation(float(pose_theta), _d(lp), _d(pos), float(lookahead_distance),
                            float(wheelbase), _d(out))
    return out[0], out[1]


def pure_pursuit_batch(wpts, poses, lookahead, wheelbase=0.33, max_reacquire=20.0, n_threads=1):
    wpts = _as_f64(wpts)
    poses = _as_f64(poses).reshape(-1, 3)
    b = poses.shape[0]
    out = dict(nearest=np.zeros((b, 4)), nearest_i=np.zeros(b, np.int32),
               lookahead=np.zeros((b, 4)), lookahead_i=np.zeros(b, np.int32),
               actuation=np.zeros((b, 2)), status=np.zeros(b, np.int32))
    lib().f1o_pure_pursuit_batch(_d(wpts), wpts.shape[0], wpts.shape[1], _d(poses), b,
                                 float(lookahead), float(wheelbase), float(max_reacquire),
                                 _d(out["nearest"]), out["nearest_i"].ctypes.data_as(_ip),
                                 _d(out["lookahead"]), out["lookahead_i"].ctypes.data_as(_ip),
                                 _d(out["actuation"]), out["status"].ctypes.data_as(_ip),
                                 int(n_threads))
    return out


def spiral(goal, p0=0.0, p3=0.0, n_newton=8, seed=None, m=100):
    goal = _as_f64(goal)
    q = _as_f64(seed if seed is not None else [0.0, 0.0, float(np.hypot(goal[0], goal[1]))]).copy()
    lib().f1o_spiral_solve(_d(goal), p0, p3, n_newton, _d(q))
    st = np.zeros((m, 4))
    lib().f1o_spiral_sample(_d(q), p0, p3, m, _d(st))
    return q, st


_edt_cache = {}
EDT_R = 24   # the transform is exact up to this many cells (CLEAR_R of the device library)


def edt2_of(grid):
    """Squared Euclidean distance, in cells between cell centres, from every cell to the nearest
    occupied or out-of-bounds cell (uint16; exact up to EDT_R^2, EDT_R^2 + 1 = farther) -- by
    scipy.ndimage.distance_transform_edt, independent of the device's separable two-pass build."""
    from scipy import ndimage
    g = np.asarray(grid) != 0
    pad = EDT_R + 1
    big = np.ones((g.shape[0] + 2 * pad, g.shape[1] + 2 * pad), bool)
    big[pad:-pad, pad:-pad] = g
    d = ndimage.distance_transform_edt(~big)[pad:-pad, pad:-pad]
    d2 = np.rint(d * d).astype(np.int64)
    return np.ascontiguousarray(np.minimum(d2, EDT_R * EDT_R + 1).astype(np.uint16))


class World_:
    """Keeps the numpy buffers alive behind a C f1o_world."""

    def __init__(self, wpts, lookaheads, widths, grid=None, grid_origin=(0.0, 0.0), grid_res=0.05,
                 lut=None, lut_ranges=LUT_RANGES, prev_theta=None, use_lut=True):
        self.wpts = _as_f64(wpts)
        self.lookaheads = _as_f64(lookaheads)
        self.widths = _as_f64(widths)
        self.grid = None if grid is None else np.ascontiguousarray(grid, dtype=np.uint8)
        if lut is None and use_lut:
            lut = lut_build()
        self.lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.float32)
        self.prev_theta = (None if prev_theta is None
                           else np.ascontiguousarray(prev_theta, dtype=np.float32))
        w = World()
        w.wpts = _d(self.wpts)
        w.n, w.ncols = self.wpts.shape
        if self.grid is not None:
            w.grid = self.grid.ctypes.data_as(_bp)
            w.gh, w.gw = self.grid.shape
        w.gox, w.goy, w.gres = float(grid_origin[0]), float(grid_origin[1]), float(grid_res)
        if self.lut is not None:
            w.lut = self.lut.ctypes.data_as(_fp)
            for i in range(3):
                w.lut_dims[i] = self.lut.shape[i]
            for i in range(6):
                w.lut_ranges[i] = float(lut_ranges[i])
        w.lookaheads = _d(self.lookaheads)
        w.widths = _d(self.widths)
        w.n_lookaheads = self.lookaheads.shape[0]
        w.n_widths = self.widths.shape[0]
        if self.prev_theta is not None:
            w.prev_theta = self.prev_theta.ctypes.data_as(_fp)
        self.edt2 = None   # built on first use by a collision_mode = 1 query (need_edt)
        self.c = w

    def need_edt(self):
        if self.edt2 is None and self.grid is not None:
            import zlib
            key = (self.grid.shape, zlib.crc32(self.grid))
            if key not in _edt_cache:
                _edt_cache[key] = edt2_of(self.grid)
            self.edt2 = _edt_cache[key]
            self.c.edt2 = self.edt2.ctypes.data_as(C.POINTER(C.c_uint16))

    def set_prev(self, prev_theta):
        self.prev_theta = (None if prev_theta is None
                           else np.ascontiguousarray(prev_theta, dtype=np.float32))
        self.c.prev_theta = (self.prev_theta.ctypes.data_as(_fp) if self.prev_theta is not None
                             else _fp())

    @property
    def n_candidates(self):
        return self.c.n_lookaheads * self.c.n_widths


def plan(cfg, world, pose, opp=None, goals=None, c_begin=0, c_end=0, want_states=False):
    if cfg.collision_mode == 1:
        world.need_edt()
    pose = _as_f64(pose)
    M = cfg.n_samples
    if goals is not None:
        goals = _as_f64(goals).reshape(-1, 3)
        Cn = goals.shape[0]
    else:
        Cn = world.n_candidates
    opp_a = None if opp is None else _as_f64(opp).reshape(-1, 3)
    k = 0 if opp_a is None else opp_a.shape[0]
    res = Result()
    out = dict(best_traj=np.zeros((M, 4)), costs=np.full(Cn, np.inf), terms=np.zeros((Cn, N_TERMS)),
               flags=np.zeros(Cn, np.uint8), goals=np.zeros((Cn, 3)), params=np.zeros((Cn, 4)),
               margins=np.full((Cn, 2), np.inf), best_traj_map=np.zeros((M, 4)))
    res.best_traj_map = _d(out["best_traj_map"])
    if want_states:
        out["states"] = np.zeros((Cn, M, 4))
        res.states = _d(out["states"])
    res.best_traj = _d(out["best_traj"])
    res.costs = _d(out["costs"])
    res.terms = _d(out["terms"])
    res.flags = out["flags"].ctypes.data_as(_bp)
    res.goals = _d(out["goals"])
    res.params = _d(out["params"])
    res.margins = _d(out["margins"])
    lib().f1o_plan(C.byref(cfg), C.byref(world.c), _d(pose), _d(opp_a) if k else _dp(), k,
                   _d(goals) if goals is not None else _dp(), Cn if goals is not None else 0,
                   int(c_begin), int(c_end), C.byref(res))
    out.update(steer=res.steer, speed=res.speed, best_idx=res.best_idx,
               no_feasible=bool(res.no_feasible), tracker_found=bool(res.tracker_found),
               best_cost=res.best_cost, n_candidates=res.n_candidates)
    return out


def plan_batch(cfg, world, poses, opp=None, n_opp=None, n_threads=1, want_traj=True,
               want_costs=True):
    if cfg.collision_mode == 1:
        world.need_edt()
    poses = _as_f64(poses).reshape(-1, 4)
    S = poses.shape[0]
    Cn = world.n_candidates
    M = cfg.n_samples
    max_opp = 0
    opp_a = None
    if opp is not None:
        opp_a = _as_f64(opp)
        max_opp = opp_a.shape[1]
    n_opp_a = None if n_opp is None else np.ascontiguousarray(n_opp, dtype=np.int32)
    out = dict(best_idx=np.zeros(S, np.int32), best_cost=np.zeros(S),
               steer_speed=np.zeros((S, 2)))
    if want_traj:
        out["best_traj"] = np.zeros((S, M, 4))
    if want_costs:
        out["costs"] = np.zeros((S, Cn))
        out["flags"] = np.zeros((S, Cn), np.uint8)
    n = lib().f1o_plan_batch(
        C.byref(cfg), C.byref(world.c), _d(poses), _d(opp_a) if opp_a is not None else _dp(),
        n_opp_a.ctypes.data_as(_ip) if n_opp_a is not None else _ip(), S, max_opp,
        out["best_idx"].ctypes.data_as(_ip), _d(out["best_cost"]),
        _d(out["best_traj"]) if want_traj else _dp(), _d(out["costs"]) if want_costs else _dp(),
        out["flags"].ctypes.data_as(_bp) if want_costs else _bp(), _d(out["steer_speed"]),
        int(n_threads))
    out["n_evaluated"] = int(n)
    return out


def max_threads():
    return int(lib().f1o_max_threads())


def collide_f32(states, headings, opp_local, n_opp, grid_xf, grid_i0, grid, half_l, half_w, rc2):
    """float32 mirror of the device collision predicate on the device's own states -> flags [C]"""
    st = np.ascontiguousarray(states, dtype=np.float32)
    hd = np.ascontiguousarray(headings, dtype=np.float32)
    op = np.ascontiguousarray(opp_local, dtype=np.float32)
    xf = np.ascontiguousarray(grid_xf, dtype=np.float32)
    i0 = np.ascontiguousarray(grid_i0, dtype=np.int32)
    Cn, M = st.shape[0], st.shape[1]
    out = np.zeros(Cn, np.uint8)
    g = None if grid is None else np.ascontiguousarray(grid, dtype=np.uint8)
    lib().f1o_collide_f32(st.ctypes.data_as(_fp), hd.ctypes.data_as(_fp), Cn, M,
                          op.ctypes.data_as(_fp), int(n_opp), xf.ctypes.data_as(_fp),
                          i0.ctypes.data_as(_ip), g.ctypes.data_as(_bp) if g is not None else _bp(),
                          g.shape[0] if g is not None else 0, g.shape[1] if g is not None else 0,
                          np.float32(half_l), np.float32(half_w), np.float32(rc2),
                          out.ctypes.data_as(_bp))
    return out


def front_axle_batch(wpts, states, wheelbase=0.33, k_path=5.0):
    """stanley.py:57-112 / lqr.py:60-102 for B states [B,4] -> (front [B,6], target_index [B])"""
    wpts = _as_f64(wpts)
    st = _as_f64(states).reshape(-1, 4)
    out = np.zeros((st.shape[0], 6))
    idx = np.zeros(st.shape[0], np.int32)
    for k in range(st.shape[0]):
        i = C.c_int32()
        lib().f1o_front_axle(_d(wpts), wpts.shape[0], wpts.shape[1], _d(st[k]), float(wheelbase),
                             float(k_path), _d(out[k]), C.byref(i))
        idx[k] = i.value
    return out, idx


def clothoid(goal, n_newton=8, m=100):
    """G1 Hermite clothoid (0,0,0) -> goal: ((kappa0, dkappa, L), states [m,4], converged)"""
    goal = _as_f64(goal)
    kdl = np.zeros(3)
    ok = lib().f1o_clothoid_g1(_d(goal), int(n_newton), _d(kdl))
    st = np.zeros((m, 4))
    lib().f1o_clothoid_sample(_d(kdl), m, _d(st))
    return kdl, st, bool(ok)
