"""Lattice pipeline parity (through the C-ABI) vs the float64 oracle on the same seeded inputs."""
import numpy as np
import pytest

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _run(eng, cfg, world, pose, opp, verbose=False):
    d = eng.plan(pose, opp, update_prev=False, want_states=True)
    o = co.plan(cfg, world, pose, opp, want_states=True)
    return d, o, H.compare_plan(d, o, cfg, verbose=verbose)


def test_lut_matches_oracle():
    from f1tenth_planning_b200.engine import Engine
    eng = Engine()
    lut, ranges = eng.get_lut()
    ref = co.lut_build()
    assert lut.shape == ref.shape
    assert (lut[..., 3] == ref[..., 3]).mean() > 0.999
    ok = (lut[..., 3] == 1) & (ref[..., 3] == 1)
    assert ok.mean() > 0.5
    np.testing.assert_allclose(lut[ok], ref[ok], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("seed", [1001, 7, 8, 9])
def test_config1_default_grid(ellipse, seed):
    """C1: 4x7 goals, M=100, 1 opponent, exact raceline window (W = N-1), kappa limit off so that
    most of the 28 candidates are feasible."""
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd, window=0, kappa_max=0.0)
    pose, opp = H.scenario(ellipse, seed, 1)
    d, o, st = _run(eng, cfg, world, pose, opp)
    assert st["n_both_valid"] >= 10


def test_config1_kappa_limit_and_grid(ellipse, corridor):
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, window=0)
    for seed in (11, 12, 13):
        pose, opp = H.scenario(ellipse, seed, 1)
        _run(eng, cfg, world, pose, opp)


def test_config3_4096_candidates(ellipse):
    """C3: 64x64 goals, M=100, 8 opponents + occupancy grid (1 m half-width corridor so that the
    outer goal columns hit the walls), W=128."""
    la, wd = synth.goal_grid(3)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=synth.corridor_grid(half_width=1.0))
    pose, opp = H.scenario(ellipse, 1003, 8)
    d, o, st = _run(eng, cfg, world, pose, opp, verbose=True)
    assert st["n_both_valid"] > 1000
    assert st["collide_map_count"] > 0 and st["collide_opp_count"] > 0
    assert st["valid_mismatch"] <= 4 and st["collide_map_mismatch"] + st["collide_opp_mismatch"] <= 8
    H.record_parity("c3_device_lut_seed", st)
    # the same query with the oracle seeded from its OWN float64 LUT (nearest cell), not the
    # device's: the converged spirals, costs, flags and the argmin must still agree
    eng2, cfg2, world2 = H.make_pair(ellipse, la, wd, grid=synth.corridor_grid(half_width=1.0),
                                     use_device_lut=False)
    for seed in (1003, 1013):
        pose, opp = H.scenario(ellipse, seed, 8)
        d2 = eng2.plan(pose, opp, update_prev=False, want_states=True, want_map=True)
        o2 = co.plan(cfg2, world2, pose, opp, want_states=True)
        st2 = H.compare_plan(d2, o2, cfg2, record="c3_oracle_own_lut_seed%d" % seed)
        assert st2["valid_mismatch"] <= 4 and st2["collide_map_mismatch"] + st2["collide_opp_mismatch"] <= 8


def test_oracle_lut_seed_gives_same_answers(ellipse):
    """Seeding Newton from the oracle's own float64 LUT instead of the device LUT must not move
    the converged spirals beyond the tolerance."""
    la, wd = synth.goal_grid(3)
    eng, cfg, world = H.make_pair(ellipse, la[::4], wd[::4], use_device_lut=False)
    pose, opp = H.scenario(ellipse, 5, 4)
    _run(eng, cfg, world, pose, opp)


def test_similarity_term_and_prev_path(ellipse):
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd, window=0, kappa_max=0.0)
    pose, opp = H.scenario(ellipse, 21, 1)
    d0 = eng.plan(pose, opp, update_prev=True, want_states=True)
    assert (d0.terms[:, 3] == 0).all()
    world.set_prev(d0.best_traj[:, 2])
    pose2 = pose + np.array([0.05, 0.02, 0.01, 0.0])
    d1 = eng.plan(pose2, opp, update_prev=False, want_states=True)
    o1 = co.plan(cfg, world, pose2, opp, want_states=True)
    H.compare_plan(d1, o1, cfg)
    assert (d1.terms[(d1.flags & 1) != 0, 3] > 0).any()


def test_m200_and_other_sample_counts(ellipse, corridor):
    la, wd = np.linspace(0.6, 3.5, 12), np.linspace(-1.0, 1.0, 9)
    for m in (200, 64, 33, 128, 256):
        eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, n_samples=m,
                                      n_shift=2, n_cull=3)
        pose, opp = H.scenario(ellipse, 30 + m, 3)
        _run(eng, cfg, world, pose, opp)


def test_spielberg_track(golden_spielberg):
    wp = golden_spielberg["waypoints"]
    la, wd = np.linspace(0.8, 3.0, 8), np.linspace(-0.8, 0.8, 9)
    eng, cfg, world = H.make_pair(wp, la, wd)
    for pose in ([0.0, -0.84, 3.40, 4.0], [-20.0, -6.0, 3.4, 5.0]):
        d, o, st = _run(eng, cfg, world, np.array(pose), None)
        assert st["n_both_valid"] > 20


def test_no_feasible_and_off_track(ellipse):
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd)
    pose = np.array([0.0, 0.0, 0.3, 3.0])   # centre of the oval: no lookahead intersection
    d, o, st = _run(eng, cfg, world, pose, None)
    assert d.no_feasible and d.best_idx == 0 and not np.isfinite(d.costs).any()
    assert ((d.flags & 8) != 0).all()


def test_lookahead_circles_clear_of_the_raceline(ellipse, corridor):
    """Poses 0.45 .. 0.95 m off the raceline with lookaheads 0.4 .. 1.0 m: the rows whose circle
    stays clear of the raceline skip the sampler's track scan (only the closing segment is
    tested) and must come out exactly like the oracle's full intersect_point scan -- no centre,
    flag 8 on all their candidates -- while the rows that do cross keep their goals.  Single
    queries and the batch sampler (K1's distance) take the same shortcut."""
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor)
    i0 = [10, 500, 1234, 1999]
    poses = []
    for k, off in zip(i0, (0.45, 0.65, 0.85, 0.95)):
        x, y, psi = ellipse[k, 0], ellipse[k, 1], ellipse[k, 3]
        sgn = 1.0 if k % 2 else -1.0
        poses.append([x - sgn * off * np.sin(psi), y + sgn * off * np.cos(psi), psi + 0.05, 4.0])
    poses = np.array(poses)
    n_missing = []
    for pose in poses:
        d, o, st = _run(eng, cfg, world, pose, None)
        miss = (d.flags & 8) != 0
        assert miss.tolist() == ((o["flags"] & 8) != 0).tolist()
        n_missing.append(int(miss.sum()) // len(wd))
    assert n_missing == [1, 2, 3, 3]   # lookaheads 0.4, 0.6, 0.8, 1.0 against the four offsets
    b = eng.plan_batch(poses, want_flags=True)
    for k, pose in enumerate(poses):
        o = co.plan(cfg, world, pose, None)
        assert ((b.flags[k] & 8) != 0).tolist() == ((o["flags"] & 8) != 0).tolist()
        assert int(b.best_idx[k]) == int(o["best_idx"]) or not np.isfinite(o["best_cost"])


def test_explicit_goals_and_generate(ellipse):
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd, kappa_max=0.0)
    pose, opp = H.scenario(ellipse, 41, 2)
    gx, gy = np.meshgrid(np.linspace(0.8, 3.5, 10), np.linspace(-1.0, 1.0, 11), indexing="ij")
    goals = np.stack([gx.ravel(), gy.ravel(), 0.2 * gy.ravel()], axis=1)
    d = eng.plan_goals(pose, goals, opp, update_prev=False, want_states=True)
    o = co.plan(cfg, world, pose, opp, goals=goals, want_states=True)
    H.compare_plan(d, o, cfg)
    states, params, valid = eng.generate(goals)
    both = valid & ((o["flags"] & 1) != 0)
    assert both.sum() > 80
    assert H.close(states[both], o["states"][both], scale=H.traj_scale(o["states"][both])).all()


def test_candidate_sharding_matches_unsharded(ellipse, corridor):
    la, wd = synth.goal_grid(3)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor)
    pose, opp = H.scenario(ellipse, 77, 8)
    full = eng.plan(pose, opp, update_prev=False)
    C = eng.n_candidates
    best = (np.inf, C)
    assert np.isfinite(full.best_cost)
    for g in range(4):
        lo, hi = g * C // 4, (g + 1) * C // 4
        part = eng.plan(pose, opp, update_prev=False, shard=(lo, hi))
        assert np.array_equal(part.costs[lo:hi], full.costs[lo:hi])
        assert not np.isfinite(part.costs[:lo]).any() and not np.isfinite(part.costs[hi:]).any()
        if (part.best_cost, part.best_idx) < best:
            best = (float(part.best_cost), int(part.best_idx))
    assert best[1] == full.best_idx


def test_batch_matches_single_and_oracle(ellipse, corridor):
    la, wd = synth.goal_grid(4)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, kappa_max=3.0)
    S, K = 96, 8
    poses, opp, n_opp = synth.scenario_batch(ellipse, S, K, 1004)
    b = eng.plan_batch(poses, opp, n_opp, want_flags=True)
    o = co.plan_batch(cfg, world, poses, opp, n_opp, n_threads=co.max_threads())
    fin = np.isfinite(o["costs"]) & np.isfinite(b.costs)
    assert fin.sum() > 200
    assert H.close(b.costs[fin], o["costs"][fin]).all()
    assert (np.isfinite(o["costs"]) != np.isfinite(b.costs)).mean() < 0.01
    agree = b.best_idx == o["best_idx"]
    for s in np.nonzero(~agree)[0]:
        gap = abs(float(o["costs"][s, b.best_idx[s]]) - float(o["best_cost"][s]))
        assert gap < 1e-5, (s, gap)
    ok = agree & np.isfinite(o["best_cost"])
    assert ok.sum() > 50
    assert H.close(b.best_traj[ok], o["best_traj"][ok], scale=H.traj_scale(o["best_traj"][ok])).all()
    assert H.close(b.steer_speed[ok], o["steer_speed"][ok], 1e-4, 1e-4).all()
    for s in (0, 17, 95):
        d = eng.plan(poses[s], opp[s, :n_opp[s]], update_prev=False)
        assert d.best_idx == b.best_idx[s]
        assert np.array_equal(d.costs, b.costs[s])
        assert np.array_equal(d.best_traj, b.best_traj[s])


def test_batch_device_pointer_api(ellipse):
    import torch
    la, wd = synth.goal_grid(4)
    eng, cfg, world = H.make_pair(ellipse, la, wd, kappa_max=3.0)
    S, K = 64, 4
    poses, opp, n_opp = synth.scenario_batch(ellipse, S, K, 5)
    dev = torch.device("cuda", eng.device)
    tp, to = torch.from_numpy(poses).to(dev), torch.from_numpy(opp).to(dev)
    tn = torch.from_numpy(n_opp).to(dev)
    idx = torch.empty(S, dtype=torch.int32, device=dev)
    cost = torch.empty(S, dtype=torch.float32, device=dev)
    costs = torch.empty(S, eng.n_candidates, dtype=torch.float32, device=dev)
    eng.plan_batch_dev(tp, to, tn, best_idx=idx, best_cost=cost, costs=costs)
    torch.cuda.synchronize(dev)
    b = eng.plan_batch(poses, opp, n_opp)
    assert np.array_equal(idx.cpu().numpy(), b.best_idx)
    assert np.array_equal(costs.cpu().numpy(), b.costs)


def test_collision_flags_bit_exact_on_device_states(ellipse):
    """Teacher-forced: the float32 mirror of the collision predicate, fed the device's own float32
    states / headings / per-query constants, reproduces the opponent and map flags bit for bit
    (the end-to-end comparison against the float64 oracle above tolerates boundary cases)."""
    la, wd = synth.goal_grid(3)
    grid = synth.corridor_grid(half_width=1.0)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=grid)
    n_checked = 0
    for seed in (1003, 2, 3):
        pose, opp = H.scenario(ellipse, seed, 8)
        d = eng.plan(pose, opp, update_prev=False, want_states=True, want_headings=True)
        f, i = eng.debug_query_ctx()
        hl, hw = np.float32(0.5 * cfg.car_length), np.float32(0.5 * cfg.car_width)
        rc2 = np.float32(4.0 * ((0.5 * cfg.car_length) ** 2 + (0.5 * cfg.car_width) ** 2))
        mirror = co.collide_f32(d.states, d.headings, f[8:].reshape(16, 4), int(i[5]), f[2:8],
                                i[0:2], grid[0], hl, hw, rc2)
        valid = (d.flags & 1) != 0
        assert valid.sum() > 1000
        assert np.array_equal(mirror[valid], d.flags[valid] & 6)
        assert (mirror[valid] & 2).any() and (mirror[valid] & 4).any() and (mirror[valid] == 0).any()
        n_checked += int(valid.sum())
    assert n_checked > 5000


def test_g1_clothoid_generator(ellipse, corridor):
    """generator=1: the reference's G1-clothoid generator through the same
    cost / collision / argmin pipeline."""
    la, wd = np.linspace(0.6, 3.8, 16), np.linspace(-1.1, 1.1, 15)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, generator=1, kappa_max=3.0)
    n_valid = 0
    for seed in (51, 52, 53):
        pose, opp = H.scenario(ellipse, seed, 4)
        d, o, st = _run(eng, cfg, world, pose, opp)
        n_valid += st["n_both_valid"]
        both = ((d.flags & 1) != 0) & ((o["flags"] & 1) != 0)
        # linear curvature: kappa_i = kappa0 + dkappa * s_i
        s = np.linspace(0, 1, cfg.n_samples)[None, :] * d.params[both, 2:3]
        np.testing.assert_allclose(d.states[both][:, :, 3], d.params[both, 0:1] + d.params[both, 1:2] * s,
                                   rtol=1e-4, atol=1e-4)
    assert n_valid > 300
    goals = np.array([[1.0, 1.0, 0.0], [2.0, 0.3, 0.2], [3.0, -0.5, -0.3]])
    eng.configure(kappa_max=0.0)   # (1, 1, 0) peaks at 3.4 rad/m, above the 3.0 limit used above
    states, params, valid = eng.generate(goals)
    assert valid.all()
    for g, stt in zip(goals, states):
        ref = co.clothoid(g, n_newton=cfg.n_newton, m=cfg.n_samples)[1]
        assert H.close(stt, ref, scale=H.traj_scale(ref)).all()


def test_row_interleaved_shards_match_unsharded(ellipse, corridor):
    """f1l_plan_rows: rank r of W evaluates lookahead rows r, r + W, ...; the shards' costs and
    flags are the unsharded ones at the same global indices (bit for bit), untouched candidates
    stay +inf, and the shard winners reduce to the unsharded winner."""
    la, wd = np.linspace(0.5, 3.6, 21), np.linspace(-1.1, 1.1, 9)     # 21 rows: ragged over W = 2, 4
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, kappa_max=0.0)
    nL, nW = len(la), len(wd)
    for seed in (61, 62):
        pose, opp = H.scenario(ellipse, seed, 5)
        whole = eng.plan(pose, opp, update_prev=False)
        for W in (2, 4, 21):
            best = []
            seen = np.zeros(nL * nW, bool)
            for r in range(W):
                d = eng.plan(pose, opp, update_prev=False, rows=(r, W))
                mine = np.zeros((nL, nW), bool)
                mine[r::W] = True
                mine = mine.ravel()
                assert np.array_equal(d.costs[mine], whole.costs[mine])
                assert np.array_equal(d.flags[mine], whole.flags[mine])
                assert np.isinf(d.costs[~mine]).all() and (d.flags[~mine] == 0).all()
                assert mine[d.best_idx]
                if np.isfinite(whole.costs[mine]).any():
                    assert d.best_idx == int(np.argmin(np.where(mine, whole.costs, np.inf)))
                else:   # all +inf: the first evaluated candidate
                    assert d.best_idx == r * nW and d.no_feasible
                best.append((float(d.best_cost), int(d.best_idx)))
                seen |= mine
            assert seen.all()
            assert min(best)[1] == whole.best_idx
    # a block and a row shard with the same (c_begin, c_end) must not share a captured graph
    la2 = np.linspace(0.5, 3.6, 20)
    eng.set_goal_grid(la2, wd)
    whole = eng.plan(pose, opp, update_prev=False)
    blk = eng.plan(pose, opp, update_prev=False, shard=(0, 10 * nW))
    row = eng.plan(pose, opp, update_prev=False, rows=(0, 2))
    mine = np.zeros((20, nW), bool)
    mine[0::2] = True
    assert np.array_equal(row.costs[mine.ravel()], whole.costs[mine.ravel()])
    assert np.isinf(row.costs[~mine.ravel()]).all()
    assert np.array_equal(blk.costs[:10 * nW], whole.costs[:10 * nW]) and np.isinf(blk.costs[10 * nW:]).all()
    with pytest.raises(Exception):
        eng.plan(pose, opp, update_prev=False, rows=(3, 2))       # row_begin >= row_step


def test_generators_reproduce_analytic_known_answers(ellipse):
    """Known answers that need no oracle: the G1 clothoid joining the two poses of a circular arc
    is that arc (kappa = 1/R, no curvature rate, length R phi), and both generators return the
    straight line for a goal straight ahead.  FP32 device states within 1e-4 (north_star)."""
    arcs = [(R, phi) for R in (0.8, 2.0, 5.0, 40.0) for phi in (0.05, 0.3, 1.0, np.pi / 2, -0.7, -1.4)
            if R * abs(phi) <= 6.0]
    goals = np.array([[R * np.sin(abs(p)), np.sign(p) * R * (1 - np.cos(p)), p] for R, p in arcs])
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd, generator=1, kappa_max=0.0)
    states, params, valid = eng.generate(goals)
    assert valid.all()
    for (R, phi), st, pr in zip(arcs, states, params):
        s = np.linspace(0, R * abs(phi), cfg.n_samples)
        ref = np.stack([R * np.sin(s / R), np.sign(phi) * R * (1 - np.cos(s / R)), np.sign(phi) * s / R,
                        np.full_like(s, np.sign(phi) / R)], axis=1)
        assert H.close(st, ref, scale=H.traj_scale(ref)).all(), (R, phi, np.abs(st - ref).max(axis=0))
        assert abs(pr[2] - R * abs(phi)) <= 1e-4 * R * abs(phi)            # length
        assert abs(pr[0] - np.sign(phi) / R) <= 1e-4 * max(1.0 / R, 1.0)   # kappa0
    line = np.array([[0.4, 0.0, 0.0], [2.0, 0.0, 0.0], [4.0, 0.0, 0.0]])
    for gen in (0, 1):
        eng.configure(generator=gen)
        states, params, valid = eng.generate(line)
        assert valid.all()
        for g, st in zip(line, states):
            np.testing.assert_allclose(st[:, 0], np.linspace(0, g[0], cfg.n_samples), atol=1e-4 * g[0] + 1e-6)
            assert np.abs(st[:, 1:]).max() <= 1e-5


def test_real_track_and_map_fixtures(golden_spielberg):
    """Spielberg raceline + its ROS map and the Levine raceline + SLAM map (reference fixtures,
    carried as tests/golden/maps.npz): plan parity against the oracle along the tracks."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "maps.npz"))

    def unpack(name):
        h, w = g[name + "_shape"]
        return np.unpackbits(g[name + "_bits"], axis=1)[:, :w]
    cases = [(golden_spielberg["waypoints"], unpack("spielberg"), tuple(g["spielberg_origin"]),
              float(g["spielberg_res"]), np.linspace(0.8, 3.5, 10), np.linspace(-1.5, 1.5, 13)),
             (g["levine_raceline"], unpack("levine"), tuple(g["levine_origin"]), float(g["levine_res"]),
              np.linspace(0.5, 2.0, 8), np.linspace(-0.8, 0.8, 9))]
    for wp, occ, origin, res, la, wd in cases:
        eng, cfg, world = H.make_pair(wp, la, wd, grid=(occ, origin, res), kappa_max=2.5)
        rng = np.random.default_rng(4)
        n_map = 0
        for k in rng.integers(0, wp.shape[0] - 1, 6):
            pose = np.array([wp[k, 0], wp[k, 1], wp[k, 3] + rng.normal(0, 0.05), 4.0])
            d, o, st = _run(eng, cfg, world, pose, None)
            n_map += st["collide_map_count"]
            assert st["n_both_valid"] >= 5
        assert n_map > 0


def test_cuda_graph_replay_matches_stream_launches(ellipse, corridor):
    """the graph of the single-query chain gives bit-identical results, survives changes of pose,
    opponent count, requested outputs, previous path and configuration"""
    la, wd = synth.goal_grid(3)
    eng, cfg, world = H.make_pair(ellipse, la[::2], wd[::2], grid=corridor)
    poses, opp, n_opp = synth.scenario_batch(ellipse, 6, 8, 9)
    outs = {}
    for mode in (True, False):
        eng.set_graph(mode)
        eng.set_prev_path(None)
        res = []
        for s in range(6):
            k = int(n_opp[s]) if s % 2 else 0
            d = eng.plan(poses[s], opp[s, :k] if k else None, update_prev=(s != 3), detail=(s % 3 != 0),
                         want_states=(s == 4))
            res.append(d)
        eng.configure(kappa_max=2.0)
        res.append(eng.plan(poses[0], opp[0], update_prev=False))
        eng.configure(kappa_max=cfg.kappa_max)
        outs[mode] = res
    for a, b in zip(outs[True], outs[False]):
        assert a.best_idx == b.best_idx and a.best_cost == b.best_cost
        assert a.steer == b.steer and a.speed == b.speed
        assert np.array_equal(a.best_traj, b.best_traj)
        if a.costs is not None:
            assert np.array_equal(a.costs, b.costs) and np.array_equal(a.flags, b.flags)
        if a.states is not None:
            assert np.array_equal(a.states, b.states)


@pytest.mark.parametrize("config,window", [(1, 128), (1, 0), (3, 128), (3, 512)])
def test_pruned_window_is_bit_identical(ellipse, corridor, config, window):
    """prune_window = 1 drops window segments that provably cannot be nearest to any sample of a
    candidate: every cost term, flag, cost and the argmin must be bit-identical to the full scan,
    with fewer (candidate, segment) pairs evaluated."""
    la, wd = synth.goal_grid(config)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, window=window, kappa_max=0.0)
    seeds = (21, 22, 23, 24) if config == 1 else (25,)
    eng.set_stats(True)
    full, work_full = [], None
    for seed in seeds:
        pose, opp = H.scenario(ellipse, seed, 4)
        full.append(eng.plan(pose, opp, update_prev=False))
    work_full = eng.stats()
    eng.configure(prune_window=1)
    for seed, f in zip(seeds, full):
        pose, opp = H.scenario(ellipse, seed, 4)
        p = eng.plan(pose, opp, update_prev=False)
        assert np.array_equal(p.flags, f.flags)
        assert np.array_equal(p.terms, f.terms)
        assert np.array_equal(p.costs, f.costs)
        assert p.best_idx == f.best_idx and p.best_cost == f.best_cost
    work_pruned = eng.stats()
    assert work_full[1] == work_pruned[1] > 0
    assert work_pruned[0] < 0.6 * work_full[0], (work_full, work_pruned)


def test_pruned_window_batch_and_hairpin(golden_spielberg):
    """pruned vs full scan on the Spielberg raceline (hairpins: the kept segments may form two
    runs that one index range must cover) in batch mode."""
    wp = golden_spielberg["waypoints"]
    la, wd = np.linspace(0.5, 3.0, 6), np.linspace(-1.0, 1.0, 9)
    from f1tenth_planning_b200.engine import Engine
    eng = Engine(window=256, kappa_max=0.0)
    eng.set_track(wp)
    eng.set_goal_grid(la, wd)
    poses, opp, n_opp = synth.scenario_batch(wp, 512, 4, 5)
    a = eng.plan_batch(poses, opp, n_opp, want_flags=True)
    eng.configure(prune_window=1)
    b = eng.plan_batch(poses, opp, n_opp, want_flags=True)
    assert np.isfinite(a.costs).sum() > 1000
    assert np.array_equal(a.costs, b.costs) and np.array_equal(a.flags, b.flags)
    assert np.array_equal(a.best_idx, b.best_idx) and np.array_equal(a.best_traj, b.best_traj)


def test_empty_ragged_and_extreme_inputs(ellipse, corridor):
    """the edges of the input space: empty batches, no / the maximum number of opponents, one
    candidate, a window longer than the track, a non-finite pose"""
    la, wd = synth.goal_grid(1)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=corridor, kappa_max=0.0)
    # empty batches are empty answers
    b = eng.plan_batch(np.zeros((0, 4)), np.zeros((0, 2, 3)), np.zeros(0, np.int32), want_flags=True)
    assert b.best_idx.shape == (0,) and b.costs.shape == (0, 28) and b.best_traj.shape == (0, 100, 4)
    r = eng.pure_pursuit_batch(np.zeros((0, 3)), 0.8)
    assert r.nearest.shape == (0, 4) and r.status.shape == (0,)
    assert eng.front_axle_batch(np.zeros((0, 4)), 0.33)[0].shape == (0, 6)
    assert eng.generate(np.zeros((0, 3)))[0].shape == (0, 100, 4)
    # ragged opponent counts, including zero, in one batch
    poses, opp, n_opp = synth.scenario_batch(ellipse, 64, 16, 9)
    n_opp = (np.arange(64) % 17).astype(np.int32)          # 0 .. 16 = F1L_MAX_OPP
    bb = eng.plan_batch(poses, opp, n_opp, want_flags=True)
    for s in (0, 5, 16, 33):
        d = eng.plan(poses[s], opp[s, :n_opp[s]] if n_opp[s] else None, update_prev=False)
        assert np.array_equal(d.costs, bb.costs[s]) and d.best_idx == bb.best_idx[s]
        o = co.plan(cfg, world, poses[s], opp[s, :n_opp[s]] if n_opp[s] else None, want_states=True)
        H.compare_plan(eng.plan(poses[s], opp[s, :n_opp[s]] if n_opp[s] else None,
                                update_prev=False, want_states=True), o, cfg)
    with pytest.raises(ValueError):
        eng.plan(poses[0], np.zeros((17, 3)))
    # one candidate
    eng.set_goal_grid([0.8], [0.0])
    d = eng.plan(poses[1], None, update_prev=False)
    assert d.costs.shape == (1,) and d.best_idx == 0
    eng.set_goal_grid(la, wd)
    # a window longer than the track is the whole track
    eng.configure(window=10 ** 6)
    d_big = eng.plan(poses[2], opp[2, :2], update_prev=False)
    eng.configure(window=0)
    d_all = eng.plan(poses[2], opp[2, :2], update_prev=False)
    assert np.array_equal(d_big.costs, d_all.costs)
    # a non-finite pose yields "no feasible candidate", not a hang or a crash
    bad = poses[3].copy()
    bad[0] = np.nan
    d = eng.plan(bad, None, update_prev=False)
    assert d.no_feasible and not np.isfinite(d.costs).any()
