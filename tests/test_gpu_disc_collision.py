"""collision_mode = 1 (the reference's map_collision stub, utils/utils.py:297-301):
three covering discs per footprint, one lookup each in a Euclidean distance transform of the
occupancy grid that the device builds at f1l_set_grid."""
import os

import numpy as np
import pytest

from f1tenth_planning_b200 import synth
from oracle import c_oracle as co
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _maps(golden_spielberg):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "maps.npz"))

    def unpack(name):
        h, w = g[name + "_shape"]
        return np.unpackbits(g[name + "_bits"], axis=1)[:, :w]
    return [("corridor", synth.ellipse_track(), synth.corridor_grid(half_width=1.0)),
            ("spielberg", golden_spielberg["waypoints"],
             (unpack("spielberg"), tuple(g["spielberg_origin"]), float(g["spielberg_res"]))),
            ("levine", g["levine_raceline"], (unpack("levine"), tuple(g["levine_origin"]), float(g["levine_res"])))]


def test_distance_transform_is_bit_exact(golden_spielberg):
    """the device's separable two-pass transform == scipy.ndimage.distance_transform_edt (squared,
    in cells, out of bounds counted as occupied), every cell, on three maps"""
    from f1tenth_planning_b200.engine import Engine
    for name, track, (occ, origin, res) in _maps(golden_spielberg):
        eng = Engine()
        eng.set_grid(occ, origin, res)
        got = eng.get_edt(occ.shape)
        ref = co.edt2_of(occ)
        assert got.dtype == ref.dtype == np.uint16
        assert np.array_equal(got, ref), (name, int((got != ref).sum()))
        assert (got[occ != 0] == 0).all() and (got[occ == 0] > 0).all()
        eng.close()


def test_disc_mode_matches_oracle_and_covers_probe_mode(golden_spielberg):
    counts = {}
    for name, track, grid in _maps(golden_spielberg):
        la = np.linspace(0.6, 3.5, 24) if name != "levine" else np.linspace(0.5, 2.0, 12)
        wd = np.linspace(-1.3, 1.3, 27) if name != "levine" else np.linspace(-0.9, 0.9, 19)
        eng, cfg, world = H.make_pair(track, la, wd, grid=grid, collision_mode=1, kappa_max=0.0,
                                      use_device_lut=False)
        eng0, cfg0, world0 = H.make_pair(track, la, wd, grid=grid, collision_mode=0, kappa_max=0.0,
                                         use_device_lut=False)
        rng = np.random.default_rng(3)
        n_disc = n_probe = n_both = n_probe_only = n_valid = 0
        for k in rng.integers(0, track.shape[0] - 1, 5):
            lat = rng.normal(0.0, 0.25)
            pose = np.array([track[k, 0] - lat * np.sin(track[k, 3]), track[k, 1] + lat * np.cos(track[k, 3]),
                             track[k, 3] + rng.normal(0, 0.08), 4.0])
            d = eng.plan(pose, None, update_prev=False, want_states=True)
            o = co.plan(cfg, world, pose, None, want_states=True)
            st = H.compare_plan(d, o, cfg)
            for key in ("valid_mismatch", "collide_map_mismatch", "collide_map_count", "n_both_valid"):
                counts[name + ":" + key] = counts.get(name + ":" + key, 0) + st[key]
            d0 = eng0.plan(pose, None, update_prev=False)
            valid = ((d.flags & 1) != 0) & ((d0.flags & 1) != 0)
            disc, probe = (d.flags & 4) != 0, (d0.flags & 4) != 0
            n_valid += int(valid.sum())
            n_disc += int((disc & valid).sum())
            n_probe += int((probe & valid).sum())
            n_both += int((disc & probe & valid).sum())
            n_probe_only += int((probe & ~disc & valid).sum())
            # everything but the map flag (and what depends on it) is the same in both modes
            assert np.array_equal(d.flags & ~np.uint8(4), d0.flags & ~np.uint8(4))
            free = valid & ~disc & ~probe
            assert np.array_equal(d.costs[free], d0.costs[free])
        counts.update({name + ":valid": n_valid, name + ":disc_hits": n_disc, name + ":probe_hits": n_probe,
                       name + ":both": n_both, name + ":probe_only": n_probe_only})
        assert n_disc > 0 and n_probe > 0
        # the discs cover the rectangle and the threshold allows for cell quantisation, the nine
        # probes only sample the rectangle: every probe collision is a disc collision
        assert n_probe_only == 0, counts
        assert n_disc >= n_probe
    # how much more conservative the discs are: up to r - W/2 = 2.8 cm to the side and
    # L/3 + r - L/2 = 8.6 cm at the ends, plus sqrt(2) cells of quantisation allowance
    cfgd = co.default_config()
    r = float(np.hypot(cfgd.car_length / 6.0, cfgd.car_width / 2.0))
    counts["disc_radius_m"] = r
    counts["over_approximation_side_m"] = r - cfgd.car_width / 2.0
    counts["over_approximation_end_m"] = cfgd.car_length / 3.0 + r - cfgd.car_length / 2.0
    H.record_parity("disc_collision_mode", counts)


def test_disc_mode_batch_and_too_fine_map(ellipse):
    from f1tenth_planning_b200.engine import Engine, F1LError
    grid = synth.corridor_grid(half_width=1.0)
    la, wd = synth.goal_grid(4)
    eng, cfg, world = H.make_pair(ellipse, la, wd, grid=grid, collision_mode=1, kappa_max=0.0,
                                  use_device_lut=False)
    poses, opp, n_opp = synth.scenario_batch(ellipse, 512, 4, 11)
    b = eng.plan_batch(poses, opp, n_opp, want_flags=True)
    counts = H.compare_batch(b, np.arange(512), poses, opp, n_opp, cfg, world)
    H.record_parity("disc_collision_mode_batch", counts)
    assert ((b.flags & 4) != 0).any()
    # a map finer than the transform's exact radius (24 cells) cannot serve the disc mode
    fine = Engine(collision_mode=1)
    fine.set_track(ellipse)
    fine.set_goal_grid(la, wd)
    fine.set_grid(np.zeros((64, 64), np.uint8), (0.0, 0.0), 0.005)
    with pytest.raises(F1LError):
        fine.plan(poses[0], None)
