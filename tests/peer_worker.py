"""One rank of the peer-memory exchange test (tests/test_gpu_peer_exchange.py):

    python tests/peer_worker.py <rank> <world> <port>

Ranks use device rank % device_count, so the test also runs with all ranks on one GPU (the IPC
mapping and the system-scope atomics are the same; the kernels of the ranks then time-slice)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import torch
    import torch.distributed as dist
    from f1tenth_planning_b200 import sharding, synth
    from f1tenth_planning_b200.engine import Engine

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    track = synth.ellipse_track(n=600, a=24.0, b=12.0)
    grid = synth.corridor_grid(a=24.0, b=12.0)
    eng = Engine(device=dev, n_samples=100)
    eng.set_track(track)
    eng.set_grid(*grid)
    eng.set_goal_grid(np.linspace(0.5, 3.5, 24), np.linspace(-1.2, 1.2, 21))
    C = eng.n_candidates
    lo, hi = sharding.block(C, rank, world)
    poses, opp, n_opp = synth.scenario_batch(track, 12, 4, 77)    # same on every rank

    # expected: the unsharded query, and the host-side reduction of the shard winners
    whole = [eng.plan(poses[s], opp[s, :n_opp[s]], update_prev=False, detail=False) for s in range(12)]
    for s in range(12):
        d = eng.plan(poses[s], opp[s, :n_opp[s]], update_prev=False, detail=False, shard=(lo, hi))
        cost, idx = sharding.reduce_best(d.best_cost, d.best_idx)
        assert idx == whole[s].best_idx and np.float32(cost) == np.float32(whole[s].best_cost), (s, idx, cost)

    eng.configure(kappa_max=1e-9)   # (almost) every candidate invalid: the all-+inf rule
    whole_inf = eng.plan(poses[0], opp[0, :1], update_prev=False, detail=False)
    eng.configure(kappa_max=0.0)

    eng.attach_peers()
    for rep in range(4):
        eng.set_graph(rep != 1)                                   # graph replay and plain launches
        for s in range(12):
            if rep < 2:   # contiguous candidate blocks
                d = eng.plan(poses[s], opp[s, :n_opp[s]], update_prev=False, detail=False, shard=(lo, hi))
            else:         # row-interleaved shards
                d = eng.plan(poses[s], opp[s, :n_opp[s]], update_prev=False, detail=False, rows=(rank, world))
            w = whole[s]
            assert d.best_idx == w.best_idx, (rank, rep, s, d.best_idx, w.best_idx)
            assert np.float32(d.best_cost) == np.float32(w.best_cost)
            assert np.array_equal(d.best_traj, w.best_traj)
            assert d.steer == w.steer and d.speed == w.speed
    # an infeasible query: every candidate +inf -> the first candidate of the whole grid everywhere
    eng.configure(kappa_max=1e-9)
    d = eng.plan(poses[0], opp[0, :1], update_prev=False, detail=False, shard=(lo, hi))
    assert whole_inf.no_feasible and whole_inf.best_idx == 0
    assert d.no_feasible == whole_inf.no_feasible and d.best_idx == whole_inf.best_idx, (d.no_feasible, d.best_idx)
    # more ranks than lookahead rows (advisor r1): a rank without rows evaluates nothing, still
    # joins the exchange and returns the global winner -- nobody times out
    eng.configure(kappa_max=0.0)
    eng.set_goal_grid(np.linspace(1.0, 3.0, world - 1) if world > 2 else [2.0], np.linspace(-1.2, 1.2, 21))
    few = eng.plan(poses[2], opp[2, :n_opp[2]], update_prev=False, detail=False)
    for graph in (True, False):
        eng.set_graph(graph)
        d = eng.plan(poses[2], opp[2, :n_opp[2]], update_prev=False, detail=False, rows=(rank, world))
        assert d.best_idx == few.best_idx and np.float32(d.best_cost) == np.float32(few.best_cost), \
            (rank, d.best_idx, few.best_idx)
        assert np.array_equal(d.best_traj, few.best_traj) and d.steer == few.steer
    eng.set_goal_grid(np.linspace(0.5, 3.5, 24), np.linspace(-1.2, 1.2, 21))
    eng.detach_peers()
    # detached again: plan(shard=...) is local
    eng.configure(kappa_max=0.0)
    d = eng.plan(poses[1], opp[1, :n_opp[1]], update_prev=False, detail=False, shard=(lo, hi))
    assert lo <= d.best_idx < hi
    eng.close()

    # the reference-facing planner: shard_across() makes plan() collective
    from f1tenth_planning_b200 import LatticePlanner
    pl = LatticePlanner(waypoints=track, device=dev)
    pl.set_goal_grid(np.linspace(0.5, 3.5, 24), np.linspace(-1.2, 1.2, 21))
    pl.set_map(*grid)
    ref = [pl.plan_detailed(*poses[s], opponent_poses=opp[s, :n_opp[s]]) for s in range(3)]
    pl2 = LatticePlanner(waypoints=track, device=dev)
    pl2.set_goal_grid(np.linspace(0.5, 3.5, 24), np.linspace(-1.2, 1.2, 21))
    pl2.set_map(*grid)
    pl2.shard_across()
    for s in range(3):
        steer, speed, traj = pl2.plan(*poses[s], opponent_poses=opp[s, :n_opp[s]])
        # the same answers as the unsharded planner, similarity-to-previous-plan term included
        # (every rank stores the global winner as its previous path)
        assert pl2.last.best_idx == ref[s].best_idx, (s, pl2.last.best_idx, ref[s].best_idx)
        assert np.float32(pl2.last.best_cost) == np.float32(ref[s].best_cost)
        assert steer == ref[s].steer and speed == ref[s].speed
        assert np.array_equal(pl2.last.best_traj, ref[s].best_traj)
    pl2.unshard()
    dist.barrier()
    dist.destroy_process_group()
    print("peer worker %d/%d ok (device %d, candidates [%d, %d))" % (rank, world, dev, lo, hi))


if __name__ == "__main__":
    main()
