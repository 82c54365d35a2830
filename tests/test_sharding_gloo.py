"""world_size-2 gloo run of the N>1 host logic on the CPU: scenario blocks and candidate blocks
computed per rank (by the oracle here; by the CUDA path on the GPU box) reassemble to the
unsharded answer, and the (cost, idx) gather reproduces np.argmin's first-minimum rule."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from f1tenth_planning_b200 import sharding, synth
from oracle import c_oracle as co


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        track = synth.ellipse_track(n=600, a=20.0, b=10.0)
        la, wd = np.linspace(0.6, 2.5, 6), np.linspace(-0.8, 0.8, 5)
        w = co.World_(track, la, wd)
        cfg = co.default_config(kappa_max=0.0)
        poses, opp, n_opp = synth.scenario_batch(track, 9, 3, 42)
        # scenario sharding (config 4)
        lo, hi = sharding.block(9, rank, world)
        part = co.plan_batch(cfg, w, poses[lo:hi], opp[lo:hi], n_opp[lo:hi])
        # candidate sharding of one query (config 5)
        clo, chi = sharding.block(30, rank, world)
        one = co.plan(cfg, w, poses[0], opp[0, :n_opp[0]], c_begin=clo, c_end=chi)
        best = sharding.reduce_best(one["best_cost"], one["best_idx"])
        tot, mx = sharding.gather_stats([hi - lo, float(rank)])
        q.put((rank, lo, hi, part["best_idx"], part["costs"], best, tot.tolist(), mx.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    track = synth.ellipse_track(n=600, a=20.0, b=10.0)
    la, wd = np.linspace(0.6, 2.5, 6), np.linspace(-0.8, 0.8, 5)
    w = co.World_(track, la, wd)
    cfg = co.default_config(kappa_max=0.0)
    poses, opp, n_opp = synth.scenario_batch(track, 9, 3, 42)
    whole = co.plan_batch(cfg, w, poses, opp, n_opp)
    idx = np.concatenate([r[3] for r in res])
    costs = np.concatenate([r[4] for r in res])
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 9)]
    assert np.array_equal(idx, whole["best_idx"]) and np.array_equal(costs, whole["costs"])
    full = co.plan(cfg, w, poses[0], opp[0, :n_opp[0]])
    for r in res:
        assert r[5][1] == full["best_idx"] and r[5][0] == full["best_cost"]
        assert r[6] == [9.0, 1.0] and r[7] == [5.0, 1.0]


def test_block_partition():
    for n in (1, 7, 8, 100000, 65536):
        for world in (1, 2, 4, 8):
            blocks = [sharding.block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
