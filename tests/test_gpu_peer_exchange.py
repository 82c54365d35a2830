"""Candidate-sharded dense query with the argmin exchanged through peer memory (f1l_xchg_*,
SURVEY 8e): every rank must return the winner of the UNSHARDED query, bit for bit.  Ranks are
separate processes; with one GPU they share it, with several each takes its own."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_query_over_peer_memory(world):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "peer_worker.py"),
                               str(r), str(world), str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, cwd=ROOT) for r in range(world)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=240)[0])
    finally:
        for p in procs:          # exactly the processes started here
            if p.poll() is None:
                p.kill()
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d:\n%s" % (r, out[-3000:])
        assert "peer worker %d/%d ok" % (r, world) in out
