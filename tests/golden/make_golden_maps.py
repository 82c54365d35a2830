"""Real-track fixtures derived from the reference's example data (build container only):
the Spielberg occupancy (examples/control/Spielberg_map.{yaml,png}) bit-packed, and the Levine
raceline in LatticePlanner column order.    python tests/golden/make_golden_maps.py
"""
import os
import sys

import numpy as np

REF = "/root/reference/examples/control"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from f1tenth_planning_b200 import io  # noqa: E402


def main():
    occ, origin, res = io.load_map(os.path.join(REF, "Spielberg_map.yaml"))
    lv = io.load_raceline(os.path.join(REF, "levine_raceline.csv"))
    locc, lorigin, lres = io.load_map(os.path.join(REF, "levine_slam.yaml"))
    np.savez_compressed(os.path.join(HERE, "maps.npz"),
                        spielberg_bits=np.packbits(occ, axis=1), spielberg_shape=occ.shape,
                        spielberg_origin=origin, spielberg_res=res,
                        levine_raceline=lv, levine_bits=np.packbits(locc, axis=1),
                        levine_shape=locc.shape, levine_origin=lorigin, levine_res=lres)
    print("maps.npz written", os.path.getsize(os.path.join(HERE, "maps.npz")))


if __name__ == "__main__":
    main()
