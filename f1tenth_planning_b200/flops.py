"""Algorithmic work model of the lattice hot path.  One
module shared by bench.py, DESIGN.md's numbers and the tests.  FMA = 2 FLOP; MUFU ops separate.
"""

Q_NEWTON = 32   # Simpson intervals of the Newton quadrature
P_PROBES = 9    # occupancy-grid probes per footprint


def candidate_flops(M=100, W=128, K=8, P=P_PROBES, I=8, Q=Q_NEWTON, sat_hits=0, full=True):
    """FLOPs of one candidate.  full=False: a candidate that failed the validity test stops after
    generation (no deviation / collision work)."""
    f = 27 + I * (44 * (Q + 1) + 110) + 30
    if full:
        f += M * (68 + 17 * W + 6 * K + 16 * P) + 90 * sat_hits
    else:
        f += M * 68
    return f


def candidate_mufu(M=100, I=8, Q=Q_NEWTON):
    return I * (2 * Q + 1) + 5 * M


def candidate_hbm_bytes(n_terms=0, flags=True):
    """algorithmic HBM bytes per candidate: cost f32 (+ flag byte, optional per-term costs)."""
    return 4 + (1 if flags else 0) + 4 * n_terms


def pose_flops(n_waypoints=2000):
    """batched nearest_point + pure pursuit (BASELINE config 2), per pose"""
    return 17 * (n_waypoints - 1) + 60


def pose_hbm_bytes():
    return 24 + 32 + 4 + 32 + 4 + 16 + 4
