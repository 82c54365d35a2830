"""f1tenth_planning_b200 -- B200-native (sm_100a CUDA) lattice-planner hot path of
f1tenth/f1tenth_planning behind the reference's Python API.

    from f1tenth_planning_b200 import LatticePlanner, PurePursuitPlanner
    from f1tenth_planning_b200.utils import nearest_point, intersect_point, get_actuation

The CUDA library (lib/libf1l.so, C-ABI in include/f1l.h) is loaded on first use; there is no CPU
fallback -- a missing library or GPU raises.
"""
__version__ = "0.1.0"

_LAZY = {
    "LatticePlanner": ("lattice_planner", "LatticePlanner"),
    "sample_lookahead_square": ("lattice_planner", "sample_lookahead_square"),
    "PurePursuitPlanner": ("pure_pursuit", "PurePursuitPlanner"),
    "StanleyPlanner": ("stanley", "StanleyPlanner"),
    "LQRPlanner": ("lqr", "LQRPlanner"),
    "Engine": ("engine", "Engine"),
    "F1LError": ("_lib", "F1LError"),
    "nearest_point": ("utils", "nearest_point"),
    "intersect_point": ("utils", "intersect_point"),
    "get_actuation": ("utils", "get_actuation"),
    "get_rotation_matrix": ("utils", "get_rotation_matrix"),
    "pi_2_pi": ("utils", "pi_2_pi"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module("." + mod, __name__), attr)
    raise AttributeError(name)
