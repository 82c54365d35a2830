"""Loaders for the on-disk formats either side of the path: raceline CSVs
and ROS map_server maps as shipped under the reference's examples/ (host-side I/O only).

    load_raceline(path) -> waypoints [N,5] float64 (x, y, v, psi, kappa)   -- LatticePlanner layout
    load_map(yaml_path) -> (occupancy uint8 [H,W] (0 free / 1 occupied), (origin_x, origin_y), res)
"""
import os

import numpy as np


def load_raceline(path):
    """Reads the two raceline layouts of the reference's fixtures:
      examples/control/Spielberg_raceline.csv  '# x_m; y_m; vx_mps; psi_rad; kappa_radpm'
      examples/control/levine_raceline.csv     3 comment lines, 's_m; x_m; y_m; psi_rad;
                                               kappa_radpm; vx_mps; ax_mps2'
    and returns [N,5] (x, y, v, psi, kappa), the column order lattice_planner.py:251 indexes
    (waypoints[i, [0, 1, 3]] = x, y, psi) and pure_pursuit.py:78 uses (column 2 = speed)."""
    names = None
    with open(path) as f:
        for line in f:
            if not line.startswith("#"):
                break
            cols = [c.strip() for c in line.lstrip("#").split(";")]
            if len(cols) >= 5 and any(c.startswith("x_m") for c in cols):
                names = cols
    data = np.loadtxt(path, delimiter=";", comments="#", ndmin=2)
    if names is None:
        names = {5: ["x_m", "y_m", "vx_mps", "psi_rad", "kappa_radpm"],
                 7: ["s_m", "x_m", "y_m", "psi_rad", "kappa_radpm", "vx_mps", "ax_mps2"]}.get(data.shape[1])
        if names is None:
            raise ValueError("unrecognised raceline layout with %d columns" % data.shape[1])
    idx = {n: i for i, n in enumerate(names)}
    try:
        order = [idx["x_m"], idx["y_m"], idx["vx_mps"], idx["psi_rad"], idx["kappa_radpm"]]
    except KeyError as e:
        raise ValueError("raceline header lacks column %s" % e)
    return np.ascontiguousarray(data[:, order], dtype=np.float64)


def _parse_yaml(path):
    try:
        import yaml
        with open(path) as f:
            return yaml.safe_load(f)
    except ImportError:   # tiny fallback for the flat key: value files map_server writes
        out = {}
        with open(path) as f:
            for line in f:
                line = line.split("#")[0]
                if ":" not in line:
                    continue
                k, v = line.split(":", 1)
                v = v.strip()
                if v.startswith("["):
                    out[k.strip()] = [float(x) for x in v.strip("[]").split(",")]
                else:
                    try:
                        out[k.strip()] = float(v)
                    except ValueError:
                        out[k.strip()] = v
        return out


def load_map(yaml_path):
    """ROS map_server convention (examples/control/Spielberg_map.yaml:1-6): occupancy probability
    p = (255 - pixel)/255 (pixel/255 when `negate`), occupied iff p > occupied_thresh; cells that
    are neither free nor occupied (unknown) count as occupied for collision checking.  The image
    is flipped vertically so that row 0 lies at origin_y (the map origin is the lower-left
    pixel).  Rotated maps (origin yaw != 0) are not supported."""
    from PIL import Image
    meta = _parse_yaml(yaml_path)
    img_path = meta["image"]
    if not os.path.isabs(img_path):
        img_path = os.path.join(os.path.dirname(os.path.abspath(yaml_path)), img_path)
    img = np.asarray(Image.open(img_path).convert("L"), dtype=np.float64)
    origin = meta["origin"]
    if len(origin) > 2 and abs(float(origin[2])) > 1e-12:
        raise ValueError("rotated map origins are not supported")
    p = img / 255.0 if int(meta.get("negate", 0)) else (255.0 - img) / 255.0
    free = p < float(meta.get("free_thresh", 0.196))
    occ = np.where(free, 0, 1).astype(np.uint8)
    occ[p > float(meta["occupied_thresh"])] = 1
    return np.ascontiguousarray(occ[::-1]), (float(origin[0]), float(origin[1])), float(meta["resolution"])
