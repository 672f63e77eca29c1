"""Generate tests/golden/panocam.npz from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_panocam.py

``PanoCam.get_up_general`` / ``get_lat_general`` (utils/panocam.py:451-556) for a fixed list of camera parameters
(incl. the elevation == 0 branch, negative elevation, off-centre principal points, non-square images)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402

# (focal_rel, im_w, im_h, elevation, roll, cx_rel, cy_rel)
CASES = [
    (0.80, 32, 24, 0.30, -0.20, 0.10, -0.05),
    (1.20, 40, 30, -0.45, 0.60, 0.00, 0.00),
    (0.55, 24, 36, 0.00, 0.35, -0.12, 0.08),     # elevation == 0 branch
    (2.00, 17, 11, 1.10, -1.30, 0.20, 0.20),
    (0.70, 33, 33, -0.05, 0.00, 0.00, 0.30),
    (1.00, 48, 20, 0.75, 3.00, -0.30, -0.10),
    (0.35, 21, 29, -1.40, -2.50, 0.05, 0.00),
    (1.50, 2, 3, 0.20, 0.10, 0.00, 0.00),         # tiny image: linspace end points
]


def main():
    load_reference()
    from perspective2d.utils.panocam import PanoCam
    out = {"cases": np.array(CASES, np.float64)}
    for i, (f, w, h, el, roll, cx, cy) in enumerate(CASES):
        out[f"up{i}"] = PanoCam.get_up_general(f, int(w), int(h), el, roll, cx, cy)
        out[f"lat{i}"] = PanoCam.get_lat_general(f, int(w), int(h), el, roll, cx, cy)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "panocam.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
