"""Generates tests/golden/*.npz from the reference's own test assets and test constants.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources: test_assets/{multiexp_decay,weighted_multiexp_decay}/*.raw (little-endian f64,
produced by the reference's python/*.py lmfit scripts) -- the golden vectors of
tests/integration_tests/main.rs:554-688.
"""
import os

import numpy as np

REF = "/root/reference/test_assets"
OUT = os.path.dirname(os.path.abspath(__file__))

for name in ("multiexp_decay", "weighted_multiexp_decay"):
    d = os.path.join(REF, name)
    np.savez_compressed(
        os.path.join(OUT, f"lmfit_{name}.npz"),
        x=np.fromfile(os.path.join(d, "xdata_1000_64bit.raw"), "<f8"),
        y=np.fromfile(os.path.join(d, "ydata_1000_64bit.raw"), "<f8"),
        conf=np.fromfile(os.path.join(d, "conf_1000_64bit.raw"), "<f8"),
        covmat=np.fromfile(os.path.join(d, "covmat_5x5_64bit.raw"), "<f8").reshape(5, 5),
    )
    print("wrote", name)
