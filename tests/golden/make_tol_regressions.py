"""TEST INFRASTRUCTURE: writes tests/golden/tol_band_regressions.npz -- the rlDisney samples on which the ulp-perturbed host
build of the tolerance policy once disagreed with the reference on a FLAG without listing the sample for the bit-exact
re-run (each one a missing term in the band tracker's estimate of the reference's own rounding noise):

  0  GTR1 lobe, ry = 1 - 2^-24: 1 - powf(a2, 1 - ry) cancels, cos(theta) is 0 or 4e-4 by rounding alone;
  1  GTR1 lobe, a2 = 0.998, ry = 9e-4: the same absolute error in sin^2 = 1 - cos^2 (cos -> 1);
  2  GTR2 lobe, rx_s / gtr2Weight = 1 - 3e-5: sqrt(rx / (1 - rx)) amplifies one ulp of the quotient 3e4 times.

The samples are addressed by (repetition, index) of the hunt's input recipe (parity.disney_inputs over 2^22 samples)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity  # noqa: E402

WHERE = [(6, 1125145), (65, 953096), (86, 2838569)]
n = 1 << 22
cols = {}
for rep, i in WHERE:
    seed = 0x5EED0002 + 7919 * (rep + 100) + 1
    sg, kw, u = parity.disney_inputs(n, seed)
    for k, v in sg.items():
        if v is not None:
            cols.setdefault("sg_" + k, []).append(v[i])
    for k, v in kw.items():
        if isinstance(v, tuple):
            for j, c in enumerate(v):
                cols.setdefault(f"p_{k}_{j}", []).append(c[i])
        else:
            cols.setdefault("p_" + k, []).append(v[i])
    for j, x in enumerate(u):
        cols.setdefault(f"u_{j}", []).append(x[i])
np.savez(os.path.join(ROOT, "tests", "golden", "tol_band_regressions.npz"), **{k: np.asarray(v, np.float32) for k, v in cols.items()})
print({k: np.asarray(v).shape for k, v in cols.items()})
