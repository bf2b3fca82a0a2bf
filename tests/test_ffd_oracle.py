"""oracle/ffd_oracle.py against the golden vectors produced by the reference's own data/ffd.py."""
import os

import numpy as np

from oracle import ffd_oracle
from tests import util


def load():
    z = np.load(os.path.join(util.GOLDEN, "ffd.npz"))
    return {k: z[k] for k in z.files}


def test_weights_and_widths_match_reference():
    g = load()
    for n, d in enumerate(g["d"]):
        w, width = ffd_oracle.ffd_weights(float(d), int(g["T"]), float(g["thres"]))
        assert width == g["widths"][n]
    ic = list(g["names"]).index("close")
    w, width = ffd_oracle.ffd_weights(float(g["d"][ic]), int(g["T"]), float(g["thres"]))
    np.testing.assert_array_equal(w[:width].flip(0).numpy(), g["weights_close"])       # sequential fp32 cumprod: bit-exact


def test_transform_matches_reference():
    g = load()
    out, widths, mw = ffd_oracle.ffd_transform(g["x"], g["d"], float(g["thres"]))
    assert mw == int(g["max_width"])
    np.testing.assert_array_equal(widths, g["widths"])
    # fp32 conv1d's summation order depends on the host's oneDNN/ISA path: the golden (one machine) and this run
    # (another) are both within ~1 ulp of max|x| of the fp64 filter, so that is the bound between them
    atol = 4 * np.finfo(np.float32).eps * float(np.abs(g["x"]).max())
    np.testing.assert_allclose(out, g["out"], rtol=1e-6, atol=atol)


def test_fp64_yardstick_close_to_fp32_reference():
    g = load()
    out64 = ffd_oracle.ffd_transform_f64(g["x"], g["d"], float(g["thres"]))
    np.testing.assert_allclose(out64, g["out"], rtol=2e-4, atol=2e-3)
