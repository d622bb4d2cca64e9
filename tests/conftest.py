import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def model_cases():
    """Names of every golden fixture that holds a full GP-GRIEF model evaluation."""
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        n = os.path.basename(f)[:-4]
        if n.startswith(("syn_", "ref_test_gp_grief", "c1_")):
            out.append(n)
    return out


def case_inputs(g):
    """Unpack a model fixture into the arguments of the oracle / the product (input-dim order)."""
    d = int(g["n_grid_dims"])
    names = [str(g["kernel_name"])] * d
    xg = [g["xg_%d" % i] for i in range(d)]
    return dict(d=d, names=names, variances=list(g["variances"]), lengthscales=list(g["lengthscales"]),
                xg=xg, n_eigs=int(g["n_eigs"]), x=g["x"], y=g["y"], w=g["w"], noise_var=float(g["noise_var"]))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
