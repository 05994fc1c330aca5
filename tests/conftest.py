import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def small_casing_mesh(nth=1, nx_fine=6, nz_core=10, npad=5):
    """A scaled-down casing mesh: fine radial cells through a thin casing wall,
    geometric padding in r and z (same construction as mesh.py:32-63, 439-453)."""
    from oracle.cyl_oracle import mesh_tensor
    hx = mesh_tensor([(0.01, nx_fine), (0.01, 6, 1.5), (0.2, 3), (0.2, npad, 1.5)])
    hz = mesh_tensor([(0.5, npad, -1.5), (0.5, nz_core), (0.5, npad, 1.5)])
    hy = np.r_[2 * np.pi] if nth == 1 else np.ones(nth) * 2 * np.pi / nth
    z0 = -np.sum(hz[:npad + nz_core - 3])
    return hx, hy, hz, z0


CASING = dict(sigma_back=1e-2, sigma_air=1e-6, sigma_casing=5.5e6, sigma_inside=1.0, mur_back=1.0,
              mur_casing=50.0, mur_inside=1.0, casing_a=0.02, casing_b=0.04, casing_z=(-3.0, 0.0), surface_z=0.0)


def paint_params(d, halfspace=True):
    return np.array([d["sigma_back"], 1.0 if halfspace else 0.0, d["sigma_air"], d["surface_z"], 1.0,
                     d["sigma_casing"], d["sigma_inside"], d["casing_a"], d["casing_b"], d["casing_z"][0],
                     d["casing_z"][1], d["mur_back"], d["mur_casing"], d["mur_inside"], 0.0, 0.0, 0.0, 0.0])


@pytest.fixture(scope="session")
def gpu_ctx():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from casingsimulations_b200 import _lib
    ctx = _lib.Context(0)
    yield ctx
    ctx.close()
