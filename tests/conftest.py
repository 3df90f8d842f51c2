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


def load_mesh(name):
    """Mesh fixtures re-encoded from the reference's bundled meshes by tests/golden/make_golden.py."""
    from chm_b200.mesh import TriMesh
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    params = {k[6:]: d[k] for k in d.files if k.startswith("param_")}
    ls = d["local_sizes"] if "local_sizes" in d.files else None
    return TriMesh(d["vertex"], d["elem"], d["neigh"], params, local_sizes=ls)


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / nb if nb > 0 else np.linalg.norm(a)


def max_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


@pytest.fixture(scope="session")
def granger():
    return load_mesh("granger1m")


@pytest.fixture(scope="session")
def slope():
    return load_mesh("slope")


@pytest.fixture(scope="session")
def slope_metis():
    return load_mesh("slope_metis")


def functest_kw(nLayer=10):
    """functional_tests/mesh_versioning/json_mesh.json:82-99 as C-ABI config overrides."""
    return dict(nLayer=nLayer, smooth_coeff=6500, do_fixed_settling=1, settling_velocity=0.5, use_R94_lambda=0)
