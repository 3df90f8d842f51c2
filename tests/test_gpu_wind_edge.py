"""-m gpu: fetchr's nearest-centre search for queries far outside the mesh (faces near the up-wind domain edge) and on
a tiny mesh (grid of one cell) — the cases where the ring search has to rely on its projection bound."""
import numpy as np
import pytest

from chm_b200 import capi, synthetic
from chm_b200.mesh import TriMesh
from oracle import wind_oracle as wo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2, 7, 40])
def test_fetch_with_queries_outside_the_domain(n):
    mesh = synthetic.uniform_mesh(n, n, h=25.0)  # 50 m .. 1 km wide: most of the 10 x 100 m steps leave the mesh
    vz = mesh.vertex.copy()
    vz[:, 2] = 1000.0 + 0.08 * (vz[:, 0] - vz[:, 0].min()) + 3.0 * np.sin(vz[:, 1] / 40.0)
    m = TriMesh(vz, mesh.elem, mesh.neigh, {})
    geo = m.geometry()
    rng = np.random.default_rng(n)
    h = capi.Handle(capi.default_config(nLayer=5), m)
    for trial in range(3):
        vw = rng.uniform(0.0, 360.0, m.n_local)
        assert np.array_equal(h.fetchr(vw), wo.fetchr(vw, geo.cx, geo.cy, geo.cz, None))
    wc = capi.default_wind_config(fetch_steps=25, fetch_max_distance=5000.0, fetch_I=0.01)
    vw = rng.uniform(0.0, 360.0, m.n_local)
    assert np.array_equal(h.fetchr(vw, wc), wo.fetchr(vw, geo.cx, geo.cy, geo.cz, None, steps=25, max_distance=5000.0, I=0.01))
    h.close()
