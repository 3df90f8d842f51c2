"""Numpy model: the down-wind ordered multiplicative tile iteration on the VARIABLE-RESOLUTION mesh (irregular adjacency, greedy
colours), tiles = runs of consecutive faces in Morton (CHM) order, ordered by the projection of their centroid on the mean wind."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chm_b200 import synthetic
from oracle.pbsm3d_oracle import Config, PBSM3DOracle


def greedy(neigh):
    T = neigh.shape[0]; col = -np.ones(T, int)
    for i in range(T):
        used = {col[n] for n in neigh[i] if n >= 0 and col[n] >= 0}
        c = 0
        while c in used: c += 1
        col[i] = c
    return col


def run(ntri, tile, ks, L=10, strips=False):
    mesh = synthetic.variable_mesh(ntri); T = mesh.n_local
    geo = mesh.geometry()
    o = PBSM3DOracle(Config.functional_test(L), mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    F = synthetic.forcing(geo.cx, geo.cy)
    asm = o.assemble(F, 3600.0)
    diag, lat, below, above, rhs = asm.diag, asm.lat, asm.below, asm.above, asm.rhs
    col = greedy(mesh.neigh); nc = col.max() + 1
    nb = mesh.neigh; has = nb >= 0; nbs = np.where(has, nb, 0)
    A = o.suspension_csr(asm); b = rhs.reshape(-1); bn = np.linalg.norm(b)
    resid = lambda x: np.linalg.norm(b - A @ x.reshape(-1)) / bn

    def thomas(cols, g):
        d = diag[:, cols].copy(); lo = below[:, cols]; up = above[:, cols]; y = g.copy()
        cp = np.zeros_like(d); cp[0] = up[0] / d[0]; y[0] = y[0] / d[0]
        for z in range(1, L):
            den = d[z] - lo[z] * cp[z - 1]; cp[z] = up[z] / den; y[z] = (y[z] - lo[z] * y[z - 1]) / den
        for z in range(L - 2, -1, -1): y[z] -= cp[z] * y[z + 1]
        return y

    def sweep(cols_all, x):
        for cc in range(nc):
            cols = cols_all[col[cols_all] == cc]
            if len(cols) == 0: continue
            acc = np.zeros((L, len(cols)))
            for j in range(3):
                acc += lat[j][:, cols] * np.where(has[cols, j][None, :], x[:, nbs[cols, j]], 0.0)
            x[:, cols] = thomas(cols, rhs[:, cols] - acc)

    x = np.zeros((L, T)); kg = None
    allf = np.arange(T)
    for k in range(300):
        sweep(allf, x)
        if resid(x) <= 1e-8: kg = k + 1; break
    # tiles: runs of consecutive faces in Morton order, or (strips) bins of the centroid's projection on the mean wind
    tid = (np.argsort(np.argsort(geo.cx[:T])) // tile) if strips else (np.arange(T) // tile)
    ntile = tid.max() + 1
    # mean wind blows towards (-cos, -sin)(450 - 270) = +x: order tiles by mean centroid x
    order = np.argsort([geo.cx[:T][tid == t].mean() for t in range(ntile)])
    members = [np.where(tid == t)[0] for t in range(ntile)]
    print(f"variable mesh T={T} colours {nc}: global multicolour GS sweeps {kg}; {ntile} {'wind-perpendicular strips' if strips else 'Morton-run tiles'} of {tile} faces")
    for kin in ks:
        x = np.zeros((L, T)); kout = None
        for it in range(80):
            for t in order:
                for kk in range(kin): sweep(members[t], x)
            if resid(x) <= 1e-8: kout = it + 1; break
        print(f"   {kin} inner sweeps: outer (down-wind ordered) iterations {kout}")


if __name__ == "__main__":
    run(30000, 400, (2, 4, 8, 12))
    run(30000, 400, (4, 8), strips=True)
