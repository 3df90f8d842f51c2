"""Numpy model of the proposed GPU schedule: a 2-D grid of tiles in (along-wind, cross-wind) coordinates on the variable-resolution
mesh; columns are visited in down-wind order (multiplicative), the tiles of one column are updated CONCURRENTLY (each sees the
other tiles of its column as they were before the column started), k multicolour line-GS sweeps inside a tile."""
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


def run(ntri, ncols, nrows, ks, L=10):
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

    # equal-count bins: along the wind (x), then across it (y) inside each column
    ci = np.argsort(np.argsort(geo.cx[:T])) * ncols // T
    rj = np.zeros(T, int)
    for c in range(ncols):
        m = np.where(ci == c)[0]
        rj[m] = np.argsort(np.argsort(geo.cy[:T][m])) * nrows // len(m)
    tile_of = ci * nrows + rj
    print(f"variable mesh T={T}: {ncols} columns x {nrows} rows of ~{T // (ncols * nrows)} faces")
    for kin in ks:
        x = np.zeros((L, T)); kout = None
        for it in range(80):
            for c in range(ncols):
                frozen = x.copy()  # what the other tiles of this column look like while the column runs
                newx = x.copy()
                for r in range(nrows):
                    m = np.where(tile_of == c * nrows + r)[0]
                    xt = frozen.copy()
                    for kk in range(kin):
                        for cc in range(nc):
                            cols = m[col[m] == cc]
                            if len(cols) == 0: continue
                            acc = np.zeros((L, len(cols)))
                            for j in range(3):
                                acc += lat[j][:, cols] * np.where(has[cols, j][None, :], xt[:, nbs[cols, j]], 0.0)
                            xt[:, cols] = thomas(cols, rhs[:, cols] - acc)
                    newx[:, m] = xt[:, m]
                x = newx
            if resid(x) <= 1e-8: kout = it + 1; break
        print(f"   {kin} inner sweeps: outer iterations {kout}")


if __name__ == "__main__":
    run(30000, 10, 8, (4, 8))     # ~375-face tiles
    run(30000, 20, 8, (4, 8))     # ~190-face tiles
