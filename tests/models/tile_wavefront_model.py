"""Numpy model: tiles of the suspension system updated in DOWN-WIND order (multiplicative between tiles, k multicolour line-GS
sweeps inside a tile).  Question: how many passes over the coefficients does a wavefront-over-tiles solver need?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chm_b200 import synthetic
from oracle.pbsm3d_oracle import Config, PBSM3DOracle


def colour2(neigh):
    T = neigh.shape[0]; col = -np.ones(T, int)
    for s in range(T):
        if col[s] >= 0: continue
        col[s] = 0; st = [s]
        while st:
            i = st.pop()
            for n in neigh[i]:
                if n >= 0 and col[n] < 0: col[n] = 1 - col[i]; st.append(n)
    return col


def run(n, side, ks, L=10):
    mesh = synthetic.uniform_mesh(n, n); T = mesh.n_local
    geo = mesh.geometry()
    o = PBSM3DOracle(Config.functional_test(L), mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    F = synthetic.forcing(geo.cx, geo.cy)
    asm = o.assemble(F, 3600.0)
    diag, lat, below, above, rhs = asm.diag, asm.lat, asm.below, asm.above, asm.rhs
    col = colour2(mesh.neigh)
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

    # square tiles of side x side squares; mean wind blows from 270 deg (towards +x): order tiles by column, then row
    h = 30.0
    ix = ((geo.cx[:T] - geo.cx[:T].min()) / h).astype(int) // side
    iy = ((geo.cy[:T] - geo.cy[:T].min()) / h).astype(int) // side
    ncol = ix.max() + 1
    tiles = {}
    for c in range(ncol):
        for r in range(iy.max() + 1):
            m = np.where((ix == c) & (iy == r))[0]
            if len(m): tiles[(c, r)] = m
    print(f"n={n} T={T} tiles {len(tiles)} of ~{np.mean([len(v) for v in tiles.values()]):.0f} faces, {ncol} tile columns along the wind")
    for kin in ks:
        x = np.zeros((L, T)); kout = None
        for it in range(60):
            for c in range(ncol):                      # down-wind order of the tile columns
                for key, m in tiles.items():
                    if key[0] != c: continue
                    for kk in range(kin):
                        for cc in (0, 1):
                            cols = m[col[m] == cc]
                            acc = np.zeros((L, len(cols)))
                            for j in range(3):
                                acc += lat[j][:, cols] * np.where(has[cols, j][None, :], x[:, nbs[cols, j]], 0.0)
                            x[:, cols] = thomas(cols, rhs[:, cols] - acc)
            if resid(x) <= 1e-8: kout = it + 1; break
        print(f"   {kin} inner sweeps: outer (down-wind ordered) iterations {kout}")


if __name__ == "__main__":
    run(120, 8, (1, 2, 4, 8))
    run(120, 16, (2, 4, 8))
