"""Numpy model: line Gauss-Seidel sweeps with x STORED in fp32 for the first n32 sweeps (arithmetic fp64, coefficients rounded to
fp32 as the device does), then fp64 x.  Does the sweep count to 1e-8 change?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chm_b200 import synthetic
from oracle.pbsm3d_oracle import Config, PBSM3DOracle


def run(n, L=10):
    mesh = synthetic.uniform_mesh(n, n); T = mesh.n_local
    geo = mesh.geometry()
    o = PBSM3DOracle(Config.functional_test(L), mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    asm = o.assemble(synthetic.forcing(geo.cx, geo.cy), 3600.0)
    diag, lat, below, above, rhs = asm.diag, asm.lat, asm.below, asm.above, asm.rhs
    col = np.zeros(T, int)
    seen = -np.ones(T, int); seen[0] = 0; st = [0]
    while st:
        i = st.pop()
        for nn in mesh.neigh[i]:
            if nn >= 0 and seen[nn] < 0: seen[nn] = 1 - seen[i]; st.append(nn)
    col = seen
    nb = mesh.neigh; has = nb >= 0; nbs = np.where(has, nb, 0)
    A = o.suspension_csr(asm); b = rhs.reshape(-1); bn = np.linalg.norm(b)
    resid = lambda x: np.linalg.norm(b - A @ x.reshape(-1).astype(np.float64)) / bn
    # Thomas factors, row-scaled streams (as the assembly kernel stores them), fp64 and fp32-rounded
    inv = np.zeros((L, T)); cp = np.zeros((L, T))
    prev = np.zeros(T)
    for z in range(L):
        inv[z] = 1.0 / (diag[z] - below[z] * prev); cp[z] = above[z] * inv[z]; prev = cp[z]
    latS = lat * inv[None]; belowS = below * inv; rhsS0 = rhs[0] * inv[0]
    r32 = lambda a: a.astype(np.float32).astype(np.float64)

    def sweep(x, c32):
        lS, bS, cP = (r32(latS), r32(belowS), r32(cp)) if c32 else (latS, belowS, cp)
        for cc in (0, 1):
            cols = np.where(col == cc)[0]
            g = np.zeros((L, len(cols)))
            for j in range(3):
                g -= lS[j][:, cols] * np.where(has[cols, j][None, :], x[:, nbs[cols, j]].astype(np.float64), 0.0)
            g[0] += rhsS0[cols]
            y = g.copy()
            for z in range(1, L): y[z] = g[z] - bS[z][cols] * y[z - 1]
            for z in range(L - 2, -1, -1): y[z] = y[z] - cP[z][cols] * y[z + 1]
            x[:, cols] = y.astype(x.dtype)
        return x

    def solve(n_x32, n_c32):
        x = np.zeros((L, T), dtype=np.float32 if n_x32 > 0 else np.float64)
        for k in range(200):
            if k == n_x32 and x.dtype == np.float32: x = x.astype(np.float64)
            x = sweep(x, k < n_c32)
            if k + 1 >= max(n_x32, n_c32) and resid(x) <= 1e-8: return k + 1
        return None

    base = solve(0, 0)
    print(f"n={n} T={T}: all fp64: {base} sweeps")
    for n_c32, n_x32 in ((base - 4, 0), (base - 4, base - 10), (base - 4, base - 7), (base - 4, base - 5)):
        print(f"   fp32 coefficient streams for {n_c32} sweeps, fp32 x for the first {n_x32}: {solve(n_x32, n_c32)} sweeps")


if __name__ == "__main__":
    run(100)
