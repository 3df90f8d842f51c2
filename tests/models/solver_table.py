"""Reproduces the iteration-count table of DESIGN.md §3 ("Why these solvers"): the suspension system of one step, solved from
x0 = 0 to ||b - Ax||2 <= 1e-8 ||b||2 by the candidate schemes, on the oracle's own assembly (oracle/pbsm3d_oracle.py).

    python tests/models/solver_table.py [side=100] [variable_faces=20000]

Schemes (T = the vertical tridiagonal blocks of A, A_lat = the three lateral couplings per row):
  jacobi      x <- T^-1 (b - A_lat x)                                   column-block Jacobi = line relaxation
  gs          the same, colour by colour (proper colouring of the faces)  multicolour line Gauss-Seidel  [shipped]
  gs w=1.1    line SOR with over-relaxation
  bicgstab    right-preconditioned BiCGStab, M = T                       [shipped fallback]
  gmres30     right-preconditioned GMRES(30), M = T
  gmres30+gs  right-preconditioned GMRES(30), M = one multicolour line-GS sweep   (the sweep as a Krylov preconditioner)
  aa(5)+gs    Anderson acceleration, depth 5, of the line-GS fixed-point map
Cost model beside the counts (bytes per row per iteration on the device, 58 B = one fp64-stream sweep, 16 B per extra vector pass):
a Krylov / Anderson iteration pays for its work vectors, so it must cut the count by more than its traffic ratio to win.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from chm_b200 import synthetic  # noqa: E402
from oracle.pbsm3d_oracle import Config, PBSM3DOracle, gmres_right  # noqa: E402

TOL = 1e-8


def colouring(neigh):
    T = neigh.shape[0]
    col = -np.ones(T, dtype=np.int64)
    for f in range(T):  # greedy first-fit in face order (what the library does when the dual graph is not bipartite)
        used = {col[n] for n in neigh[f] if n >= 0}
        c = 0
        while c in used:
            c += 1
        col[f] = c
    return col


def system(mesh, cfg, L, seed=7):
    geo = mesh.geometry()
    o = PBSM3DOracle(cfg, mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    asm = o.assemble(synthetic.forcing(geo.cx, geo.cy, seed=seed), 3600.0)
    A = o.suspension_csr(asm).tocsr()
    T = mesh.n_local
    b = np.zeros(L * T)
    b[:T] = asm.rhs[0]
    row = np.arange(L * T)
    same_col = sp.csr_matrix((np.ones(L * T), (row, row % T)), shape=(L * T, T))
    mask = (same_col @ same_col.T).tocsr()            # entries within one face column
    Tm = A.multiply(mask).tocsc()                     # vertical tridiagonal blocks
    Alat = (A - Tm.tocsr()).tocsr()
    return A, b, Tm, Alat, T


def run_schemes(mesh, cfg, L):
    A, b, Tm, Alat, T = system(mesh, cfg, L)
    bn = np.linalg.norm(b)
    lu = spla.splu(Tm)
    res = lambda x: np.linalg.norm(b - A @ x) / bn
    out = {}
    # Jacobi line relaxation
    x, k = np.zeros_like(b), 0
    while res(x) > TOL and k < 2000:
        x = lu.solve(b - Alat @ x)
        k += 1
    out["jacobi"] = k
    # multicolour line GS / SOR
    col = colouring(mesh.neigh)
    ncol = int(col.max()) + 1
    rows_of = [np.concatenate([z * T + np.flatnonzero(col == c) for z in range(L)]) for c in range(ncol)]
    lus = [spla.splu(Tm[r][:, r].tocsc()) for r in rows_of]
    Alat_c = [Alat[r] for r in rows_of]

    def gs_sweep(x, omega=1.0, rhs=b):
        for c in range(ncol):
            r = rows_of[c]
            xn = lus[c].solve(rhs[r] - Alat_c[c] @ x)
            x[r] = xn if omega == 1.0 else x[r] + omega * (xn - x[r])
        return x

    for name, om in (("gs", 1.0), ("gs w=1.1", 1.1)):
        x, k = np.zeros_like(b), 0
        while res(x) > TOL and k < 600:
            x = gs_sweep(x, om)
            k += 1
        out[name] = k if k < 600 else ">600"
    out["colours"] = ncol
    # BiCGStab, M = T (right preconditioning)
    its = [0]
    M = spla.LinearOperator(A.shape, matvec=lu.solve)

    def cb(_):
        its[0] += 1
    AM = spla.LinearOperator(A.shape, matvec=lambda v: A @ lu.solve(v))
    y, info = spla.bicgstab(AM, b, x0=np.zeros_like(b), rtol=TOL * 0.5, atol=0.0, maxiter=2000, callback=cb)
    out["bicgstab"] = its[0] if res(lu.solve(y)) <= TOL * 1.01 else f"{its[0]} (res {res(lu.solve(y)):.1e})"
    # GMRES(30), M = T and M = one GS sweep from zero
    _, k = gmres_right(A, b, lu.solve, tol=TOL, restart=30, maxiter=1000)
    out["gmres30"] = k
    _, k = gmres_right(A, b, lambda v: gs_sweep(np.zeros_like(v), 1.0, v), tol=TOL, restart=30, maxiter=1000)
    out["gmres30+gs"] = k
    # Anderson acceleration (depth 5) of the GS map
    m = 5
    x, k = np.zeros_like(b), 0
    dX, dF = [], []
    g_prev = f_prev = None
    while res(x) > TOL and k < 600:
        g = gs_sweep(x.copy())
        f = g - x
        if f_prev is not None:
            dF.append(f - f_prev)
            dX.append(g - g_prev)
            dF, dX = dF[-m:], dX[-m:]
        f_prev, g_prev = f, g
        if dF:
            Fm = np.stack(dF, axis=1)
            gamma = np.linalg.lstsq(Fm, f, rcond=None)[0]
            x = g - np.stack(dX, axis=1) @ gamma
        else:
            x = g
        k += 1
    out["aa(5)+gs"] = k if k < 600 else ">600"
    return out


# device bytes per row per iteration (fp64 streams): sweep 58; + 16 per extra read+write pass over a vector, 8 per read
COST = {"jacobi": 58 + 16, "gs": 58, "gs w=1.1": 58, "bicgstab": 2 * 58 + 2 * 40 + 6 * 16, "gmres30": 58 + 40 + 15 * 16,
        "gmres30+gs": 2 * 58 + 15 * 16, "aa(5)+gs": 58 + 16 + (2 * 5 + 2) * 8 + 5 * 16}


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nvar = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    L = 10
    cases = [(f"uniform {side}^2, functional-test block", synthetic.uniform_mesh(side, side), Config.functional_test(L)),
             (f"uniform {side}^2, code-default block", synthetic.uniform_mesh(side, side), Config(nLayer=L)),
             (f"variable-resolution {nvar}, functional-test block", synthetic.variable_mesh(nvar), Config.functional_test(L))]
    names = ["jacobi", "gs", "gs w=1.1", "bicgstab", "gmres30", "gmres30+gs", "aa(5)+gs"]
    print("| scheme | B/row/iteration | " + " | ".join(c[0] for c in cases) + " |")
    print("|---|---|" + "---|" * len(cases))
    res = [run_schemes(m, cfg, L) for _, m, cfg in cases]
    for n in names:
        cells = []
        for r in res:
            k = r[n]
            cells.append(f"{k} ({k * COST[n] / 1000:.1f} kB/row)" if isinstance(k, (int, np.integer)) else str(k))
        print(f"| {n} | {COST[n]} | " + " | ".join(cells) + " |")
    print("colours:", [r["colours"] for r in res])


if __name__ == "__main__":
    main()
