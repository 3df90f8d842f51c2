"""Numpy model of the ACTIVE-SET line Gauss-Seidel sweep (gs_persistent_kernel with SolvePlan::use_live).

Every solve starts from x0 = 0 and the right-hand side is non-zero only on layer 0 of saltating faces.  A column update is
    x_p <- T_p^-1 (b_p - A_lat[p,:] x)
so a column with b_p = 0 whose three neighbour columns are still identically zero is updated to exactly zero: it can be skipped
without changing a single bit of any iterate.  The device tracks a superset of the non-zero columns with one byte per face:
    live[p] = 1   if b_p != 0 (written by the assembly), or once p has been updated while any neighbour was live;
    a column is updated  iff  live[p] | live[n0] | live[n1] | live[n2]   (read BEFORE the pass: neighbours have another colour).
The live set grows by one ring of faces per colour pass.  This script runs the sweep with and without the skip on the oracle's own
assembly, asserts that the iterates are IDENTICAL, and prints the share of column updates that were executed.

    python tests/models/active_set_model.py [n] [step]      (n x n squares -> 2 n^2 triangles; default 236)
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chm_b200 import synthetic
from oracle.pbsm3d_oracle import Config, PBSM3DOracle


def two_colouring(neigh):
    T = len(neigh)
    seen = -np.ones(T, int); seen[0] = 0; st = [0]
    while st:
        i = st.pop()
        for nn in neigh[i]:
            if nn >= 0 and seen[nn] < 0: seen[nn] = 1 - seen[i]; st.append(nn)
    return seen


def build(n, L=10, step=0, forcing=None):
    mesh = synthetic.uniform_mesh(n, n)
    geo = mesh.geometry()
    o = PBSM3DOracle(Config.functional_test(L), mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    f = forcing(geo.cx, geo.cy) if forcing else synthetic.forcing(geo.cx, geo.cy, step=step)
    asm = o.assemble(f, 3600.0)
    T = mesh.n_local
    inv = np.zeros((L, T)); cp = np.zeros((L, T)); prev = np.zeros(T)
    for z in range(L):
        inv[z] = 1.0 / (asm.diag[z] - asm.below[z] * prev); cp[z] = asm.above[z] * inv[z]; prev = cp[z]
    return dict(T=T, L=L, neigh=mesh.neigh, col=two_colouring(mesh.neigh), latS=asm.lat * inv[None], belowS=asm.below * inv, cp=cp,
                rhsS0=asm.rhs[0] * inv[0], rhs0=asm.rhs[0], A=o.suspension_csr(asm), b=asm.rhs.reshape(-1))


def solve(sy, skip, tol=1e-8, maxit=200):
    """Returns (sweeps, x, executed column updates per sweep)."""
    T, L = sy["T"], sy["L"]
    nb = sy["neigh"]; has = nb >= 0; nbs = np.where(has, nb, np.arange(T)[:, None])  # a missing neighbour points to the face itself
    x = np.zeros((L, T))
    live = (sy["rhs0"] != 0.0)
    bn = np.linalg.norm(sy["b"])
    done = []
    for k in range(maxit):
        n_upd = 0
        for cc in (0, 1):
            cols = np.where(sy["col"] == cc)[0]
            if skip:
                act = live[cols] | live[nbs[cols]].any(1)
                cols = cols[act]
            n_upd += len(cols)
            g = np.zeros((L, len(cols)))
            for j in range(3):
                g -= sy["latS"][j][:, cols] * np.where(has[cols, j][None, :], x[:, nbs[cols, j]], 0.0)
            g[0] += sy["rhsS0"][cols]
            y = g.copy()
            for z in range(1, L): y[z] = g[z] - sy["belowS"][z][cols] * y[z - 1]
            for z in range(L - 2, -1, -1): y[z] = y[z] - sy["cp"][z][cols] * y[z + 1]
            x[:, cols] = y
            if skip: live[cols] = True
        done.append(n_upd)
        if np.linalg.norm(sy["b"] - sy["A"] @ x.reshape(-1)) <= tol * bn:
            return k + 1, x, done
    return None, x, done


def patchy(cx, cy):
    """Saltation on wind-exposed patches only, as on a real winter day (synthetic.patchy_forcing)."""
    return synthetic.patchy_forcing(cx, cy)


def run(n=236, step=0):
    for name, fo in (("bench forcing", None), ("patchy saltation", patchy)):
        sy = build(n, step=step, forcing=fo)
        k0, x0, _ = solve(sy, False)
        k1, x1, done = solve(sy, True)
        assert k0 == k1 and np.array_equal(x0, x1), "the skip changed an iterate"
        share = np.array(done) / sy["T"]
        print(f"{name}: T={sy['T']}, rhs != 0 on {np.mean(sy['rhs0'] != 0):.3f} of the faces, {k0} sweeps, iterates identical; "
              f"column updates executed: {share.mean():.3f} of all (first sweep {share[0]:.3f}, last {share[-1]:.3f})")


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 236, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
