"""Numpy model: two-grid cycle for the deposition system (V + eps*L) q = b with piecewise-constant aggregates (runs of `agg`
consecutive faces in Morton order), multicolour Gauss-Seidel/SOR smoothing, exact coarse solve; stationary iteration and as a CG
preconditioner (symmetric cycle).  Question: how many fine-level passes against the 124-134 SOR sweeps of the device solver?"""
import os, sys
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chm_b200 import synthetic


def dep_matrix(mesh, eps=6500.0):
    geo = mesh.geometry(); T = mesh.n_local
    rows, cols, vals = [], [], []
    diag = geo.area.copy()
    for j in range(3):
        n = mesh.neigh[:, j]; has = n >= 0
        dx = np.hypot(geo.cx[has] - geo.cx[n[has]], geo.cy[has] - geo.cy[n[has]])
        c = eps * geo.elen[j][has] / dx
        diag[has] += c
        rows.append(np.where(has)[0]); cols.append(n[has]); vals.append(-c)
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(T, T)) + sp.diags(diag)
    return A.tocsr(), diag, geo


def greedy(neigh):
    T = neigh.shape[0]; col = -np.ones(T, int)
    for i in range(T):
        used = {col[n] for n in neigh[i] if n >= 0 and col[n] >= 0}
        c = 0
        while c in used: c += 1
        col[i] = c
    return col


def run(ntri, aggs, omega_s=1.0):
    mesh = synthetic.variable_mesh(ntri); T = mesh.n_local
    A, d, geo = dep_matrix(mesh)
    rng = np.random.default_rng(0)
    b = np.sin(geo.cx / 700.0) * np.cos(geo.cy / 500.0) * geo.area + 0.1 * rng.standard_normal(T)
    bn = np.linalg.norm(b)
    col = greedy(mesh.neigh); nc = col.max() + 1
    masks = [np.where(col == c)[0] for c in range(nc)]
    Ac = [A[m] for m in masks]
    Dinv = 1.0 / d

    def gs(q, rhs, reverse=False):
        order = range(nc - 1, -1, -1) if reverse else range(nc)
        for c in order:
            m = masks[c]
            q[m] += omega_s * Dinv[m] * (rhs[m] - Ac[c] @ q)
        return q

    print(f"variable mesh T={T}, {nc} colours")
    for agg in aggs:
        P = sp.csr_matrix((np.ones(T), (np.arange(T), np.arange(T) // agg)), shape=(T, (T + agg - 1) // agg))
        Acoarse = (P.T @ A @ P).tocsc()
        lu = spla.splu(Acoarse)

        def cycle(rhs):  # symmetric two-grid cycle from a zero guess: pre-smooth, coarse correction, post-smooth (reversed)
            q = gs(np.zeros(T), rhs)
            q += P @ lu.solve(P.T @ (rhs - A @ q))
            return gs(q, rhs, reverse=True)

        # stationary
        q = np.zeros(T); ks = None
        for k in range(300):
            q += cycle(b - A @ q)
            if np.linalg.norm(b - A @ q) <= 1e-8 * bn: ks = k + 1; break
        # preconditioned CG
        q = np.zeros(T); r = b.copy(); z = cycle(r); p = z.copy(); rz = r @ z; kc = None
        for k in range(300):
            Ap = A @ p; al = rz / (p @ Ap); q += al * p; r -= al * Ap
            if np.linalg.norm(r) <= 1e-8 * bn: kc = k + 1; break
            z = cycle(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
        print(f"   aggregates of {agg:3d} faces (coarse n = {P.shape[1]}): stationary cycles {ks}, CG iterations {kc}"
              f"  (each = 2 smoothing passes + 2 residual/transfer passes on the fine level)")


if __name__ == "__main__":
    run(80000, (8, 16, 32, 64))
