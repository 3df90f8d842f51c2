"""N>1 host logic on CPU: two processes (gloo) partition the mesh by CHM's rule, negotiate the halo the way
setup_nearest_neighbor_communication does (each rank sends the owner the global ids it needs), exchange one
per-face variable owner→ghost, and check every ghost holds its owner's value (triangulation.cpp:1845-2079)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_mesh


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from chm_b200.mesh import partition_mesh
        mesh = load_mesh("slope_metis")
        p = partition_mesh(mesh, rank, world)
        T = p.n_local
        start = int(p.global_id[0])
        # 1. everyone learns how many ghosts each rank needs from each owner
        need = np.bincount(p.ghost_owner, minlength=world).astype(np.int64)
        allneed = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allneed, torch.from_numpy(need))
        M = torch.stack(allneed).numpy()  # M[r][q]
        # 2. send the owner the global ids we need; receive the ids others need from us
        ghosts = p.global_id[T:]
        reqs, give = [], {}
        for q_ in range(world):
            if q_ == rank:
                continue
            if need[q_]:
                reqs.append(dist.isend(torch.from_numpy(ghosts[p.ghost_owner == q_].copy()), q_))
            if M[q_][rank]:
                give[q_] = torch.zeros(int(M[q_][rank]), dtype=torch.int64)
                reqs.append(dist.irecv(give[q_], q_))
        for r in reqs:
            r.wait()
        send_idx = {q_: (g.numpy() - start) for q_, g in give.items()}
        for idx in send_idx.values():
            assert (idx >= 0).all() and (idx < T).all()
        # 3. owner→ghost exchange of a variable whose value encodes the global id
        var = (np.arange(start, start + T) * 1.5 + 0.25)
        ghost_val = np.full(p.n_ghost, -9999.0)
        reqs, bufs = [], {}
        for q_ in range(world):
            if q_ == rank:
                continue
            if q_ in send_idx:
                reqs.append(dist.isend(torch.from_numpy(var[send_idx[q_]].copy()), q_))
            if need[q_]:
                bufs[q_] = torch.zeros(int(need[q_]), dtype=torch.float64)
                reqs.append(dist.irecv(bufs[q_], q_))
        for r in reqs:
            r.wait()
        for q_, b in bufs.items():
            ghost_val[p.ghost_owner == q_] = b.numpy()
        ok = bool(np.array_equal(ghost_val, ghosts * 1.5 + 0.25))
        # 4. a global reduction the way the Krylov dots / rhs max are reduced
        t = torch.tensor([float(var.sum()), float(T)], dtype=torch.float64)
        dist.all_reduce(t)
        G = mesh.n_local
        ok &= abs(t[0].item() - (np.arange(G) * 1.5 + 0.25).sum()) < 1e-6 and int(t[1].item()) == G
        q.put((rank, ok, p.n_ghost))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_halo_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(ok for _, ok, _ in res), res
    assert all(ng > 0 for _, _, ng in res)


# ------------------------------------------------------------------------------------------------------------------
# The deposition solve across ranks (sor_pass_halo_kernel): rank-local colours, global sweep order = the key
# (colour, rank), a ghost with a smaller key is read at this sweep's value, one with a larger key at the previous sweep's.
def _greedy_colours(neigh_local):
    T = neigh_local.shape[0]
    col = -np.ones(T, dtype=np.int64)
    for i in range(T):
        used = {col[n] for n in neigh_local[i] if 0 <= n < T and col[n] >= 0}
        c = 0
        while c in used:
            c += 1
        col[i] = c
    return col


def _dep_rows(p, geo_cx, geo_cy, elen, area, eps=6500.0):
    """Jacobi-scaled deposition rows of the owned faces: offS [T,3] (0 where no neighbour), from PBSM3D.cpp:1609-1628."""
    T = p.n_local
    off = np.zeros((T, 3))
    diag = area[:T].copy()
    for j in range(3):
        n = p.neigh[:, j]
        has = n >= 0
        dx = np.hypot(geo_cx[:T][has] - geo_cx[n[has]], geo_cy[:T][has] - geo_cy[n[has]])
        c = eps * elen[j][has] / dx
        diag[has] += c
        off[has, j] = -c
    return off / diag[:, None], diag


def _sor_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from chm_b200.mesh import partition_mesh
        mesh = load_mesh("slope_metis")
        G = mesh.n_local
        p = partition_mesh(mesh, rank, world)
        T, nG = p.n_local, p.n_ghost
        start = int(p.global_id[0])
        geo = p.geometry()  # owned + ghost centres
        offS, diag = _dep_rows(p, geo.cx, geo.cy, geo.elen, geo.area)
        gid = p.global_id
        b_all = np.sin(np.arange(G) * 0.37) + 0.3  # a right-hand side that depends on the global id only
        bS = b_all[start:start + T] / diag
        col = _greedy_colours(np.where(p.neigh < T, p.neigh, -1))
        key = col * world + rank
        # ghost keys + the send lists, negotiated as in test_two_rank_halo_over_gloo
        need = np.bincount(p.ghost_owner, minlength=world).astype(np.int64)
        allneed = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allneed, torch.from_numpy(need))
        M = torch.stack(allneed).numpy()
        other = 1 - rank
        req = [dist.isend(torch.from_numpy(gid[T:].copy()), other)]
        give = torch.zeros(int(M[other][rank]), dtype=torch.int64)
        req.append(dist.irecv(give, other))
        for r in req:
            r.wait()
        send_idx = give.numpy() - start

        def exchange(vals):  # owner -> ghost of one per-face array
            out = torch.zeros(nG, dtype=torch.float64)
            rq = [dist.isend(torch.from_numpy(np.ascontiguousarray(vals[send_idx], dtype=np.float64)), other), dist.irecv(out, other)]
            for r in rq:
                r.wait()
            return out.numpy()

        ghost_key = exchange(key.astype(np.float64)).astype(np.int64)
        ncmax = int(max(key.max(), ghost_key.max())) // world + 1
        omega, nsweeps = 1.7, 12
        qv = np.zeros(T)
        ghost_new = np.zeros(nG)   # value of this sweep (valid once its owner has passed that key)
        ghost_old = np.zeros(nG)   # value of the previous sweep
        for e in range(nsweeps):
            for c in range(ncmax):
                for r in range(world):  # the global order: key (c, r) ascending; only the rank whose turn it is updates
                    if r == rank:
                        for i in np.where(col == c)[0]:
                            z = bS[i] - qv[i]
                            for j in range(3):
                                n = p.neigh[i, j]
                                if n < 0:
                                    continue
                                if n < T:
                                    v = qv[n]
                                else:
                                    g = n - T
                                    v = ghost_new[g] if ghost_key[g] < c * world + rank else ghost_old[g]
                                z -= offS[i, j] * v
                            qv[i] += omega * z
                    # what the tagged stores do: the updated boundary values reach the partner before the next key starts
                    fresh = exchange(qv)
                    upd = ghost_key == c * world + r if r != rank else np.zeros(nG, dtype=bool)
                    ghost_new[upd] = fresh[upd]
            ghost_old = ghost_new.copy()
        q.put((rank, qv, key, start, T))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sor_order_equals_the_global_sweep():
    """(colour, rank)-ordered SOR with per-key ghost refresh on two ranks == sequential SOR over the global mesh in that
    order, bit for bit (the rule sor_pass_halo_kernel implements with tagged ghost entries)."""
    ctx = mp.get_context("spawn")
    qq = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sor_worker, args=(r, 2, port, qq)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([qq.get(timeout=280) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(30)
    mesh = load_mesh("slope_metis")
    G = mesh.n_local
    geo = mesh.geometry()
    off, diag = _dep_rows(mesh, geo.cx, geo.cy, geo.elen, geo.area)
    bS = (np.sin(np.arange(G) * 0.37) + 0.3) / diag
    key = np.concatenate([r[2] for r in res])
    part = np.concatenate([r[1] for r in res])
    # the keys define a proper sequential order: no two neighbours share one
    for j in range(3):
        n = mesh.neigh[:, j]
        has = n >= 0
        assert (key[has] != key[n[has]]).all()
    qv = np.zeros(G)
    for e in range(12):
        for k in np.unique(key):
            for i in np.where(key == k)[0]:
                z = bS[i] - qv[i]
                for j in range(3):
                    n = mesh.neigh[i, j]
                    if n >= 0:
                        z -= off[i, j] * qv[n]
                qv[i] += 1.7 * z
    assert np.array_equal(part, qv)
