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
