"""Worker for tests/test_gpu_multi.py: one rank per GPU under torch.distributed.run.

Each rank partitions the global mesh by CHM's rule, runs `nsteps` PBSM3D steps through the C-ABI with NCCL halos
and reductions, and rank 0 gathers the per-face outputs in global order into an npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from chm_b200 import capi, synthetic
    from chm_b200.mesh import partition_mesh
    from conftest import functest_kw, load_mesh

    out_path, meshname, L, solver, nsteps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())
    if meshname.startswith("alpine"):  # snow_slide across ranks: steep terrain, every rank runs init + run on its partition
        from oracle import slide_oracle as so
        n = int(meshname[6:])
        gmesh = synthetic.with_elevation(synthetic.uniform_mesh(n, n))
        ggeo = gmesh.geometry()
        slope = so.face_slope(gmesh.face_vertices().reshape(-1, 3, 3))
        sd, sdv, swe = so.synthetic_snow(ggeo.cx, ggeo.cy, slope, seed=n, deep=float(L) / 2)
        p = partition_mesh(gmesh, rank, world)
        T, s = p.n_local, int(p.global_id[0])
        h = capi.Handle(capi.default_config(nLayer=2), p, device=local, rank=rank, n_ranks=world, unique_id=uid)
        h.slide_init()
        gathered = {}
        for k in range(nsteps):
            o, st = h.slide_run(sd[s:s + T], sdv[s:s + T], swe[s:s + T])
            o["cos_slope"] = h.slide_constants()[1]
            for name, a in o.items():
                sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
                dist.all_gather(sizes, torch.tensor([T], dtype=torch.int64, device="cuda"))
                parts = [torch.zeros(int(n_.item()), dtype=torch.float64, device="cuda") for n_ in sizes]
                dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(a)).cuda())
                gathered[f"{name}_{k}"] = torch.cat(parts).cpu().numpy()
            gathered[f"stats_{k}"] = np.array([st["iterations"], st["wavefront_rounds"], st["faces_fired"]])
        if rank == 0:
            np.savez(out_path, **gathered)
        h.close()
        dist.destroy_process_group()
        return
    if meshname.startswith("uniform"):
        n = int(meshname[7:])
        gmesh = synthetic.uniform_mesh(n, n)
    else:
        gmesh = load_mesh(meshname)
    ggeo = gmesh.geometry()
    p = partition_mesh(gmesh, rank, world)
    T = p.n_local
    s = int(p.global_id[0])
    h = capi.Handle(capi.default_config(solver=solver, tolerance=1e-10, **functest_kw(L)), p, device=local, rank=rank,
                    n_ranks=world, unique_id=uid)
    gathered = {}
    for k in range(nsteps):
        Fg = synthetic.forcing(ggeo.cx, ggeo.cy, seed=7, step=k, calm=(k == 1))
        F = {n: v[s:s + T] for n, v in Fg.items()}
        outs, st = h.step(3600.0, F)
        x = h.solution()
        pieces = dict(outs, **{f"c{z}": x[z] for z in range(L)})
        for name, a in pieces.items():
            sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([T], dtype=torch.int64, device="cuda"))
            parts = [torch.zeros(int(n_.item()), dtype=torch.float64, device="cuda") for n_ in sizes]
            dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(a)).cuda())
            gathered[f"{name}_{k}"] = torch.cat(parts).cpu().numpy()
        gathered[f"iters_{k}"] = np.array([st["suspension_iterations"], st["deposition_iterations"], st["suspension_present"],
                                           st["deposition_present"]])
    # scale_wind_vert in domain mode on the partition: its neighbour spline needs the partners' point-scaled values (halo)
    Fg = synthetic.forcing(ggeo.cx, ggeo.cy, seed=7, step=0)
    u2 = h.scale_wind_vert(Fg["U_R"][s:s + T], Fg["snowdepthavg"][s:s + T])
    import time
    calls = []
    for _ in range(7):  # the halo of the point-scaled wind must not stall at any rank count (round 1 reported 12 ms at 8 ranks)
        dist.barrier()
        t0 = time.perf_counter()
        h.scale_wind_vert(Fg["U_R"][s:s + T], Fg["snowdepthavg"][s:s + T])
        calls.append(time.perf_counter() - t0)
    tmed = torch.tensor([float(np.median(calls[2:]))], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmed, op=dist.ReduceOp.MAX)
    gathered["scale_wind_vert_call_ms"] = np.array(1e3 * float(tmed.item()))
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([T], dtype=torch.int64, device="cuda"))
    parts = [torch.zeros(int(n_.item()), dtype=torch.float64, device="cuda") for n_ in sizes]
    dist.all_gather(parts, torch.from_numpy(u2).cuda())
    gathered["u2_domain"] = torch.cat(parts).cpu().numpy()
    gathered["halo_transport"] = np.array(st["halo_transport"])
    gathered["halo_fused"] = np.array(st["halo_fused"])
    gathered["halo_exchanges"] = np.array(st["halo_exchanges"])
    if rank == 0:
        np.savez(out_path, **gathered)
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
