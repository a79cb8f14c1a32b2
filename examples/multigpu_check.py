"""Multi-GPU parity check (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 examples/multigpu_check.py [--case cylinder|synth3d]

cylinder: every rank builds its `.ma2`-rule partition of the 2-D cylinder mesh; synth3d: z-slab partition of a window of the
extruded mesh (lx1 = 8, lxd = 12).  Over NCCL the ranks run dssum, one preconditioner application (fused Schwarz with its
ghost exchange + coarse), a short exptA matvec and a short Krylov-Schur run (multi-GPU Arnoldi); rank 0 compares with the
same computation on one GPU (all elements on rank 0's device).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neklab_b200 import api, build  # noqa: E402


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--case", default="cylinder", choices=["cylinder", "synth3d"]); a = ap.parse_args()
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        build.build()
    dist.barrier()
    case = bench.cylinder_inputs()
    if a.case == "cylinder":
        coords, vel, vertex, cbc, lxd = case["coords"], case["vel"], case["vertex"], case["cbc"], 9
        gllnid = api.partition(case["pid"], world)
        tau = 0.05
    else:
        L = 2 * world
        coords, vel, vertex, cbc = bench.extrude(bench.window_case(case, 3.0), 8, max(L, 3))
        lxd = 12
        E2 = coords.shape[0] // max(L, 3)
        gllnid = (np.arange(coords.shape[0]) // (E2 * max(L, 3) // world)).clip(0, world - 1).astype(np.int32)
        tau = 0.02
    d = coords.shape[1]
    sel = np.where(gllnid == rank)[0]
    mesh = api.Mesh(coords[sel], vertex, cbc, lxd, gllnid=gllnid, rank=rank, nranks=world)
    buf = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(buf, src=0)
    prm = dict(viscosity=1.0 / 50.0, torder=3, vtol=1e-12, ptol=1e-11, gmres_maxit=1000, pr_proj=0, rst_mode=1)
    ctx = api.Context(mesh, api.default_params(**prm), device=local, nccl_id=buf[0])
    # --- dssum on integer data (bit-exact)
    rng = np.random.default_rng(5)
    u_all = rng.integers(-100, 100, size=coords[:, 0].shape).astype(np.float64)
    mine = ctx.dssum(u_all[sel])
    # --- one preconditioner application on a rough residual
    q = coords.shape[-1] - 2
    r_all = np.random.default_rng(6).standard_normal((coords.shape[0], q if d == 3 else 1, q, q))
    z_mine = ctx.precond(r_all[sel])
    # --- short matvec
    bf = ctx.vec(); bf.upload([vel[sel, c] for c in range(d)])
    x_all = [np.sin(0.7 * coords[:, 0] + 0.3 * c) * np.cos(0.9 * coords[:, 1] - 0.2 * c) * (np.cos(2 * np.pi * coords[:, 2] / (0.5 * max(2 * world, 3))) if d == 3 else 1.0) for c in range(d)]
    x = ctx.vec(); x.upload([f[sel] for f in x_all])
    A = api.exptA_linop(ctx, tau, bf)
    y = A.matvec(x)
    v, pr, _ = y.download()
    nrm_loc = y.norm()                       # global (allreduced) norm
    st = A.stats()
    # --- short multi-GPU Arnoldi (device-resident basis partitioned like the mesh; dots by NCCL allreduce)
    x0 = ctx.vec(); x0.rand(True, 12345)
    re = api.eigs(A, 1, 8, tol=1e-7, x0=x0)
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(sel=sel, dssum=mine, z=z_mine, v=v, pr=pr, nrm=nrm_loc, stats=st, lam=complex(re["lam"][0]), neigh=[(r, len(g)) for r, g in mesh.neighbors()]))
    ok = True
    if rank == 0:
        ds = np.zeros_like(u_all); V = [np.zeros_like(u_all) for _ in range(d)]; Z = np.zeros_like(r_all)
        for g in gathered:
            ds[g["sel"]] = g["dssum"]; Z[g["sel"]] = g["z"]
            for c in range(d):
                V[c][g["sel"]] = g["v"][c]
        # single-GPU reference on this rank's device
        m1 = api.Mesh(coords, vertex, cbc, lxd)
        c1 = api.Context(m1, api.default_params(**prm), device=local)
        ds1 = c1.dssum(u_all)
        z1 = c1.precond(r_all)
        b1 = c1.vec(); b1.upload([vel[:, c] for c in range(d)])
        x1 = c1.vec(); x1.upload(x_all)
        A1 = api.exptA_linop(c1, tau, b1)
        y1 = A1.matvec(x1)
        v1, _, _ = y1.download()
        x01 = c1.vec(); x01.rand(True, 12345)
        re1 = api.eigs(A1, 1, 8, tol=1e-7, x0=x01)
        bm1 = m1.field("bm1")
        err = np.sqrt(sum(((V[c] - v1[c]) ** 2 * bm1).sum() for c in range(d)) / sum((v1[c] ** 2 * bm1).sum() for c in range(d)))
        zerr = float(np.abs(Z - z1).max() / np.abs(z1).max())
        lam_err = abs(gathered[0]["lam"] - complex(re1["lam"][0])) / abs(re1["lam"][0])
        res = dict(case=a.case, world=world, elements=int(coords.shape[0]), dssum_bit_exact=bool(np.array_equal(ds, ds1)), precond_rel_err=zerr, matvec_rel_err=float(err),
                   arnoldi_ritz_rel_err=float(lam_err), norm_multi=gathered[0]["nrm"], norm_single=y1.norm(),
                   steps=gathered[0]["stats"]["steps"], neighbours={i: g["neigh"] for i, g in enumerate(gathered)},
                   ms_multi=gathered[0]["stats"]["ms_total"], ms_single=A1.stats()["ms_total"])
        ok = res["dssum_bit_exact"] and err < 1e-9 and zerr < 1e-10 and lam_err < 1e-8
        res["ok"] = bool(ok)
        print(json.dumps(res), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open(f"gpurun_out/multigpu_check_{a.case}_{world}.json", "w"), indent=1)
        c1.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
