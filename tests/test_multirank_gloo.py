"""N>1 host logic on CPU: two gloo ranks build their partitions with the C-ABI mesh setup and run the direct-stiffness
exchange plan (local segmented sum + neighbour send/recv + unpack-add) in numpy; result must equal the single-rank
oracle dssum.  (The device path does the same with pack kernel -> ncclSend/Recv -> unpack kernel.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from neklab_b200 import api
    from tests.util import cylinder_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    om, bf, prm, z = cylinder_case()
    gllnid = api.partition(z["pid"], world)
    sel = np.where(gllnid == rank)[0]
    m = api.Mesh(om.coords[sel], om.vertex, om.cbc_v, 9, gllnid=gllnid, rank=rank, nranks=world)
    glo = m.glo_num()
    assert np.array_equal(glo, om.glo[sel])                         # local numbering == slice of the global numbering
    rng = np.random.default_rng(7)
    u_all = rng.integers(-50, 50, size=om.bm1.shape).astype(np.float64)
    u = u_all[sel].copy()
    # local gather-scatter
    uq, inv = np.unique(glo.ravel(), return_inverse=True)
    s = np.bincount(inv, weights=u.ravel())
    # neighbour exchange on the plan's shared global ids (same order on both sides)
    nbs = m.neighbors()
    recv = {}
    reqs = []
    for r, gids in nbs:
        pos = np.searchsorted(uq, gids)
        assert np.array_equal(uq[pos], gids)
        send = torch.from_numpy(s[pos].copy())
        buf = torch.zeros(len(gids), dtype=torch.float64)
        recv[r] = (pos, buf)
        reqs.append(dist.isend(send, dst=r)); reqs.append(dist.irecv(buf, src=r))
    for rq in reqs:
        rq.wait()
    for r in sorted(recv):
        pos, buf = recv[r]
        np.add.at(s, pos, buf.numpy())
    out = s[inv].reshape(u.shape)
    ref = om.dssum(u_all)[sel]
    ok = bool(np.array_equal(out, ref))
    q.put((rank, ok, len(nbs), int(sum(len(g) for _, g in nbs))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_dssum_matches_single_rank(nlk_lib, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert all(nn >= 1 for _, _, nn, _ in res)


def _worker_zslab(rank, world, port, q):
    """The strong-scaling layout of bench.py: z-slabs of a periodic extruded mesh (every rank generates only its own layers)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bench
    from neklab_b200 import api
    from oracle.mesh import SEMesh
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, n = 8, 4
    case = bench.window_case(bench.cylinder_inputs(), 2.0)
    E2 = case["coords"].shape[0]
    bounds, gllnid = bench.zslab_partition(E2, L, world)
    l0, l1 = int(bounds[rank]), int(bounds[rank + 1])
    coords, _, vertex, cbc = bench.extrude(case, n, L, layer_range=(l0, l1))
    full, _, _, _ = bench.extrude(case, n, L)
    om = SEMesh(full, vertex, cbc, 6)
    sel = np.where(gllnid == rank)[0]
    assert np.array_equal(full[sel], coords)                        # the rank's own layers, in global element order
    m = api.Mesh(coords, vertex, cbc, 6, gllnid=gllnid, rank=rank, nranks=world)
    glo = m.glo_num()
    u_all = np.random.default_rng(7).integers(-50, 50, size=om.bm1.shape).astype(np.float64)
    uq, inv = np.unique(glo.ravel(), return_inverse=True)
    s = np.bincount(inv, weights=u_all[sel].ravel())
    nbs = m.neighbors()
    recv = {}; reqs = []
    for r, gids in nbs:
        pos = np.searchsorted(uq, gids)
        assert np.array_equal(uq[pos], gids)
        buf = torch.zeros(len(gids), dtype=torch.float64)
        recv[r] = (pos, buf)
        reqs.append(dist.isend(torch.from_numpy(s[pos].copy()), dst=r)); reqs.append(dist.irecv(buf, src=r))
    for rq in reqs:
        rq.wait()
    for r in sorted(recv):
        pos, buf = recv[r]
        np.add.at(s, pos, buf.numpy())
    ok = bool(np.array_equal(s[inv].reshape(u_all[sel].shape), om.dssum(u_all)[sel]))
    want = sorted({(rank - 1) % world, (rank + 1) % world} - {rank})      # periodic in z: the first and the last slab are neighbours
    q.put((rank, ok and sorted(r for r, _ in nbs) == want, len(nbs), [int(b) for b in bounds]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_zslab_partition_exchange_plan(nlk_lib, world):
    """bench.py's strong-scaling partition (uneven slabs for world = 3, one layer per rank for world = 8, periodic wrap)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_zslab, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert sum(1 for _ in res) == world and res[0][3][-1] == 8 and res[0][3][0] == 0
