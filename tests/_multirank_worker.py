"""Worker for tests/test_multirank_cpu.py: one process per rank over gloo.  Each rank
advances the ORACLE's per-phase functions on its own brick, while every exchange is
performed with the PRODUCT's halo plan (lulesh_b200_halo_plan_*) and real message
passing -- the CPU twin of the NCCL path in lulesh_b200/csrc/api.cu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def exchange(dist, torch, rank, send_parts, recv_sizes, peers):
    """post all irecv/isend for one exchange; returns list of received numpy arrays"""
    reqs, recv = [], []
    for peer, n in zip(peers, recv_sizes):
        t = torch.empty(n, dtype=torch.float64)
        recv.append(t)
        reqs.append(dist.irecv(t, src=int(peer)))
    for peer, part in zip(peers, send_parts):
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(part)), dst=int(peer)))
    for r in reqs:
        r.wait()
    return [t.numpy() for t in recv]


def canonical_sum(plan, own3, recvs):
    """own3: (3, nb) partials; returns (3, nb) totals summed in ascending-rank order."""
    nb = own3.shape[1]
    halo = np.concatenate([own3.reshape(-1)] + recvs)
    out = np.zeros((3, nb))
    start, src = plan["bsum_start"], plan["bsum_src"].reshape(-1, 2)
    for b in range(nb):
        for a in range(3):
            s = 0.0
            for k in range(start[b], start[b + 1]):
                s += halo[src[k, 0] + a * src[k, 1]]
            out[a, b] = s
    return out


def worker(rank, world, port, decomp, sizes, cycles, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import lulesh_b200 as lb
    import oracle as ora

    dist.init_process_group("gloo", rank=rank, world_size=world)
    dom = lb.Domain(sizes[0], 11, 1, 1, num_ranks=world, rank=rank, decomp=decomp, sizes=sizes)
    plan = lb.halo_plan(dom)
    o = ora.OracleDomain(sizes[0], 11, 1, 1, num_ranks=world, rank=rank, decomp=decomp, sizes=sizes)
    bnode, nb = plan["bnode"], len(plan["bnode"])
    peers, counts = plan["msg_rank"], plan["msg_count"]

    def node_exchange(own3):
        sendbuf = own3.reshape(-1)[plan["pack_idx"]]
        parts = [sendbuf[o_:o_ + 3 * c] for o_, c in zip(plan["msg_send_off"], counts)]
        recvs = exchange(dist, torch, rank, parts, [3 * c for c in counts], peers)
        return canonical_sum(plan, own3, recvs)

    # lulesh.cc:2720-2729: nodal mass (shipped as three identical planes, like the device path)
    m = o.field("nodalMass")
    m[bnode] = node_exchange(np.stack([m[bnode]] * 3))[0]

    for _ in range(cycles):
        g = torch.tensor([o.lib.ora_dt_candidate(o._p)], dtype=torch.float64)
        dist.all_reduce(g, op=dist.ReduceOp.MIN)            # lulesh.cc:186
        o.lib.ora_time_increment(o._p, float(g.item()))
        assert o.calc_force() == 0
        f = [o.field(n) for n in ("fx", "fy", "fz")]
        tot = node_exchange(np.stack([a[bnode] for a in f]))
        for a in range(3):
            f[a][bnode] = tot[a]
        o.node_update()
        assert o.kinematics() == 0
        o.monoq_gradients()
        dv = [o.field(n) for n in ("delv_xi", "delv_eta", "delv_zeta")]
        all_elem = dv[0].size
        flat = np.concatenate(dv)
        mq = flat[plan["mq_idx"]]
        parts = [mq[o_:o_ + 3 * c] for o_, c in zip(plan["face_send_off"], plan["face_count"])]
        recvs = exchange(dist, torch, rank, parts, [3 * c for c in plan["face_count"]], plan["face_rank"])
        for got, c, goff in zip(recvs, plan["face_count"], plan["face_ghost_off"]):
            for a in range(3):
                dv[a][goff:goff + c] = got[a * c:(a + 1) * c]
        assert all_elem == dom.numElem + 2 * (sizes[0] * sizes[1] + sizes[0] * sizes[2] + sizes[1] * sizes[2])
        assert o.monoq_regions() == 0
        assert o.material() == 0
        o.time_constraints()

    s = o.scalars
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), cycle=s.cycle, time=s.time, dt=s.deltatime,
             **{n: o.field(n).copy() for n in "x y z xd yd zd e p q v nodalMass".split()})
    dist.barrier()
    dist.destroy_process_group()
