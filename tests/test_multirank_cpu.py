"""N>1 path on CPU.  (1) Structural checks of the product's halo plan
(lulesh_b200_halo_plan_*, the host logic behind the NCCL exchanges) for 2/4/8-rank
layouts incl. non-cubic bricks.  (2) world_size-2 and -4 gloo runs in which real
messages are packed/unpacked with that plan around the oracle's per-phase functions,
checked against the oracle's in-process emulation of the reference's MPI semantics
and for bit-identical shared nodes (the property that replaces CommSyncPosVel)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def all_plans(lb, decomp, sizes):
    n = decomp[0] * decomp[1] * decomp[2]
    doms = [lb.Domain(sizes[0], num_ranks=n, rank=r, decomp=decomp, sizes=sizes) for r in range(n)]
    return doms, [lb.halo_plan(d) for d in doms]


@pytest.mark.parametrize("decomp,sizes", [((1, 1, 2), (4, 4, 2)), ((1, 2, 2), (4, 3, 2)),
                                          ((2, 2, 2), (3, 3, 3)), ((3, 3, 3), (2, 2, 2)),
                                          ((2, 1, 3), (2, 5, 3)), ((1, 3, 2), (1, 1, 1)), ((2, 2, 2), (1, 2, 1))])
def test_halo_plan_is_pairwise_consistent(lb, decomp, sizes):
    doms, plans = all_plans(lb, decomp, sizes)
    for r, (d, p) in enumerate(zip(doms, plans)):
        xyz = np.stack([d.field("x"), d.field("y"), d.field("z")], 1)
        nb = len(p["bnode"])
        assert len(set(p["msg_rank"])) == len(p["msg_rank"]) and r not in p["msg_rank"]
        for peer, cnt, soff in zip(p["msg_rank"], p["msg_count"], p["msg_send_off"]):
            q = plans[peer]
            j = list(q["msg_rank"]).index(r)
            assert q["msg_count"][j] == cnt
            # the same global nodes, in the same order, on both sides
            mine = p["bnode"][p["pack_idx"][soff:soff + cnt]]
            theirs = q["bnode"][q["pack_idx"][q["msg_send_off"][j]:q["msg_send_off"][j] + cnt]]
            pxyz = np.stack([doms[peer].field(n) for n in "xyz"], 1)
            assert np.array_equal(xyz[mine], pxyz[theirs])
            # field-major packing of three planes
            for a in range(3):
                seg = p["pack_idx"][soff + a * cnt: soff + (a + 1) * cnt]
                assert np.all(seg // nb == a)
        # every boundary node: own slot exactly once, sources in ascending rank order
        start, src = p["bsum_start"], p["bsum_src"].reshape(-1, 2)
        recv_rank = np.full(p["msg_recv_off"][-1] + 3 * p["msg_count"][-1] if len(p["msg_rank"]) else 0, -1)
        for peer, cnt, roff in zip(p["msg_rank"], p["msg_count"], p["msg_recv_off"]):
            recv_rank[roff:roff + 3 * cnt] = peer
        for b in range(nb):
            ranks = []
            for k in range(start[b], start[b + 1]):
                base, stride = src[k]
                ranks.append(r if base < 3 * nb else recv_rank[base])
                if base < 3 * nb:
                    assert base == b and stride == nb
            assert ranks == sorted(ranks) and len(set(ranks)) == len(ranks) and r in ranks
            assert 2 <= len(ranks) <= 8
        # MonoQ: my ghost slots are exactly what lzetam/letam/lxim ... point to
        ghosts = np.concatenate([d.ints(n) for n in "lxim lxip letam letap lzetam lzetap".split()])
        ghosts = np.unique(ghosts[ghosts >= d.numElem])
        want = np.concatenate([np.arange(g, g + c) for g, c in zip(p["face_ghost_off"], p["face_count"])]) \
            if len(p["face_rank"]) else np.zeros(0, int)
        assert np.array_equal(ghosts, want)


def test_halo_plan_rejects_bad_view(lb):
    d = lb.Domain(3)
    v = d.refresh_view()
    v.numRanks = 2   # inconsistent with px*py*pz == 1
    import ctypes as C
    p = C.c_void_p()
    assert lb._lib.lulesh_b200_halo_plan_create(C.byref(v), C.byref(p)) == lb.EINVAL


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("decomp,sizes,cycles", [((1, 1, 2), (6, 6, 3), 25), ((1, 2, 2), (4, 2, 2), 12)])
def test_gloo_ranks_with_product_halo_plan(lb, oracle_mod, tmp_path, decomp, sizes, cycles):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _multirank_worker import worker
    world = decomp[0] * decomp[1] * decomp[2]
    mp.spawn(worker, args=(world, _free_port(), decomp, sizes, cycles, str(tmp_path)), nprocs=world, join=True)
    got = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world)]

    ref = oracle_mod.OracleMulti(decomp, sizes)          # reference MPI semantics, in-process
    assert ref.run(cycles) == 0
    for r in range(world):
        o = ref.rank(r)
        assert int(got[r]["cycle"]) == o.scalars.cycle == cycles
        assert abs(float(got[r]["time"]) - o.scalars.time) <= 1e-14 * o.scalars.time
        for n in "x y z xd yd zd e p q v nodalMass".split():
            scale = max(np.max(np.abs(o.field(n))), 1e-300)
            assert np.max(np.abs(got[r][n] - o.field(n))) <= 1e-10 * scale, (r, n)
    # time control identical on every rank (F10) and shared nodes bit-identical across ranks
    assert len({float(g["time"]) for g in got}) == 1 and len({float(g["dt"]) for g in got}) == 1
    doms, plans = all_plans(lb, decomp, sizes)
    for r, p in enumerate(plans):
        for peer, cnt, soff in zip(p["msg_rank"], p["msg_count"], p["msg_send_off"]):
            q = plans[peer]
            j = list(q["msg_rank"]).index(r)
            mine = p["bnode"][p["pack_idx"][soff:soff + cnt]]
            theirs = q["bnode"][q["pack_idx"][q["msg_send_off"][j]:q["msg_send_off"][j] + cnt]]
            for n in "x y z xd yd zd nodalMass".split():
                assert np.array_equal(got[r][n][mine], got[peer][n][theirs]), (r, peer, n)
