"""Pins the CPU oracle (oracle/lulesh_oracle.c) to the UNMODIFIED reference.

The goldens in tests/golden/ were produced by oracle/_ref (the reference compiled
from /root/reference by oracle/Makefile, see tests/golden/make_goldens.py).  The
oracle keeps the reference's floating-point association and, like the reference
Makefile:24 build, forms no FMA, so every comparison here is BIT-EXACT."""
import numpy as np
import pytest

from conftest import load_npz

CHECK_KEYS = ("cycles e0 time dt dtcourant dthydro sum_e sum_p sum_q sum_v sum_ss sum_xyz "
              "sum_absvel max_abs_diff total_abs_diff max_rel_diff").split()


def run_oracle(mod, nx, its=9999999, r=11, b=1, c=1):
    d = mod.OracleDomain(nx, r, b, c)
    assert d.run(its) == 0
    s = d.scalars
    sym = d.symmetry(nx)
    f = d.field
    rec = {"cycles": s.cycle, "e0": f("e")[0], "time": s.time, "dt": s.deltatime,
           "dtcourant": s.dtcourant, "dthydro": s.dthydro}
    # the reference shim sums sequentially in index order (oracle/ref_util_wrap.cc)
    seq = lambda a: float(np.add.accumulate(a)[-1]) if False else float(_seqsum(a))
    rec.update(sum_e=seq(f("e")), sum_p=seq(f("p")), sum_q=seq(f("q")), sum_v=seq(f("v")),
               sum_ss=seq(f("ss")), sum_xyz=seq(f("x") + f("y") + f("z")),
               sum_absvel=seq(np.abs(f("xd")) + np.abs(f("yd")) + np.abs(f("zd"))),
               max_abs_diff=sym[0], total_abs_diff=sym[1], max_rel_diff=sym[2],
               regions=[len(d.region_list(i)) for i in range(r)])
    return rec


def _seqsum(a):
    s = 0.0
    for v in a.tolist():
        s += v
    return s


@pytest.mark.parametrize("key,args", [
    ("lulesh_omp -s 30 -i 100", dict(nx=30, its=100)),
    ("lulesh_omp -s 5", dict(nx=5)),
    ("lulesh_omp -s 10", dict(nx=10)),
    ("lulesh_omp -s 8 -i 10", dict(nx=8, its=10)),
    ("lulesh_omp -s 12 -i 40 -r 16 -b 1 -c 8", dict(nx=12, its=40, r=16, b=1, c=8)),
    ("lulesh_omp -s 12 -i 40 -r 1 -c 0", dict(nx=12, its=40, r=1, c=0)),
    ("lulesh_omp -s 12 -i 40 -r 21 -b 2 -c 3", dict(nx=12, its=40, r=21, b=2, c=3)),
])
def test_oracle_bit_identical_to_reference(oracle_mod, goldens, key, args):
    gold = goldens[key]
    rec = run_oracle(oracle_mod, **args)
    for k in CHECK_KEYS:
        assert rec[k] == gold[k], f"{key}: {k} oracle {rec[k]!r} != reference {gold[k]!r}"
    assert rec["regions"] == gold["regions"]


def test_threaded_reference_differs_from_serial_only_by_roundoff(goldens):
    a, b = goldens["lulesh_omp -s 30 -i 100"], goldens["lulesh_serial -s 30 -i 100"]
    assert a["cycles"] == b["cycles"] == 100
    assert abs(a["e0"] - b["e0"]) / a["e0"] < 1e-14


def test_region_flags_never_change_the_answer(goldens):
    base = goldens["lulesh_omp -s 12 -i 40 -r 1 -c 0"]
    for k in ("lulesh_omp -s 12 -i 40 -r 16 -b 1 -c 8", "lulesh_omp -s 12 -i 40 -r 21 -b 2 -c 3"):
        for name in ("e0", "sum_e", "sum_p", "sum_q"):
            assert goldens[k][name] == base[name]


def test_oracle_arrays_bit_identical_after_9_and_10_cycles(oracle_mod):
    for cyc in (9, 10):
        ref = load_npz(f"ref_s8_c{cyc}.npz")
        d = oracle_mod.OracleDomain(8)
        assert d.run(cyc) == 0
        for name in ("x y z xd yd zd xdd ydd zdd fx fy fz nodalMass e p q ql qq v volo vnew delv "
                     "vdov arealg ss elemMass").split():
            assert np.array_equal(d.field(name), ref[name]), f"cycle {cyc}: {name}"
        for name in "lxim lxip letam letap lzetam lzetap elemBC regNumList nodelist symmX symmY symmZ".split():
            assert np.array_equal(d.ints(name), ref[name]), name


def test_one_cycle_from_reference_state(oracle_mod):
    """Inject the reference's cycle-9 state, advance one cycle phase by phase, compare with
    the reference's cycle-10 arrays (per-kernel fixtures)."""
    c9, c10 = load_npz("ref_s8_c9.npz"), load_npz("ref_s8_c10.npz")
    d = oracle_mod.OracleDomain(8)
    for name in "x y z xd yd zd e p q v ss".split():
        d.field(name)[:] = c9[name]
    s = d.scalars
    s.time, s.deltatime, s.dtcourant, s.dthydro = c9["scalars"][:4]
    s.cycle = 9
    d.time_increment()
    assert (s.time, s.deltatime) == tuple(c10["scalars"][:2])
    assert d.calc_force() == 0
    for name in ("fx", "fy", "fz"):
        assert np.array_equal(d.field(name), c10[name])
    d.node_update()
    for name in "xdd ydd zdd xd yd zd x y z".split():
        assert np.array_equal(d.field(name), c10[name]), name
    assert d.kinematics() == 0
    for name in "vnew delv vdov arealg".split():
        assert np.array_equal(d.field(name), c10[name]), name
    d.monoq_gradients()
    assert d.monoq_regions() == 0
    for name in ("ql", "qq"):
        assert np.array_equal(d.field(name), c10[name]), name
    assert d.material() == 0
    d.time_constraints()
    for name in "e p q ss v".split():
        assert np.array_equal(d.field(name), c10[name]), name
    assert (s.dtcourant, s.dthydro) == tuple(c10["scalars"][2:4])


def test_multi_rank_emulation_matches_single_domain(oracle_mod):
    """8 ranks x 5^3 == one 10^3 domain (same mesh, energy, dt0); differs only by halo add order."""
    single = oracle_mod.OracleDomain(10, 1, 1, 0)
    assert single.run(60) == 0
    for decomp, sizes in (((2, 2, 2), (5, 5, 5)), ((1, 1, 2), (10, 10, 5)), ((1, 2, 2), (10, 5, 5))):
        m = oracle_mod.OracleMulti(decomp, sizes, 1, 1, 0)
        assert m.run(60) == 0
        r0 = m.rank(0)
        assert r0.scalars.cycle == single.scalars.cycle == 60
        assert abs(r0.scalars.time - single.scalars.time) <= 1e-15 * single.scalars.time
        assert abs(r0.field("e")[0] - single.field("e")[0]) <= 1e-12 * single.field("e")[0]


@pytest.mark.parametrize("key,decomp,n,its", [
    ("lulesh_mpi -np 8 -s 5", (2, 2, 2), 5, 9999999), ("lulesh_mpi -np 8 -s 6", (2, 2, 2), 6, 9999999),
    ("lulesh_mpi -np 8 -s 8 -i 100", (2, 2, 2), 8, 100), ("lulesh_mpi -np 8 -s 10 -i 60", (2, 2, 2), 10, 60),
    ("lulesh_mpi -np 27 -s 3 -i 60", (3, 3, 3), 3, 60)])
def test_multi_rank_emulation_bit_identical_to_reference_mpi_build(oracle_mod, goldens, key, decomp, n, its):
    """The reference's USE_MPI=1 build (its own CommSBN / CommSyncPosVel / CommMonoQ and
    MPI_Allreduce code, compiled unmodified against oracle/mpishim) versus the oracle's
    in-process emulation of a rank grid: rank 0's scalars and checksums, bit for bit.
    The reference derives dt0 per rank (lulesh-init.cc:192), hence use_reference_dt0()."""
    gold = goldens[key]
    m = oracle_mod.OracleMulti(decomp, (n, n, n))
    for r in range(m.n):
        m.rank(r).use_reference_dt0()
    assert m.run(its) == 0
    d = m.rank(0)
    s, f = d.scalars, d.field
    sym = d.symmetry(n)
    got = {"cycles": s.cycle, "e0": f("e")[0], "time": s.time, "dt": s.deltatime,
           "dtcourant": s.dtcourant, "dthydro": s.dthydro, "sum_e": _seqsum(f("e")),
           "sum_p": _seqsum(f("p")), "sum_q": _seqsum(f("q")), "sum_v": _seqsum(f("v")),
           "sum_ss": _seqsum(f("ss")), "sum_xyz": _seqsum(f("x") + f("y") + f("z")),
           "sum_absvel": _seqsum(np.abs(f("xd")) + np.abs(f("yd")) + np.abs(f("zd"))),
           "max_abs_diff": sym[0], "total_abs_diff": sym[1], "max_rel_diff": sym[2]}
    for k in CHECK_KEYS:
        assert got[k] == gold[k], f"{key}: {k} emulation {got[k]!r} != reference MPI {gold[k]!r}"
    assert [len(d.region_list(i)) for i in range(11)] == gold["regions"]


def test_oracle_error_codes(oracle_mod):
    d = oracle_mod.OracleDomain(4)
    d.field("v")[3] = -1.0
    assert d.step() == -1          # VolumeError, lulesh.h:42
    d = oracle_mod.OracleDomain(4)
    d.field("q")[5] = 2.0e12       # q > qstop (1e12); p = -q keeps the stress, hence the mesh, intact
    d.field("p")[5] = -2.0e12
    assert d.step() == -2          # QStopError
