"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU restatement
(oracle/lulesh_oracle.c).  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by lulesh_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liblulesh_oracle.so")
CLI = os.path.join(HERE, "_build", "lulesh_oracle")
REF_OMP = os.path.join(HERE, "_ref", "lulesh_omp")
REF_SERIAL = os.path.join(HERE, "_ref", "lulesh_serial")

FIELDS = ("x y z xd yd zd xdd ydd zdd fx fy fz nodalMass e p q ql qq v volo vnew delv vdov "
          "arealg ss elemMass delv_xi delv_eta delv_zeta delx_xi delx_eta delx_zeta").split()
F = {n: i for i, n in enumerate(FIELDS)}


class Scalars(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "dtcourant dthydro dtfixed time deltatime deltatimemultlb deltatimemultub dtmax "
        "stoptime").split()] + [("cycle", C.c_int32), ("error", C.c_int32)]


def build():
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def load():
    if not os.path.exists(LIB):
        build()
    lib = C.CDLL(LIB)
    vp, pd, pi = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.ora_new.restype = vp
    lib.ora_new.argtypes = [C.c_int] * 11
    lib.ora_free.argtypes = [vp]
    lib.ora_use_reference_dt0.argtypes = [vp]
    lib.ora_real.restype = pd
    lib.ora_real.argtypes = [vp, C.c_int]
    lib.ora_real_count.restype = C.c_size_t
    lib.ora_real_count.argtypes = [vp, C.c_int]
    lib.ora_int.restype = pi
    lib.ora_int.argtypes = [vp, C.c_char_p, pi]
    lib.ora_region_list.restype = pi
    lib.ora_region_list.argtypes = [vp, C.c_int, pi]
    lib.ora_scalars.restype = C.POINTER(Scalars)
    lib.ora_scalars.argtypes = [vp]
    lib.ora_dt_candidate.restype = C.c_double
    lib.ora_dt_candidate.argtypes = [vp]
    lib.ora_time_increment.argtypes = [vp, C.c_double]
    for fn in ("ora_calc_force", "ora_kinematics", "ora_monoq_regions", "ora_material", "ora_step"):
        getattr(lib, fn).restype = C.c_int
        getattr(lib, fn).argtypes = [vp]
    for fn in ("ora_node_update", "ora_monoq_gradients", "ora_time_constraints"):
        getattr(lib, fn).restype = None
        getattr(lib, fn).argtypes = [vp]
    lib.ora_run.restype = C.c_int
    lib.ora_run.argtypes = [vp, C.c_int]
    lib.ora_multi_new.restype = vp
    lib.ora_multi_new.argtypes = [C.c_int] * 9
    lib.ora_multi_free.argtypes = [vp]
    lib.ora_multi_rank.restype = vp
    lib.ora_multi_rank.argtypes = [vp, C.c_int]
    lib.ora_multi_step.restype = C.c_int
    lib.ora_multi_step.argtypes = [vp]
    lib.ora_multi_run.restype = C.c_int
    lib.ora_multi_run.argtypes = [vp, C.c_int]
    lib.ora_symmetry.argtypes = [vp, C.c_int, pd]
    return lib


class OracleDomain:
    """One rank's Domain inside the oracle (owning, or borrowed from an OracleMulti)."""

    def __init__(self, nx=30, num_reg=11, balance=1, cost=1, *, num_ranks=1, rank=0,
                 decomp=(1, 1, 1), sizes=None, _borrow=None, _lib=None):
        self.lib = _lib or load()
        sx, sy, sz = sizes if sizes else (nx, nx, nx)
        self.sizes = (sx, sy, sz)
        if _borrow is not None:
            self._p, self._own = _borrow, False
        else:
            self._p = self.lib.ora_new(num_ranks, rank, *decomp, sx, sy, sz, num_reg, balance, cost)
            self._own = True
            if not self._p:
                raise ValueError("bad oracle domain arguments")

    def field(self, name):
        n = self.lib.ora_real_count(self._p, F[name])
        return np.ctypeslib.as_array(self.lib.ora_real(self._p, F[name]), shape=(n,))

    def ints(self, name):
        n = C.c_int()
        p = self.lib.ora_int(self._p, name.encode(), C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)) if n.value else np.zeros(0, np.int32)

    def region_list(self, r):
        n = C.c_int()
        p = self.lib.ora_region_list(self._p, r, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)) if n.value else np.zeros(0, np.int32)

    @property
    def scalars(self):
        return self.lib.ora_scalars(self._p).contents

    def time_increment(self):
        self.lib.ora_time_increment(self._p, self.lib.ora_dt_candidate(self._p))

    def calc_force(self): return self.lib.ora_calc_force(self._p)
    def node_update(self): self.lib.ora_node_update(self._p)
    def kinematics(self): return self.lib.ora_kinematics(self._p)
    def monoq_gradients(self): self.lib.ora_monoq_gradients(self._p)
    def monoq_regions(self): return self.lib.ora_monoq_regions(self._p)
    def material(self): return self.lib.ora_material(self._p)
    def time_constraints(self): self.lib.ora_time_constraints(self._p)
    def step(self): return self.lib.ora_step(self._p)
    def run(self, max_cycles=9999999): return self.lib.ora_run(self._p, max_cycles)

    def use_reference_dt0(self):
        self.lib.ora_use_reference_dt0(self._p)

    def symmetry(self, n=None):
        out = (C.c_double * 3)()
        self.lib.ora_symmetry(self._p, n or min(self.sizes[0], self.sizes[1]), out)
        return tuple(out)

    def __del__(self):
        if getattr(self, "_own", False) and self._p:
            self.lib.ora_free(self._p)
            self._p = None


class OracleMulti:
    """In-process emulation of a (px,py,pz) rank grid with the reference's halo semantics."""

    def __init__(self, decomp, sizes, num_reg=11, balance=1, cost=1):
        self.lib = load()
        self.decomp, self.sizes = tuple(decomp), tuple(sizes)
        self._p = self.lib.ora_multi_new(*decomp, *sizes, num_reg, balance, cost)
        if not self._p:
            raise ValueError("bad oracle multi arguments")
        self.n = decomp[0] * decomp[1] * decomp[2]

    def rank(self, r):
        return OracleDomain(sizes=self.sizes, _borrow=self.lib.ora_multi_rank(self._p, r), _lib=self.lib)

    def step(self): return self.lib.ora_multi_step(self._p)
    def run(self, max_cycles=9999999): return self.lib.ora_multi_run(self._p, max_cycles)

    def __del__(self):
        if getattr(self, "_p", None):
            self.lib.ora_multi_free(self._p)
            self._p = None
