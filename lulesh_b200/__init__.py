"""lulesh_b200 -- B200-native Lagrange-leapfrog step of LULESH 2.0.

Python is only a thin ctypes veneer over the C ABI (include/lulesh_b200.h,
include/lulesh_host.h) for tests and bench.py; the product is
``lib/liblulesh_b200.so`` (hand-written sm_100a FP64 kernels + runtime + the C++
host ``Domain``) and ``bin/lulesh_b200`` (the drop-in driver).

There is no CPU fallback: creating a :class:`Device` without a usable sm_100
GPU raises :class:`LuleshError`; importing the package without the built
library raises ImportError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LULESH_B200_LIB") or os.path.join(_HERE, "lib", "liblulesh_b200.so")
BIN_PATH = os.path.join(_HERE, "bin", "lulesh_b200")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make` at the repo root "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "lulesh_b200 has no pure-python or CPU fallback.")

_lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

ABI_VERSION = 1
UNIQUE_ID_BYTES = 128
NUM_KERNELS = 5
KERNEL_NAMES = ("time_increment", "force_elem", "node_update", "kinematics_grad", "material")
# slots of lulesh_b200_timeline (enum lulesh_b200_timeline_slot)
TIMELINE_NAMES = ("time_increment", "k1_force", "k2_node", "node_join_wait", "k3_kinematics", "k45_interior",
                  "monoq_join_wait", "k45_tail", "cycle", "comm_dt", "comm_node", "comm_monoq")

# status codes (lulesh.h:42 + infrastructure)
OK, VOLUME_ERROR, QSTOP_ERROR, EINVAL, ECUDA, ENCCL = 0, -1, -2, -10, -11, -12

# field ids, in the order of enum lulesh_b200_field
FIELDS = ("x y z xd yd zd xdd ydd zdd fx fy fz nodalMass e p q ql qq v volo vnew delv vdov "
          "arealg ss elemMass delv_xi delv_eta delv_zeta delx_xi delx_eta delx_zeta").split()
F = {name: i for i, name in enumerate(FIELDS)}


class Constants(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "e_cut p_cut q_cut v_cut u_cut hgcoef ss4o3 qstop monoq_max_slope monoq_limiter_mult "
        "qlc_monoq qqc_monoq qqc eosvmax eosvmin pmin emin dvovmax refdens").split()]


class Scalars(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "dtcourant dthydro dtfixed time deltatime deltatimemultlb deltatimemultub dtmax "
        "stoptime").split()] + [("cycle", C.c_int32), ("error", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)


class HostView(C.Structure):
    _fields_ = (
        [("abi_version", C.c_int32)]
        + [(n, C.c_int32) for n in "sizeX sizeY sizeZ numElem numNode numRanks rank px py pz "
                                   "colLoc rowLoc planeLoc".split()]
        + [(n, _pd) for n in "x y z xd yd zd nodalMass".split()]
        + [(n, _pi) for n in "symmX symmY symmZ".split()]
        + [(n, C.c_int32) for n in "numSymmX numSymmY numSymmZ".split()]
        + [(n, _pi) for n in "nodelist lxim lxip letam letap lzetam lzetap elemBC".split()]
        + [(n, _pd) for n in "e p q v volo ss elemMass".split()]
        + [("numReg", C.c_int32), ("cost", C.c_int32), ("regElemSize", _pi),
           ("regElemlist", C.POINTER(_pi)), ("nodeElemStart", _pi), ("nodeElemCornerList", _pi),
           ("constants", Constants), ("scalars", Scalars)])


class SedovParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in "abi_version numRanks rank px py pz sx sy sz numReg balance cost".split()]


PROGRESS_CB = C.CFUNCTYPE(None, C.c_int32, C.c_double, C.c_double, C.c_void_p)

# every symbol declared in include/lulesh_b200.h and include/lulesh_host.h
ABI_SYMBOLS = (
    "lulesh_b200_get_unique_id lulesh_b200_create lulesh_b200_create_sedov lulesh_b200_sum_nodal_mass lulesh_b200_run "
    "lulesh_b200_step lulesh_b200_get_scalars lulesh_b200_set_scalars lulesh_b200_download "
    "lulesh_b200_upload lulesh_b200_field_count lulesh_b200_set_debug "
    "lulesh_b200_kernel_time_increment lulesh_b200_kernel_force lulesh_b200_kernel_node "
    "lulesh_b200_kernel_kinematics lulesh_b200_kernel_material lulesh_b200_time_cycles lulesh_b200_timeline "
    "lulesh_b200_device_bytes lulesh_b200_upload_bytes lulesh_b200_last_error lulesh_b200_halo_mode "
    "lulesh_b200_halo_plan_create lulesh_b200_halo_plan_query lulesh_b200_halo_plan_destroy "
    "lulesh_b200_destroy "
    "lulesh_host_domain_new lulesh_host_domain_free lulesh_host_domain_view "
    "lulesh_host_domain_field lulesh_host_domain_ints lulesh_host_domain_scalars "
    "lulesh_host_domain_write_vtk lulesh_host_decompose lulesh_host_main").split()


def _sig(name, restype, *argtypes):
    fn = getattr(_lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_vp = C.c_void_p
_sig("lulesh_b200_get_unique_id", C.c_int, _vp)
_sig("lulesh_b200_create", C.c_int, C.POINTER(HostView), C.c_int, _vp, C.POINTER(_vp))
_sig("lulesh_b200_create_sedov", C.c_int, C.POINTER(SedovParams), C.c_int, _vp, C.POINTER(_vp))
_sig("lulesh_b200_sum_nodal_mass", C.c_int, _vp)
_sig("lulesh_b200_run", C.c_int, _vp, C.c_int32, C.c_int32, PROGRESS_CB, _vp)
_sig("lulesh_b200_step", C.c_int, _vp)
_sig("lulesh_b200_get_scalars", C.c_int, _vp, C.POINTER(Scalars))
_sig("lulesh_b200_set_scalars", C.c_int, _vp, C.POINTER(Scalars))
_sig("lulesh_b200_download", C.c_int, _vp, C.c_int, _pd, C.c_size_t)
_sig("lulesh_b200_upload", C.c_int, _vp, C.c_int, _pd, C.c_size_t)
_sig("lulesh_b200_field_count", C.c_size_t, _vp, C.c_int)
_sig("lulesh_b200_set_debug", C.c_int, _vp, C.c_int)
for _k in ("time_increment", "force", "kinematics", "material"):
    _sig(f"lulesh_b200_kernel_{_k}", C.c_int, _vp)
_sig("lulesh_b200_kernel_node", C.c_int, _vp, C.c_int)
_sig("lulesh_b200_time_cycles", C.c_int, _vp, C.c_int32, C.POINTER(C.c_float),
     C.POINTER(C.c_float), C.POINTER(C.c_int64))
_sig("lulesh_b200_timeline", C.c_int, _vp, C.c_int32, C.POINTER(C.c_float))
_sig("lulesh_b200_device_bytes", C.c_size_t, _vp)
_sig("lulesh_b200_upload_bytes", C.c_size_t, _vp)
_sig("lulesh_b200_last_error", C.c_char_p)
_sig("lulesh_b200_halo_mode", C.c_char_p, _vp)
_sig("lulesh_b200_destroy", None, _vp)
_sig("lulesh_b200_halo_plan_create", C.c_int, C.POINTER(HostView), C.POINTER(_vp))
_sig("lulesh_b200_halo_plan_query", C.c_int, _vp, C.c_char_p, C.POINTER(_pi), C.POINTER(C.c_size_t))
_sig("lulesh_b200_halo_plan_destroy", None, _vp)
_sig("lulesh_host_domain_new", _vp, *([C.c_int] * 11))
_sig("lulesh_host_domain_free", None, _vp)
_sig("lulesh_host_domain_view", None, _vp, C.POINTER(HostView))
_sig("lulesh_host_domain_field", _pd, _vp, C.c_int, C.POINTER(C.c_size_t))
_sig("lulesh_host_domain_ints", _pi, _vp, C.c_char_p, C.POINTER(C.c_size_t))
_sig("lulesh_host_domain_scalars", C.POINTER(Scalars), _vp)
_sig("lulesh_host_domain_write_vtk", C.c_int, _vp, C.c_int, C.c_char_p)
_sig("lulesh_host_decompose", C.c_int, C.c_int, *([C.POINTER(C.c_int)] * 3))
_sig("lulesh_host_main", C.c_int, C.c_int, C.POINTER(C.c_char_p))


class LuleshError(RuntimeError):
    def __init__(self, code, what):
        msg = _lib.lulesh_b200_last_error().decode() if code <= EINVAL else ""
        names = {VOLUME_ERROR: "VolumeError", QSTOP_ERROR: "QStopError"}
        super().__init__(f"{what}: status {code} {names.get(code, '')} {msg}".strip())
        self.code = code


def decompose(num_ranks: int):
    """(px, py, pz) for `num_ranks` (lulesh-init.cc:676-738 generalised to 2 and 4)."""
    px, py, pz = C.c_int(), C.c_int(), C.c_int()
    if _lib.lulesh_host_decompose(num_ranks, C.byref(px), C.byref(py), C.byref(pz)) != 0:
        raise ValueError(f"unsupported rank count {num_ranks}")
    return px.value, py.value, pz.value


def get_unique_id() -> bytes:
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    rc = _lib.lulesh_b200_get_unique_id(buf)
    if rc:
        raise LuleshError(rc, "get_unique_id")
    return buf.raw


class Domain:
    """Host Domain (C++ class behind include/lulesh_host.h), reference lulesh.h:148-595."""

    def __init__(self, nx=30, num_reg=11, balance=1, cost=1, *, num_ranks=1, rank=0,
                 decomp=None, sizes=None):
        px, py, pz = decomp if decomp else decompose(num_ranks)
        sx, sy, sz = sizes if sizes else (nx, nx, nx)
        self._p = _lib.lulesh_host_domain_new(num_ranks, rank, px, py, pz, sx, sy, sz,
                                              num_reg, balance, cost)
        if not self._p:
            raise ValueError("invalid Domain arguments")
        self.view = HostView()
        _lib.lulesh_host_domain_view(self._p, C.byref(self.view))
        self.sizes, self.decomp = (sx, sy, sz), (px, py, pz)

    numElem = property(lambda s: s.view.numElem)
    numNode = property(lambda s: s.view.numNode)

    def field(self, name) -> np.ndarray:
        """numpy view (no copy) of a host field."""
        n = C.c_size_t()
        p = _lib.lulesh_host_domain_field(self._p, F[name], C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def ints(self, name) -> np.ndarray:
        n = C.c_size_t()
        p = _lib.lulesh_host_domain_ints(self._p, name.encode(), C.byref(n))
        if not p:
            return np.zeros(0, dtype=np.int32)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def region_list(self, r) -> np.ndarray:
        n = self.view.regElemSize[r]
        return np.ctypeslib.as_array(self.view.regElemlist[r], shape=(n,)) if n else np.zeros(0, np.int32)

    @property
    def scalars(self) -> Scalars:
        return _lib.lulesh_host_domain_scalars(self._p).contents

    def refresh_view(self):
        _lib.lulesh_host_domain_view(self._p, C.byref(self.view))
        return self.view

    def write_vtk(self, path, rank=0):
        """The `-v` dump of this Domain's host arrays (lulesh-viz.cc:56-258, VTK instead of Silo)."""
        if _lib.lulesh_host_domain_write_vtk(self._p, rank, os.fsencode(path)) != 0:
            raise OSError(f"cannot write {path}")

    def __del__(self):
        if getattr(self, "_p", None):
            _lib.lulesh_host_domain_free(self._p)
            self._p = None


HALO_ARRAYS = ("bnode bsum_start bsum_src pack_idx msg_rank msg_count msg_send_off msg_recv_off "
               "mq_idx face_rank face_count face_send_off face_ghost_off").split()


def halo_plan(domain: Domain) -> dict:
    """Host-only halo description of one rank (lulesh_b200_halo_plan_*), as numpy arrays."""
    p = _vp()
    rc = _lib.lulesh_b200_halo_plan_create(C.byref(domain.refresh_view()), C.byref(p))
    if rc:
        raise LuleshError(rc, "halo_plan_create")
    out = {}
    for name in HALO_ARRAYS:
        data, n = _pi(), C.c_size_t()
        rc = _lib.lulesh_b200_halo_plan_query(p, name.encode(), C.byref(data), C.byref(n))
        if rc:
            raise LuleshError(rc, "halo_plan_query")
        out[name] = (np.ctypeslib.as_array(data, shape=(n.value,)).copy() if n.value
                     else np.zeros(0, np.int32))
    _lib.lulesh_b200_halo_plan_destroy(p)
    return out


class Device:
    """Per-GPU handle of the C ABI (include/lulesh_b200.h)."""

    def __init__(self, domain: Domain, device=0, unique_id: bytes | None = None):
        self._h = _vp()
        self.domain = domain
        view = domain.refresh_view()
        uid = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES) if unique_id else None
        rc = _lib.lulesh_b200_create(C.byref(view), device, uid, C.byref(self._h))
        if rc:
            self._h = None
            raise LuleshError(rc, "lulesh_b200_create")

    @classmethod
    def sedov(cls, nx=30, num_reg=11, balance=1, cost=1, *, num_ranks=1, rank=0, decomp=None, sizes=None,
              device=0, unique_id: bytes | None = None):
        """Device-side setup (lulesh_b200_create_sedov): no host Domain is built."""
        px, py, pz = decomp if decomp else decompose(num_ranks)
        sx, sy, sz = sizes if sizes else (nx, nx, nx)
        p = SedovParams(ABI_VERSION, num_ranks, rank, px, py, pz, sx, sy, sz, num_reg, balance, cost)
        self = cls.__new__(cls)
        self._h, self.domain = _vp(), None
        uid = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES) if unique_id else None
        rc = _lib.lulesh_b200_create_sedov(C.byref(p), device, uid, C.byref(self._h))
        if rc:
            self._h = None
            raise LuleshError(rc, "lulesh_b200_create_sedov")
        return self

    def _check(self, rc, what):
        if rc:
            raise LuleshError(rc, what)

    def sum_nodal_mass(self):
        self._check(_lib.lulesh_b200_sum_nodal_mass(self._h), "sum_nodal_mass")

    def run(self, max_cycles=9999999, sync_every=64, progress=None):
        cb = PROGRESS_CB(lambda c, t, dt, u: progress(c, t, dt)) if progress else PROGRESS_CB()
        self._check(_lib.lulesh_b200_run(self._h, max_cycles, sync_every, cb, None), "run")

    def step(self):
        self._check(_lib.lulesh_b200_step(self._h), "step")

    def set_debug(self, on=True):
        self._check(_lib.lulesh_b200_set_debug(self._h, int(on)), "set_debug")

    @property
    def scalars(self) -> Scalars:
        s = Scalars()
        self._check(_lib.lulesh_b200_get_scalars(self._h, C.byref(s)), "get_scalars")
        return s

    @scalars.setter
    def scalars(self, s: Scalars):
        self._check(_lib.lulesh_b200_set_scalars(self._h, C.byref(s)), "set_scalars")

    def count(self, name):
        return _lib.lulesh_b200_field_count(self._h, F[name])

    def download(self, name, out: np.ndarray | None = None) -> np.ndarray:
        n = self.count(name)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        self._check(_lib.lulesh_b200_download(self._h, F[name], out.ctypes.data_as(_pd), out.size),
                    f"download {name}")
        return out

    def upload(self, name, src: np.ndarray):
        src = np.ascontiguousarray(src, dtype=np.float64)
        self._check(_lib.lulesh_b200_upload(self._h, F[name], src.ctypes.data_as(_pd), src.size),
                    f"upload {name}")

    def kernel(self, which, debug=1):
        """Run one kernel synchronously: time_increment|force|node|kinematics|material."""
        fn = getattr(_lib, f"lulesh_b200_kernel_{which}")
        rc = fn(self._h, debug) if which == "node" else fn(self._h)
        self._check(rc, f"kernel_{which}")

    def time_cycles(self, cycles, per_kernel=False):
        """(total_ms, per_kernel_ms or None, launches) for `cycles` cycles, CUDA-event timed."""
        total, launches = C.c_float(), C.c_int64()
        pk = (C.c_float * NUM_KERNELS)() if per_kernel else None
        rc = _lib.lulesh_b200_time_cycles(self._h, cycles, C.byref(total), pk, C.byref(launches))
        self._check(rc, "time_cycles")
        return total.value, (list(pk) if per_kernel else None), launches.value

    def timeline(self, cycles):
        """Per-cycle timeline in ms (dict), measured on the shipped two-stream schedule."""
        out = (C.c_float * len(TIMELINE_NAMES))()
        self._check(_lib.lulesh_b200_timeline(self._h, cycles, out), "timeline")
        return dict(zip(TIMELINE_NAMES, (float(v) for v in out)))

    halo_mode = property(lambda s: _lib.lulesh_b200_halo_mode(s._h).decode())
    device_bytes = property(lambda s: _lib.lulesh_b200_device_bytes(s._h))
    upload_bytes = property(lambda s: _lib.lulesh_b200_upload_bytes(s._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lulesh_b200_destroy(self._h)
            self._h = None

    __del__ = close


def main(argv) -> int:
    """The drop-in driver's main() (lulesh.cc:2650-2792) in-process."""
    args = [a.encode() for a in ["lulesh_b200"] + list(argv)]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    return _lib.lulesh_host_main(len(args), arr)
