"""Host side above the C ABI: the C++ Domain (lulesh_b200/csrc/host) against the
reference's own setup (fixtures dumped from the reference Domain constructor) and
against the oracle's independent setup; library loading and ABI surface; the
driver's command-line behaviour (lulesh-util.cc:63-171)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_npz

INT_FIELDS = "nodelist lxim lxip letam letap lzetam lzetap elemBC symmX symmY symmZ".split()


def test_library_exports_every_declared_symbol(lb):
    declared = set()
    for hdr in ("lulesh_b200.h", "lulesh_host.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        declared |= set(re.findall(r"\b(lulesh_(?:b200|host)_[a-z_0-9]+)\s*\(", text))
    declared -= {"lulesh_b200_progress_cb"}
    assert declared == set(lb.ABI_SYMBOLS)
    lib = ctypes.CDLL(lb.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_single_rank_setup_matches_reference_dump(lb):
    ref = load_npz("ref_s8_c9.npz")   # connectivity and volo do not change with the cycle
    d = lb.Domain(8)
    for name in INT_FIELDS + ["regNumList"]:
        assert np.array_equal(d.ints(name), ref[name]), name
    for name in ("volo", "elemMass", "nodalMass"):
        assert np.array_equal(d.field(name), ref[name]), name
    assert list(d.ints("regElemSize")) == list(ref["regElemSize"])


@pytest.mark.parametrize("rank", range(8))
def test_multi_rank_setup_matches_reference_constructor(lb, rank):
    """tp=2, nx=4: the reference Domain built for every rank location in a non-MPI
    process (oracle/ref_setup_dump.cc).  Region lists are excluded: the non-MPI
    reference does not rotate them by rank (lulesh-init.cc:408-409)."""
    ref = load_npz(f"ref_setup_tp2_nx4_r{rank}.npz")
    d = lb.Domain(4, num_ranks=8, rank=rank)
    for name in INT_FIELDS:
        got, want = d.ints(name), ref.get(name, np.zeros(0, np.int32))
        assert np.array_equal(got, want), name
    for name in ("x", "y", "z", "volo", "nodalMass", "e"):
        assert np.array_equal(d.field(name), ref[name]), name
    assert d.scalars.deltatime == ref["scalars"][1]   # dt0; identical on every rank here (F10)


@pytest.mark.parametrize("kw", [
    dict(nx=6), dict(nx=5, num_reg=16, balance=1, cost=8), dict(nx=7, num_reg=1, cost=0),
    dict(nx=4, num_ranks=2, rank=1), dict(nx=4, num_ranks=4, rank=2),
    dict(nx=3, num_ranks=8, rank=5), dict(nx=4, num_ranks=2, rank=0, sizes=(6, 4, 3)),
])
def test_host_domain_equals_oracle_setup(lb, oracle_mod, kw):
    d = lb.Domain(**kw)
    nr = kw.get("num_ranks", 1)
    o = oracle_mod.OracleDomain(kw["nx"], kw.get("num_reg", 11), kw.get("balance", 1),
                                kw.get("cost", 1), num_ranks=nr, rank=kw.get("rank", 0),
                                decomp=d.decomp, sizes=kw.get("sizes"))
    for name in INT_FIELDS + ["regNumList", "regElemSize", "nodeElemStart", "nodeElemCornerList"]:
        assert np.array_equal(d.ints(name), o.ints(name)), name
    for name in "x y z xd yd zd nodalMass e p q v volo ss elemMass".split():
        assert np.array_equal(d.field(name), o.field(name)), name
    for r in range(kw.get("num_reg", 11)):
        assert np.array_equal(d.region_list(r), o.region_list(r))
    for n, _ in lb.Scalars._fields_:
        assert getattr(d.scalars, n) == getattr(o.scalars, n), n


def test_default_region_sizes_match_reference(lb, goldens):
    d = lb.Domain(30)
    assert list(d.ints("regElemSize")) == goldens["lulesh_omp -s 30 -i 100"]["regions"]
    d = lb.Domain(12, 16, 1, 8)
    assert list(d.ints("regElemSize")) == goldens["lulesh_omp -s 12 -i 40 -r 16 -b 1 -c 8"]["regions"]


def test_corner_list_is_ascending_elements(lb):
    d = lb.Domain(5)
    start, corners = d.ints("nodeElemStart"), d.ints("nodeElemCornerList")
    nodelist = d.ints("nodelist")
    assert start[-1] == 8 * d.numElem
    for n in range(d.numNode):
        seg = corners[start[n]:start[n + 1]]
        assert 1 <= len(seg) <= 8 and np.all(np.diff(seg // 8) > 0)
        assert np.all(nodelist[seg] == n)


def test_decompose(lb):
    assert lb.decompose(1) == (1, 1, 1) and lb.decompose(8) == (2, 2, 2)
    assert lb.decompose(2) == (1, 1, 2) and lb.decompose(4) == (1, 2, 2)
    assert lb.decompose(27) == (3, 3, 3)
    with pytest.raises(ValueError):
        lb.decompose(3)


def test_invalid_domain_arguments(lb):
    with pytest.raises(ValueError):
        lb.Domain(0)
    with pytest.raises(ValueError):
        lb.Domain(4, num_ranks=2, rank=2)


def test_create_fails_loudly_without_gpu(lb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lb.LuleshError) as e:
        lb.Device(lb.Domain(4))
    assert e.value.code == lb.ECUDA and "no CPU fallback" in str(e.value)


def _cli(lb, *args):
    return subprocess.run([lb.BIN_PATH, *args], capture_output=True, text=True)


def test_cli_help_and_errors(lb):
    p = _cli(lb, "-h")
    assert p.returncode == 0 and p.stdout.startswith("Usage: ") and " -q              : quiet mode" in p.stdout
    p = _cli(lb, "-z")
    assert p.returncode == 255 and "ERROR: Unknown command line argument: -z" in p.stdout
    p = _cli(lb, "-s")
    assert p.returncode == 255 and "Missing integer argument to -s" in p.stdout
    p = _cli(lb, "-i", "2x")
    assert p.returncode == 255
    assert "Parse Error on option -i integer value required after argument" in p.stdout
    assert " -v              : Output viz file" in _cli(lb, "-h").stdout


def test_create_rejects_bad_views_before_touching_the_gpu(lb):
    """Argument validation happens before any CUDA call, so it is testable here."""
    import ctypes as C
    h = C.c_void_p()
    d = lb.Domain(4)
    assert lb._lib.lulesh_b200_create(None, 0, None, C.byref(h)) == lb.EINVAL
    v = lb.HostView.from_buffer_copy(d.refresh_view())
    v.abi_version = 99
    assert lb._lib.lulesh_b200_create(C.byref(v), 0, None, C.byref(h)) == lb.EINVAL
    assert b"abi_version" in lb._lib.lulesh_b200_last_error()
    v = lb.HostView.from_buffer_copy(d.refresh_view())
    v.numNode += 1
    assert lb._lib.lulesh_b200_create(C.byref(v), 0, None, C.byref(h)) == lb.EINVAL
    v = lb.HostView.from_buffer_copy(d.refresh_view())
    v.numRanks = 3
    assert lb._lib.lulesh_b200_create(C.byref(v), 0, None, C.byref(h)) == lb.EINVAL
    v = lb.HostView.from_buffer_copy(d.refresh_view())
    v.nodelist = None
    assert lb._lib.lulesh_b200_create(C.byref(v), 0, None, C.byref(h)) == lb.EINVAL
    assert not h.value
    # null handles are rejected, not dereferenced
    assert lb._lib.lulesh_b200_step(None) == lb.EINVAL
    assert lb._lib.lulesh_b200_run(None, 1, 1, lb.PROGRESS_CB(), None) == lb.EINVAL
    assert lb._lib.lulesh_b200_field_count(None, 0) == 0
    lb._lib.lulesh_b200_destroy(None)


@pytest.mark.parametrize("array,index,value,message", [
    ("nodelist", 5, -1, b"nodelist entry out of range"),
    ("nodelist", 17, 10**6, b"nodelist entry out of range"),
    ("lxip", 3, 10**6, b"face neighbour out of range"),
    ("lzetam", 0, -2, b"face neighbour out of range"),
    ("symmY", 1, 10**6, b"symmetry node out of range"),
    ("nodeElemCornerList", 9, 8 * 64, b"corner list entry out of range"),
    ("nodeElemStart", 4, 10**6, b"corners"),
])
def test_create_rejects_out_of_range_indices(lb, array, index, value, message):
    """The kernels index with these arrays unchecked; create() must refuse them (no GPU needed)."""
    import ctypes as C
    d = lb.Domain(4)
    a = d.ints(array)
    saved = int(a[index])
    a[index] = value
    try:
        h = C.c_void_p()
        assert lb._lib.lulesh_b200_create(C.byref(d.refresh_view()), 0, None, C.byref(h)) == lb.EINVAL
        assert message in lb._lib.lulesh_b200_last_error()
        assert not h.value
    finally:
        a[index] = saved


@pytest.mark.parametrize("num_ranks", [1, 2, 4, 8, 27])
def test_valid_views_pass_validation_and_then_need_a_gpu(lb, num_ranks):
    """Every rank's view of every supported decomposition is accepted by the GPU-free checks;
    in this container create() then stops at the first CUDA call (no CPU fallback)."""
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    for rank in range(num_ranks):
        d = lb.Domain(3, 5, 1, 2, num_ranks=num_ranks, rank=rank)
        h = C.c_void_p()
        uid = C.create_string_buffer(lb.UNIQUE_ID_BYTES)
        rc = lb._lib.lulesh_b200_create(C.byref(d.refresh_view()), 0, uid, C.byref(h))
        assert rc == lb.ECUDA, lb._lib.lulesh_b200_last_error()
        assert b"no CPU fallback" in lb._lib.lulesh_b200_last_error() and not h.value


def test_create_rejects_region_lists_that_are_not_a_partition(lb):
    import ctypes as C
    d = lb.Domain(4)
    h = C.c_void_p()
    nonempty = [r for r in range(d.view.numReg) if d.view.regElemSize[r] > 0]
    r0 = d.region_list(nonempty[0])
    saved = int(r0[0])
    other = int(d.region_list(nonempty[1])[0])
    r0[0] = other                                   # element listed twice (and one missing)
    assert lb._lib.lulesh_b200_create(C.byref(d.refresh_view()), 0, None, C.byref(h)) == lb.EINVAL
    assert b"more than one region list" in lb._lib.lulesh_b200_last_error()
    r0[0] = 64                                      # past the last element
    assert lb._lib.lulesh_b200_create(C.byref(d.refresh_view()), 0, None, C.byref(h)) == lb.EINVAL
    assert b"region list entry out of range" in lb._lib.lulesh_b200_last_error()
    r0[0] = saved
    v = lb.HostView.from_buffer_copy(d.refresh_view())
    sizes = (C.c_int32 * v.numReg)(*[v.regElemSize[r] for r in range(v.numReg)])
    sizes[nonempty[0]] -= 1                         # one element in no region
    v.regElemSize = C.cast(sizes, type(v.regElemSize))
    assert lb._lib.lulesh_b200_create(C.byref(v), 0, None, C.byref(h)) == lb.EINVAL
    assert b"region lists cover 63 of 64 elements" in lb._lib.lulesh_b200_last_error()
    assert not h.value


def read_vtk(path):
    """Minimal reader of the binary legacy-VTK files written by vizdump.cc."""
    raw = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        end = raw.index(b"\n", pos)
        out = raw[pos:end].decode()
        pos = end + 1
        return out

    def block(dtype, count):
        nonlocal pos
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=pos)
        pos += a.nbytes
        assert raw[pos:pos + 1] == b"\n"
        pos += 1
        return a

    out = {"header": [line() for _ in range(4)]}
    tok = line().split()
    assert tok[0] == "POINTS" and tok[2] == "double"
    nn = int(tok[1])
    out["points"] = block(">f8", 3 * nn).reshape(nn, 3)
    tok = line().split()
    assert tok[0] == "CELLS"
    ne = int(tok[1])
    assert int(tok[2]) == 9 * ne
    out["cells"] = block(">i4", 9 * ne).reshape(ne, 9)
    assert line().split() == ["CELL_TYPES", str(ne)]
    out["cell_types"] = block(">i4", ne)
    for section, n in (("CELL_DATA", ne), ("POINT_DATA", nn)):
        assert line().split() == [section, str(n)]
        while pos < len(raw) and raw[pos:pos + 7] == b"SCALARS":
            _, name, typ, _one = line().split()
            assert line() == "LOOKUP_TABLE default"
            out[name] = block(">f8" if typ == "double" else ">i4", n)
    assert pos == len(raw)
    return out


def test_viz_dump_holds_the_reference_field_set(lb, tmp_path):
    """`-v` (lulesh-viz.cc:121-258): mesh, connectivity, regions, e p v q, speed xd yd zd."""
    dom = lb.Domain(5, 4, 1, 1, num_ranks=8, rank=7)
    dom.field("xd")[:] = np.linspace(-1.0, 2.0, dom.numNode)
    dom.field("yd")[:] = 0.5
    dom.field("p")[:] = np.arange(dom.numElem) * 0.25
    dom.scalars.cycle = 42
    path = tmp_path / "lulesh_plot_c42.007.vtk"
    dom.write_vtk(path, rank=7)
    vtk = read_vtk(path)
    assert vtk["header"][0] == "# vtk DataFile Version 3.0" and "cycle 42" in vtk["header"][1]
    assert "rank 7" in vtk["header"][1] and vtk["header"][2:] == ["BINARY", "DATASET UNSTRUCTURED_GRID"]
    for c, name in enumerate("xyz"):
        assert np.array_equal(vtk["points"][:, c], dom.field(name))
    assert np.all(vtk["cells"][:, 0] == 8) and np.all(vtk["cell_types"] == 12)   # VTK_HEXAHEDRON
    assert np.array_equal(vtk["cells"][:, 1:].ravel(), dom.ints("nodelist"))
    for name in "e p v q".split():
        assert np.array_equal(vtk[name], dom.field(name)), name
    assert np.array_equal(vtk["regions"], dom.ints("regNumList"))
    for name in "xd yd zd".split():
        assert np.array_equal(vtk[name], dom.field(name)), name
    speed = np.sqrt(dom.field("xd") ** 2 + dom.field("yd") ** 2 + dom.field("zd") ** 2)
    assert np.allclose(vtk["speed"], speed, rtol=1e-15)
    with pytest.raises(OSError):
        dom.write_vtk(tmp_path / "missing_dir" / "x.vtk")


def test_reference_binding_view_against_reference_domain():
    """include/lulesh_b200_reference_binding.h on the CPU: the view it builds from the REFERENCE's own
    Domain (oracle/_ref/binding_check, compiled from /root/reference by oracle/Makefile) points at the
    reference's arrays, and the node -> corner lists it rebuilds from nodelist equal the ones the
    reference builds itself when threaded (lulesh-init.cc:272-337).  No GPU call is made."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "binding_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/binding_check was not built (needs /root/reference at build time)")
    for args, entries in ((["6"], 8 * 6 ** 3), (["9", "16", "1", "8"], 8 * 9 ** 3)):
        p = subprocess.run([exe] + args, capture_output=True, text=True, timeout=120,
                           env=dict(os.environ, OMP_NUM_THREADS="2"))
        assert p.returncode == 0, p.stdout + p.stderr
        assert f"corner_entries_compared={entries}" in p.stdout, p.stdout
