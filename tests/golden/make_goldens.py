#!/usr/bin/env python
"""Regenerates the golden fixtures from the UNMODIFIED reference build.

Run in the build container only (needs oracle/_ref, which oracle/Makefile
compiles from /root/reference):

    make -C oracle ref && python tests/golden/make_goldens.py [--long]

The `lulesh_mpi -np N ...` records come from the reference's own MPI code paths
(lulesh-comm.cc) compiled unmodified against oracle/mpishim (no MPI exists in this image).

Outputs (committed):
  tests/golden/ref_goldens.json   REFJSON records (cycles, %.17g e0, checksums,
                                  symmetry triple, region sizes) per run
  tests/golden/ref_s8_c{9,10}.npz every accessor-reachable Domain array of the
                                  reference after 9 and 10 cycles of -s 8
                                  (input and expected output of one cycle)
  tests/golden/ref_setup_tp2_nx4_r*.npz  reference Domain setup for all 8 rank
                                  locations of a 2x2x2 layout (lulesh-init.cc)
`--long` also runs -s 45/-s 60 to stoptime (minutes).  The -s 90 and -s 128
records were produced with the same binary (`-r 1 -c 0`, results are region
independent) and are merged from /tmp/gold/*.out when present, and so
are the config-3-size records (16.8 M elements, ~16 GB, 6 s per cycle on 6 cores):

    OMP_NUM_THREADS=6 oracle/_ref/lulesh_omp -s 256 -i 10 -r 1 -c 0 > /tmp/gold/s256_i10.out
    OMP_NUM_THREADS=6 oracle/_ref/lulesh_omp -s 256 -i 40 -r 1 -c 0 > /tmp/gold/s256_i40.out
    OMP_NUM_THREADS=8 oracle/_ref/lulesh_omp -s 384 -i 10 -r 1 -c 0 > /tmp/gold/s384_i10.out   # config 4's mesh, 42 GB
"""
import io, json, os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            line = f.readline()
            if not line:
                break
            name, dtype, count = line.decode().split()
            dt = np.dtype("<" + dtype)
            out[name] = np.frombuffer(f.read(int(count) * dt.itemsize), dtype=dt).copy()
    return out


def run(binary, args, threads=4, dump=None):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    if dump:
        env["LULESH_REF_DUMP"] = dump
    p = subprocess.run([os.path.join(REF, binary)] + args.split(), env=env,
                       capture_output=True, text=True, check=True)
    for line in p.stdout.splitlines():
        if line.startswith("REFJSON "):
            rec = json.loads(line[len("REFJSON "):])
            rec.pop("elapsed", None)
            return rec
    raise RuntimeError("no REFJSON line")


def run_mpi(np_, args, threads=2):
    """The reference's USE_MPI=1 build under the single-node MPI stand-in (oracle/mpishim).
    OMP_NUM_THREADS >= 2 selects the reference's threaded force path (lulesh.cc:514), the
    canonical summation order of this project."""
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    p = subprocess.run([os.path.join(REF, "mpirun_shim"), "-np", str(np_), os.path.join(REF, "lulesh_mpi")]
                       + args.split(), env=env, capture_output=True, text=True, check=True)
    for line in p.stdout.splitlines():
        if line.startswith("REFJSON "):
            rec = json.loads(line[len("REFJSON "):])
            rec.pop("elapsed", None)
            return rec
    raise RuntimeError("no REFJSON line")


def main():
    runs = [
        ("lulesh_omp", "-s 30 -i 100"),
        ("lulesh_serial", "-s 30 -i 100"),
        ("lulesh_omp", "-s 5"),
        ("lulesh_omp", "-s 10"),
        ("lulesh_omp", "-s 20"),
        ("lulesh_omp", "-s 30 -r 1 -c 0"),
        ("lulesh_omp", "-s 8 -i 10"),
        ("lulesh_omp", "-s 12 -i 40 -r 16 -b 1 -c 8"),
        ("lulesh_omp", "-s 12 -i 40 -r 1 -c 0"),
        ("lulesh_omp", "-s 12 -i 40 -r 21 -b 2 -c 3"),
        ("lulesh_omp", "-s 48 -i 20"),
    ]
    if "--long" in sys.argv:
        runs += [("lulesh_omp", "-s 45 -r 1 -c 0"), ("lulesh_omp", "-s 60 -r 1 -c 0")]
    path = os.path.join(HERE, "ref_goldens.json")
    gold = json.load(open(path)) if os.path.exists(path) else {}
    for binary, args in runs:
        key = f"{binary} {args}"
        gold[key] = run(binary, args)
        print(key, gold[key]["cycles"], gold[key]["e0"])
    for np_, args in ((8, "-s 5"), (8, "-s 6"), (8, "-s 8 -i 100"), (8, "-s 10 -i 60"), (8, "-s 12 -i 120"),
                      (27, "-s 3 -i 60")):
        key = f"lulesh_mpi -np {np_} {args}"
        gold[key] = run_mpi(np_, args)
        print(key, gold[key]["cycles"], gold[key]["e0"])
    for tag in ("s90", "s128"):
        p = f"/tmp/gold/{tag}.out"
        if os.path.exists(p):
            for line in open(p):
                if line.startswith("REFJSON "):
                    rec = json.loads(line[len("REFJSON "):])
                    rec.pop("elapsed", None)
                    gold[f"lulesh_omp -s {rec['nx']} -r 1 -c 0"] = rec
    for tag in ("s256_i10", "s256_i40", "s384_i10"):
        p = f"/tmp/gold/{tag}.out"
        if os.path.exists(p):
            for line in open(p):
                if line.startswith("REFJSON "):
                    rec = json.loads(line[len("REFJSON "):])
                    rec.pop("elapsed", None)
                    gold[f"lulesh_omp -s {rec['nx']} -i {rec['cycles']} -r 1 -c 0"] = rec
    json.dump(gold, open(path, "w"), indent=1, sort_keys=True)

    for cyc in (9, 10):
        tmp = f"/tmp/ref_s8_c{cyc}.bin"
        run("lulesh_omp", f"-s 8 -i {cyc}", dump=tmp)
        np.savez_compressed(os.path.join(HERE, f"ref_s8_c{cyc}.npz"), **read_dump(tmp))
    for r in range(8):
        col, row, plane = r % 2, (r // 2) % 2, r // 4
        tmp = f"/tmp/ref_setup_{r}.bin"
        subprocess.run([os.path.join(REF, "ref_setup_dump"), "2", "4", str(col), str(row),
                        str(plane), "11", "1", "1", tmp], check=True)
        d = read_dump(tmp)
        keep = {k: d[k] for k in ("header", "scalars", "x", "y", "z", "nodalMass", "volo", "e",
                                  "lxim", "lxip", "letam", "letap", "lzetam", "lzetap", "elemBC",
                                  "nodelist", "symmX", "symmY", "symmZ") if k in d}
        np.savez_compressed(os.path.join(HERE, f"ref_setup_tp2_nx4_r{r}.npz"), **keep)


if __name__ == "__main__":
    main()
