#!/usr/bin/env python
"""bench.py -- LULESH FOM (zone-cycles/s) of the B200-native Lagrange-leapfrog step.

A "step" is one cycle (TimeIncrement + LagrangeLeapFrog, lulesh.cc:2747-2748) of a synthetic
Sedov mesh.  Default workload: `-s 256` per GPU with the default regions (-r 11 -b 1 -c 1), the
size the north-star quotes its roofline target on; N GPUs = N ranks of a (px,py,pz)
decomposition, weak scaling (fixed elements per GPU).  `--global G` switches to strong scaling
(one G^3 mesh split over the ranks, BASELINE config 4 with G = 384).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S | --global G]
                  [--regions R --balance B --cost C] [--impl b200|reference] [--no-extras]

One JSON line on stdout (rank 0).  `value` is whole-job zone-cycles/s with the Domain resident
in HBM (CUDA events on the stream the kernels are launched on, max over ranks); `e2e` is the same
metric through the reference-facing C-ABI call sequence with HOST buffers (upload of the Domain
state, lulesh_b200_run, download of e()) inside the timed region; `roofline` is the dominant
kernel against the measured HBM bandwidth; `cpu_baseline` is the UNMODIFIED reference
(oracle/_ref, OpenMP on all host cores) on a bounded sample of the same workload.  Secondary
blocks (`extras`, skipped with --no-extras) carry the other BASELINE configs: config 2 (-s 128,
short window, developed window and the whole run to stoptime with the FOM lulesh-util.cc:185-227
would print), config 3 (-s 256 -r 16 -b 1 -c 8), config 4 (global 384^3, strong) and config 5
(-s 320 per GPU).  At N > 1 the line also carries `parity` (a physics self-check of the
multi-GPU path against the committed reference goldens, for both halo back ends; a failure makes
the exit code non-zero), `nccl_fallback` (the same workload with NCCL send/recv instead of
peer-to-peer stores) and `timeline` (per-cycle intervals on both streams).
`--impl reference` times the reference's own CPU build instead (rank 0 only); it never loads the
product library.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per zone-cycle per kernel (DESIGN.md "Kernels")
B_ALG = {"time_increment": 0, "force_elem": 320, "node_update": 332, "kinematics_grad": 176,
         "material": 212}
B_ALG_STEP = 1040   # 1072 canonical, 1040 with K4+K5 fused (ql,qq stay in registers)
NOMINAL_HBM_GBS = 8000.0   # the north-star's yardstick


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def decompose(n):
    """(px, py, pz) for n ranks: cubes as the reference (lulesh-init.cc:684-734), plus 1x1x2 and
    1x2x2.  Same rule as lulesh_host_decompose; restated here so that the reference arm does
    not have to load the product library."""
    c = round(n ** (1.0 / 3.0))
    if c ** 3 == n:
        return c, c, c
    if n == 2:
        return 1, 1, 2
    if n == 4:
        return 1, 2, 2
    raise SystemExit(f"unsupported rank count {n}")


def rank_sizes(args, n):
    px, py, pz = decompose(n)
    if args.glob:
        if args.glob % px or args.glob % py or args.glob % pz:
            raise SystemExit(f"--global {args.glob} is not divisible by {px}x{py}x{pz}")
        return (px, py, pz), (args.glob // px, args.glob // py, args.glob // pz)
    return (px, py, pz), (args.size,) * 3


def workload_config(args, n):
    """The named workload.  A pure function of the command line: both arms print the same dict."""
    (px, py, pz), (sx, sy, sz) = rank_sizes(args, n)
    regions = f"-r {args.regions} -b {args.balance} -c {args.cost}"
    if args.glob:
        what = (f"global {args.glob}^3 Sedov blast {regions} on {n} GPU(s), {sx}x{sy}x{sz} elements per GPU "
                f"(BASELINE config 4 when 384)")
    else:
        named = {(128, 11, 1, 1): "BASELINE config 2 mesh", (256, 11, 1, 1): "the north-star's roofline size",
                 (256, 16, 1, 8): "BASELINE config 3", (320, 11, 1, 1): "BASELINE config 5 mesh"}
        what = (f"-s {args.size} {regions} Sedov blast, {args.size}^3 elements per GPU ("
                + named.get((args.size, args.regions, args.balance, args.cost), "size override") + ")")
    return {"workload": what, "elements_per_gpu": sx * sy * sz, "decomposition": f"{px}x{py}x{pz}",
            "global_elements": n * sx * sy * sz,
            "timed_window": f"cycles {args.warmup + 1}..{args.warmup + args.steps} from the initial state",
            "l2_policy": "working set per cycle (>= 1.1 GB at -s 128, 9 GB at -s 256) exceeds the 126 MB L2; no flush needed",
            "eos_work_list": "b200 arm: the region lists of one repetition class are merged and sorted by element id "
                             "(indices still read from memory, every repetition executes); reference arm: per-region lists"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_ready(self, timeout=8.0):
        """nvidia-smi start-up (NVML init touches every GPU of the box) must be over before the
        timed region begins, or it perturbs the lock-stepped multi-GPU cycles."""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4)
                          if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# the reference's own CPU build (oracle/_ref: compiled unmodified from /root/reference)
# --------------------------------------------------------------------------------------------
def reference_binary():
    exe = os.path.join(ROOT, "oracle", "_ref", "lulesh_omp")
    if not os.path.exists(exe) and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=False)
    return exe if os.path.exists(exe) else None


def reference_cycle_budget(size, steps, seconds=60.0, nranks=1):
    """How many cycles of -s `size` the OpenMP reference gets through in `seconds` on this box
    (it does ~0.55e6 zone-cycles/s per host core: 8.5-11.9e6 measured on the 16-core GPU boxes,
    2.5e6 on 8 slower cores in BASELINE.md), capped at `steps`."""
    per_cycle = nranks * float(size) ** 3 / (0.55e6 * (os.cpu_count() or 1))
    return max(2, min(steps, int(seconds / per_cycle)))


def reference_sample_size(size, nranks, cycles, seconds=150.0):
    """The per-rank edge the reference arm actually runs: the requested one when `cycles` cycles of it
    fit the time budget and the host memory (~0.75 kB per zone and rank), else the largest smaller
    one that does.  zones/s of the reference is flat in the mesh size once it is out of cache."""
    try:
        avail = int(re.search(r"MemAvailable:\s+(\d+)", open("/proc/meminfo").read()).group(1)) * 1024
    except Exception:
        avail = 32 << 30
    for s in [size] + [c for c in (192, 160, 128, 96, 64, 48, 32) if c < size]:
        secs = cycles * nranks * float(s) ** 3 / (0.55e6 * (os.cpu_count() or 1))
        if secs <= seconds and nranks * 750.0 * s ** 3 <= 0.5 * avail:
            return s
    return min(size, 32)


def run_reference(size, cycles, regions=(11, 1, 1), nranks=1):
    """Runs the unmodified reference on the host cores for `cycles` cycles; returns
    (elapsed_s, cycles_done, cores, kind, what).  nranks == 1: the OpenMP build on all cores.
    nranks a cube (8, 27): the reference's USE_MPI=1 build, one process per rank, under the
    single-node MPI stand-in of oracle/mpishim ("MPI+OpenMP -np N"; this image has no MPI).
    Falls back to the oracle port if the reference binary did not travel."""
    cores = os.cpu_count() or 1
    r, b, c = regions
    exe, kind = reference_binary(), "reference"
    refdir = os.path.join(ROOT, "oracle", "_ref")
    use_mpi = nranks > 1 and os.path.exists(os.path.join(refdir, "lulesh_mpi")) \
        and os.path.exists(os.path.join(refdir, "mpirun_shim"))
    threads = max(1, cores // nranks) if use_mpi else cores
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false" if use_mpi else "close")
    if use_mpi:
        env["OMP_WAIT_POLICY"] = "passive"   # ranks x threads == cores: do not spin against each other
    else:
        env["OMP_PLACES"] = "cores"
    if exe is None:
        exe, kind = os.path.join(ROOT, "oracle", "_build", "lulesh_oracle"), "port"
        if not os.path.exists(exe):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    args = ["-s", str(size), "-i", str(cycles), "-r", str(r), "-b", str(b), "-c", str(c)]
    if use_mpi:
        cmd = [os.path.join(refdir, "mpirun_shim"), "-np", str(nranks), os.path.join(refdir, "lulesh_mpi")] + args
    else:
        cmd = [exe] + args
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True).stdout
    elapsed = float(re.search(r'"elapsed": ([0-9.eE+-]+)', out).group(1))
    done = int(re.search(r'"cycles": (\d+)', out).group(1))
    what = (f"mpirun_shim -np {nranks} lulesh_mpi (reference USE_MPI=1 build), OMP_NUM_THREADS={threads}"
            if use_mpi else f"lulesh_omp, OMP_NUM_THREADS={threads}")
    return elapsed, done, cores, kind, what


def impl_reference(args):
    """The reference arm: W untimed + K timed cycles of the reference's own code on the host cores.
    The reference binary times a whole run, so the K timed cycles are the difference of a run
    of W+K cycles and a run of W cycles."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    n = args.gpus
    cube = round(n ** (1.0 / 3.0)) ** 3 == n
    nranks = n if (n > 1 and cube and not args.glob) else 1
    (px, py, pz), sizes = rank_sizes(args, n)
    want = max(sizes) if not args.glob else args.glob
    total = 2 * args.warmup + args.steps
    size = reference_sample_size(want, nranks, total)
    regions = (args.regions, args.balance, args.cost)
    t0 = time.time()
    ea, ca, cores, kind, what = run_reference(size, args.warmup, regions, nranks)
    eb, cb, _, _, _ = run_reference(size, args.warmup + args.steps, regions, nranks)
    if cb > ca and eb > ea:
        secs, cycles = eb - ea, cb - ca
    else:            # tiny meshes: the run ended (stoptime) or the clock did not resolve the difference
        secs, cycles = eb, cb
    zones = nranks * float(size) ** 3
    zcs = zones * cycles / secs
    r, b, c = regions
    sample = (f"-s {size} -r {r} -b {b} -c {c}: cycles {ca + 1}..{cb} of a run ({secs:.2f} s = run of {cb} minus run of "
              f"{ca} cycles; {what})")
    if size != want:
        sample += f"; bounded sample: -s {size} instead of -s {want} to fit the time/memory budget of this box"
    layout = (f"{nranks} MPI ranks {px}x{py}x{pz} of -s {size}" if nranks > 1 else
              f"single domain -s {size}" + ("" if n == 1 else
                                            f" (the reference has no {n}-rank layout: cubic rank counts only, "
                                            "lulesh-init.cc:684-692)"))
    line = {
        "impl": "reference", "metric": "LULESH FOM (zone-cycles/s)", "value": zcs, "unit": "zones/s",
        "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(cycles, 1), "higher_is_better": True,
        "scaling": "strong" if args.glob else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n),
        "reference_layout": layout,
        "cpu_baseline": {"value": zcs, "unit": "zones/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": zcs, "unit": "zones/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
        "note": "reference's own CPU implementation on the host cores of this box: OpenMP build for one "
                "domain; for cubic rank counts its USE_MPI=1 build under a single-node MPI stand-in "
                "(oracle/mpishim; the image has no MPI)",
    }
    emit(line)
    return 0


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  NCCL (version banner, NCCL_DEBUG output) and other
    native libraries write to file descriptor 1 directly, so fd 1 is pointed at stderr for the
    whole run and the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# --------------------------------------------------------------------------------------------
# the b200 arm
# --------------------------------------------------------------------------------------------
class Job:
    """One process of the (possibly multi-rank) bench: torch.distributed plumbing + helpers."""

    def __init__(self, n):
        import lulesh_b200 as lb
        self.lb, self.n = lb, n
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world != n and n != 1:
            raise SystemExit(f"--gpus {n} needs a torchrun launch with WORLD_SIZE={n} (got {world})")
        self.dist = None
        if n > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def uid(self):
        """NCCL unique id for one communicator: made on rank 0, broadcast over torch.distributed."""
        if self.dist is None:
            return None
        import torch
        buf = torch.zeros(self.lb.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            buf.copy_(torch.frombuffer(bytearray(self.lb.get_unique_id()), dtype=torch.uint8))
        self.dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def allmax(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, obj):
        if self.dist is None:
            return [obj]
        out = [None] * self.n
        self.dist.all_gather_object(out, obj)
        return out

    def sedov(self, decomp, sizes, regions=(11, 1, 1)):
        """Device-side setup (lulesh_b200_create_sedov) of this rank's brick."""
        dev = self.lb.Device.sedov(sizes[0], *regions, num_ranks=self.n, rank=self.rank, decomp=decomp,
                                   sizes=sizes, device=self.local, unique_id=self.uid())
        dev.sum_nodal_mass()
        return dev

    def timed(self, dev, warmup, steps):
        """(ms, launches) of exactly `steps` cycles after `warmup` untimed ones: CUDA events on the
        launching stream, barrier on both sides, max over ranks.  Fails if the device-side loop
        did not execute every timed cycle (they would be no-op kernels)."""
        dev.time_cycles(warmup)
        c0 = dev.scalars.cycle
        self.barrier()
        ms, _, launches = dev.time_cycles(steps)
        ms = self.allmax(ms)
        self.barrier()
        done = dev.scalars.cycle - c0
        if done != steps:
            raise SystemExit(f"timed region ran {done} real cycles, not {steps}: the run reached stoptime; "
                             "use fewer --steps or a larger --size")
        return ms, launches


def per_kernel_block(lb, dev, cycles, ne):
    _, pk_ms, _ = dev.time_cycles(cycles, per_kernel=True)
    pk = {k: v / cycles for k, v in zip(lb.KERNEL_NAMES, pk_ms)}
    gbs = {k: (B_ALG[k] * ne / (pk[k] * 1e-3) / 1e9 if pk[k] > 0 else 0.0) for k in pk}
    return pk, gbs


def parity_check(job, goldens):
    """Physics self-check of the multi-GPU path (both halo back ends) against the committed
    reference goldens: N = 8 -> 2x2x2 ranks of -s 64 to stoptime == the reference's -s 128 run;
    N = 2 / 4 -> a global 90^3 mesh on 1x1x2 / 1x2x2 == the reference's -s 90 run.  Checks:
    identical cycle count, |e0 - e0_ref| / e0_ref <= 1e-8, global sum of e within 1e-8,
    time / dt / cycle bit-identical on all ranks, shared nodes bit-identical on both sides of
    every cut (this design has no CommSyncPosVel pass)."""
    import numpy as np
    lb, n = job.lb, job.n
    decomp = decompose(n)
    if n == 8:
        sizes, key = (64, 64, 64), "lulesh_omp -s 128 -r 1 -c 0"
    elif n in (2, 4):
        sizes, key = tuple(90 // p for p in decomp), "lulesh_omp -s 90 -r 1 -c 0"
    else:
        return {"status": "skipped", "why": f"no golden chain for {n} ranks"}
    gold = goldens[key]
    out = {"golden": key, "layout": f"{decomp[0]}x{decomp[1]}x{decomp[2]} ranks of {sizes[0]}x{sizes[1]}x{sizes[2]}",
           "tolerance_e0": 1e-8, "modes": {}}
    ok_all = True
    saved = os.environ.get("LULESH_B200_HALO")
    for mode in ("p2p", "nccl"):
        if mode == "nccl":
            os.environ["LULESH_B200_HALO"] = "nccl"
        else:
            os.environ.pop("LULESH_B200_HALO", None)
        dom = lb.Domain(sizes[0], 11, 1, 1, num_ranks=n, rank=job.rank, decomp=decomp, sizes=sizes)
        dev = lb.Device(dom, device=job.local, unique_id=job.uid())
        dev.sum_nodal_mass()
        got_mode = dev.halo_mode
        t0 = time.perf_counter()
        dev.run()
        secs = job.allmax(time.perf_counter() - t0)
        s = dev.scalars
        e = dev.download("e")
        fields = {f: dev.download(f) for f in "x y z xd yd zd".split()}
        dev.close()
        sum_e = job.allsum(float(np.sum(e)))
        scal = job.gather((s.cycle, s.time, s.deltatime, float(e[0])))
        plan = lb.halo_plan(dom)
        mine = {}
        for peer, cnt, soff in zip(plan["msg_rank"], plan["msg_count"], plan["msg_send_off"]):
            nodes = plan["bnode"][plan["pack_idx"][soff:soff + cnt]]
            mine[int(peer)] = np.stack([fields[f][nodes] for f in fields]).tobytes()
        everyone = job.gather(mine)
        shared_ok = all(everyone[peer][job.rank] == blob for peer, blob in mine.items())
        shared_ok = job.allsum(0.0 if shared_ok else 1.0) == 0.0
        e0 = scal[0][3]
        rec = {"halo": got_mode, "cycles": scal[0][0], "cycles_ref": gold["cycles"], "e0": e0, "e0_ref": gold["e0"],
               "e0_rel_err": abs(e0 - gold["e0"]) / gold["e0"],
               "sum_e_rel_err": abs(sum_e - gold["sum_e"]) / gold["sum_e"],
               "scalars_bit_identical_on_all_ranks": all(t[:3] == scal[0][:3] for t in scal),
               "shared_nodes_bit_identical": shared_ok, "seconds": secs,
               "zones_per_s": n * dom.numElem * scal[0][0] / secs}
        # a box without peer access falls back to NCCL inside the library: still a valid run of that back end
        mode_ok = got_mode == mode or (mode == "p2p" and got_mode == "nccl")
        rec["ok"] = bool(mode_ok and rec["cycles"] == gold["cycles"] and rec["e0_rel_err"] <= 1e-8 and
                         rec["sum_e_rel_err"] <= 1e-8 and rec["scalars_bit_identical_on_all_ranks"] and shared_ok)
        ok_all = ok_all and rec["ok"]
        out["modes"][mode] = rec
    if saved is None:
        os.environ.pop("LULESH_B200_HALO", None)
    else:
        os.environ["LULESH_B200_HALO"] = saved
    out["status"] = "ok" if ok_all else "FAILED"
    return out


def extras_block(job, args, goldens):
    """The other BASELINE configs, measured in the same run (device-side setup, short windows)."""
    import numpy as np
    lb, n = job.lb, job.n
    decomp = decompose(n)
    ex = {}

    def window(dev, ne_total, warm, steps, pk_cycles=0):
        ms, _ = job.timed(dev, warm, steps)
        rec = {"value": ne_total * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps, "warmup": warm}
        if pk_cycles:
            pk, _ = per_kernel_block(lb, dev, pk_cycles, ne_total // n)
            rec["per_kernel_ms"] = pk
        rec["step_frac_of_measured_hbm"] = B_ALG_STEP * rec["value"] / n / 1e9 / measured_peak_gbs()[0]
        rec["step_frac_of_nominal_8tbs"] = B_ALG_STEP * rec["value"] / n / 1e9 / NOMINAL_HBM_GBS
        return rec

    # config 2 mesh, -s 128 per GPU: early window (what round 1 reported)
    dev = job.sedov(decomp, (128,) * 3)
    ex["s128"] = dict(window(dev, n * 128 ** 3, 20, 200, 20), workload="-s 128 -r 11 -b 1 -c 1 per GPU, cycles 21..220")
    if n == 1:
        # ... and a window well into the blast (the fast paths for undisturbed zones no longer apply
        # to most of the mesh; cycle 3000 of 4561)
        dev.run(3000)
        ex["s128_developed"] = dict(window(dev, 128 ** 3, 0, 200, 20),
                                    workload="-s 128, cycles 3001..3200 of 4561 (blast wave well developed)")
    dev.close()
    if n == 1:
        # config 2 proper: the whole run to stoptime, timed like lulesh.cc:2737-2767, and the FOM
        # lulesh-util.cc:185-227 would print (thousands of zone-cycles per second)
        gold = goldens["lulesh_omp -s 128 -r 1 -c 0"]
        dev = job.sedov(decomp, (128,) * 3)
        t0 = time.perf_counter()
        dev.run()
        s = dev.scalars
        e = dev.download("e")
        secs = time.perf_counter() - t0
        dev.close()
        zps = 128 ** 3 * s.cycle / secs
        plane = e[:128 * 128].reshape(128, 128)
        iu = np.triu_indices(128, 1)
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.abs(plane[iu] - plane.T[iu]) / plane.T[iu]
        ex["config2_to_stoptime"] = {
            "workload": "-s 128 -r 11 -b 1 -c 1 run to stoptime (BASELINE config 2), host clock around "
                        "lulesh_b200_run + download(e)",
            "cycles": s.cycle, "cycles_ref": gold["cycles"], "seconds": secs, "zones_per_s": zps,
            "fom_as_printed": zps / 1000.0, "e0": float(e[0]), "e0_ref": gold["e0"],
            "e0_rel_err": abs(float(e[0]) - gold["e0"]) / gold["e0"],
            "max_rel_diff": float(np.nanmax(rel)), "max_rel_diff_ref": gold["max_rel_diff"],
            "step_frac_of_measured_hbm": B_ALG_STEP * zps / 1e9 / measured_peak_gbs()[0]}
        # config 3: region load-imbalanced EOS
        dev = job.sedov(decomp, (256,) * 3, (16, 1, 8))
        ex["config3"] = dict(window(dev, 256 ** 3, 5, 40, 10), workload="-s 256 -r 16 -b 1 -c 8, cycles 6..45")
        dev.close()
    # config 4: global 384^3, strong scaling (the driver's per-N lines give the curve)
    if not args.glob:
        sizes = tuple(384 // p for p in decomp)
        dev = job.sedov(decomp, sizes)
        ex["config4_global384"] = dict(window(dev, 384 ** 3, 3, 20),
                                       workload=f"global 384^3 on {n} GPU(s), {sizes[0]}x{sizes[1]}x{sizes[2]} per GPU",
                                       scaling="strong")
        dev.close()
    # config 5: -s 320 per GPU (weak); N = 1 is the denominator of its efficiency
    if n in (1, 8) and args.size != 320:
        dev = job.sedov(decomp, (320,) * 3)
        ex["config5_s320"] = dict(window(dev, n * 320 ** 3, 3, 20), workload="-s 320 -r 11 -b 1 -c 1 per GPU",
                                  scaling="weak")
        dev.close()
    return ex


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--size", type=int, default=256, help="elements per edge per GPU (-s)")
    ap.add_argument("--global", dest="glob", type=int, default=0,
                    help="strong scaling: edge of the GLOBAL mesh, split over the GPUs (384 = BASELINE config 4)")
    ap.add_argument("--regions", type=int, default=11)
    ap.add_argument("--balance", type=int, default=1)
    ap.add_argument("--cost", type=int, default=1)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary config blocks and nccl_fallback")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU physics self-check")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return impl_reference(args)

    import numpy as np
    import torch
    job = Job(args.gpus)
    lb, n, rank = job.lb, job.n, job.rank
    decomp, sizes = rank_sizes(args, n)
    regions = (args.regions, args.balance, args.cost)
    goldens = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_goldens.json")))

    dom = lb.Domain(sizes[0], *regions, num_ranks=n, rank=rank, decomp=decomp, sizes=sizes)
    ne_total = n * dom.numElem

    # ---------------- device-resident throughput (`value`)
    dev = lb.Device(dom, device=job.local, unique_id=job.uid())
    dev.sum_nodal_mass()
    sampler = ClockSampler(job.local) if rank == 0 else None
    if sampler:
        sampler.wait_ready()
    job.barrier()   # rank 0 may have waited seconds for nvidia-smi; the ranks of a cycle must start together
    ms, launches = job.timed(dev, args.warmup, args.steps)
    clocks = sampler.stop() if sampler else None
    value = ne_total * args.steps / (ms * 1e-3)
    halo_mode = dev.halo_mode

    # ---------------- per-kernel times, live, CUDA events on the launching streams (the shipped
    # two-stream schedule at several ranks; see lulesh_b200_timeline)
    pk_cycles = min(args.steps, 30)
    per_kernel, per_kernel_gbs = per_kernel_block(lb, dev, pk_cycles, dom.numElem)
    timeline = dev.timeline(min(args.steps, 30)) if n > 1 else None
    s_end = dev.scalars
    dev.close()

    peak, peak_src = measured_peak_gbs()
    dom_k = max((k for k in per_kernel if B_ALG[k] > 0), key=lambda k: per_kernel[k])
    k_bytes = B_ALG[dom_k] * dom.numElem
    achieved = k_bytes / (per_kernel[dom_k] * 1e-3) / 1e9
    step_gbs = B_ALG_STEP * (value / n) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_k, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k_bytes, "avg_launch_ms": per_kernel[dom_k],
                "per_kernel_ms": per_kernel, "per_kernel_gbs": per_kernel_gbs,
                "step": {"algorithmic_bytes_per_zone_cycle": B_ALG_STEP, "achieved": step_gbs,
                         "frac": step_gbs / peak, "frac_of_nominal_8tbs": step_gbs / NOMINAL_HBM_GBS}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            t = json.load(open(traffic_file)).get(f"s{sizes[0]}", {})
            roofline["traffic"] = t.get(dom_k)
            roofline["traffic_source"] = ("committed constant from profiles/traffic.json (ncu dram__bytes of one "
                                          "--set full capture of this kernel at this size), not measured in this run")
        except Exception:
            pass

    # ---------------- end to end through the C ABI with host buffers (`e2e`)
    # A second handle is created untimed (mesh tables and the NCCL communicator are setup, as
    # Domain construction and MPI_Init are in the reference, lulesh.cc:2663-2741).  The timed
    # region then does what a host driver does per run: copy the Domain STATE from pinned host
    # memory to the device through lulesh_b200_upload/set_scalars, run K cycles with
    # lulesh_b200_run (polling the control block every 64 cycles), and read back e() and the
    # scalars that VerifyAndWriteFinalOutput needs.
    state_fields = "x y z xd yd zd e p q v ss".split()
    pinned = {}
    for name in state_fields:
        src = dom.field(name)
        t = torch.empty(src.size, dtype=torch.float64).pin_memory()
        t.numpy()[:] = src
        pinned[name] = t
    e_out = torch.empty(dom.numElem, dtype=torch.float64).pin_memory()
    s0 = lb.Scalars.from_buffer_copy(dom.scalars)
    dev2 = lb.Device(dom, device=job.local, unique_id=job.uid())
    dev2.sum_nodal_mass()
    dev2.run(3)                                              # warm the graph / NCCL channels
    job.barrier()
    t0 = time.perf_counter()
    for name in state_fields:
        dev2.upload(name, pinned[name].numpy())              # H2D
    dev2.scalars = s0
    dev2.run(args.steps)                                     # the reference's timed loop
    dev2.download("e", e_out.numpy())                        # D2H of what the final report reads
    sc = dev2.scalars
    t1 = time.perf_counter()
    e2e_s = job.allmax(t1 - t0)
    h2d = sum(t.numel() * 8 for t in pinned.values()) + 96
    d2h = e_out.numel() * 8 + 96 * (1 + args.steps // 64)
    dev2.close()
    del pinned, e_out
    e2e = {"value": ne_total * sc.cycle / e2e_s, "unit": "zones/s",
           "h2d_bytes_per_step": h2d / max(sc.cycle, 1), "d2h_bytes_per_step": d2h / max(sc.cycle, 1),
           "seconds": e2e_s, "cycles": sc.cycle,
           "what": "upload of the 11 state arrays from pinned host memory + set_scalars + "
                   "lulesh_b200_run + download(e) + get_scalars, host clock, max over ranks"}

    # ---------------- several ranks: the NCCL back end on the same workload, and the physics check
    nccl_fallback = parity = None
    if n > 1 and not args.no_extras:
        os.environ["LULESH_B200_HALO"] = "nccl"
        dev3 = lb.Device(dom, device=job.local, unique_id=job.uid())
        os.environ.pop("LULESH_B200_HALO", None)
        dev3.sum_nodal_mass()
        ms3, _ = job.timed(dev3, args.warmup, args.steps)
        nccl_fallback = {"halo": dev3.halo_mode, "value": ne_total * args.steps / (ms3 * 1e-3),
                         "ms_per_step": ms3 / args.steps,
                         "what": "same workload with ncclSend/ncclRecv halo exchange and ncclAllReduce(min) for dt "
                                 "(LULESH_B200_HALO=nccl), cycles launched eagerly on two streams"}
        dev3.close()
    del dom
    if n > 1 and not args.no_parity:
        parity = parity_check(job, goldens)
    extras = None if args.no_extras else extras_block(job, args, goldens)

    rc = 0
    if parity is not None and parity.get("status") == "FAILED":
        rc = 3
    if rank != 0:
        if job.dist is not None:
            job.dist.destroy_process_group()
        return rc

    cpu = None   # reported at N=1 only (the reference arm, --impl reference, covers every N)
    if not args.no_cpu_baseline and n == 1:
        try:
            size = reference_sample_size(sizes[0], 1, 6, 30.0)
            cyc = reference_cycle_budget(size, 10 ** 9, 15.0)
            secs, done, cores, kind, what = run_reference(size, cyc, regions)
            cpu = {"value": float(size) ** 3 * done / secs, "unit": "zones/s", "cores": cores, "kind": kind,
                   "sample": f"-s {size} -i {done} -r {regions[0]} -b {regions[1]} -c {regions[2]} ({secs:.2f} s, {what})"}
        except Exception as ex:  # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "zones/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"failed: {ex}"}

    line = {
        "metric": "LULESH FOM (zone-cycles/s)", "value": value, "unit": "zones/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.glob else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, n), "halo": halo_mode,
        "fom_reference_units": value / 1000.0,
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "clocks": clocks,
        "state": {"cycle": s_end.cycle, "time": s_end.time, "dt": s_end.deltatime},
    }
    if timeline is not None:
        line["timeline_ms"] = timeline
    if nccl_fallback is not None:
        line["nccl_fallback"] = nccl_fallback
    if parity is not None:
        line["parity"] = parity
    if extras is not None:
        line["extras"] = extras
    emit(line)
    if job.dist is not None:
        job.dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
