#!/usr/bin/env python
"""bench.py -- LULESH FOM (zone-cycles/s) of the B200-native Lagrange-leapfrog step.

A "step" is one cycle (TimeIncrement + LagrangeLeapFrog, lulesh.cc:2747-2748) of a
synthetic Sedov mesh with `--size`^3 elements per GPU (default: BASELINE config 2,
-s 128, default regions -r 11 -b 1 -c 1).  N GPUs = N ranks of a (px,py,pz)
decomposition with NCCL halo exchange, weak scaling (fixed elements per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl b200|reference]

One JSON line on stdout (rank 0).  `value` is whole-job zone-cycles/s with the
Domain resident in HBM (CUDA events on the stream the kernels are launched on, max
over ranks); `e2e` is the same metric through the reference-facing C-ABI call
sequence with HOST buffers (lulesh_b200_create = H2D of the Domain, run, download of
e(), destroy) inside the timed region; `roofline` is the dominant kernel against the
measured HBM bandwidth; `cpu_baseline` is the UNMODIFIED reference (oracle/_ref,
OpenMP on all host cores) on a bounded sample of the same workload.
`--impl reference` times that reference build instead of the GPU path.
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


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def reference_binary():
    exe = os.path.join(ROOT, "oracle", "_ref", "lulesh_omp")
    if not os.path.exists(exe) and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=False)
    return exe if os.path.exists(exe) else None


def run_reference(size, cycles, regions=(11, 1, 1), nranks=1):
    """Times the unmodified reference on the host cores for `cycles` cycles; returns
    (zone_cycles_per_s, cores, kind, sample).  nranks == 1: the OpenMP build on all cores.
    nranks a cube (8, 27): the reference's USE_MPI=1 build, one process per rank, under the
    single-node MPI stand-in of oracle/mpishim ("MPI+OpenMP -np N"; this image has no MPI).
    Other rank counts do not exist in the reference (cubic layouts only, lulesh-init.cc:684):
    the single-domain OpenMP run of the same per-rank size is reported instead.
    Falls back to the oracle port if the reference binary did not travel."""
    cores = os.cpu_count() or 1
    r, b, c = regions
    exe, kind = reference_binary(), "reference"
    refdir = os.path.join(ROOT, "oracle", "_ref")
    cube = round(nranks ** (1.0 / 3.0)) ** 3 == nranks
    use_mpi = nranks > 1 and cube and os.path.exists(os.path.join(refdir, "lulesh_mpi")) \
        and os.path.exists(os.path.join(refdir, "mpirun_shim"))
    threads = max(1, cores // nranks) if use_mpi else cores
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false" if use_mpi else "close")
    if use_mpi:
        env["OMP_WAIT_POLICY"] = "passive"   # ranks x threads == cores: do not spin against each other
    if not use_mpi:
        env["OMP_PLACES"] = "cores"
    if exe is None:
        exe, kind = os.path.join(ROOT, "oracle", "_build", "lulesh_oracle"), "port"
        if not os.path.exists(exe):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    args = ["-s", str(size), "-i", str(cycles), "-r", str(r), "-b", str(b), "-c", str(c)]
    if use_mpi:
        cmd = [os.path.join(refdir, "mpirun_shim"), "-np", str(nranks), os.path.join(refdir, "lulesh_mpi")] + args
        ranks_done = nranks
    else:
        cmd = [exe] + args
        ranks_done = 1
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True).stdout
    m = re.search(r'"elapsed": ([0-9.eE+-]+)', out)
    n = re.search(r'"cycles": (\d+)', out)
    elapsed, done = float(m.group(1)), int(n.group(1))
    zcs = ranks_done * float(size) ** 3 * done / elapsed
    what = (f"mpirun_shim -np {nranks} lulesh_mpi (reference USE_MPI=1 build), OMP_NUM_THREADS={threads}"
            if use_mpi else f"lulesh_omp, OMP_NUM_THREADS={threads}")
    return zcs, cores, kind, f"-s {size} -i {done} -r {r} -b {b} -c {c} ({elapsed:.2f} s, {what})"


def reference_cycle_budget(size, steps, seconds=60.0):
    # the OpenMP reference does ~0.55e6 zone-cycles/s per host core (8.5-11.9e6 measured on the
    # 16-core GPU boxes, 2.5e6 on 8 slower cores in BASELINE.md)
    per_cycle = float(size) ** 3 / (0.55e6 * (os.cpu_count() or 1))
    return max(2, min(steps, int(seconds / per_cycle)))


def impl_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cube = round(args.gpus ** (1.0 / 3.0)) ** 3 == args.gpus
    nranks = args.gpus if (args.gpus > 1 and cube) else 1
    cycles = reference_cycle_budget(args.size, args.steps + args.warmup, 60.0 / nranks)
    t0 = time.time()
    zcs, cores, kind, sample = run_reference(args.size, cycles, (args.regions, args.balance, args.cost), nranks)
    line = {
        "impl": "reference", "metric": "LULESH FOM (zone-cycles/s)", "value": zcs, "unit": "zones/s",
        "n_gpus": args.gpus, "steps": cycles, "warmup": 0,
        "ms_per_step": 1e3 * nranks * float(args.size) ** 3 / zcs, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, nranks),
        "cpu_baseline": {"value": zcs, "unit": "zones/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": zcs, "unit": "zones/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
        "note": "reference's own CPU implementation on the host cores of this box: OpenMP build for one "
                "domain; for cubic rank counts its USE_MPI=1 build under a single-node MPI stand-in "
                "(oracle/mpishim; the image has no MPI); 2 and 4 ranks do not exist in the reference",
    }
    emit(line)
    return 0


def workload_config(args, n):
    import lulesh_b200 as lb
    px, py, pz = lb.decompose(n)
    return {"workload": f"-s {args.size} -r {args.regions} -b {args.balance} -c {args.cost} "
                        f"Sedov blast, {args.size}^3 elements per GPU, fixed -i (BASELINE config "
                        f"{'2' if args.size == 128 else 'size override'})",
            "elements_per_gpu": args.size ** 3, "decomposition": f"{px}x{py}x{pz}",
            "global_elements": n * args.size ** 3,
            "l2_policy": "working set per cycle (>= 1.1 GB at -s 128) exceeds the 126 MB L2; no flush needed"}


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


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--size", type=int, default=128, help="elements per edge per GPU (-s)")
    ap.add_argument("--regions", type=int, default=11)
    ap.add_argument("--balance", type=int, default=1)
    ap.add_argument("--cost", type=int, default=1)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return impl_reference(args)

    import numpy as np
    import lulesh_b200 as lb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.gpus
    if world != n:
        if n == 1:
            world = 1
        else:
            raise SystemExit(f"--gpus {n} needs a torchrun launch with WORLD_SIZE={n} (got {world})")

    dist = None
    if n > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def fresh_uid():
        """NCCL unique id for one communicator: made on rank 0, broadcast over torch.distributed."""
        if dist is None:
            return None
        import torch
        buf = torch.zeros(lb.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(lb.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dom = lb.Domain(args.size, args.regions, args.balance, args.cost, num_ranks=n, rank=rank)
    ne_total = n * dom.numElem

    # ---------------- device-resident throughput (`value`)
    dev = lb.Device(dom, device=local, unique_id=fresh_uid())
    dev.sum_nodal_mass()
    sampler = ClockSampler(local) if rank == 0 else None
    dev.time_cycles(args.warmup)
    if sampler:
        sampler.wait_ready()
    barrier()
    ms, _, launches = dev.time_cycles(args.steps)
    ms = allmax(ms)
    barrier()
    clocks = sampler.stop() if sampler else None
    value = ne_total * args.steps / (ms * 1e-3)
    halo_mode = dev.halo_mode

    # ---------------- per-kernel times, live, CUDA events on the launching stream
    pk_cycles = min(args.steps, 50)
    _, pk_ms, _ = dev.time_cycles(pk_cycles, per_kernel=True)
    per_kernel = {k: v / pk_cycles for k, v in zip(lb.KERNEL_NAMES, pk_ms)}
    s_end = dev.scalars
    dev.close()

    peak, peak_src = measured_peak_gbs()
    dom_k = max((k for k in per_kernel if B_ALG[k] > 0), key=lambda k: per_kernel[k])
    k_bytes = B_ALG[dom_k] * dom.numElem
    achieved = k_bytes / (per_kernel[dom_k] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_k, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k_bytes, "avg_launch_ms": per_kernel[dom_k],
                "per_kernel_ms": per_kernel,
                "per_kernel_gbs": {k: (B_ALG[k] * dom.numElem / (per_kernel[k] * 1e-3) / 1e9 if per_kernel[k] > 0 else 0.0)
                                   for k in per_kernel},
                "step": {"algorithmic_bytes_per_zone_cycle": B_ALG_STEP,
                         "achieved": B_ALG_STEP * (value / n) / 1e9, "frac": B_ALG_STEP * (value / n) / 1e9 / peak}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            t = json.load(open(traffic_file)).get(f"s{args.size}", {})
            roofline["traffic"] = t.get(dom_k)
        except Exception:
            pass

    # ---------------- end to end through the C ABI with host buffers (`e2e`)
    # A second handle is created untimed (mesh tables and the NCCL communicator are setup, as
    # Domain construction and MPI_Init are in the reference, lulesh.cc:2663-2741).  The timed
    # region then does what a host driver does per run: copy the Domain STATE from pinned host
    # memory to the device through lulesh_b200_upload/set_scalars, run K cycles with
    # lulesh_b200_run (polling the control block every 64 cycles), and read back e() and the
    # scalars that VerifyAndWriteFinalOutput needs.
    import torch
    state_fields = "x y z xd yd zd e p q v ss".split()
    pinned = {}
    for name in state_fields:
        src = dom.field(name)
        t = torch.empty(src.size, dtype=torch.float64).pin_memory()
        t.numpy()[:] = src
        pinned[name] = t
    e_out = torch.empty(dom.numElem, dtype=torch.float64).pin_memory()
    s0 = lb.Scalars.from_buffer_copy(dom.scalars)
    dev2 = lb.Device(dom, device=local, unique_id=fresh_uid())
    dev2.sum_nodal_mass()
    dev2.run(3)                                              # warm the graph / NCCL channels
    barrier()
    t0 = time.perf_counter()
    for name in state_fields:
        dev2.upload(name, pinned[name].numpy())              # H2D
    dev2.scalars = s0
    dev2.run(args.steps)                                     # the reference's timed loop
    dev2.download("e", e_out.numpy())                        # D2H of what the final report reads
    sc = dev2.scalars
    t1 = time.perf_counter()
    e2e_s = allmax(t1 - t0)
    h2d = sum(t.numel() * 8 for t in pinned.values()) + 96
    d2h = e_out.numel() * 8 + 96 * (1 + args.steps // 64)
    dev2.close()
    e2e = {"value": ne_total * sc.cycle / e2e_s, "unit": "zones/s",
           "h2d_bytes_per_step": h2d / max(sc.cycle, 1), "d2h_bytes_per_step": d2h / max(sc.cycle, 1),
           "seconds": e2e_s, "cycles": sc.cycle,
           "what": "upload of the 11 state arrays from pinned host memory + set_scalars + "
                   "lulesh_b200_run + download(e) + get_scalars, host clock, max over ranks"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    cpu = None   # reported at N=1 only (the reference arm, --impl reference, covers every N)
    if not args.no_cpu_baseline and n == 1:
        try:
            cyc = reference_cycle_budget(args.size, 10**9, 15.0)
            zcs, cores, kind, sample = run_reference(args.size, cyc, (args.regions, args.balance, args.cost))
            cpu = {"value": zcs, "unit": "zones/s", "cores": cores, "kind": kind, "sample": sample}
        except Exception as ex:  # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "zones/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"failed: {ex}"}

    line = {
        "metric": "LULESH FOM (zone-cycles/s)", "value": value, "unit": "zones/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": dict(workload_config(args, n), halo=halo_mode),
        "fom_reference_units": value / 1000.0,
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "clocks": clocks,
        "state": {"cycle": s_end.cycle, "time": s_end.time, "dt": s_end.deltatime},
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
