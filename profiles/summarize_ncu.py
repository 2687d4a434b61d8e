#!/usr/bin/env python
"""Turns an `ncu --set full` report (brought back in gpurun_out/) into the small, tracked
summaries under profiles/: one markdown table + one json per capture, and profiles/traffic.json
(measured DRAM bytes per launch per kernel, read by bench.py for `roofline.traffic`).

    python profiles/summarize_ncu.py gpurun_out/prof_s128_v4.ncu-rep r01_s128 [--size 128]
"""
import csv, io, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
KNAME = {"k_force": "force_elem", "k_node": "node_update", "k_kinematics": "kinematics_grad",
         "k_material": "material", "k_time_increment": "time_increment"}
METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "dadd_pc"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "dmul_pc"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "dfma_pc"),
    ("smsp__cycles_elapsed.max", "cycles"),
]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val) * mult.get(unit, 1)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    size = int(sys.argv[sys.argv.index("--size") + 1]) if "--size" in sys.argv else 128
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in data:
        short = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
        name = KNAME.get(short, short)
        rec = {}
        for m, key in METRICS:
            if m not in ix:
                continue
            v, u = r[ix[m]], units[ix[m]]
            if key in ("dram_read", "dram_write"):
                rec[key + "_bytes"] = to_bytes(v, u)
            elif key == "time":
                rec["time_us"] = float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
            else:
                rec[key] = float(v)
        rec["dram_bytes"] = rec.get("dram_read_bytes", 0) + rec.get("dram_write_bytes", 0)
        out.setdefault(name, rec)   # first launch of each kernel
    json.dump(out, open(os.path.join(HERE, f"{tag}_ncu_summary.json"), "w"), indent=1)
    ne = size ** 3
    lines = [f"# ncu --set full summary `{os.path.basename(rep)}` (-s {size}, {ne} elements, one launch each)", "",
             "| kernel | time us | DRAM read MB | DRAM write MB | DRAM B/zone | DRAM % | FP64 pipe % | issue % | warps % | regs | grid x block | warp inst | DADD+DMUL+DFMA / zone | L1 hit % | L2 hit % | stall long_sb / wait / math |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for k, r in out.items():
        lines.append(f"| {k} | {r['time_us']:.1f} | {r['dram_read_bytes']/1e6:.1f} | {r['dram_write_bytes']/1e6:.1f} | "
                     f"{r['dram_bytes']/ne:.0f} | {r.get('dram_pct',0):.1f} | {r.get('fp64_pipe_pct',0):.1f} | "
                     f"{r.get('issue_active_pct',0):.1f} | {r.get('warps_active_pct',0):.1f} | {r.get('regs',0):.0f} | "
                     f"{r.get('grid',0):.0f} x {r.get('block',0):.0f} | {r.get('warp_inst',0):.3g} | "
                     f"{(r.get('dadd_pc',0) + r.get('dmul_pc',0) + r.get('dfma_pc',0)) * r.get('cycles',0) / ne:.0f} | "
                     f"{r.get('l1_hit_pct',0):.1f} | {r.get('l2_hit_pct',0):.1f} | "
                     f"{r.get('stall_long_sb',0):.2f} / {r.get('stall_wait',0):.2f} / {r.get('stall_math',0):.2f} |")
    lines += ["", "Times are under the profiler (cold caches, serialised); bench.py reports the live CUDA-event times.", ""]
    open(os.path.join(HERE, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines))
    tpath = os.path.join(HERE, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    traffic[f"s{size}"] = {k: r["dram_bytes"] for k, r in out.items()}
    traffic[f"s{size}"]["_source"] = f"{tag}: {os.path.basename(rep)}"
    json.dump(traffic, open(tpath, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
