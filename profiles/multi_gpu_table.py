#!/usr/bin/env python
"""Builds profiles/r02_multi_gpu.md from the bench lines of the 2- and 8-GPU boxes
(tools/gpu_job_multi.sh) and the one-GPU line of the final build.

    python profiles/multi_gpu_table.py
"""
import json, os

HERE = os.path.dirname(os.path.abspath(__file__))
L = lambda f: json.load(open(os.path.join(HERE, f)))


def main():
    n1 = L("r02_bench_n1_driver_like.json")
    n1_8, n8s, n8n = L("r02_8gpubox_bench_n1_s128.json"), L("r02_8gpubox_bench_n8_s128.json"), L("r02_8gpubox_bench_n8_s128_nccl.json")
    n8, n4 = L("r02_8gpubox_bench_n8.json"), L("r02_8gpubox_bench_n4.json")
    n1_2, n2s, n2n, n2 = (L("r02_2gpubox_bench_n1_s128.json"), L("r02_2gpubox_bench_n2_s128.json"),
                          L("r02_2gpubox_bench_n2_s128_nccl.json"), L("r02_2gpubox_bench_n2.json"))
    out = ["# Round-2 multi-GPU measurements (`tools/gpu_job_multi.sh`; `bench.py` under torchrun, one rank per GPU)", "",
           "All multi-rank parity tests green on the same boxes (`r02_8gpubox_pytest_multirank.log`: 14 passed; "
           "`r02_2gpubox_pytest_multirank.log`: 8 passed, 6 skipped for lack of GPUs).", "",
           "## Weak scaling", "",
           "| per GPU | N | decomposition | halo | G z/s | ms/cycle | efficiency vs N=1 | `parity` block |", "|---|---|---|---|---|---|---|---|"]

    def row(tag, n, d, base, par=True):
        p = d.get("parity") or {}
        out.append(f"| {tag} | {n} | {d['config']['decomposition']} | {d['halo']} | {d['value']/1e9:.2f} | "
                   f"{d['ms_per_step']:.4f} | {base/d['ms_per_step']:.3f} | {p.get('status', '-') if par else '-'} |")

    b256 = n1["ms_per_step"]
    row("-s 256", 1, n1, b256, False); row("-s 256", 2, n2, b256); row("-s 256", 4, n4, b256); row("-s 256", 8, n8, b256)
    for n, d in ((2, n2), (8, n8)):
        nf = d["nccl_fallback"]
        out.append(f"| -s 256 | {n} | {d['config']['decomposition']} | nccl | {nf['value']/1e9:.2f} | {nf['ms_per_step']:.4f} | "
                   f"{b256/nf['ms_per_step']:.3f} | ok (same run) |")
    row("-s 128 (2-GPU box)", 1, n1_2, n1_2["ms_per_step"], False); row("-s 128", 2, n2s, n1_2["ms_per_step"], False)
    row("-s 128", 2, n2n, n1_2["ms_per_step"], False)
    row("-s 128 (8-GPU box)", 1, n1_8, n1_8["ms_per_step"], False); row("-s 128", 8, n8s, n1_8["ms_per_step"], False)
    row("-s 128", 8, n8n, n1_8["ms_per_step"], False)
    e1, e8, e2 = n1["extras"], n8["extras"], n2["extras"]
    c4 = lambda e: e["config4_global384"]
    out += ["", "The `-s 256` one-GPU line is the final build on a one-GPU box; the `-s 128` ones ran on the same box as their N > 1 lines.", "",
            "## BASELINE configs 4 and 5 (`extras` blocks of the N = 1 / 2 / 8 lines)", "",
            "| config | N | G z/s | ms/cycle | efficiency |", "|---|---|---|---|---|",
            f"| 4: global 384^3 (strong) | 1 | {c4(e1)['value']/1e9:.2f} | {c4(e1)['ms_per_step']:.3f} | 1 |",
            f"| 4: global 384^3 (strong) | 2 | {c4(e2)['value']/1e9:.2f} | {c4(e2)['ms_per_step']:.3f} | {c4(e2)['value']/c4(e1)['value']/2:.3f} |",
            f"| 4: global 384^3 (strong) | 8 | {c4(e8)['value']/1e9:.2f} | {c4(e8)['ms_per_step']:.3f} | {c4(e8)['value']/c4(e1)['value']/8:.3f} |",
            f"| 5: -s 320 per GPU (weak) | 1 | {e1['config5_s320']['value']/1e9:.2f} | {e1['config5_s320']['ms_per_step']:.3f} | 1 |",
            f"| 5: -s 320 per GPU (weak) | 8 | {e8['config5_s320']['value']/1e9:.2f} | {e8['config5_s320']['ms_per_step']:.3f} | "
            f"{e8['config5_s320']['value']/e1['config5_s320']['value']/8:.3f} |",
            "", "## Per-cycle timeline at N = 8 (ms; `lulesh_b200_timeline`, eager launches on the shipped two-stream schedule)", "",
            "| workload | halo | K1 | K2 interior | wait node chain | K3 | K45 interior | wait MonoQ | cycle (eager) | cycle (timed run) | comm: dt chain | comm: node chain | comm: MonoQ + K45 face layer |",
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for tag, d in (("-s 256", n8), ("-s 128", n8s), ("-s 128", n8n)):
        t = d["timeline_ms"]
        out.append(f"| {tag} | {d['halo']} | {t['k1_force']:.4f} | {t['k2_node']:.4f} | {t['node_join_wait']:.4f} | {t['k3_kinematics']:.4f} | "
                   f"{t['k45_interior']:.4f} | {t['monoq_join_wait']:.4f} | {t['cycle']:.4f} | {d['ms_per_step']:.4f} | {t['comm_dt']:.4f} | "
                   f"{t['comm_node']:.4f} | {t['comm_monoq']:.4f} |")
    pk = n1_8["roofline"]["per_kernel_ms"]
    out += ["", f"One GPU of the same box at -s 128, for comparison: K1 {pk['force_elem']:.4f}, K2 {pk['node_update']:.4f}, "
                f"K3 {pk['kinematics_grad']:.4f}, K45 {pk['material']:.4f}, cycle (graph) {n1_8['ms_per_step']:.4f} ms.",
            "", "Reading: all three exchanges are hidden (the main stream waits 2.7 us at each join, the latency of an event edge); the dt chain",
            "(0.16 ms at -s 128, mostly waiting for the slowest rank's previous cycle) ends before K1 does, the shared-node chain before K2,",
            "MonoQ + the face-layer K45 before the interior K45.  What remains of the 60 us at -s 128 on 8 GPUs is (i) kernels running 3-7 us",
            "longer next to the exchange kernels that share their SMs (K2 +7, K1 +3, K45 +3), (ii) lock-step jitter: every cycle ends with the",
            "slowest of 8 ranks, and (iii) the two-stream graph itself: its replay is no faster than eager launches (0.568 vs 0.569 ms), while",
            "the one-stream graph of a single rank saves 33 us per cycle over eager launches (0.508 vs 0.541 ms).", ""]
    open(os.path.join(HERE, "r02_multi_gpu.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
