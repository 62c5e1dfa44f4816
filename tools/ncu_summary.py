"""Condense ncu output brought back in gpurun_out/ into small tracked summaries.

    python tools/ncu_summary.py launches gpurun_out/launches_cfg5.csv > profiles/r01_launches_cfg5.md
    python tools/ncu_summary.py full gpurun_out/prof_cfg5.ncu-rep > profiles/r01_full_cfg5.md

`launches`: per-kernel count / total / mean device time and share of the run
(ncu serialises launches with cold caches: compare shares, not absolutes).
`full`: the handful of `--set full` metrics the roofline discussion uses.
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("void ", "")
    cut = name.find("(")
    return name[:cut] if cut > 0 else name


def launches(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    for r in rd:
        rows.append((short(r[ik]), float(r[iv].replace(",", "")), r[ig], r[ib]))
    agg = OrderedDict()
    for k, ns, g, b in rows:
        a = agg.setdefault(k, [0, 0.0, g, b])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | grid | block | total ms | mean ms | share |")
    print("|---|---:|---|---|---:|---:|---:|")
    for k, (n, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %s | %s | %.3f | %.4f | %.1f%% |" % (k[:90], n, g, b, ns / 1e6, ns / n / 1e6, 100 * ns / total))
    print("\n%d launches, %.3f ms of device time in total" % (len(rows), total / 1e6))


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("### `%s`\n" % short(d["Kernel Name"][1]))
        print("| metric | unit | value |\n|---|---|---:|")
        for k in WANT:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
        fp64_pct = float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][1])
        cyc = float(d["smsp__cycles_active.avg"][1].replace(",", ""))
        nsm = 148
        fp64_inst = fp64_pct / 100.0 * cyc / 2.0 * 4 * nsm
        tot = float(d["smsp__inst_executed.sum"][1].replace(",", ""))
        print("\nderived: ~%.3e FP64-pipe warp instructions (%.1f%% of the %.3e executed); "
              "DRAM traffic %.3f GB\n" % (fp64_inst, 100 * fp64_inst / tot, tot,
                                          _gb(d["dram__bytes_read.sum"]) + _gb(d["dram__bytes_write.sum"])))


def _gb(uv):
    u, v = uv
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u]


def facts(path):
    """One JSON object per captured kernel: the figures bench.py attaches to its `roofline`
    (profiles/ncu_facts.json is assembled from these, keyed by workload)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        f = lambda k: float(d[k][1].replace(",", ""))
        nsm = 148
        cyc = f("smsp__cycles_active.avg")
        n_all = f("smsp__inst_executed.sum")
        n_fp64 = f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100.0 * cyc / 2.0 * 4 * nsm
        # issue model measured by tools/exp_probe.cu: an FP64 warp instruction takes two issue cycles
        bound_cyc = (2.0 * n_fp64 + (n_all - n_fp64)) / (4 * nsm)
        print(json.dumps({
            "kernel": short(d["Kernel Name"][1]),
            "source": "ncu --set full --clock-control none, one launch (profiles/)",
            "dram_bytes_per_launch": (_gb(d["dram__bytes_read.sum"]) + _gb(d["dram__bytes_write.sum"])) * 1e9,
            "gpu_time_ms": _ms(d["gpu__time_duration.sum"]),
            "fp64_pipe_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_slot_pct": f("sm__inst_issued.avg.pct_of_peak_sustained_active"),
            "dram_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "warp_instructions": n_all,
            "fp64_warp_instructions": n_fp64,
            "issue_bound_frac": bound_cyc / cyc,
            "issue_bound_note": "(2*N_fp64 + N_other) issue cycles per scheduler / active cycles; see DESIGN.md 7",
            "registers": f("launch__registers_per_thread"),
            "smem_bank_conflicts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")}))


def _ms(uv):
    u, v = uv
    v = float(v.replace(",", ""))
    return v * {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6,
                "second": 1e3}[u]


if __name__ == "__main__":
    {"launches": launches, "full": full, "facts": facts}[sys.argv[1]](sys.argv[2])
