"""Runs ON THE GPU BOX: one `ncu --set full` capture per dominant kernel of the bench legs, condensed
into gpurun_out/ncu_facts.json (-> profiles/ncu_facts.json) together with the hash of the CUDA
sources the library was built from, plus a markdown summary per kernel (-> profiles/).
    python tools/ncu_capture.py [tag] [key ...]
A capture only runs after the same command has exited 0 without ncu."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

import bench  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
# facts key -> (profile_leg key, kernel regex, launches of that kernel to skip)
CAPTURES = [
    ("cfg5", "cfg5", "loglike_delta_kernel", 2),
    ("default_model", "default_model", "loglike_delta_kernel", 2),
    ("ens", "ens", "ens_resident_kernel", 2),
    ("cfg2", "cfg2", "loglike_nodes_kernel", 2),
    ("cfg2_gauss", "cfg2_gauss", "loglike_gauss_thread_kernel", 2),
    ("cfg5p", "cfg5p", "loglike_nodes_kernel", 2),
    ("cfg5p_gauss", "cfg5p_gauss", "loglike_gauss_thread_kernel", 2),
    ("cfg3", "cfg3", "loglike_nodes_kernel", 2),
    ("cfg3_gauss", "cfg3_gauss", "loglike_nodes_kernel", 2),
    ("chain_dedupe", "chain", "chain_dedupe_kernel", 1),
    ("chain_unique", "chain", "chain_unique_kernel", 1),
    ("chain_lir_qags", "chain", "chain_lir_qags_kernel", 1),
    ("chain_lir", "chain", "chain_lir_kernel", 1),
]
EXTRA = {"local_ld": "smsp__sass_inst_executed_op_local_ld.sum", "local_st": "smsp__sass_inst_executed_op_local_st.sum",
         "long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
         "barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
         "wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
         "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}


def facts_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    f = lambda k: float(d[k][1].replace(",", ""))         # noqa: E731
    nsm = 148
    cyc = f("smsp__cycles_active.avg")
    n_all = f("smsp__inst_executed.sum")
    n_fp64 = f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100.0 * cyc / 2.0 * 4 * nsm
    res = {"kernel": ncu_summary.short(d["Kernel Name"][1]),
           "dram_bytes_per_launch": (ncu_summary._gb(d["dram__bytes_read.sum"]) +
                                     ncu_summary._gb(d["dram__bytes_write.sum"])) * 1e9,
           "gpu_time_ms": ncu_summary._ms(d["gpu__time_duration.sum"]),
           "fp64_pipe_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
           "issue_slot_pct": f("sm__inst_issued.avg.pct_of_peak_sustained_active"),
           "dram_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "warp_instructions": n_all, "fp64_warp_instructions": n_fp64,
           "issue_bound_frac": (2.0 * n_fp64 + (n_all - n_fp64)) / (4 * nsm) / cyc,
           "registers": f("launch__registers_per_thread")}
    for k, m in EXTRA.items():
        if m in d:
            res[k] = f(m)
    return res


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    want = set(sys.argv[2:])
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "ncu_facts.json")
    try:
        facts = json.load(open(path))
    except Exception:
        facts = {}
    sha = bench.csrc_sha()
    if facts.get("csrc_sha") != sha:
        facts = {"csrc_sha": sha, "source": "ncu --set full --clock-control none --import-source on, one launch "
                                            "per kernel (tools/ncu_capture.py on the GPU box)", "kernels": {}}
    units_done = {}
    for fkey, leg, rx, skip in CAPTURES:
        if want and fkey not in want:
            continue
        ujson = os.path.join(OUT, "units_%s.json" % leg)
        if leg not in units_done:
            rc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_leg.py"), leg, ujson],
                                capture_output=True, text=True)
            if rc.returncode != 0:
                print("leg %s failed without ncu:\n%s" % (leg, rc.stderr[-1500:]))
                continue
            units_done[leg] = json.load(open(ujson))
        rep = os.path.join(OUT, "%s_prof_%s" % (tag, fkey))
        rc = subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k",
                             "regex:" + rx, "-s", str(skip), "-c", "1", "-o", rep, "-f", sys.executable,
                             os.path.join(ROOT, "tools", "profile_leg.py"), leg], capture_output=True, text=True)
        rep += ".ncu-rep"
        if rc.returncode != 0 or not os.path.exists(rep):
            print("ncu failed for %s:\n%s" % (fkey, (rc.stdout + rc.stderr)[-1500:]))
            continue
        f = facts_of(rep)
        f["units_per_launch"] = units_done[leg].get(fkey, units_done[leg].get(leg))
        facts["kernels"][fkey] = f
        with open(os.path.join(OUT, "%s_full_%s.md" % (tag, fkey)), "w") as fh:
            so = sys.stdout
            sys.stdout = fh
            try:
                ncu_summary.full(rep)
            finally:
                sys.stdout = so
        if os.environ.get("KEEP_SRC") == fkey:
            with open(os.path.join(OUT, "%s_src_%s.csv" % (tag, fkey)), "w") as fh:
                fh.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                                        text=True).stdout)
        os.remove(rep)
        json.dump(facts, open(path, "w"), indent=1)
        print("%-16s %-40s %8.3f ms  fp64 %5.1f%%  issue %5.1f%%  units %d" %
              (fkey, f["kernel"][:40], f["gpu_time_ms"], f["fp64_pipe_pct"], f["issue_slot_pct"],
               f["units_per_launch"] or 0), flush=True)


if __name__ == "__main__":
    main()
