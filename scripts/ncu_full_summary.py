"""Summarise an `ncu --set full` report (one column per captured launch) as a markdown table.

    python scripts/ncu_full_summary.py gpurun_out/x.ncu-rep profiles/x.md "title" "command / context paragraph"
"""
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
]


def main():
    rep, out, title, ctx = sys.argv[1:5]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = [{h: (v[i], units[i]) for i, h in enumerate(hdr)} for v in rows[2:] if len(v) == len(hdr)]
    names = [re.sub(r"\(.*", "", r["Kernel Name"][0]).replace("<unnamed>::", "") for r in recs]
    with open(out, "w") as f:
        f.write(f"# {title}\n\n{ctx}\nDurations under ncu are cold-cache and serialised (`--clock-control none`).\n\n")
        f.write("| metric | " + " | ".join(f"`{n}`" for n in names) + " | unit |\n|---|" + "---:|" * len(names) + "---|\n")
        for k in KEYS:
            if any(k in r for r in recs):
                f.write(f"| `{k}` | " + " | ".join(r.get(k, ("", ""))[0] for r in recs) + f" | {next(r[k][1] for r in recs if k in r)} |\n")
        f.write("\nTop warp stall reasons per issue-active cycle:\n\n")
        for n, r in zip(names, recs):
            st = sorted(((float(v[0].replace(",", "")), k.split("issue_stalled_")[1].split("_per")[0]) for k, v in r.items()
                         if "issue_stalled" in k and "per_issue_active" in k and v[0]), reverse=True)[:5]
            f.write(f"* `{n}`: " + ", ".join(f"{nm} {v:.2f}" for v, nm in st) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
