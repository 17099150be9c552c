#!/usr/bin/env python
"""Turn the raw outputs of scripts/gpu_profiles.sh (gpurun_out/prof/) into the committed summaries
under profiles/ (round tag given on the command line, default r01)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "prof")
DST = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def launches():
    rows = []
    with open(os.path.join(SRC, "launches.csv")) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((int(row["ID"]), re.sub(r"\(.*", "", row["Kernel Name"]), float(row["Metric Value"].replace(",", ""))))
    til = [i for i, r in enumerate(rows) if "tile_letterbox_kernel" in r[1]]
    start = til[-1]  # last timed step
    stop = max(i for i, r in enumerate(rows) if "column_peaks_kernel" in r[1]) + 1  # what follows is the status check
    agg = collections.OrderedDict()
    for i in range(start, stop):
        agg[rows[i][1]] = agg.get(rows[i][1], 0.0) + rows[i][2]
    total = sum(agg.values())
    mine = {k: v for k, v in agg.items() if not k.startswith("void at::")}
    out = ["| kernel (one step = 64 pages, last timed step) | device time (us) | share of step |", "|---|---:|---:|"]
    for k, v in agg.items():
        out.append(f"| `{k[:70]}` | {v / 1e3:.1f} | {100 * v / total:.1f} % |")
    out.append(f"| **total** | {total / 1e3:.1f} | 100 % |")
    return "\n".join(out), agg, total, len(rows), sum(mine.values()) / total


def raw_rows(name):
    """One dict {metric: (value, unit)} per captured launch of gpurun_out/prof/<name>.ncu-rep ([] if absent)."""
    rep = os.path.join(SRC, name + ".ncu-rep")
    if not os.path.exists(rep):
        return []
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{h: (v[i], units[i]) for i, h in enumerate(hdr)} for v in rows[2:] if len(v) == len(hdr)]


def raw_metrics():
    return raw_rows("tiler_full")[0]


MERGE_KEYS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
]


def merge_profile():
    """cfg4 merge kernels (ncu --set full of one pg_nms_merge call): SM / L1 throughput per kernel."""
    rows = raw_rows("merge_full")
    if not rows:
        return
    plain = None
    try:
        plain = json.load(open(os.path.join(SRC, "merge_plain.json")))
    except Exception:
        pass
    with open(os.path.join(DST, f"{TAG}_merge_ncu_full.md"), "w") as f:
        f.write(f"# {TAG}: cfg4 merge (`pg_nms_merge`, 8 pages x 100 000 boxes) — ncu --set full --clock-control none, one call\n\n")
        f.write("Command: `python scripts/bench_merge_stress.py` (the six kernels of the fourth call; the three before it are warm-up).\n"
                "The merge is bound by SM issue (fp64 compares in the mask kernel) and shared-memory/L1 traffic, not by HBM: the\n"
                "whole input is 4.8 MB per page.  Durations under ncu are cold-cache and serialised.\n\n")
        if plain:
            f.write(f"Same command without ncu: `{json.dumps(plain)}`\n\n")
        names = [re.sub(r"\(.*", "", r["Kernel Name"][0]) for r in rows]
        f.write("| metric | " + " | ".join(f"`{n}`" for n in names) + " | unit |\n|---|" + "---:|" * len(names) + "---|\n")
        for k in MERGE_KEYS:
            if any(k in r for r in rows):
                f.write(f"| `{k}` | " + " | ".join(r.get(k, ("", ""))[0] for r in rows) + f" | {next(r[k][1] for r in rows if k in r)} |\n")
        f.write("\nTop warp stall reasons per issue-active cycle:\n\n")
        for n, r in zip(names, rows):
            st = sorted(((float(v[0].replace(",", "")), k.split("issue_stalled_")[1].split("_per")[0]) for k, v in r.items()
                         if "issue_stalled" in k and "per_issue_active" in k and v[0]), reverse=True)[:5]
            f.write(f"* `{n}`: " + ", ".join(f"{nm} {v:.2f}" for v, nm in st) + "\n")


def main():
    os.makedirs(DST, exist_ok=True)
    merge_profile()
    table, agg, total, n_launch, _ = launches()
    m = raw_metrics()

    def g(name):
        v, u = m[name]
        return float(v.replace(",", "")), u

    unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    rd, ru = g("dram__bytes_read.sum")
    wr, wu = g("dram__bytes_write.sum")
    dram = rd * unit_scale[ru] + wr * unit_scale[wu]
    dur, du = g("gpu__time_duration.sum")
    dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(du, 1e-6)
    bench = {k: json.load(open(os.path.join(SRC, f"bench_{k}.json"))) for k in ("default", "no_overlap", "tiler_only", "reference")}
    pages = bench["tiler_only"]["config"]["pages_per_gpu"]
    alg = bench["tiler_only"]["roofline"]["algorithmic_bytes_per_launch"]
    with open(os.path.join(DST, "tiler_traffic.json"), "w") as f:
        json.dump({"workload": "cfg3", "pages_per_launch": pages, "dram_bytes_per_launch": int(dram),
                   "algorithmic_bytes_per_launch": alg, "source": f"profiles/{TAG}_tiler_ncu_full.md (ncu --set full, 1 launch)"}, f, indent=1)
    keys = [
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    ]
    stalls = sorted(((float(v[0]), k.split("issue_stalled_")[1].split("_per")[0]) for k, v in m.items()
                     if "issue_stalled" in k and "per_issue_active" in k and v[0]), reverse=True)[:7]
    with open(os.path.join(DST, f"{TAG}_tiler_ncu_full.md"), "w") as f:
        f.write(f"# {TAG}: `tile_letterbox_kernel<2,3>` — ncu --set full --clock-control none, one launch\n\n")
        f.write(f"Command: `python bench.py --tiler-only --steps 1 --warmup 3 --no-cpu-baseline` ({pages} pages of 8000x6000 per launch, "
                f"4x4 grid).  Numbers under ncu are cold-cache and replayed; the bench values are in `{TAG}_bench.md`.\n\n")
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for k in keys:
            if k in m:
                f.write(f"| `{k}` | {m[k][0]} | {m[k][1]} |\n")
        f.write(f"\nDRAM traffic per launch = read + write = **{dram / 1e9:.3f} GB**; algorithmic bytes per launch = "
                f"{alg / 1e9:.3f} GB (ratio {dram / alg:.3f}).  Rows that no output row samples are skipped (-27 % of source rows) "
                f"while tile overlaps that miss L2 are re-read; the two nearly cancel.\n\n")
        f.write(f"Under ncu: {dram / dur_s / 1e12:.2f} TB/s of DRAM traffic over {dur_s * 1e3:.3f} ms.\n\n")
        f.write("Top warp stall reasons (per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + "\n")
    with open(os.path.join(DST, f"{TAG}_launches.md"), "w") as f:
        f.write(f"# {TAG}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("Command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-overlap` (default cfg3 workload).\n"
                "ncu serialises kernels, so this is the single-stream picture: compare shares with the `--no-overlap` bench line; in the\n"
                "default two-stream run the box kernels execute under the tiler and the step is as long as the tiler alone.\n\n")
        f.write(table + f"\n\n{n_launch} launches captured in total (page generation + warm-up + 2 timed steps).\n")
        til = next(v for k, v in agg.items() if "tile_letterbox" in k)
        f.write(f"\nTiler share of the serialised step: {100 * til / total:.1f} %; bench `--no-overlap` kernel_share_of_step: "
                f"{100 * bench['no_overlap']['roofline']['kernel_share_of_step']:.1f} %.\n")
    with open(os.path.join(DST, f"{TAG}_launches.csv"), "w") as f:
        f.write(open(os.path.join(SRC, "launches.csv")).read())
    # clocks during the default bench
    clk = []
    with open(os.path.join(SRC, "clocks.csv")) as f:
        rd_ = csv.reader(f)
        next(rd_, None)
        for r in rd_:
            try:
                clk.append((int(r[1].split()[0]), int(r[2].split()[0]), float(r[3].split()[0]), [x.strip() for x in r[4:]]))
            except Exception:
                pass
    with open(os.path.join(DST, f"{TAG}_bench.md"), "w") as f:
        f.write(f"# {TAG}: bench lines measured on a B200 (one `gpurun` box, same call as the ncu captures)\n\n")
        for k, title in (("default", "python bench.py"), ("no_overlap", "python bench.py --no-overlap --no-cpu-baseline"),
                         ("tiler_only", "python bench.py --tiler-only --no-cpu-baseline"),
                         ("reference", "python bench.py --impl reference --steps 3 --warmup 1")):
            f.write(f"## `{title}`\n\n```json\n{json.dumps(bench[k])}\n```\n\n")
        if clk:
            sm = sorted(c[0] for c in clk)
            f.write(f"nvidia-smi during the default run ({len(clk)} samples at 200 ms, idle gaps included): SM clock median {sm[len(sm) // 2]} MHz, "
                    f"max {max(sm)} MHz (limit {clk[0][1]} MHz), power max {max(c[2] for c in clk):.0f} W; "
                    f"hw_slowdown / hw_thermal / sw_thermal active in {sum(1 for c in clk if any('Active' == x for x in c[3][1:4]))} samples.\n")
    print("profiles written:", sorted(os.listdir(DST)))


if __name__ == "__main__":
    main()
