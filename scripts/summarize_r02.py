"""profiles/r02_*.md from the files scripts/gpu_r2_final.sh (and gpu_r2_two.sh / gpu_r2_eight.sh) leave in gpurun_out/.

    python scripts/summarize_r02.py
"""
import collections
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

REF_CLI = [  # scripts/ref_cli_baseline.py --pages 2 in the build container (needs /root/reference)
    {"reference_stage": 2, "pages": 2, "boxes_per_page_in": 10000, "seconds": 16.23, "seconds_per_page": 8.11, "json_files_written": 8, "cores": 1, "host": "build container, 8 vCPU"},
    {"reference_stage": 3, "pages": 2, "boxes_per_page_in": 10000, "seconds": 29.63, "seconds_per_page": 14.81, "json_files_written": 2, "cores": 1, "host": "build container, 8 vCPU"},
    {"reference_stage": 4, "pages": 2, "boxes_per_page_in": 10000, "seconds": 2.11, "seconds_per_page": 1.05, "json_files_written": 2, "cores": 1, "host": "build container, 8 vCPU"},
    {"reference_stage": 5, "pages": 2, "boxes_per_page_in": 10000, "seconds": 6.43, "seconds_per_page": 3.22, "json_files_written": 2, "cores": 1, "host": "build container, 8 vCPU"},
]


def last_json(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ui, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    out = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum" and r[ui] in ("ns", "nsecond"):
            v /= 1e3
        name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        out.setdefault((r[ii], name), {})[r[mi]] = v
    return out


def configs():
    out = ["# r02: bench lines and command-line timings (B200, this round)\n",
           "Raw JSON lines as printed by the commands named; one B200 unless stated.  Reference-side numbers that need",
           "`/root/reference` were taken in the build container (8 vCPU) and say so.\n"]
    d = last_json(os.path.join(G, "bench_default.json"))
    out += ["## `python bench.py` (default: cfg3, N = 1)\n", "```json\n" + json.dumps(d) + "\n```\n"]
    su = d["roofline"]["sustained"]
    out.append(f"* value {d['value']:.0f} pages/s ({d['ms_per_step']:.3f} ms per 64-page step); tiler in-step {d['roofline']['frac']:.3f} of the copy "
               f"peak, isolated {d['roofline']['isolated']['frac']:.3f}, sustained ({su['seconds']:.1f} s, {su['clocks']['sm_mhz']} MHz, "
               f"{su['clocks']['reasons']}) {su['frac']:.3f}")
    e = d["e2e"]
    out.append(f"* e2e {e['value']:.0f} pages/s on {e['h2d_bytes_per_step'] / 1e6:.0f} MB of JPEG files + detections per {e['pages_per_step']}-page "
               f"step ({e['h2d_gb_per_s_per_gpu']:.1f} GB/s host->device); three traced steps [copy start, copy end, compute start, compute end] ms: "
               f"{e['timeline_ms']['steps']}")
    out.append(f"* cpu_baseline {d['cpu_baseline']['value']:.2f} pages/s on {d['cpu_baseline']['cores']} cores; corpus sha {d['corpus']['hist_sha256']}\n")
    r = last_json(os.path.join(G, "bench_reference.json"))
    out += ["## `python bench.py --impl reference --steps 2 --warmup 1`\n", "```json\n" + json.dumps(r) + "\n```\n"]
    out.append("## Multi-GPU (`torchrun ... bench.py --gpus N`; earlier commits of this round where noted)\n")
    out.append("| N | value (pages/s) | e2e (pages/s) | H2D per GPU (GB/s) | corpus sha | exchange |")
    out.append("|---:|---:|---:|---:|---|---|")
    for n, f, note in ((1, "bench_default.json", ""), (2, "bench_n2.json", ""), (4, "bench_n4.json", ""), (8, "bench_n8.json", "")):
        path = os.path.join(G, f)
        if os.path.exists(path):
            x = last_json(path)
            out.append(f"| {n} | {x['value']:.0f} | {x['e2e']['value']:.0f}{note} | {x['e2e']['h2d_gb_per_s_per_gpu']:.1f} | "
                       f"{x['corpus']['hist_sha256']} | {x['corpus']['exchange']}, {x['corpus']['exchange_ms']:.3f} ms |")
            if n > 1:
                shutil.copy(path, os.path.join(P, f"r02_{f}"))
    p8 = os.path.join(G, "bench_n8.json")
    if os.path.exists(p8):
        x = last_json(p8)
        e8 = x["e2e"]
        agg = e8["h2d_gb_per_s_per_gpu"] * 8
        out.append(f"\nAt N = 8 the end-to-end leg is bound by the box's aggregate host->device rate (8 x {e8['h2d_gb_per_s_per_gpu']:.1f} = "
                   f"{agg:.0f} GB/s against 51-54 GB/s for one GPU alone): 17.3 MB per page x {e8['value']:.0f} pages/s = "
                   f"{17.3e-3 * e8['value']:.0f} GB/s.\n")
    out.append("## Command lines (`scripts/bench_cli_stage1.py`, `scripts/bench_cli_stages.py`)\n")
    c1 = json.loads(open(os.path.join(G, "cli_stage1.json")).read())
    cs = [json.loads(l) for l in open(os.path.join(G, "cli_stages.jsonl")) if l.startswith("{")]
    out.append("```json\n" + json.dumps(c1) + "\n" + "\n".join(json.dumps(c) for c in cs) + "\n```\n")
    out.append("## The reference's own `main()`s on the same kind of tree (`scripts/ref_cli_baseline.py --pages 2`, build container, 1 core)\n")
    out.append("```json\n" + "\n".join(json.dumps(x) for x in REF_CLI) + "\n```\n")
    out.append("| stage | reference `main()` (s/page, 1 core, incl. its PNG decode + visualisation JPEGs) | this repository's command line (s/page, B200 box) |")
    out.append("|---|---:|---:|")
    for ref, c in zip(REF_CLI, cs):
        out.append(f"| {ref['reference_stage']} | {ref['seconds_per_page']} | {c['seconds_device_json'] / c['pages']:.4f} |")
    out.append(f"| 1 (decode + tiles + files, trivial detector; the reference's stage 1 cannot run without weights) | — | "
               f"{c1['seconds_device_decode'] / c1['pages']:.4f} (host decode: {c1['seconds_host_decode'] / c1['pages']:.4f}) |")
    others = [("cfg2", "python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --sustained-seconds 0"),
              ("cfg5", "python bench.py --workload cfg5 --total-pages 20480 --no-cpu-baseline --no-e2e --sustained-seconds 0"),
              ("cfg4", "python scripts/bench_merge_stress.py")]
    if all(os.path.exists(os.path.join(G, f"r02_{n}.json")) for n, _ in others):  # scripts/gpu_r2_cfgs.sh
        out.append("## Other BASELINE configurations at the round's last build (one B200, `scripts/gpu_r2_cfgs.sh`)\n")
        got = {}
        for n, cmd in others:
            got[n] = last_json(os.path.join(G, f"r02_{n}.json"))
            out.append(f"`{cmd}`\n\n```json\n{json.dumps(got[n])[:1500]}\n```\n")
        out.append(f"* cfg2 (19 page sizes, heterogeneous batches): {got['cfg2']['value']:.0f} pages/s, {got['cfg2']['ms_per_step']:.3f} ms per step")
        out.append(f"* cfg5 (corpus streamed in 64-page batches with histogram accumulation, 20 480 pages): {got['cfg5']['value']:.0f} pages/s "
                   f"sustained, {got['cfg5']['ms_per_step']:.3f} ms per batch")
        out.append(f"* cfg4 (8 pages x 100 000 boxes, stage 3 only): {got['cfg4']['ms_per_launch']:.3f} ms per launch\n")
    open(os.path.join(P, "r02_configs.md"), "w").write("\n".join(out) + "\n")


def jpeg():
    out = ["# r02: device JPEG decoder (D1-D9) and the one-channel tiler\n",
           "Command: `python scripts/bench_jpeg.py 8` on one B200 (8 newspaper-like 8000x6000 grey pages, `synth.newspaper_page`;",
           "decode of the whole batch timed with CUDA events over 5 repetitions, input already in HBM; `equal_cv2` compares page 0",
           "with `cv2.imdecode`).  Launch list: `ncu --metrics gpu__time_duration.sum,... --clock-control none` of one decode of the",
           "quality-95 batch (`python scripts/bench_jpeg.py 8 once`; raw: `r02_jpeg_launches.csv`).\n",
           "| quality | restart interval | chunk bytes | MB/page | ms per 8 pages | pages/s | chunks | redone in round 1 | equal cv2 |",
           "|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
    lines = [json.loads(l) for l in open(os.path.join(G, "bench_jpeg.log")) if l.startswith("{")]
    for d in lines:
        if "chunk_bytes" in d:
            out.append(f"| {d['quality']} | {d['restart_interval']} | {d['chunk_bytes']} | {d['mb_per_page']} | {d['ms_per_batch']} | {d['pages_per_s']} | "
                       f"{d['chunks']} | {d['states_replaced_round1']} | {d['equal_cv2']} |")
    out.append("")
    for d in lines:
        if "h2d_gb_per_s" in d:
            out.append(f"* host->device copy of the quality-{d['quality']} files (pinned): {d['h2d_gb_per_s']} GB/s")
        if "tiler" in d:
            out.append(f"* tiler, {d['tiler']} pages ({d['pages']} pages of 8000x6000, 4x4 grid): {d['ms']} ms = {d['pages_per_s']} pages/s, "
                       f"{d['algorithmic_gb_per_s']} GB/s algorithmic")
    cv = [d["cv2_imdecode_s_per_page"] for d in lines if "cv2_imdecode_s_per_page" in d]
    out.append(f"* `cv2.imdecode` of one such file on a host core of the GPU box: {min(cv)}-{max(cv)} s")
    out += ["\n## Launch list (one decode of 8 pages, quality 95, 17.3 MB each; ncu serialises and runs cold)\n",
            "| kernel | device time (us) | threads active per instruction (of 32) | issue slots busy | warps active |", "|---|---:|---:|---:|---:|"]
    tot = 0.0
    for (_, k), m in launches(os.path.join(G, "jpeg_launches.csv")).items():
        if k.startswith("jpeg") or "jpeg_" in k:
            t = m.get("gpu__time_duration.sum", 0)
            tot += t
            out.append(f"| `{k}` | {t:.1f} | {m.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):.1f} | "
                       f"{m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} % | "
                       f"{m.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f} % |")
    out.append(f"| **total** | {tot:.1f} | | | |")
    shutil.copy(os.path.join(G, "jpeg_launches.csv"), os.path.join(P, "r02_jpeg_launches.csv"))
    open(os.path.join(P, "r02_jpeg.md"), "w").write("\n".join(out) + "\n")


def step():
    r01 = {"tile_letterbox_kernel": 1988.2, "edge_filter_kernel": 22.4, "nms_bin_kernel": 48.5, "nms_super_kernel": 5.0, "nms_cand_kernel": 13.7,
           "nms_mask_kernel": 93.3, "nms_resolve_kernel": 53.7, "nms_emit_kernel": 57.0, "class_flags_kernel": 4.7, "width_median_kernel": 67.8,
           "column_prep_kernel": 40.4, "column_density_kernel": 121.8, "column_peaks_kernel": 91.2}
    per = collections.OrderedDict()
    for (_, k), m in launches(os.path.join(G, "step_launches.csv")).items():
        if not k.startswith("void at::") and "synth" not in k:
            per[k] = m.get("gpu__time_duration.sum", 0)  # the last launch of each kernel wins: the last timed step
    tot = sum(per.values())
    out = ["# r02: ncu launch list of one step (`--metrics gpu__time_duration.sum --clock-control none`)\n",
           "Command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-overlap --no-corpus --sustained-seconds 0`",
           "(default cfg3 workload; ncu serialises kernels, so this is the single-stream picture — raw list: `r02_step_launches.csv`).\n",
           "| kernel (one step = 64 pages, last timed step) | device time (us) | share | r01 |", "|---|---:|---:|---:|"]
    for k, v in per.items():
        key = next((x for x in r01 if x in k), None)
        out.append(f"| `{k}` | {v:.1f} | {100 * v / tot:.1f} % | {r01.get(key, '')} |")
    out.append(f"| **total** | {tot:.1f} | | 2607.8 |")
    out.append("\n`column_peaks_kernel` 91 -> 22 us: distance selection as parallel rounds instead of the serial walk, prominence on one warp per peak.")
    out.append("`nms_mask_kernel` 93 -> 72 us is the predicate-chain prefilter of the end of round 1 (its r01 capture predated it).")
    shutil.copy(os.path.join(G, "step_launches.csv"), os.path.join(P, "r02_step_launches.csv"))
    open(os.path.join(P, "r02_launches.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    configs()
    jpeg()
    step()
    print("profiles/r02_configs.md, r02_jpeg.md, r02_launches.md written")
