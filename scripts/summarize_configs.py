#!/usr/bin/env python
"""gpurun_out/configs/*.json(l) (scripts/gpu_configs_all.sh) -> profiles/<tag>_configs.md: one row per run plus
the raw JSON lines, so that every number quoted in DESIGN.md §6 can be traced to a measured line."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "configs")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def lines_of(path):
    out = []
    with open(path) as f:
        for ln in f:
            ln = ln.strip()
            if ln.startswith("{"):
                try:
                    out.append(json.loads(ln))
                except ValueError:
                    pass
    return out


def row(name, d):
    if "metric" in d:
        r = d.get("roofline") or {}
        c = d.get("corpus") or {}
        extra = []
        if r:
            extra.append(f"tiler {r.get('kernel_ms_per_launch', 0):.3f} ms, frac {r.get('frac', 0):.3f}"
                         + (f" (alone {r['isolated']['frac']:.3f})" if r.get("isolated") else ""))
        if d.get("e2e") and d["e2e"].get("h2d_bytes_per_step"):
            extra.append(f"e2e {d['e2e']['value']:.0f} pages/s")
        if c:
            extra.append(f"corpus {c['pages']} pages, median {c['median_plain_text_width_px']} px, sha {c['hist_sha256']}")
        if d.get("clocks"):
            extra.append(f"SM {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons'] or ''}".strip())
        return f"| `{name}` | {d['n_gpus']} | {d['value']:.0f} pages/s | {d['ms_per_step']:.3f} | {'; '.join(extra)} |"
    if "grids" in d:
        return (f"| `{name}` | 1 | {d['pages_per_s']:.0f} pages/s | {d['ms_per_launch']:.3f} | {d['grids']}: {d['tiles_per_page']} tiles/page, "
                f"{d['achieved_gbs']:.0f} GB/s algorithmic = {d['frac_of_measured_hbm_peak']:.3f} of peak |")
    if "seconds_device_json" in d:
        return (f"| `{name}` | 1 | {d['pages_per_s_device_json']:.1f} pages/s | {1e3 * d['seconds_device_json']:.0f} | {d['what']}: "
                f"{d['input_json_mb']:.1f} MB in, {d['output_json_mb']:.1f} MB out; CPython json {d['pages_per_s_cpython_json']:.1f} "
                f"pages/s ({d['speedup']:.1f}x slower), identical files |")
    if "what" in d:
        return (f"| `{name}` | 1 | {d['pages_per_s_device']:.0f} pages/s | {d['device_ms_per_step']:.3f} | {d['what']}: "
                f"{d['json_bytes_per_step'] / 1e6:.1f} MB of text, CPython "
                f"{d.get('cpython_json_dumps_ms_per_page', d.get('cpython_json_loads_ms_per_page', 0)):.1f} ms/page |")
    if "ms_per_page" in d:
        return f"| `{name}` | 1 | {d['pages_per_s']:.0f} pages/s | {d['ms_per_launch']:.3f} | {d.get('workload', '')} |"
    if "impl" in d:
        return f"| `{name}` | {d.get('n_gpus', 1)} | {d.get('value', 0):.2f} pages/s | {d.get('ms_per_step', 0):.0f} | reference arm (CPU) |"
    return f"| `{name}` | | | | {json.dumps(d)[:120]} |"


def main():
    files = sorted(glob.glob(os.path.join(SRC, "*.json")) + glob.glob(os.path.join(SRC, "*.jsonl")))
    out = [f"# {TAG}: every configuration, measured lines (`scripts/gpu_configs_all.sh` on B200 boxes)", "",
           "| run | GPUs | throughput | ms per step / launch | notes |", "|---|---:|---:|---:|---|"]
    raw = []
    for path in files:
        name = os.path.basename(path)
        for d in lines_of(path):
            out.append(row(name, d))
            raw.append((name, d))
    out += ["", "## Raw lines", ""]
    for name, d in raw:
        out += [f"`{name}`", "", "```json", json.dumps(d), "```", ""]
    dst = os.path.join(ROOT, "profiles", f"{TAG}_configs.md")
    with open(dst, "w") as f:
        f.write("\n".join(out))
    print(dst, len(raw), "lines")


if __name__ == "__main__":
    main()
