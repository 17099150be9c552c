"""SASS evidence for the kernels of libpagegeom.so (runs anywhere cuobjdump exists; no GPU needed):
per kernel, the instruction total and the counts of the mnemonics that matter for the claims in DESIGN.md —
bulk TMA copies (UBLKCP), mbarrier traffic (SYNCS), cluster barriers (UCGABAR / CGA), the 2-way byte dot product
(IDP.2A), shared-memory loads, fp64 arithmetic — plus a short excerpt of the tiler's producer and consumer code.

    python scripts/sass_summary.py > profiles/r02_sass.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal_embeddings_b200", "libpagegeom.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "UCGABAR", "IDP.2A", "IDP.4A", "LDS", "STS", "LDG", "STG", "PRMT", "SHF", "IMAD", "DFMA",
         "DADD", "DMUL", "DSETP", "HFMA2", "HMUL2", "SHFL", "ATOM", "RED", "BAR", "MUFU", "HMMA", "UTCMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[name] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and name:
            kernels[name].append(m.group(1).strip())
    print("# r02: SASS summary of libpagegeom.so\n")
    print(f"`cuobjdump -sass` of the in-tree library; architectures in the fatbin: {', '.join(arch)} (no PTX for other targets).")
    print("Counts are static instruction counts per kernel (all template instances listed separately).\n")
    cols = ["UBLKCP", "UTMALDG", "SYNCS", "UCGABAR", "IDP.2A", "LDS", "PRMT", "SHF", "IMAD", "HFMA2", "DFMA", "DADD", "DSETP", "SHFL", "HMMA"]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for k, ins in kernels.items():
        cnt = collections.Counter()
        for i in ins:
            op = re.sub(r"^@!?U?P\d+\s+", "", i).split()[0]
            for w in WATCH:
                if op == w or op.startswith(w + ".") or op.startswith(w + "_"):
                    cnt[w] += 1
        short = re.sub(r"\(anonymous namespace\)::", "", k)
        short = re.sub(r"\(.*", "", short)
        print(f"| `{short[:70]}` | {len(ins)} | " + " | ".join(str(cnt[c]) if cnt[c] else "" for c in cols) + " |")
    # excerpt: the tiler's bulk copies and the consumer's inner pixel arithmetic
    key = next(k for k in kernels if "tile_letterbox_kernel<2, 3, 3>" in k)
    ins = kernels[key]
    print(f"\n## Excerpt: `{key}`\n")
    print("Producer (the bulk-TMA copies and their mbarrier bookkeeping):\n\n```")
    for idx, i in enumerate(ins):
        if "UBLKCP" in i or ("SYNCS" in i and ("ARRIVE" in i or "EXPECT" in i or "TRYWAIT" in i)):
            print(f"{idx:5d}  {i}")
    print("```\n\nConsumer (first 40 instructions after the first `IDP.2A`):\n\n```")
    first = next(idx for idx, i in enumerate(ins) if "IDP.2A" in i)
    for idx in range(max(0, first - 12), min(len(ins), first + 28)):
        print(f"{idx:5d}  {ins[idx]}")
    print("```")


if __name__ == "__main__":
    main()
