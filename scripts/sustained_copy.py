"""Context for the sustained roofline: what a plain device copy reaches when it runs back to back for
seconds (power-capped clocks), next to its burst figure.  Prints one JSON line.

    python scripts/sustained_copy.py [seconds]
"""
import json
import subprocess
import sys
import threading
import time

import torch

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty_like(a)
a.fill_(1.0)
nbytes = 2 * a.numel() * a.element_size()


def timed(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for _ in range(3):
    timed(1)
burst = min(timed(1) for _ in range(10))
clocks = []
stop = False


def sample():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout.strip().split(",")
        try:
            clocks.append((int(out[0]), float(out[1])))
        except (ValueError, IndexError):
            pass
        time.sleep(0.05)


th = threading.Thread(target=sample)
th.start()
iters = max(10, int(secs * 1e3 / burst))
sustained = timed(iters)
stop = True
th.join()
mhz = sorted(c[0] for c in clocks)
print(json.dumps({"copy_bytes": nbytes, "burst_gbs": nbytes / burst / 1e6, "sustained_gbs": nbytes / sustained / 1e6,
                  "sustained_iters": iters, "sm_mhz_median": mhz[len(mhz) // 2] if mhz else None,
                  "power_w_max": max((c[1] for c in clocks), default=None)}))
