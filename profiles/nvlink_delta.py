"""Difference of two `nvidia-smi nvlink -gt d` dumps -> NVLink bytes per GPU (and per optimizer step).
    python profiles/nvlink_delta.py before.txt after.txt <steps run in between>
Counters are cumulative KiB per link ("Data Tx" / "Data Rx")."""
import re
import sys
from collections import defaultdict


def parse(path):
    out = defaultdict(lambda: [0, 0])
    gpu = None
    for line in open(path):
        m = re.match(r"GPU (\d+):", line)
        if m:
            gpu = int(m.group(1))
            continue
        m = re.search(r"Link (\d+): Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None:
            out[gpu][0 if m.group(2) == "Tx" else 1] += int(m.group(3)) * 1024
    return out


a, b = parse(sys.argv[1]), parse(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
print("| GPU | NVLink Tx MB | NVLink Rx MB | Tx MB / step | Rx MB / step |")
print("|---|---|---|---|---|")
for g in sorted(b):
    tx, rx = b[g][0] - a[g][0], b[g][1] - a[g][1]
    print(f"| {g} | {tx / 1e6:.1f} | {rx / 1e6:.1f} | {tx / 1e6 / steps:.1f} | {rx / 1e6 / steps:.1f} |")
print(f"\n(all links of a GPU summed; {steps} optimizer steps incl. warm-up between the two dumps; the bench's set-up "
      "traffic -- IPC handle exchange, NCCL init -- is included and negligible)")
