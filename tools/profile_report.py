"""Turn the CSV of geneo_profile_dump() (GENEO_PROFILE=1) into a per-kernel table: launch sites are mapped to the kernel
named on that source line."""
import csv
import os
import re
import sys

root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "geneo4petsc_b200", "csrc")
rows = list(csv.DictReader(open(sys.argv[1])))
out = {}
for r in rows:
    fn, ln = r["site"].split(":")
    try:
        line = open(os.path.join(root, fn)).read().splitlines()[int(ln) - 1]
    except (OSError, IndexError):
        line = ""
    m = re.search(r"(k_\w+(?:<[^<>]*>)?)\s*<<<", line)
    name = m.group(1) if m else r["site"]
    a = out.setdefault(name, [0, 0.0, 0.0])
    a[0] += int(r["launches"]); a[1] += float(r["gpu_ms_until_next_launch"]); a[2] += float(r["host_ms_until_next_launch"])
tot = sum(a[1] for a in out.values())
print("%-34s %9s %12s %7s %12s %10s" % ("kernel", "launches", "gpu_ms", "share", "host_ms", "us/launch"))
for k, a in sorted(out.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %9d %12.2f %6.1f%% %12.2f %10.1f" % (k, a[0], a[1], 100 * a[1] / max(tot, 1e-9), a[2], 1e3 * a[1] / max(a[0], 1)))
print("%-34s %9d %12.2f" % ("total", sum(a[0] for a in out.values()), tot))
