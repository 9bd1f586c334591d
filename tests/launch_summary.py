"""Helper (not a test): summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tests/launch_summary.py launches.csv [marker-substring]   -- aggregates the region between the last two
launches whose name contains the marker (default: philox_planes = one loss/train step)."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]
marker = sys.argv[2] if len(sys.argv) > 2 else "absmax_kernel"
need = sys.argv[3] if len(sys.argv) > 3 else None          # region must contain a kernel with this substring
marks = [i for i, (k, _) in enumerate(seq) if marker in k]
region = None
for a, b in reversed(list(zip(marks[:-1], marks[1:]))):
    if need is None or any(need in k for k, _ in seq[a:b]):
        region = seq[a:b]
        break
if region is None:
    region = seq
agg = collections.OrderedDict()
for k, v in region:
    k2 = re.sub(r"\(.*", "", k)
    k2 = re.sub(r"<unnamed>::", "", k2)[:110]
    agg.setdefault(k2, [0.0, 0])
    agg[k2][0] += v
    agg[k2][1] += 1
print(f"{len(region)} launches, {sum(v for _, v in region) / 1000:.1f} us")
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{v / 1000:9.1f} us x{n:3d}  {k}")
