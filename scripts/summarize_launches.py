"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown share table."""
import collections
import csv
import re
import sys

src, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}[row["Metric Unit"]]
    tot[k] += v
    cnt[k] += 1
T = sum(tot.values())
print("# " + title + "\n")
print("%d launches, %.3f ms of kernel time (serialised, cold cache: compare shares, not absolutes).\n" % (sum(cnt.values()), T / 1e6))
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k, cnt[k], v / 1e6, 100 * v / T, v / cnt[k] / 1e3))
