"""ncu --metrics gpu__time_duration.sum CSV -> per-kernel share table (profiles/*.txt)."""
import csv, collections, re, sys
path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = row.get("Metric Unit", "")
    v *= {"usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9}.get(unit, 1.0)
    short = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    tot[short] += v
    cnt[short] += 1
T = sum(tot.values())
print("# " + title)
print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
print("total_ms %.2f launches %d" % (T / 1e6, sum(cnt.values())))
for k, v in tot.most_common(45):
    print("%9.3f ms %5.1f%% %5d  %s" % (v / 1e6, 100 * v / T, cnt[k], k))
