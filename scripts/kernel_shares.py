"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch log into per-kernel time shares."""
import csv, sys, collections, re, json
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.Counter(); cnt = collections.Counter()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    name = re.sub(r"<.*", "", name)[:70]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    elif unit in ("ms", "msecond"): v *= 1e6
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
out = [{"kernel": k, "launches": cnt[k], "total_us": round(v / 1e3, 1), "share": round(v / s, 4)} for k, v in tot.most_common(40)]
print(json.dumps({"total_us": round(s / 1e3, 1), "kernels": out}, indent=1))
