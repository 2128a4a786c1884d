"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and grid size."""
import collections, csv, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); det = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    val = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    ms = val / 1e6 if unit.startswith("n") else (val / 1e3 if unit.startswith("u") else val)
    name = re.sub(r"\(.*", "", row["Kernel Name"]); name = re.sub(r"^void ", "", name)
    blk = row.get("Block Size", "")
    agg[name][0] += 1; agg[name][1] += ms
    det[(name, row["Grid Size"], blk)][0] += 1; det[(name, row["Grid Size"], blk)][1] += ms
tot = sum(v[1] for v in agg.values())
print("total %.2f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%5d %10.2f ms %5.1f%%" % (k[:60], v[0], v[1], 100 * v[1] / tot))
print()
for k, v in sorted(det.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-34s grid=%-16s blk=%-14s n=%5d tot=%9.3f ms avg=%9.1f us" % (k[0][:34], k[1], k[2], v[0], v[1], v[1] / v[0] * 1000))
