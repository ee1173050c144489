"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
tot, cnt = collections.Counter(), collections.Counter()
for x in rows:
  nm = re.sub(r"\(.*", "", x["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
  v = float(x["Metric Value"].replace(",", ""))
  v = v / 1e3 if x["Metric Unit"] == "ns" else (v * 1e3 if x["Metric Unit"] == "ms" else v)
  tot[nm] += v; cnt[nm] += 1
T = sum(tot.values())
print("launches %d, total %.1f ms (cold-cache, serialised: compare SHARES)" % (len(rows), T / 1e3))
for k, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
  print("%9.1f us %5.1f%%  n=%5d  avg %8.1f us  %s" % (v, 100 * v / T, cnt[k], v / cnt[k], k[:90]))
