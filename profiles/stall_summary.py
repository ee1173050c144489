"""Summarise the warp-stall sampling of one captured launch:  ncu -i REP --page source --csv --launch-skip K --launch-count 1
| python profiles/stall_summary.py [top]   (SASS view; the samples per instruction and the dominant stall reason)."""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 14
rows = list(csv.reader(sys.stdin))
name = rows[0][1]
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[col["# Samples"]].isdigit()]
tot = sum(int(r[col["# Samples"]]) for r in data)
by = {k: sum(int(r[col[k]]) for r in data) for k in reasons}
print("== %s: %d samples over %d SASS instructions ==" % (name[:70], tot, len(data)))
print("   by reason: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:7]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]]))[:top]
for i in order:
  r = data[i]
  n = int(r[col["# Samples"]])
  why = max(reasons, key=lambda k: int(r[col[k]]))
  print("  %7d %5.1f%%  #%4d  %-58s %s" % (n, 100.0 * n / tot, i, r[col["Source"]].strip()[:58], why[6:]))
