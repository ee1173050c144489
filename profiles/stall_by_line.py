"""Warp-stall samples per CUDA source line of one captured launch:
ncu -i REP --page source --print-source cuda,sass --csv --launch-skip K --launch-count 1 | python profiles/stall_by_line.py [top]"""
import collections, csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rows = list(csv.reader(sys.stdin))
hdr = next(r for r in rows if "# Samples" in r)
isamp, iline, isrc = hdr.index("# Samples"), 0, 1
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); why = collections.defaultdict(collections.Counter); text = {}
fname = ""
for r in rows:
  if r and r[0] == "File Name":
    fname = r[1].split("/")[-1]
  if len(r) < len(hdr) or not r[isamp].isdigit() or not r[iline].isdigit():
    continue
  key = (fname, int(r[iline]))
  tot[key] += int(r[isamp]); text.setdefault(key, r[isrc].strip())
  for i, h in reasons:
    why[key][h[6:]] += int(r[i])
T = sum(tot.values())
print("total samples", T)
for key, n in tot.most_common(top):
  w = ", ".join("%s %d%%" % (k, 100 * v / max(n, 1)) for k, v in why[key].most_common(2))
  print("%7d %5.1f%%  %s:%d  %-90s %s" % (n, 100.0 * n / T, key[0], key[1], text[key][:90], w))
