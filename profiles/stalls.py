import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
items = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    n = int(r[ix["# Samples"]] or 0)
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
    items.append((n, r[ix["Source"]].strip()[:90], {s: int(r[ix[s]] or 0) for s in stalls}))
all_n = sum(n for n, _, _ in items)
print("total samples", all_n)
print("stall reasons:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(all_n, 1)) for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
for n, src, st in sorted(items, key=lambda t: -t[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 22]:
    top = max(st, key=st.get)
    print("%6.2f%%  %-90s  %s" % (100.0 * n / all_n, src, top[6:]))
