"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the run)."""
import csv, sys, collections

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(r[ui], 1e-6)
    tot[r[ki]] += v
    cnt[r[ki]] += 1
allms = sum(tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes)")
print(f"{'kernel':92s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k, v in tot.most_common():
    print(f"{k[:92]:92s} {cnt[k]:8d} {v:10.3f} {v / allms * 100:6.1f}%")
