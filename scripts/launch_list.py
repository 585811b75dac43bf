"""ncu launch-list CSV (--metrics gpu__time_duration.sum) -> per-kernel table.
usage: python scripts/launch_list.py LAUNCHES.csv"""
import csv, io, re, sys, collections
txt = open(sys.argv[1]).read()
txt = txt[txt.index('"ID"'):]
agg = collections.OrderedDict()
for r in csv.DictReader(io.StringIO(txt)):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"void |<unnamed>::|\(RsArgs, RsArgsCold\)", "", r["Kernel Name"])
    name = re.sub(r"\(bool\)([01])", r"\1", name)
    key = (name, r["Grid Size"], r["Block Size"])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':80s} {'grid':>13s} {'block':>12s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for (name, g, b), (n, ms) in agg.items():
    print(f"{name[:80]:80s} {g:>13s} {b:>12s} {n:8d} {ms:10.2f} {100 * ms / tot:6.1f}%")
print(f"{'total':80s} {'':>13s} {'':>12s} {sum(v[0] for v in agg.values()):8d} {tot:10.2f}")
