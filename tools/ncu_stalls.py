"""Summarise the source page of an ncu report: stall reasons (sampled) per kernel, and the top SASS lines.
usage: python tools/ncu_stalls.py <report.ncu-rep> <kernel-id 1-based> [top N]"""
import csv, subprocess, sys, io, collections
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:160])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
lines = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break  # the filter matched a second launch: keep the first
    if len(r) < len(hdr) - 1:
        continue
    n = int(r[ix["# Samples"]] or 0)
    per = {s: int(r[ix[s]] or 0) for s in stalls}
    for s, v in per.items():
        tot[s] += v
    lines.append((n, r[ix["Source"]], per, r[ix["Instructions Executed"]]))
allsum = sum(tot.values())
print("stall reasons (all samples):")
for s, v in tot.most_common(10):
    print(f"  {s:28s} {v:8d} {100.0 * v / max(allsum, 1):5.1f}%")
print("top SASS lines by samples:")
lines.sort(key=lambda x: -x[0])
for n, src, per, ie in lines[:top]:
    main = max(per, key=per.get)
    print(f"  {n:7d} {100.0 * n / max(allsum, 1):5.1f}%  {main:18s} {src[:110]}")
