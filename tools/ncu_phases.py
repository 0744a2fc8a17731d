"""Splits a kernel's SASS listing (ncu --page source --csv --print-source sass) into phases at BAR.SYNC and sums executed
warp instructions + stall samples per phase.  usage: python tools/ncu_phases.py src.csv <kernel-substring> [occurrence]"""
import csv, sys
path, key = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and key in r[1]]
i0 = starts[occ]
hdr = rows[i0 + 1]
ci, cs, csamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
phase, acc, tot = 0, [0, 0, 0], 0
out = []
for r in rows[i0 + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    try:
        n, smp = int(r[ci]), int(r[csamp])
    except ValueError:
        continue
    acc[0] += n; acc[1] += smp; acc[2] += 1; tot += n
    if "BAR.SYNC" in r[cs] or "EXIT" in r[cs] and n > 0:
        out.append((phase, *acc, r[cs].strip()[:30])); phase += 1; acc = [0, 0, 0]
out.append((phase, *acc, "end"))
for p, n, smp, k, what in out:
    if n: print(f"phase {p:2d}: {n:12d} warp-instr ({100*n/tot:5.1f}%)  samples {smp:7d}  sass {k:4d}  ends at {what}")
print("total", tot)
