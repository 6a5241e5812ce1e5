"""Per-segment warp-instruction counts of one kernel from an ncu report (SASS page): consecutive SASS instructions with the
same execution count are one segment (a loop body / phase).  Usage: python tools/ncu_segments.py report.ncu-rep kernel [--sass lo hi]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = next(x for x in r if x and x[0] == "Address")
ie, isrc, ith = h.index("Instructions Executed"), h.index("Source"), h.index("Thread Instructions Executed")
rows = [x for x in r if len(x) > ie and x[ie].isdigit()]
first = rows[0][0]
n1 = next((i for i in range(1, len(rows)) if rows[i][0] == first), len(rows))   # a report may hold the kernel more than once
rows = rows[:n1]
tot = sum(int(x[ie]) for x in rows)
print("total warp instructions", tot, "SASS instructions", len(rows))
if "--sass" in sys.argv:
    lo, hi = int(sys.argv[sys.argv.index("--sass") + 1]), int(sys.argv[sys.argv.index("--sass") + 2])
    for i in range(lo, hi):
        print(i, rows[i][ie].rjust(9), rows[i][ith].rjust(10), rows[i][isrc].strip())
    sys.exit(0)
cur, start, acc, segs = None, 0, 0, []
for i, x in enumerate(rows):
    n = int(x[ie])
    if cur is None or abs(n - cur) > 0.02 * max(cur, 1):
        if cur is not None:
            segs.append((start, i - 1, cur, acc))
        cur, start, acc = n, i, 0
    acc += n
segs.append((start, len(rows) - 1, cur, acc))
for s in segs:
    if s[3] > 0.004 * tot:
        ops = {}
        thr = sum(int(x[ith]) for x in rows[s[0]:s[1] + 1])
        for x in rows[s[0]:s[1] + 1]:
            t = x[isrc].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda t: -t[1])[:7]
        print(f"{s[0]:5d}-{s[1]:5d} n={s[1]-s[0]+1:4d} exec/inst={s[2]:9d} share={100*s[3]/tot:5.1f}% lanes={thr/max(s[3],1):4.1f}  {top}")
