"""Per-source-line share of executed warp instructions + stall samples for one kernel of an .ncu-rep
(all source files of the first profiled launch)."""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = []
fname = ""
seen_funcs = 0
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        idx = {h: i for i, h in enumerate(hdr)}
        ie, sa = idx["Instructions Executed"], idx["# Samples"]
        continue
    if hdr is None or len(r) <= ie or r[0] == "":
        continue
    try:
        lines.append((int(r[ie]), int(r[sa]), fname, r[0], r[1].strip()[:95]))
    except ValueError:
        pass
# several launches repeat the tables: keep the first occurrence of each (file, line)
seen, uniq = set(), []
for l in lines:
    k = (l[2], l[3])
    if k in seen:
        continue
    seen.add(k); uniq.append(l)
tot = sum(l[0] for l in uniq); ts = sum(l[1] for l in uniq)
print("total warp-instr", tot, "samples", ts)
for n, s, fn, ln, src in sorted(uniq, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/max(ts,1):5.1f}% samp  {fn[:18]:18s} L{ln:>4s} {src}")
