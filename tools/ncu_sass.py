"""SASS-level stats of one kernel from an .ncu-rep: lane utilisation, opcode mix, top stalled instructions."""
import csv, subprocess, sys, io
from collections import Counter
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hdr_i]
end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
idx = {h: i for i, h in enumerate(hdr)}
ie, te, sa, src = idx["Instructions Executed"], idx["Thread Instructions Executed"], idx["# Samples"], idx["Source"]
tot = tt = ts = 0
ops, opsamp, recs = Counter(), Counter(), []
for k, r in enumerate(rows[hdr_i + 1:end]):
    try:
        n, t, s = int(r[ie]), int(r[te]), int(r[sa])
    except (ValueError, IndexError):
        continue
    tot += n; tt += t; ts += s
    parts = r[src].split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    ops[op] += n; opsamp[op] += s
    recs.append((s, n, k, r[src].strip()[:70]))
print(f"warp-instr {tot}  thread-instr {tt}  lanes/instr {tt/max(tot,1):.1f}  samples {ts}")
print("opcodes (inst% / stall-sample%):", ", ".join(f"{o} {100*n/tot:.1f}/{100*opsamp[o]/max(ts,1):.1f}" for o, n in ops.most_common(18)))
print("top stalled instructions:")
for s, n, k, txt in sorted(recs, reverse=True)[:top]:
    print(f"  {100*s/max(ts,1):5.1f}% samp  {100*n/tot:4.1f}% inst  #{k:4d} {txt}")
