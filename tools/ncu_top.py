"""Top source lines by stall samples / executed instructions per profiled kernel of an .ncu-rep
(needs -lineinfo).  usage: ncu_top.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; hdr = None; data = []; block = 0; first = None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        fname = r[1].split('/')[-1]
        if first is None: first = fname
        if fname == first: block += 1
        continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Address": hdr = None; continue
    if hdr and r[0].isdigit() and len(r) > 5:
        ie = hdr.index("Instructions Executed"); isx = hdr.index("# Samples")
        try: data.append((block, int(r[isx]), int(r[ie]), fname, int(r[0]), r[1].strip()[:86]))
        except ValueError: pass
for b in sorted(set(x[0] for x in data)):
    v = [x for x in data if x[0] == b]
    ts = sum(x[1] for x in v); tot = sum(x[2] for x in v)
    print("== kernel", b, "samples", ts, "instr", tot)
    for _, s, n, fn, ln, src in sorted(v, key=lambda x: -x[1])[:top]:
        print(f"{100*s/max(ts,1):5.1f}% samp {100*n/max(tot,1):5.1f}% inst {fn[:12]:12s} L{ln:4d} {src}")
