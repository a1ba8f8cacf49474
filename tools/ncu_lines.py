"""Per-source-line share of warp-stall samples and executed warp instructions for one kernel of an
.ncu-rep captured with --import-source on (source page, CUDA + SASS), plus the hottest SASS
instructions.  usage: ncu_lines.py report.ncu-rep "<substring of the demangled kernel name>" [lines] [sass]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_fn=None; fname=""; hdr=None; lines=[]; sass=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": fname=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": cur_fn=r[1]; continue
    if r[0]=="Line No":
        hdr=r; ie=hdr.index("Instructions Executed"); sa=hdr.index("# Samples"); continue
    if hdr is None or kern not in (cur_fn or ""): continue
    if r[0]!="" :
        try: lines.append((int(r[sa]), int(r[ie]), fname, r[0], r[1].strip()[:100])); lastline=(fname,r[0])
        except ValueError: pass
    elif r[2].startswith("0x"):
        try: sass.append((int(r[sa]), int(r[ie]), lastline, r[3].strip()[:80]))
        except ValueError: pass
seen=set(); u=[]
for l in lines:
    k=(l[2],l[3])
    if k in seen: continue
    seen.add(k); u.append(l)
ts=sum(l[0] for l in u); ti=sum(l[1] for l in u)
print("samples",ts,"warp instr",ti)
for l in sorted(u,reverse=True)[:top]:
    print(f"{l[0]/ts*100:5.1f}% smp {l[1]/ti*100:5.1f}% ins  {l[2]}:{l[3]}  {l[4]}")
if len(sys.argv)>4:
    print("--- top SASS")
    for s in sorted(sass,reverse=True)[:int(sys.argv[4])]:
        print(f"{s[0]/ts*100:5.1f}% {s[1]:9d} {s[2][0]}:{s[2][1]}  {s[3]}")
