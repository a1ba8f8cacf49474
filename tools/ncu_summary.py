"""Summarise an .ncu-rep (ncu --set full) as a small text table: per profiled launch the duration,
DRAM bytes, throughput percentages, occupancy, issue statistics and the executed SASS opcode mix.
Usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex] > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu wavefronts %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads/instr"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
]


def raw(rep, kern):
    cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
    if kern:
        cmd += ["--kernel-name", "regex:" + kern]
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    return rows[0], rows[1], rows[2:]


def opcode_mix(rep, kern):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
    if kern:
        cmd += ["--kernel-name", "regex:" + kern]
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    out, name, hdr, ops = [], None, None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            if ops:
                out.append((name, ops))
            name, ops = r[1], collections.Counter()
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and ops is not None and len(r) > 6 and r[0].startswith("0x"):
            parts = r[1].split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            ops[op.split(".")[0]] += int(r[hdr.index("Instructions Executed")])
    if ops:
        out.append((name, ops))
    return out


def main():
    rep = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else None
    hdr, units, rows = raw(rep, kern)
    ik = hdr.index("Kernel Name")
    print(f"# {rep}")
    for r in rows:
        print(f"\n== {r[ik]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {label:24s} {r[i]:>16s} {units[i]}")
    seen = set()
    for name, ops in opcode_mix(rep, kern):
        if name in seen:
            continue
        seen.add(name)
        tot = sum(ops.values())
        print(f"\n== executed SASS opcode mix: {name}  ({tot} warp instructions)")
        print("  " + ", ".join(f"{op} {100 * n / tot:.1f}%" for op, n in ops.most_common(18)))


if __name__ == "__main__":
    main()
