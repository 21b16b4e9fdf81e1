#!/usr/bin/env python
"""Digest of an `ncu --set full --import-source on` report into a small text summary that can be committed under profiles/.

  python scripts/ncu_digest.py REPORT.ncu-rep [--launch I] [--top N] > profiles/rNN_<kernel>.md

Part 1: headline metrics per captured launch (raw page).  Part 2 (first launch unless --launch): the SASS-level picture
from the source page - dynamic instruction mix by opcode, SIMT efficiency (threads per executed instruction), stall-sample
totals by reason, and the instructions where the samples pile up.
"""
import argparse
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

RAW = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ. limit (regs), CTAs"), ("launch__occupancy_limit_shared_mem", "occ. limit (smem), CTAs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction (of 32)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_bytes.sum", "L1 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "stall no_instruction %"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_scoreboard %"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall math_pipe_throttle %"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "stall wait %"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stall short_scoreboard %"),
    ("smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "stall not_selected %"),
    ("smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "stall branch_resolving %"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--lib", help="the .so the kernel came from: adds a per-source-line table (nvdisasm -g line info)")
    a = ap.parse_args()

    rows = list(csv.reader(io.StringIO(ncu(["-i", a.report, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu digest of `{a.report.split('/')[-1]}` ({len(data)} launch(es) captured, --set full --clock-control none)\n")
    print("Kernel:", data[0][col["Kernel Name"]], "\n")
    print("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|" + "---|" * len(data))
    for key, label in RAW:
        if key in col:
            i = col[key]
            print(f"| {label} ({units[i]}) `{key}` | " + " | ".join(r[i] for r in data) + " |")

    # ---- source page of one launch
    out = ncu(["-i", a.report, "--page", "source", "--csv", "--print-source", "sass"])
    blocks = re.split(r'(?m)^"Kernel Name",', out)[1:]
    blk = blocks[min(a.launch, len(blocks) - 1)]
    lines = blk.split("\n", 1)[1]
    rows = list(csv.reader(io.StringIO(lines)))
    h = {n: i for i, n in enumerate(rows[0])}
    ins = []
    for r in rows[1:]:
        if len(r) < len(rows[0]):
            continue
        ins.append(r)
    tot_inst = sum(int(r[h["Instructions Executed"]]) for r in ins)
    tot_thr = sum(int(r[h["Predicated-On Thread Instructions Executed"]]) for r in ins)
    tot_samp = sum(int(r[h["# Samples"]]) for r in ins)
    print(f"\n## SASS-level digest of launch {a.launch}\n")
    print(f"warp instructions {tot_inst:,}; predicated-on threads / instruction {tot_thr / max(tot_inst, 1):.2f}; stall samples {tot_samp:,}\n")
    by_op = defaultdict(lambda: [0, 0, 0])
    for r in ins:
        src = r[h["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.rstrip(";")
        base = op.split(".")[0]
        if base in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "I2F", "MUFU", "ATOMG", "RED"):
            base = ".".join(op.split(".")[:2]) if base in ("I2F", "MUFU") else base
        e = by_op[base]
        e[0] += int(r[h["Instructions Executed"]]); e[1] += int(r[h["Predicated-On Thread Instructions Executed"]]); e[2] += int(r[h["# Samples"]])
    print("| opcode | warp instr | % of instr | threads/instr | % of stall samples |")
    print("|---|---|---|---|---|")
    for op, (n, t, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print(f"| {op} | {n:,} | {100 * n / tot_inst:.1f} | {t / max(n, 1):.1f} | {100 * s / max(tot_samp, 1):.1f} |")
    stalls = [n for n in rows[0] if n.startswith("stall_") and "Not Issued" not in n]
    st = {n: sum(int(r[h[n]] or 0) for r in ins) for n in stalls}
    print("\nStall samples by reason (all samples): " + ", ".join(f"{n[6:]} {100 * v / max(tot_samp, 1):.1f}%" for n, v in sorted(st.items(), key=lambda kv: -kv[1]) if v))
    if a.lib:
        kname = data[0][col["Kernel Name"]]
        head = re.split(r"[<(]", kname.replace("<unnamed>::", ""), 1)[0]          # "void ns::kernel" -> kernel
        mang = head.split()[-1].split("::")[-1]
        by_source_line(ins, h, a.lib, mang + ("ILb0" if "<0>" in data[0][col["Kernel Name"]] or "(bool)0" in data[0][col["Kernel Name"]] else ""), 40, tot_inst, tot_samp)
    print(f"\n| # | address | instruction | samples % | executed | threads/instr | top stall |")
    print("|---|---|---|---|---|---|---|")
    order = sorted(range(len(ins)), key=lambda i: -int(ins[i][h["# Samples"]]))[: a.top]
    for k, i in enumerate(order):
        r = ins[i]
        top = max(stalls, key=lambda n: int(r[h[n]] or 0))
        n = int(r[h["Instructions Executed"]])
        print(f"| {k} | +{i * 16:#x} | `{r[h['Source']].strip()}` | {100 * int(r[h['# Samples']]) / max(tot_samp, 1):.2f} | {n:,} | "
              f"{int(r[h['Predicated-On Thread Instructions Executed']]) / max(n, 1):.1f} | {top[6:]} |")


def line_map(lib, kernel_substr):
    """instruction offset -> 'file:line' for the first kernel whose mangled name contains kernel_substr"""
    import glob, os, tempfile
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    for cubin in glob.glob(d + "/*.cubin"):
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur, inside, m = None, False, {}
        for ln in dis.split("\n"):
            if ln.startswith(".text."):
                if inside and m:
                    return m
                inside = kernel_substr in ln
                continue
            if not inside:
                continue
            f = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if f:
                cur = f"{os.path.basename(f.group(1))}:{f.group(2)}"
                continue
            a = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if a and cur:
                m[int(a.group(1), 16)] = cur
        if inside and m:
            return m
    return {}


def by_source_line(ins, h, lib, mangled, top, tot_inst, tot_samp):
    m = line_map(lib, mangled)
    if not m:
        print("\n(no line info found for", mangled, ")")
        return
    agg = defaultdict(lambda: [0, 0, 0])
    for i, r in enumerate(ins):
        e = agg[m.get(i * 16, "?")]
        e[0] += int(r[h["Instructions Executed"]]); e[1] += int(r[h["Predicated-On Thread Instructions Executed"]]); e[2] += int(r[h["# Samples"]])
    print(f"\n### by source line (top {top} by executed warp instructions)\n")
    print("| source line | warp instr | % of instr | threads/instr | % of stall samples |")
    print("|---|---|---|---|---|")
    for k, (n, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {k} | {n:,} | {100 * n / tot_inst:.1f} | {t / max(n, 1):.1f} | {100 * s / max(tot_samp, 1):.1f} |")


if __name__ == "__main__":
    main()
