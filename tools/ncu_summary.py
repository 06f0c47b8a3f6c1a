#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full --import-source on) into markdown: key raw metrics per
kernel launch, stall-reason shares and the executed-opcode histogram (from the SASS source page).
usage: python tools/ncu_summary.py report.ncu-rep [units_per_launch]   (units = data symbols)"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = rows[0]
    kn = hdr.index("Kernel Name")
    print("## raw metrics (%s)\n" % rep.split("/")[-1])
    for r in rows[2:]:
        print("### %s" % r[kn][:90])
        print("| metric | value | unit |\n|---|---|---|")
        for m in RAW:
            if m in hdr:
                print("| %s | %s | %s |" % (m, r[hdr.index(m)], rows[1][hdr.index(m)]))
        print()
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    h = None
    func = None
    seen = set()
    per = collections.OrderedDict()
    for r in src:
        if r and r[0] in ("Function Name", "Kernel Name"):
            func = r[1]
            continue
        if r and len(r) == 2 and r[0] == "File Path":
            continue
        if r and "Address" in r and "Source" in r:
            h = r
            continue
        if h is None or len(r) != len(h):
            continue
        a = r[h.index("Address")]
        if not a.startswith("0x") or (func, a) in seen:
            continue
        seen.add((func, a))
        d = per.setdefault(func, dict(ops=collections.Counter(), st=collections.Counter(), total=0.0))
        sass = r[h.index("Source")].strip()
        n = float(r[h.index("Instructions Executed")] or 0)
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
        op = m.group(2)
        base = op.split(".")[0]
        if base in ("LDS", "STS", "LDG", "STG", "LDL", "STL"):
            base = op
        d["ops"][base] += n
        d["total"] += n
        for i, name in enumerate(h):
            if name.startswith("stall_") and "Not Issued" not in name:
                try:
                    d["st"][name] += float(r[i])
                except ValueError:
                    pass
    for f, d in per.items():
        print("### executed warp-instructions: %s" % (f or "(profiled kernel)")[:90])
        print("total %.0f" % d["total"] + (" = %.1f per unit" % (d["total"] / units) if units else ""))
        print("| opcode | warp-instructions" + (" per unit |" if units else " |") + " share |\n|---|---|---|")
        for k, v in d["ops"].most_common(18):
            print("| %s | %.1f | %.1f %% |" % (k, v / units if units else v, 100 * v / d["total"]))
        # asynchronous-copy / mbarrier / tensor-memory opcodes, however rare: the evidence for (or against) TMA use
        special = [(k, v) for k, v in d["ops"].items() if re.match(r"(UBLKCP|UTMA|SYNCS|UTCBAR|UTCMMA|LDGSTS|BAR|FENCE|MEMBAR|ATOM|RED)", k)]
        if special:
            print("\nasync / barrier / atomic opcodes: " + ", ".join("%s %.0f" % (k, v) for k, v in sorted(special)))
        t = sum(d["st"].values()) or 1
        print("\nstall samples: " + ", ".join("%s %.1f %%" % (k.replace("stall_", ""), 100 * v / t) for k, v in d["st"].most_common(8)))
        print()


if __name__ == "__main__":
    main()
