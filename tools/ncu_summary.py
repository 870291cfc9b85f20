#!/usr/bin/env python
"""Summarises an .ncu-rep (one `ncu --set full` capture) into the markdown kept under profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/rNN_<kernel>.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    print("# %s\n" % title)
    print("Source: `%s` (`ncu --set full --clock-control none --import-source on`, one launch per kernel).\n" % rep)
    for V in rows[2:]:
        d = dict(zip(H, zip(U, V)))
        print("## %s\n" % d.get("Kernel Name", ("", "?"))[1])
        print("| metric | value | unit |\n|---|---:|---|")
        for key, label in WANT:
            if key in d:
                print("| %s (`%s`) | %s | %s |" % (label, key, d[key][1], d[key][0]))
        print()
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(src)) if r]
    try:
        hi = [i for i, r in enumerate(rows) if r[0] == "Address"][0]
    except IndexError:
        return
    hdr = rows[hi]
    seen, cols = set(), []
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and h not in seen:
            seen.add(h); cols.append(i)
    tot = {}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) - 5:
            continue
        for i in cols:
            try:
                tot[hdr[i]] = tot.get(hdr[i], 0.0) + float(r[i] or 0)
            except ValueError:
                pass
    s = sum(tot.values()) or 1.0
    print("## warp stall samples (first kernel)\n\n| reason | share |\n|---|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        print("| %s | %.1f %% |" % (k, 100 * v / s))


if __name__ == "__main__":
    main()
