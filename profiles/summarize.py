#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` export (and optionally the `--page source --csv` SASS export) into the
markdown committed under profiles/. Usage: summarize.py raw.csv [source.csv] > profiles/rNN_xxx.md"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak (this ncu's name in --set full)"),
    ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak, dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("FBSP.TriageCompute.dram__read_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM read throughput % of peak"),
    ("FBSP.TriageCompute.dram__write_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM write throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"### `{name[:110]}`\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in hdr:
                print(f"| {label} (`{k}`) | {r[hdr.index(k)]} {units[hdr.index(k)]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")))
                except ValueError:
                    pass
        if not stalls:  # newer ncu: stalled warps per issued instruction
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("| warp stall reasons (stalled warps per issued instruction) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]) + " |")
            print()
            continue
        stalls.sort(reverse=True)
        print("| top warp stall reasons (% of active warp cycles) | " + ", ".join(f"{n} {v:.1f}" for v, n in stalls[:6]) + " |")
        print()
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr = rows[1]
        ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        data = []
        for r in rows[2:]:
            try:
                data.append((int(r[ist] or 0), int(r[ia]), r[isrc]))
            except (ValueError, IndexError):
                pass
        tot = sum(d[0] for d in data) or 1
        print("### hottest SASS instructions by stall samples\n")
        print("| samples | share | executed | SASS |\n|---|---|---|---|")
        for s, n, src in sorted(data, reverse=True)[:12]:
            print(f"| {s} | {100 * s / tot:.1f}% | {n} | `{src[:90]}` |")


if __name__ == "__main__":
    main()
