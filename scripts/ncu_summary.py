"""Summarise an .ncu-rep: per kernel the headline counters (raw page) and, for one kernel, where the instructions and
the stall samples sit in the SASS (source page).  Usage: python scripts/ncu_summary.py rep.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, io

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__grid_size", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(r[hdr.index("Kernel Name")].split("(")[0])
        for w in WANT:
            if w in hdr:
                print("    %-88s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))


def source(rep, kern, block=100):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, isrc, ii, iss = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    num = lambda x: int(x) if x.isdigit() else 0
    data, seen = [], set()
    for r in rows[2:]:
        if len(r) > ii and r[ia] != "Address" and r[ia] not in seen:
            seen.add(r[ia]); data.append((r[isrc], num(r[ii]), num(r[iss])))
    tot, tots = sum(d[1] for d in data), sum(d[2] for d in data)
    print("# %s: %d SASS instructions, %d executed, %d stall samples" % (kern, len(data), tot, tots))
    for k in range(0, len(data), block):
        blk = data[k:k + block]
        print("%5d  inst %5.1f%%  stall %5.1f%%  %s" % (k, 100 * sum(d[1] for d in blk) / max(tot, 1), 100 * sum(d[2] for d in blk) / max(tots, 1), blk[0][0].strip()[:70]))
    print("# memory / sync instructions: index, executed, stall samples")
    for k, d in enumerate(data):
        if any(t in d[0] for t in ("LDG", "MATCH", "BAR", "STG", "ATOM", "LDS", "STS", "SHFL", "RED")) and d[1] > 0:
            print("%5d %10d %6d  %s" % (k, d[1], d[2], d[0].strip()[:90]))


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[1], sys.argv[2])
