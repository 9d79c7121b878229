"""Text summary of an .ncu-rep for profiles/: per kernel the raw-page metrics that matter on this path (time, DRAM bytes, issue and
occupancy figures, instruction count) and, when the report holds source counters, the SASS instructions with the most warp-stall
samples and what they waited on.  Usage: python tools/ncu_summary.py report.ncu-rep [title] > profiles/xyz.txt  (ncu must be on PATH)"""
import csv, io, subprocess, sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
       "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
       "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def page(rep, name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
        print()
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        kernels.append(name)
        print(f"kernel: {name}   (launch id {r[hdr.index('ID')]})")
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:92s} {r[i]} {units[i]}")
        print()
    src = page(rep, "source")
    blocks = src.split('"Kernel Name",')
    for blk in blocks[1:]:
        lines = blk.splitlines()
        kname = lines[0].strip('",').split("(")[0]
        rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        if len(rd) < 2:
            continue
        h = rd[0]
        try:
            iS, iSamp, iIns = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        except ValueError:
            continue
        st = {k: h.index(k) for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_barrier", "stall_lg", "stall_mio") if k in h}
        body = [r for r in rd[1:] if len(r) == len(h)]
        tot = sum(int(r[iSamp] or 0) for r in body) or 1
        totI = sum(int(r[iIns] or 0) for r in body) or 1
        print(f"top SASS instructions by warp-stall samples: {kname}  (samples {tot}, warp instructions {totI})")
        for r in sorted(body, key=lambda r: -int(r[iSamp] or 0))[:14]:
            s = int(r[iSamp] or 0)
            if s == 0:
                break
            why = ", ".join(f"{k[6:]} {100 * int(r[i] or 0) // max(s, 1)}%" for k, i in st.items() if int(r[i] or 0) * 10 >= s)
            print(f"  {100.0 * s / tot:5.1f}% smp  {100.0 * int(r[iIns] or 0) / totI:5.1f}% ins  {r[iS].strip():60s} [{why}]")
        print()


if __name__ == "__main__":
    main()
