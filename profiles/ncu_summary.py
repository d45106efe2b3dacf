"""Summarise an .ncu-rep (run where ncu is installed, no GPU needed):
    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]
Prints the headline counters of the first captured kernel and, with --source N, the N hottest SASS/source lines."""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("kernel:", d.get("Kernel Name", ("?",))[0])
    for w in WANT:
        if w in d:
            print("%-82s %s %s" % (w, d[w][0], d[w][1]))


def source(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Samples" in c for c in r))
    hdr = rows[hi]
    si = hdr.index("Source")
    col = next(i for i, c in enumerate(hdr) if c.strip() == "# Samples" or c.strip() == "Warp Stall Sampling (All Samples)")
    ie = next((i for i, c in enumerate(hdr) if c.strip() == "Instructions Executed"), None)
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((float(r[col] or 0), float(r[ie] or 0) if ie is not None else 0.0, r[si]))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1.0
    toti = sum(d[1] for d in data) or 1.0
    print("total samples %d, total inst executed %d" % (tot, toti))
    for smp, ins, src in sorted(data, reverse=True)[:top]:
        print("%6.2f%% smp  %6.2f%% inst  %s" % (100 * smp / tot, 100 * ins / toti, src.strip()[:150]))


if __name__ == "__main__":
    rep = sys.argv[1]
    raw(rep)
    if "--source" in sys.argv:
        source(rep, int(sys.argv[sys.argv.index("--source") + 1]))
