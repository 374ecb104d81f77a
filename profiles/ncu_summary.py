#!/usr/bin/env python3
"""Compact text summary of an .ncu-rep (read here, no GPU needed):

    python profiles/ncu_summary.py gpurun_out/sweep.ncu-rep [profiles/out.txt]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active",
]
STALL = "smsp__average_warps_issue_stalled_"
PCSAMP = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"== {d.get('Kernel Name', '?')}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        for h, u in zip(hdr, units):
            if h in KEYS:
                lines.append(f"   {h:90s} {d[h]:>18s} {u}")
        # warp-state samples of the PC sampler: share of all samples per stall reason (what DESIGN.md quotes)
        samp = {}
        for h in hdr:
            if h.startswith(PCSAMP) and not h.endswith("_not_issued"):
                try:
                    samp[h[len(PCSAMP):]] = float(d[h].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(samp.values())
        if tot > 0:
            top = sorted(samp.items(), key=lambda kv: -kv[1])[:8]
            lines.append("   stall reasons (% of warp-state samples): " + ", ".join(f"{n}={100 * v / tot:.1f}" for n, v in top if v > 0))
        ratios = []
        for h in hdr:
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    ratios.append((float(d[h].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        ratios.sort(reverse=True)
        if ratios:
            lines.append("   warps stalled per issued instruction: " + ", ".join(f"{n}={v:.2f}" for v, n in ratios[:8] if v > 0))
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
