"""Summarise an .ncu-rep (read here, no GPU needed) into a small tracked text file under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_kernel.md "title"
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", "", f"source: `ncu --set full --clock-control none` capture `{rep}` (read with `ncu -i ... --page raw --csv`)", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## {name[:110]}", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                lines.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        st = [(h, r[i]) for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]

        def f(v):
            try:
                return float(v.replace(",", ""))
            except ValueError:
                return 0.0
        lines += ["", "warp stall reasons (stalled warps per issue-active cycle, top 8):", ""]
        for h, v in sorted(st, key=lambda t: -f(t[1]))[:8]:
            lines.append(f"- {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v}")
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    main()
