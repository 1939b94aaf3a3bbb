"""Summarise an .ncu-rep (read here, no GPU needed) into the text kept under profiles/.
    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/r01_x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        un = dict(zip(hdr, units))
        print(f"kernel: {rec.get('Kernel Name')}   grid {rec.get('Grid Size')} block {rec.get('Block Size')}")
        for k in KEYS:
            if k in rec:
                print(f"  {k:75s} {rec[k]:>18s} {un.get(k, '')}")
        stalls = []
        for h, v in rec.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(v.replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("  warp stall samples: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(stalls, reverse=True)[:7]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
