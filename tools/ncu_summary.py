"""Print the headline metrics of every kernel in an .ncu-rep (read offline with `ncu -i`).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "occ lim smem"),
    ("launch__occupancy_limit_registers", "occ lim regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__cycles_elapsed.max", "sm cycles"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"== {path}")
        for r in rows[2:]:
            print(r[hdr.index("Kernel Name")][:110])
            for key, label in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"    {label:22s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
