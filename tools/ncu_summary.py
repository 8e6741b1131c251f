"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small markdown table for profiles/."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "sm__cycles_active.avg", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def main(rep: str) -> None:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print(f"### `{d['Kernel Name'][1][:140]}`\n\ngrid {d['Grid Size'][1]} block {d['Block Size'][1]}\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            for h in hdr:
                if h.endswith(k):
                    print(f"| {k} | {d[h][1]} | {d[h][0]} |")
                    break
        print()


if __name__ == "__main__":
    main(sys.argv[1])
