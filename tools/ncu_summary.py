"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV-ish text for profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_atom.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_atom.sum.per_second", "lts__t_sectors_srcunit_tex_op_atom_dot_alu.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_atom_dot_alu_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_atom_dot_alu_lookup_miss.sum",
        "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_requests_srcunit_tex_op_write.sum",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    for r in rows[2:]:
        print("---")
        for w, i in idx:
            print(f"{w} = {r[i]} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
