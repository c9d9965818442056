"""Text summary of an .ncu-rep (ncu -i ... --page raw --csv): the metrics DESIGN.md / profiles/ quote, per captured launch.
    python tools/ncu_summary.py gpurun_out/ncu/r02_conv64.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "sm__inst_executed_pipe_uniform", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "smsp__warp_issue_stalled", "smsp__average_warp", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "gpc__cycles_elapsed.max", "smsp__cycles_active.avg", "tensor", "tma", "sm__mio", "l1tex__data_pipe", "smsp__pcsamp_warps_issue_stalled"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        rows = [r for r in rows if len(r) > 10]
        if len(rows) < 3:
            print(path, "no data"); continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            print("=== %s :: %s" % (path.split("/")[-1], name[:110]))
            for i, h in enumerate(hdr):
                if any(h.startswith(k) or (k in ("tensor", "tma") and k in h.lower()) for k in KEYS):
                    v = vals[i]
                    if v in ("", "0", "n/a") and not h.startswith(("dram", "lts__t_sectors_op")):
                        continue
                    print("  %-88s %-14s %s" % (h, units[i], v))


if __name__ == "__main__":
    main()
